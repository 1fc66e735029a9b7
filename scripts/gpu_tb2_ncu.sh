#!/bin/bash
# one ncu --set full capture of the double-sweep kernel (after a plain run of the same command)
mkdir -p gpurun_out
V=${1:-20}
CMD="python bench.py --steps 1 --warmup 3 --sub-iters 4 --no-e2e --no-cpu-baseline --no-others --variant $V"
$CMD > gpurun_out/tb2_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi3d_tb2 -s 2 -c 1 -f -o gpurun_out/r1_prof_3d7_768_tb2_v$V $CMD > gpurun_out/tb2_ncu.log 2>&1
echo rc=$?; tail -3 gpurun_out/tb2_ncu.log
