#!/bin/bash
# one ncu --set full capture of a 2D temporally blocked kernel (after a plain run of the same command)
# usage: gpu_tb2_2d_ncu.sh [variant: 30 = double sweep (default), 32..37 = K-sweep kernel]
mkdir -p gpurun_out
V=${1:-30}
CMD="python bench.py --workload 2d5_8192 --steps 1 --warmup 3 --sub-iters 12 --no-e2e --no-cpu-baseline --no-others --variant $V"
$CMD > gpurun_out/tb2_2d_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi2d_tb -s 2 -c 1 -f -o gpurun_out/prof_2d5_8192_v$V $CMD > gpurun_out/tb2_2d_ncu.log 2>&1
echo rc=$?; tail -2 gpurun_out/tb2_2d_ncu.log
