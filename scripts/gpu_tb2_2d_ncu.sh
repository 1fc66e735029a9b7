#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --workload 2d5_8192 --steps 1 --warmup 3 --sub-iters 4 --no-e2e --no-cpu-baseline --no-others"
$CMD > gpurun_out/tb2_2d_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi2d_tb2 -s 2 -c 1 -f -o gpurun_out/r1_prof_2d5_8192_tb2 $CMD > gpurun_out/tb2_2d_ncu.log 2>&1
echo rc=$?; tail -2 gpurun_out/tb2_2d_ncu.log
