#!/bin/bash
# one full ncu capture of the three-sweep kernel at 768^3 (after a plain run of the same command)
mkdir -p gpurun_out
TAG=${1:-c}
CMD3="python bench.py --steps 1 --warmup 3 --sub-iters 7 --no-e2e --no-cpu-baseline --no-others"
$CMD3 > gpurun_out/tb3${TAG}_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi3d_tb3 -s 2 -c 1 -f -o gpurun_out/r2_prof_3d7_768_tb3${TAG} $CMD3 > gpurun_out/tb3${TAG}_ncu_full.log 2>&1
echo "ncu rc=$?"
