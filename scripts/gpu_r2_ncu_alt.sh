#!/bin/bash
# full ncu capture of the three-sweep kernel of an alternative build: gpu_r2_ncu_alt.sh <alt name> <tag>
mkdir -p gpurun_out
cp rnmf_b200/librnmf_b200.so /tmp/base.so
cp rnmf_b200/csrc/_alt/$1/librnmf_b200.so rnmf_b200/librnmf_b200.so
CMD3="python bench.py --steps 1 --warmup 3 --sub-iters 7 --no-e2e --no-cpu-baseline --no-others"
$CMD3 > gpurun_out/tb3$2_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi3d_tb3 -s 2 -c 1 -f -o gpurun_out/r2_prof_3d7_768_tb3$2 $CMD3 > gpurun_out/tb3$2_ncu_full.log 2>&1
echo "ncu $1 rc=$?"
cp /tmp/base.so rnmf_b200/librnmf_b200.so
