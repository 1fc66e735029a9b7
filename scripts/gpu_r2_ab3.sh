#!/bin/bash
# A/B of builds of the three-sweep kernel on one box, with the two-sweep kernel (variant 31) of the in-tree build as the box's yardstick
# usage: gpu_r2_ab3.sh name1 name2 ...   ("base" = the in-tree build; others: rnmf_b200/csrc/_alt/<name>/librnmf_b200.so)
mkdir -p gpurun_out
ARGS="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-others"
cp rnmf_b200/librnmf_b200.so /tmp/base.so
show() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['value'],1), 'GLUP/s', d['clocks']['sm_mhz'], 'MHz', round(d['clocks']['power_w_max']), 'W kernel_ms', round(d['roofline'].get('avg_launch_ms'),3), 'per sweep', round(d['roofline'].get('avg_launch_ms')/d['roofline'].get('sweeps_per_launch'),3))" $1 $2 || tail -3 ${1%.json}.err; }
for rep in 1 2; do
  timeout 200 python bench.py $ARGS --variant 31 > gpurun_out/ab3_k2_$rep.json 2> gpurun_out/ab3_k2_$rep.err; show gpurun_out/ab3_k2_$rep.json "K2-yardstick"
  for name in "$@"; do
    if [ "$name" = base ]; then cp /tmp/base.so rnmf_b200/librnmf_b200.so; else cp rnmf_b200/csrc/_alt/$name/librnmf_b200.so rnmf_b200/librnmf_b200.so; fi
    timeout 200 python bench.py $ARGS > gpurun_out/ab3_${name}_$rep.json 2> gpurun_out/ab3_${name}_$rep.err; show gpurun_out/ab3_${name}_$rep.json $name
  done
  cp /tmp/base.so rnmf_b200/librnmf_b200.so
done
