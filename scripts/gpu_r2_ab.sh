#!/bin/bash
# A/B of library builds on one box: the default in-tree build against rnmf_b200/csrc/_alt/<name>/librnmf_b200.so (scripts/build_alt.sh)
# usage: gpu_r2_ab.sh "<bench args>" name1 name2 ...   ("base" = the in-tree build)
mkdir -p gpurun_out
ARGS=$1; shift
cp rnmf_b200/librnmf_b200.so /tmp/base.so
for rep in 1 2; do
for name in "$@"; do
  if [ "$name" = base ]; then cp /tmp/base.so rnmf_b200/librnmf_b200.so; else cp rnmf_b200/csrc/_alt/$name/librnmf_b200.so rnmf_b200/librnmf_b200.so; fi
  timeout 200 python bench.py $ARGS > gpurun_out/r2ab_${name}_$rep.json 2> gpurun_out/r2ab_${name}_$rep.err
  python -c "
import json; d=json.load(open('gpurun_out/r2ab_${name}_$rep.json')); print('$name', $rep, round(d['value'],1), 'GLUP/s', round(d['ms_per_step'],2), 'ms/step', d['clocks']['sm_mhz'], d['clocks']['power_w_max'], 'kernel_ms', d['roofline'].get('avg_launch_ms'))" || tail -3 gpurun_out/r2ab_${name}_$rep.err
done
done
cp /tmp/base.so rnmf_b200/librnmf_b200.so
