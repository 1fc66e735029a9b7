#!/bin/bash
# chunk lengths of the three-sweep kernel + the other 3D shapes, against the two-sweep kernel on the same box
mkdir -p gpurun_out
show() { python -c "
import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[2], round(d['value'],1), 'GLUP/s', d['clocks']['sm_mhz'], 'MHz kernel_ms', round(d['roofline'].get('avg_launch_ms'),3))" $1 "$2" || tail -3 ${1%.json}.err; }
A="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-others"
timeout 200 python bench.py $A --variant 31 > gpurun_out/ch3_k2.json 2> gpurun_out/ch3_k2.err; show gpurun_out/ch3_k2.json "768 K2"
for c in 0 32 48 64 128; do
  timeout 200 python bench.py $A --chunk $c > gpurun_out/ch3_c$c.json 2> gpurun_out/ch3_c$c.err; show gpurun_out/ch3_c$c.json "768 K3 chunk $c"
done
for w in 3d7_512 3d7_1024; do
  for v in 31 auto; do
    timeout 300 python bench.py $A --workload $w --variant $v > gpurun_out/ch3_${w}_$v.json 2> gpurun_out/ch3_${w}_$v.err; show gpurun_out/ch3_${w}_$v.json "$w variant $v"
  done
done
