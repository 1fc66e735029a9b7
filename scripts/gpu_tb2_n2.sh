#!/bin/bash
# double-sweep kernel across two slabs: parity, then short weak-scaling benches
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/dist_worker.py nccl > gpurun_out/tb2_n2_parity.log 2>&1
echo "parity rc=$?"; grep -v "^W\|^\[W\|Warning" gpurun_out/tb2_n2_parity.log | tail -6
S="--steps 2 --warmup 2 --sub-iters 200 --no-cpu-baseline --no-others --no-e2e"
port=29580
for v in ${VARIANTS:-20}; do
  port=$((port+1))
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 $S --variant $v 2>gpurun_out/tb2_n2_v$v.err | tail -1 > gpurun_out/tb2_n2_v$v.json
  python -c "import json; d=json.load(open('gpurun_out/tb2_n2_v$v.json')); print('N=2 variant $v', round(d['value'],1), round(d['ms_per_step'],2))"
done
