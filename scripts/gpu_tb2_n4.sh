#!/bin/bash
# double-sweep kernel on 4 slabs (interior ranks have both neighbours): parity, then a short bench
mkdir -p gpurun_out
RNMF_DIST_MODES=peer_tb2 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 tests/dist_worker.py nccl > gpurun_out/tb2_n4_parity.log 2>&1
echo "parity rc=$?"; grep -v "^W\|^\[W\|Warning\|^\*\*\*\|OMP_NUM" gpurun_out/tb2_n4_parity.log | tail -6
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 4 --steps 2 --warmup 2 --sub-iters 150 --no-cpu-baseline --no-others --no-e2e 2>gpurun_out/tb2_n4.err | tail -1 > gpurun_out/tb2_n4.json
python -c "import json; d=json.load(open('gpurun_out/tb2_n4.json')); print('N=4', round(d['value'],1), round(d['ms_per_step'],2), d['roofline']['double_sweep_launches'])"
