#!/bin/bash
# round 2: 8 x B200: the default bench at N = 8 with the end-to-end leg through rnmf_jacobi_sweeps_host (time-skewed start across slabs)
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 > gpurun_out/r2n8b_bench_n8.json 2> gpurun_out/r2n8b_bench_n8.err ) 2>&1 | grep real
tail -c 300 gpurun_out/r2n8b_bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2n8b_bench_n8.json").read().strip().splitlines()[-1])
    print(8, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 1), d["e2e"] and round(d["e2e"]["ms_per_step"], 1))
    print("   others", json.dumps(d["other_configs"])[:1200])
except Exception as e:
    print("failed", e)
PY
