#!/bin/bash
# round 2: 2 x B200: dist parity (incl. the host-frame call across slabs), then the default bench at N = 2 (e2e through rnmf_jacobi_sweeps_host)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/r2n2b_pytest.log 2>&1; echo "dist pytest rc=$?"; tail -12 gpurun_out/r2n2b_pytest.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --no-others > gpurun_out/r2n2b_bench_n2.json 2> gpurun_out/r2n2b_bench_n2.err ) 2>&1 | grep real
tail -c 600 gpurun_out/r2n2b_bench_n2.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2n2b_bench_n2.json").read().strip().splitlines()[-1])
    print(2, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 1), "e2e", d["e2e"])
except Exception as e:
    print("failed", e)
PY
