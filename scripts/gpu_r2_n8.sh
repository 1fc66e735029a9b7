#!/bin/bash
# round 2: 8 x B200: the multi-GPU parity worker at world 8 (deep halo, 3D and 2D), then the default bench at N = 8
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/r2n8_pytest.log 2>&1; echo "dist pytest rc=$?"; tail -3 gpurun_out/r2n8_pytest.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 3 > gpurun_out/r2n8_bench_n8.json 2> gpurun_out/r2n8_bench_n8.err ) 2>&1 | grep real
tail -c 400 gpurun_out/r2n8_bench_n8.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2n8_bench_n8.json").read().strip().splitlines()[-1])
    print(8, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 1), d["e2e"] and round(d["e2e"]["ms_per_step"], 1), "roofline", {k: d["roofline"].get(k) for k in ("frac", "avg_launch_ms", "avg_launch_ms_from_step")}, d["clocks"])
    print("   others", json.dumps(d["other_configs"])[:1500])
    print("   checks", d["checks"])
except Exception as e:
    print("failed", e)
PY
