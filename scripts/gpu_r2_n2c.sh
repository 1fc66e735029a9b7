#!/bin/bash
# final library at 2 GPUs: the multi-GPU parity worker (world 2) and the default bench at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/r2n2c_pytest.log 2>&1; echo "dist pytest rc=$?"; tail -2 gpurun_out/r2n2c_pytest.log
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 3 > gpurun_out/r2n2c_bench_n2.json 2> gpurun_out/r2n2c_bench_n2.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2n2c_bench_n2.json").read().strip().splitlines()[-1])
print(2, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 1), {k: d["roofline"].get(k) for k in ("frac", "avg_launch_ms")})
print("   others", {k: (round(v["glups"], 1), v.get("abs_sum_psi")) for k, v in d["other_configs"].items() if "glups" in v})
PY
