#!/bin/bash
# round 2, final library: 4 x B200: the multi-GPU parity worker at world 4, then the default bench at N = 1 and N = 4 on the same box
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q > gpurun_out/r2n4_pytest.log 2>&1; echo "dist pytest rc=$?"; tail -3 gpurun_out/r2n4_pytest.log
( time timeout 600 python bench.py --steps 3 --no-cpu-baseline > gpurun_out/r2n4_bench_n1.json 2> gpurun_out/r2n4_bench_n1.err ) 2>&1 | grep real
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 4 --steps 3 > gpurun_out/r2n4_bench_n4.json 2> gpurun_out/r2n4_bench_n4.err ) 2>&1 | grep real
python - <<'PY'
import json
for n in (1, 4):
    try:
        d = json.loads(open(f"gpurun_out/r2n4_bench_n{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 1), {k: d["roofline"].get(k) for k in ("frac", "avg_launch_ms")}, d["clocks"]["sm_mhz"])
        print("   others", {k: (round(v["glups"], 1), v.get("abs_sum_psi")) for k, v in d["other_configs"].items() if "glups" in v})
    except Exception as e:
        print(n, "failed", e)
PY
