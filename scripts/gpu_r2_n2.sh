#!/bin/bash
# round 2: whole GPU suite on a 2-GPU box (the dist parity test runs with world 2), default bench at N = 1 and N = 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2n2_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2n2_pytest.log
( time timeout 600 python bench.py --steps 3 > gpurun_out/r2n2_bench_n1.json 2> gpurun_out/r2n2_bench_n1.err ) 2>&1 | grep real
tail -c 600 gpurun_out/r2n2_bench_n1.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 > gpurun_out/r2n2_bench_n2.json 2> gpurun_out/r2n2_bench_n2.err ) 2>&1 | grep real
tail -c 600 gpurun_out/r2n2_bench_n2.err
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads(open(f"gpurun_out/r2n2_bench_n{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 1), "e2e", d["e2e"] and round(d["e2e"]["value"], 1), "roofline", {k: d["roofline"].get(k) for k in ("frac", "frac_of_k_pass_ceiling", "avg_launch_ms", "avg_launch_ms_from_step", "dram_frac")}, d["clocks"])
        print("   others", json.dumps(d["other_configs"])[:1500])
        print("   checks", d["checks"])
    except Exception as e:
        print(n, "failed", e)
PY
