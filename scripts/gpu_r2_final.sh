#!/bin/bash
# end of round 2: whole GPU suite, smoke(), the default bench line and the --impl reference line of the same box (1 GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2f_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2f_smoke.log
( time timeout 900 python bench.py > gpurun_out/r2f_bench_default_n1.json 2> gpurun_out/r2f_bench_default_n1.err ) 2>&1 | grep real
( time timeout 600 python bench.py --impl reference > gpurun_out/r2f_bench_reference_n1.json 2> gpurun_out/r2f_bench_reference_n1.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2f_bench_default_n1.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 1), d["clocks"])
print("roofline", {k: v for k, v in d["roofline"].items() if k not in ("note", "peak_source", "launch_timing")})
print("cpu", d["cpu_baseline"]["value"], "others", {k: (v.get("glups") or v) for k, v in d["other_configs"].items()})
r = json.loads(open("gpurun_out/r2f_bench_reference_n1.json").read().strip().splitlines()[-1])
print("reference", r["value"], r["unit"])
PY
