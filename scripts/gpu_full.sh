#!/bin/bash
# whole GPU suite + default bench line (1 GPU)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
( time timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_default.json'))
print(d['value'], d['ms_per_step'], d['roofline'], d['e2e'], d['clocks'], d['gpu_launches'], d['other_configs'])
PY
# reference arm and the ncu launch list of the default command (short form), after the plain runs above
( time timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err ) 2>&1 | grep real
CMD="python bench.py --steps 1 --warmup 3 --sub-iters 6 --no-e2e --no-cpu-baseline --no-others"
$CMD > gpurun_out/plain3d.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r1_launches_3d7_768_tb2.csv $CMD > gpurun_out/ncu_list3d.log 2>&1
echo "launch list rc=$?"
