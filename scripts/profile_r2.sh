#!/bin/bash
# Round-2 profiling recipe (gpurun, 1 GPU).  Every ncu run follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
# the lines the driver will see (defaults)
( time timeout 900 python bench.py > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err ) 2>&1 | grep real
( time timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_default_n1.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 1), d["clocks"])
print("roofline", {k: v for k, v in d["roofline"].items() if k not in ("note", "peak_source", "launch_timing")})
print("cpu", d["cpu_baseline"]["value"], "others", {k: (v.get("glups") or v) for k, v in d["other_configs"].items()})
PY
# launch list of the bench command (short form: 6 sweeps per step = 2 double launches + singles), gpu__time_duration only
CMD="python bench.py --steps 2 --warmup 3 --sub-iters 7 --no-e2e --no-cpu-baseline --no-others"
$CMD > gpurun_out/r2_plain3d.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches_3d7_768.csv $CMD > gpurun_out/r2_ncu_list3d.log 2>&1
echo "launch list rc=$?"
# one full capture of the dominant kernel of each dimension
CMD3="python bench.py --steps 1 --warmup 3 --sub-iters 4 --no-e2e --no-cpu-baseline --no-others"
$CMD3 > gpurun_out/r2_plain3d_b.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi3d_tb2 -s 2 -c 1 -f -o gpurun_out/r2_prof_3d7_768_tb2 $CMD3 > gpurun_out/r2_ncu_full3d.log 2>&1
echo "ncu 3D rc=$?"
CMD5="python bench.py --workload 3d7_512 --steps 1 --warmup 3 --sub-iters 4 --no-e2e --no-cpu-baseline --no-others"
$CMD5 > gpurun_out/r2_plain3d_c.log 2>&1 && timeout 600 ncu --set full --clock-control none -k regex:jacobi3d_tb2 -s 2 -c 1 -f -o gpurun_out/r2_prof_3d7_512_tb2 $CMD5 > gpurun_out/r2_ncu_full3d512.log 2>&1
echo "ncu 3D 512 rc=$?"
CMD2="python bench.py --workload 2d5_8192 --steps 1 --warmup 3 --sub-iters 8 --no-e2e --no-cpu-baseline --no-others"
$CMD2 > gpurun_out/r2_plain2d.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi2d_tbk -s 2 -c 1 -f -o gpurun_out/r2_prof_2d5_8192_tbk $CMD2 > gpurun_out/r2_ncu_full2d.log 2>&1
echo "ncu 2D rc=$?"
