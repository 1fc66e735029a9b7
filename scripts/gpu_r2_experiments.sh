#!/bin/bash
# First GPU call of the next round: the opt-in kernels that were written and emulator-checked without a GPU
# (2D K-sweep kernel, variants 32..40; 3D two-rows-per-warp double sweep, variants 24/25, interior-specialised loops, variants 26/27; time-skewed e2e start): parity first, then
# short sustained benches per variant.  One GPU, about 10 minutes.  Results: gpurun_out/r2exp_*.json|log
mkdir -p gpurun_out
export RNMF_EXPERIMENTAL=1
: > gpurun_out/r2exp_pytest.log
# one pytest process per opt-in variant: a kernel that hangs costs one 150 s timeout, not the whole call
for v in 35 32 33 34 36 37 38 39 40; do
  RNMF_EXPERIMENTAL_VARIANTS=$v timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "2d_multi_sweep" > gpurun_out/r2exp_pytest_v$v.log 2>&1
  echo "parity 2D variant $v rc=$? $(tail -1 gpurun_out/r2exp_pytest_v$v.log)" | tee -a gpurun_out/r2exp_pytest.log
done
for v in 24 25 26 27 28; do
  RNMF_EXPERIMENTAL_VARIANTS=$v timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "3d_double_sweep" > gpurun_out/r2exp_pytest_v$v.log 2>&1
  echo "parity 3D variant $v rc=$? $(tail -1 gpurun_out/r2exp_pytest_v$v.log)" | tee -a gpurun_out/r2exp_pytest.log
done
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "time_skewed" > gpurun_out/r2exp_pytest_skew.log 2>&1
echo "parity time-skewed host call rc=$? $(tail -1 gpurun_out/r2exp_pytest_skew.log)" | tee -a gpurun_out/r2exp_pytest.log
unset RNMF_EXPERIMENTAL
run() { name=$1; shift; timeout 200 "$@" > gpurun_out/r2exp_$name.json 2> gpurun_out/r2exp_$name.err; python -c "
import json; d=json.load(open('gpurun_out/r2exp_$name.json')); print('$name', round(d['value'],1), 'GLUP/s', round(d['ms_per_step'],2), 'ms/step', d['clocks']['sm_mhz'], d['clocks']['reasons'], d['clocks']['power_w_max'], d['gpu_launches'])" || tail -3 gpurun_out/r2exp_$name.err; }
# 2D 8192^2: 480 sweeps per step divide by 2, 3, 4, 5 and 6; 40 steps = 2-5 s per run, long enough for the power cap to settle
S2="--workload 2d5_8192 --steps 40 --warmup 3 --sub-iters 480 --no-e2e --no-cpu-baseline --no-others"
run 2d_v20_single python bench.py $S2 --variant 20
run 2d_v30_tb2 python bench.py $S2 --variant 30
run 2d_v35_k2 python bench.py $S2 --variant 35
run 2d_v32_k3 python bench.py $S2 --variant 32
run 2d_v33_k4 python bench.py $S2 --variant 33
run 2d_v34_k4w python bench.py $S2 --variant 34
run 2d_v34_k4w_c128 python bench.py $S2 --variant 34 --chunk 128
run 2d_v34_k4w_c512 python bench.py $S2 --variant 34 --chunk 512
run 2d_v32_k3_c128 python bench.py $S2 --variant 32 --chunk 128
run 2d_v36_k3_di python bench.py $S2 --variant 36
run 2d_v37_k4w_di python bench.py $S2 --variant 37
run 2d_v38_k4_di python bench.py $S2 --variant 38
run 2d_v39_k5w_di python bench.py $S2 --variant 39
run 2d_v40_k6w_di python bench.py $S2 --variant 40
# 3D 768^3
S3="--workload 3d7_768 --steps 3 --warmup 3 --sub-iters 550 --no-e2e --no-cpu-baseline --no-others"
run 3d_v0 python bench.py $S3
run 3d_v24 python bench.py $S3 --variant 24
run 3d_v25 python bench.py $S3 --variant 25
run 3d_v26 python bench.py $S3 --variant 26
run 3d_v27 python bench.py $S3 --variant 27
run 3d_v28 python bench.py $S3 --variant 28
# end to end with the time-skewed start (chunked upload overlapped with the first double sweeps): compare `e2e` of the two lines
SE="--workload 3d7_768 --steps 2 --warmup 3 --sub-iters 550 --no-cpu-baseline --no-others"
run 3d_e2e_plain python bench.py $SE
RNMF_E2E_SKEW=1 run 3d_e2e_skew python bench.py $SE
RNMF_E2E_SKEW=2 run 3d_e2e_skew2 python bench.py $SE
python - <<'PY'
import json
for n in ("3d_e2e_plain", "3d_e2e_skew", "3d_e2e_skew2"):
    try:
        d = json.load(open(f"gpurun_out/r2exp_{n}.json")); print(n, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "GLUP/s", round(d["e2e"]["ms_per_step"], 1), "ms", d["e2e"]["checks"])
    except Exception as e:
        print(n, "failed", e)
PY
