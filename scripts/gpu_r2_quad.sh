#!/bin/bash
# round 2, call 3: the quad mapping of the 3D double-sweep kernel (variants 22 / 23) against variant 26 and the default:
# parity, sustained benches at 768^3 and 512^3, one ncu --set full capture of variant 22
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "3d_double_sweep" > gpurun_out/r2quad_pytest.log 2>&1
echo "parity rc=$? $(tail -1 gpurun_out/r2quad_pytest.log)"
run() { name=$1; shift; timeout 200 "$@" > gpurun_out/r2quad_$name.json 2> gpurun_out/r2quad_$name.err; python -c "
import json; d=json.load(open('gpurun_out/r2quad_$name.json')); print('$name', round(d['value'],1), 'GLUP/s', round(d['ms_per_step'],2), 'ms/step', d['clocks']['sm_mhz'], d['clocks']['reasons'], d['clocks']['power_w_max'], d['gpu_launches'])" || tail -3 gpurun_out/r2quad_$name.err; }
S3="--workload 3d7_768 --steps 3 --warmup 3 --sub-iters 550 --no-e2e --no-cpu-baseline --no-others"
run 768_v0 python bench.py $S3
run 768_v26 python bench.py $S3 --variant 26
run 768_v22 python bench.py $S3 --variant 22
run 768_v23 python bench.py $S3 --variant 23
run 768_v22_c128 python bench.py $S3 --variant 22 --chunk 128
run 768_v22_c192 python bench.py $S3 --variant 22 --chunk 192
S5="--workload 3d7_512 --steps 6 --warmup 3 --sub-iters 550 --no-e2e --no-cpu-baseline --no-others"
run 512_v26 python bench.py $S5 --variant 26
run 512_v22 python bench.py $S5 --variant 22
CMD="python bench.py --steps 1 --warmup 3 --sub-iters 4 --no-e2e --no-cpu-baseline --no-others --variant 22"
$CMD > gpurun_out/r2quad_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi3d_tb2 -s 2 -c 1 -f -o gpurun_out/r2_prof_3d7_768_tb2_v22 $CMD > gpurun_out/r2quad_ncu.log 2>&1
echo ncu rc=$?; tail -3 gpurun_out/r2quad_ncu.log
