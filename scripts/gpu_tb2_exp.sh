#!/bin/bash
# double-sweep kernel experiments: E/W by shuffles (22), L2 evict_first hints (23)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "double_sweep" > gpurun_out/tb2_pytest.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/tb2_pytest.log
S="--steps 3 --warmup 2 --sub-iters 200 --no-e2e --no-cpu-baseline --no-others"
run() { name=$1; shift; timeout 200 "$@" > gpurun_out/exp_$name.json 2> gpurun_out/exp_$name.err; python -c "
import json; d=json.load(open('gpurun_out/exp_$name.json')); print('$name', round(d['value'],1), 'GLUP/s', round(d['roofline']['avg_launch_ms'],3), 'ms/launch', d['clocks']['sm_mhz'], d['clocks']['reasons'], d['clocks']['power_w_max'])"; }
run v20 python bench.py $S --variant 20
run v22_shfl python bench.py $S --variant 22
RNMF_L2_HINTS=1 run v23_hint1 python bench.py $S --variant 23
RNMF_L2_HINTS=2 run v23_hint2 python bench.py $S --variant 23
run v20_c128 python bench.py $S --variant 20 --chunk 128
run v20_c96 python bench.py $S --variant 20 --chunk 96
run v21 python bench.py $S --variant 21
