#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "2d" > gpurun_out/tb2_2d_pytest.log 2>&1
echo "pytest rc=$?"; grep -v "^E  \|^$" gpurun_out/tb2_2d_pytest.log | tail -12
S="--workload 2d5_8192 --steps 3 --warmup 2 --sub-iters 400 --no-e2e --no-cpu-baseline --no-others"
run() { name=$1; shift; timeout 200 "$@" > gpurun_out/exp2d_$name.json 2> gpurun_out/exp2d_$name.err; python -c "
import json; d=json.load(open('gpurun_out/exp2d_$name.json')); print('$name', round(d['value'],1), 'GLUP/s', round(d['roofline']['avg_launch_ms'],3), 'ms/launch', d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['clocks']['power_w_max'])"; }
run v20_single python bench.py $S --variant 20
run v30 python bench.py $S --variant 30
run v30_c32 python bench.py $S --variant 30 --chunk 32
run v30_c128 python bench.py $S --variant 30 --chunk 128
run v30_c256 python bench.py $S --variant 30 --chunk 256
