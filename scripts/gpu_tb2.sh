#!/bin/bash
# tb2 (two sweeps per HBM pass) bring-up: parity tests, then short benches of the variants
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "double_sweep or (jacobi_3d_bit_exact and (20 or 21)) or full_size or resident or launch_counter" > gpurun_out/tb2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tb2_pytest.log
tail -15 gpurun_out/tb2_pytest.log
for v in ${VARIANTS:-6 20 21}; do
  timeout 200 python bench.py --steps 2 --warmup 1 --sub-iters 100 --no-e2e --no-cpu-baseline --no-others --variant $v > gpurun_out/tb2_bench_v$v.json 2> gpurun_out/tb2_bench_v$v.err
  echo "bench v$v rc=$?"; head -c 600 gpurun_out/tb2_bench_v$v.json; echo
done
