#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "double_sweep or (jacobi_3d_bit_exact and (20 or 21))" > gpurun_out/tb2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/tb2_pytest.log
grep -v "^E  \|^$" gpurun_out/tb2_pytest.log | tail -40
