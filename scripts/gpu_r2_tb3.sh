#!/bin/bash
# First hardware run of the 3D three-sweep kernel (kernels_sweep_tb3.cu): parity through the C ABI, A/B against two sweeps per launch
# (variant 31) on the same box, chunk lengths, one full ncu capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "jacobi_3d or kernel_launch or full_size_oracle or solve_poisson_host" > gpurun_out/tb3_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/tb3_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/tb3_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/tb3_smoke.log
B="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-others"
for v in auto 31 auto; do
  timeout 300 python bench.py $B --variant $v > gpurun_out/tb3_bench_$v.json 2> gpurun_out/tb3_bench_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/tb3_bench_$v.json').read().strip().splitlines()[-1]); print('variant $v', round(d['value'],1), 'GLUP/s', round(d['ms_per_step'],2), 'ms/step', d['clocks']['sm_mhz'], d['clocks']['power_w_max'], 'kernel_ms', d['roofline'].get('avg_launch_ms'), 'spl', d['roofline'].get('sweeps_per_launch'))" || tail -3 gpurun_out/tb3_bench_$v.err
done
for c in 48 64 128; do
  timeout 300 python bench.py $B --chunk $c > gpurun_out/tb3_bench_c$c.json 2> gpurun_out/tb3_bench_c$c.err
  python -c "
import json; d=json.loads(open('gpurun_out/tb3_bench_c$c.json').read().strip().splitlines()[-1]); print('chunk $c', round(d['value'],1), 'GLUP/s', 'kernel_ms', d['roofline'].get('avg_launch_ms'))" || tail -3 gpurun_out/tb3_bench_c$c.err
done
CMD3="python bench.py --steps 1 --warmup 3 --sub-iters 7 --no-e2e --no-cpu-baseline --no-others"
$CMD3 > gpurun_out/tb3_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi3d_tb3 -s 2 -c 1 -f -o gpurun_out/r2_prof_3d7_768_tb3 $CMD3 > gpurun_out/tb3_ncu_full.log 2>&1
echo "ncu rc=$?"
