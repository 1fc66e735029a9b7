#!/bin/bash
# A/B of the three-sweep kernel against two sweeps per launch (variant 31) on one box + one full ncu capture of the three-sweep kernel
mkdir -p gpurun_out
TAG=${1:-b}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "jacobi_3d_triple or jacobi_3d_bit_exact" > gpurun_out/tb3${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/tb3${TAG}_pytest.log
B="--steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-others"
for v in auto 31 auto; do
  timeout 300 python bench.py $B --variant $v > gpurun_out/tb3${TAG}_bench_$v.json 2> gpurun_out/tb3${TAG}_bench_$v.err
  python -c "
import json; d=json.loads(open('gpurun_out/tb3${TAG}_bench_$v.json').read().strip().splitlines()[-1]); print('variant $v', round(d['value'],1), 'GLUP/s', round(d['ms_per_step'],2), 'ms/step', d['clocks']['sm_mhz'], d['clocks']['power_w_max'], 'kernel_ms', d['roofline'].get('avg_launch_ms'), 'spl', d['roofline'].get('sweeps_per_launch'))" || tail -3 gpurun_out/tb3${TAG}_bench_$v.err
done
CMD3="python bench.py --steps 1 --warmup 3 --sub-iters 7 --no-e2e --no-cpu-baseline --no-others"
$CMD3 > gpurun_out/tb3${TAG}_plain.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi3d_tb3 -s 2 -c 1 -f -o gpurun_out/r2_prof_3d7_768_tb3${TAG} $CMD3 > gpurun_out/tb3${TAG}_ncu_full.log 2>&1
echo "ncu rc=$?"
