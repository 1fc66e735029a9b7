set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for v in 1 2 3 4; do
  timeout 300 python bench.py --steps 2 --warmup 3 --sub-iters 30 --variant $v --no-e2e --no-cpu-baseline --no-others > gpurun_out/bench_v$v.log 2>&1
done
for v in 1 2; do
  timeout 300 python bench.py --workload 2d5_8192 --steps 2 --warmup 3 --sub-iters 50 --variant $v --no-e2e --no-cpu-baseline --no-others > gpurun_out/bench2d_v$v.log 2>&1
done
CMD="python bench.py --steps 1 --warmup 3 --sub-iters 2 --no-e2e --no-cpu-baseline --no-others"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1a.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:jacobi3d -s 6 -c 2 -o gpurun_out/prof_r1a $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/pytest_gpu.log; grep -h -o '"value": [0-9.]*' gpurun_out/bench_v*.log gpurun_out/bench2d_v*.log
