set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for v in 5 6 7 8 10 11; do
  timeout 300 python bench.py --steps 2 --warmup 3 --sub-iters 30 --variant $v --no-e2e --no-cpu-baseline --no-others > gpurun_out/bench_v$v.log 2>&1
  echo -n "v$v: "; grep -o '"value": [0-9.]*' gpurun_out/bench_v$v.log | head -1
done
