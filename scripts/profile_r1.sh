# Round-1 profiling recipe (run under gpurun, 1 GPU). Each ncu run follows a plain run of the same command.
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
CMD="python bench.py --steps 1 --warmup 3 --sub-iters 2 --no-e2e --no-cpu-baseline --no-others"
$CMD > gpurun_out/plain3d.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1_launches_3d7_768.csv $CMD > gpurun_out/ncu_list3d.log 2>&1
$CMD > gpurun_out/plain3d_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:jacobi3d -s 6 -c 2 -o gpurun_out/r1_prof_3d7_768_tma $CMD > gpurun_out/ncu_full3d.log 2>&1
CMD2="python bench.py --workload 2d5_8192 --steps 1 --warmup 3 --sub-iters 2 --no-e2e --no-cpu-baseline --no-others"
$CMD2 > gpurun_out/plain2d.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:jacobi2d -s 6 -c 2 -o gpurun_out/r1_prof_2d5_8192 $CMD2 > gpurun_out/ncu_full2d.log 2>&1
# the line the driver will see (defaults)
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
( time python bench.py --impl reference > gpurun_out/bench_reference.json 2>&1 ) 2>&1 | grep real
