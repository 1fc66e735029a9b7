#!/bin/bash
# Round-2 profiling of the shipped 3D path (three sweeps per launch): launch list of the bench command, full captures at 768^3 / 512^3.
# Every ncu run follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --sub-iters 10 --no-e2e --no-cpu-baseline --no-others"
$CMD > gpurun_out/r2b_plain3d.log 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2b_launches_3d7_768.csv $CMD > gpurun_out/r2b_ncu_list3d.log 2>&1
echo "launch list rc=$?"
CMD3="python bench.py --steps 1 --warmup 3 --sub-iters 7 --no-e2e --no-cpu-baseline --no-others"
$CMD3 > gpurun_out/r2b_plain3d_b.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi3d_tb3 -s 2 -c 1 -f -o gpurun_out/r2b_prof_3d7_768_tb3 $CMD3 > gpurun_out/r2b_ncu_full3d.log 2>&1
echo "ncu 3D rc=$?"
CMD5="python bench.py --workload 3d7_512 --steps 1 --warmup 3 --sub-iters 7 --no-e2e --no-cpu-baseline --no-others"
$CMD5 > gpurun_out/r2b_plain3d_c.log 2>&1 && timeout 600 ncu --set full --clock-control none -k regex:jacobi3d_tb3 -s 2 -c 1 -f -o gpurun_out/r2b_prof_3d7_512_tb3 $CMD5 > gpurun_out/r2b_ncu_full3d512.log 2>&1
echo "ncu 3D 512 rc=$?"
