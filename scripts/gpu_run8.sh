timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 3 --warmup 2 --sub-iters 1200 --no-e2e --no-cpu-baseline --no-others"
run() { echo -n "$1: "; shift; timeout 400 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print(round(d['value'],1), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], round(r['copy_gbs_this_box']), round(r['copy_gbs_this_box_sustained_2s']))"; }
run "v0 sustained" $B
run "v7 sustained" $B --variant 7
run "v5 sustained" $B --variant 5
run "v0 short" python bench.py --steps 2 --warmup 3 --sub-iters 30 --no-e2e --no-cpu-baseline --no-others
CMD="python bench.py --steps 1 --warmup 3 --sub-iters 2 --no-e2e --no-cpu-baseline --no-others"
$CMD > gpurun_out/plain3d_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:jacobi3d -s 6 -c 2 -o gpurun_out/r1_prof_3d7_768_tma_lean $CMD > gpurun_out/ncu_full3d.log 2>&1
