B="python bench.py --steps 3 --warmup 2 --sub-iters 1000 --no-e2e --no-cpu-baseline --no-others"
run() { echo -n "$1: "; shift; timeout 400 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print(round(d['value'],1), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], round(r['frac'],3))"; }
RNMF_L2_HINTS=0 run "hints0" $B
RNMF_L2_HINTS=1 run "hints1" $B
RNMF_L2_HINTS=0 run "hints0 again" $B
RNMF_L2_HINTS=1 run "hints1 again" $B
RNMF_L2_HINTS=1 run "hints1 chunk96" $B --chunk 96
CMD="python bench.py --steps 1 --warmup 3 --sub-iters 2 --no-e2e --no-cpu-baseline --no-others"
export RNMF_L2_HINTS=1
$CMD > gpurun_out/plain_h1.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:jacobi3d -s 6 -c 2 --csv --log-file gpurun_out/hints_1.csv $CMD > gpurun_out/ncu_h1.log 2>&1
grep -h "jacobi3d" gpurun_out/hints_1.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
