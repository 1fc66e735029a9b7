B="python bench.py --steps 3 --warmup 2 --sub-iters 1000 --no-e2e --no-cpu-baseline --no-others"
run() { echo -n "$1: "; shift; timeout 400 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print(round(d['value'],1), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], round(r['frac'],3))"; }
for pr in 3 2 1 0; do RNMF_TMA_L2PROMO=$pr run "3d promo$pr" $B; done
run "3d v7 c48" $B --variant 7
run "3d v5" $B --variant 5
run "3d v9" $B --variant 9
