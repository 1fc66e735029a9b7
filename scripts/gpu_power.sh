# sustained (power-capped) behaviour of the sweep variants: ~10 s of sweeps each
B="python bench.py --steps 3 --warmup 2 --sub-iters 1200 --no-e2e --no-cpu-baseline --no-others"
run() { echo -n "$1: "; shift; timeout 400 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print(round(d['value'],1), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], round(r['copy_gbs_this_box']), round(r['copy_gbs_this_box_sustained_2s']))"; }
run "v0 promo3" $B
RNMF_TMA_L2PROMO=0 run "v0 promo0" $B
RNMF_TMA_L2PROMO=2 run "v0 promo2" $B
run "v7 128x16" $B --variant 7
run "v5 5stages" $B --variant 5
run "v13 128x16x3" $B --variant 13
run "v1 zmarch" $B --variant 1
run "v10 zmarch occ6" $B --variant 10
run "v0 chunk96" $B --chunk 96
