timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --workload 2d5_8192 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-others"
run() { echo -n "$1: "; shift; timeout 400 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print(round(d['value'],1), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], round(r['frac'],3))"; }
run "2d tma burst" $B --sub-iters 200
run "2d tma sustained" $B --sub-iters 8000
run "2d ymarch sustained" $B --sub-iters 8000 --variant 1
run "2d tma chunk128" $B --sub-iters 4000 --chunk 128
run "2d tma chunk256" $B --sub-iters 4000 --chunk 256
run "2d tma chunk1024" $B --sub-iters 4000 --chunk 1024
