timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
time host/bin/magnetostatics cases/ferro_droplet.lua --no-vtk --json | tail -3 | cut -c1-200
time host/bin/magnetostatics cases/ferro_droplet.lua --no-vtk --json --order jacobi | tail -3 | cut -c1-200
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-others"
run() { echo -n "$1: "; shift; timeout 400 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print(round(d['value'],1), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], round(r['frac'],3))"; }
run "2d sustained" $B --workload 2d5_8192 --sub-iters 8000
run "512 sustained" $B --workload 3d7_512 --sub-iters 3000
run "768 chunk64" $B --sub-iters 1000 --chunk 64
run "768 chunk32" $B --sub-iters 1000 --chunk 32
run "768 chunk96" $B --sub-iters 1000 --chunk 96
run "768 v7 chunk64" $B --sub-iters 1000 --chunk 64 --variant 7
