timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo -n "GS: "; host/bin/magnetostatics cases/ferro_droplet.lua --no-vtk --json | grep seconds | cut -c1-60
echo -n "jacobi order: "; host/bin/magnetostatics cases/ferro_droplet.lua --no-vtk --json --order jacobi | grep seconds | cut -c1-60
B="python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-others"
run() { echo -n "$1: "; shift; timeout 400 "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print(round(d['value'],1), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], round(r['frac'],3))"; }
run "2d default" $B --workload 2d5_8192 --sub-iters 6000
run "2d chunk32" $B --workload 2d5_8192 --sub-iters 6000 --chunk 32
run "2d chunk48" $B --workload 2d5_8192 --sub-iters 6000 --chunk 48
