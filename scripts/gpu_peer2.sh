timeout 600 python -m pytest tests/test_dist.py -x -q -m gpu 2>&1 | tail -8
S="--steps 3 --warmup 3 --sub-iters 300 --no-cpu-baseline --no-others --no-e2e"
run() { echo -n "$1: "; shift; timeout 600 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],2))"; }
run "N=1 3D" python bench.py $S
run "N=2 3D in-kernel" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 $S
RNMF_PEER_HALO=0 run "N=2 3D nccl" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 2 $S
S2="--workload 2d5_8192 --steps 3 --warmup 3 --sub-iters 1500 --no-cpu-baseline --no-others --no-e2e"
run "N=1 2D" python bench.py $S2
run "N=2 2D in-kernel" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29583 bench.py --gpus 2 $S2
