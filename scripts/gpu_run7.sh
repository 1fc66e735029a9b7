timeout 600 python -m pytest tests/test_dist.py -x -q -m gpu 2>&1 | tail -15
B="--steps 2 --warmup 3 --sub-iters 30 --no-cpu-baseline --no-others --no-e2e"
echo "== N=2 peer"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 $B 2>&1 | tail -1 | cut -c1-330
echo "== N=2 nccl"; RNMF_PEER_HALO=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 $B 2>&1 | tail -1 | cut -c1-330
echo "== N=2 strong 1024 peer"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --workload 3d7_1024 $B 2>&1 | tail -1 | cut -c1-330
echo "== N=2 2D"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --workload 2d5_8192 $B 2>&1 | tail -1 | cut -c1-330
