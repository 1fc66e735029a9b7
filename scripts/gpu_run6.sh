nvidia-smi -L
timeout 600 python -m pytest tests/test_dist.py -x -q -m gpu 2>&1 | tail -15
B="--steps 2 --warmup 3 --sub-iters 30 --no-cpu-baseline --no-others --no-e2e"
echo "== N=1"; timeout 300 python bench.py $B 2>&1 | tail -1 | cut -c1-400
echo "== N=2 overlap"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 $B 2>&1 | tail -2 | cut -c1-600
echo "== N=2 no overlap"; RNMF_HALO_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 $B 2>&1 | tail -2 | cut -c1-600
