nvidia-smi -L | wc -l
B="--steps 3 --warmup 3 --sub-iters 50 --no-cpu-baseline --no-others"
for n in 1 2 4 8; do
  echo "== N=$n peer"
  if [ $n = 1 ]; then timeout 600 python bench.py $B > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n $B > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; fi
  tail -1 gpurun_out/scale_n$n.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'] and d['e2e']['value'], d['clocks'])"
done
echo "== N=8 nccl"; RNMF_PEER_HALO=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 $B --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['ms_per_step'])"
echo "== N=8 strong 1024"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --workload 3d7_1024 $B --no-e2e 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['ms_per_step'])"
timeout 600 python -m pytest tests/test_dist.py -x -q -m gpu 2>&1 | tail -3
