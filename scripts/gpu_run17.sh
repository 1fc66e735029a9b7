B="--steps 2 --warmup 3 --sub-iters 200 --no-cpu-baseline --no-others"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 $B > gpurun_out/n2.json 2> gpurun_out/n2.err; tail -3 gpurun_out/n2.err
tail -1 gpurun_out/n2.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], round(d['value'],1), d['e2e'], d['gpu_launches'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-150
