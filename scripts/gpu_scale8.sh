# bench.py at 1/2/4/8 GPUs on one box (weak scaling, 768^3 per GPU), as the driver launches it
nvidia-smi -L | wc -l
B="--steps 3 --warmup 3 --no-cpu-baseline --no-others --e2e-steps 1"
for n in 1 2 4 8; do
  if [ $n = 1 ]; then timeout 600 python bench.py $B > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n $B > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err; fi
  tail -1 gpurun_out/scale_n$n.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],2), d['e2e'] and round(d['e2e']['value'],1), d['clocks']['sm_mhz'], d['clocks']['power_w_max'], round(d['roofline']['frac'],3), d['gpu_launches'])"
done
S="--workload 2d5_8192 --steps 3 --warmup 3 --sub-iters 1500 --no-cpu-baseline --no-others --no-e2e"
echo "== N=8 2D weak"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 8 $S 2>/dev/null | tail -1 | tee gpurun_out/scale_n8_2d.json | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print(d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],2))"
