B="python bench.py --steps 2 --warmup 3 --sub-iters 30 --no-e2e --no-cpu-baseline --no-others"
for c in 24 32 48 64 96 128 192 256; do
  echo -n "v6 chunk $c: "; timeout 300 $B --variant 6 --chunk $c 2>&1 | grep -o '"value": [0-9.]*' | head -1
done
for c in 32 64 128; do
  echo -n "v7 chunk $c: "; timeout 300 $B --variant 7 --chunk $c 2>&1 | grep -o '"value": [0-9.]*' | head -1
  echo -n "v9 chunk $c: "; timeout 300 $B --variant 9 --chunk $c 2>&1 | grep -o '"value": [0-9.]*' | head -1
  echo -n "v5 chunk $c: "; timeout 300 $B --variant 5 --chunk $c 2>&1 | grep -o '"value": [0-9.]*' | head -1
  echo -n "v11 chunk $c: "; timeout 300 $B --variant 11 --chunk $c 2>&1 | grep -o '"value": [0-9.]*' | head -1
done
CMD="python bench.py --steps 1 --warmup 3 --sub-iters 2 --no-e2e --no-cpu-baseline --no-others"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:jacobi3d -s 6 -c 2 -o gpurun_out/prof_r1b $CMD > gpurun_out/ncu_full.log 2>&1
