B="python bench.py --steps 2 --warmup 3 --sub-iters 30 --no-e2e --no-cpu-baseline --no-others"
for pr in 0 1 2 3; do
  echo -n "v0 promo $pr: "; RNMF_TMA_L2PROMO=$pr timeout 300 $B 2>&1 | grep -o '"value": [0-9.]*' | head -1
done
for v in 5 9 13 7; do
 for c in 32 48; do
  echo -n "v$v chunk $c: "; timeout 300 $B --variant $v --chunk $c 2>&1 | grep -o '"value": [0-9.]*' | head -1
 done
done
for w in 3d7_512 3d7_1024; do
  echo -n "$w v0: "; timeout 300 $B --workload $w 2>&1 | grep -o '"value": [0-9.]*' | head -1
done
