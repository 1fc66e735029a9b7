#!/bin/bash
# round 2: chunk length (rows marched per CTA) of the 2D four-sweep kernel against the sustained rate
mkdir -p gpurun_out
S2="--workload 2d5_8192 --steps 30 --warmup 3 --sub-iters 481 --no-e2e --no-cpu-baseline --no-others"
for c in 0 32 48 64 96 128 160 192 256; do
  timeout 200 python bench.py $S2 --chunk $c > gpurun_out/tbkx_c$c.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/tbkx_c$c.json')); print('chunk $c', round(d['value'],1), 'GLUP/s', 'kernel_ms', round(d['roofline']['avg_launch_ms'],4), d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"
done
for c in 0 96; do
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:jacobi2d_tbk -s 2 -c 2 --csv --log-file gpurun_out/tbkx_ncu_c$c.csv python bench.py --workload 2d5_8192 --steps 1 --warmup 1 --sub-iters 9 --no-e2e --no-cpu-baseline --no-others --chunk $c > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/tbkx_ncu_c$c.csv')) if len(r)>10]
print('ncu chunk $c', [(r[8], r[-3][-28:], r[-1]) for r in rows[1:]])
PY
done
