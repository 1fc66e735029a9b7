#!/bin/bash
# round 2, call 4: where do the 13 % of extra DRAM reads of the quad double-sweep kernel come from?  chunk length (CTA drift) and
# TMA L2 promotion against DRAM bytes per launch (ncu, two metrics only) and sustained rate
mkdir -p gpurun_out
run() { name=$1; shift; timeout 200 "$@" > gpurun_out/r2l2_$name.json 2> gpurun_out/r2l2_$name.err; python -c "
import json; d=json.load(open('gpurun_out/r2l2_$name.json')); print('$name', round(d['value'],1), 'GLUP/s', round(d['ms_per_step'],2), 'ms/step', d['clocks']['sm_mhz'], d['clocks']['power_w_max'])" || tail -3 gpurun_out/r2l2_$name.err; }
S3="--workload 3d7_768 --steps 3 --warmup 3 --sub-iters 550 --no-e2e --no-cpu-baseline --no-others --variant 22"
for c in 32 64 96 128; do run c$c python bench.py $S3 --chunk $c; done
for p in 0 1 3; do RNMF_TMA_L2PROMO=$p run promo$p python bench.py $S3 --chunk 128; done
CMD="python bench.py --steps 1 --warmup 1 --sub-iters 4 --no-e2e --no-cpu-baseline --no-others --variant 22"
for c in 32 64 128 256; do
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:jacobi3d_tb2 -c 3 --csv --log-file gpurun_out/r2l2_ncu_c$c.csv $CMD --chunk $c > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2l2_ncu_c$c.csv')) if len(r)>10]
for r in rows[1:]:
    print('chunk $c', r[-3], r[-2], r[-1])
PY
done
for p in 0 3; do
  RNMF_TMA_L2PROMO=$p timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:jacobi3d_tb2 -c 3 --csv --log-file gpurun_out/r2l2_ncu_promo$p.csv $CMD --chunk 128 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2l2_ncu_promo$p.csv')) if len(r)>10]
for r in rows[1:]:
    print('promo $p', r[-3], r[-2], r[-1])
PY
done
