CMD="python bench.py --steps 1 --warmup 3 --sub-iters 2 --no-e2e --no-cpu-baseline --no-others"
for pr in 0 2; do
export RNMF_TMA_L2PROMO=$pr
$CMD > gpurun_out/plain_p$pr.log 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:jacobi3d -s 6 -c 2 --csv --log-file gpurun_out/promo_$pr.csv $CMD > gpurun_out/ncu_p$pr.log 2>&1
done
grep -h "jacobi3d" gpurun_out/promo_0.csv gpurun_out/promo_2.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' 
