# compute-sanitizer memcheck over the small-grid parity tests (one tool per gpurun call)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 0 python -m pytest tests/test_gpu_parity.py -x -q -k "fill_bc or jacobi_2d or (jacobi_3d and (65 or 2-2 or 1-1 or 31)) or two_ghost or gs_reference or operators or varcoef or get_h or host_entry" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/memcheck.log
tail -15 gpurun_out/memcheck.log
grep -c "Invalid" gpurun_out/memcheck.log
