#!/bin/bash
# the round's last GPU seconds: whole GPU suite, 8 workers, hard 9 s limit (host binaries are prebuilt and travel)
mkdir -p gpurun_out
timeout 9 python -m pytest tests -m gpu -q -n 8 -p no:cacheprovider > gpurun_out/final_suite.log 2>&1
echo "rc=$?"; tail -15 gpurun_out/final_suite.log
