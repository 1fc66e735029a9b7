#!/bin/bash
# A/B builds for GPU experiments: the same sources with extra nvcc flags into rnmf_b200/csrc/_alt/<name>/librnmf_b200.so
# (git-ignored; ships to the GPU box).  usage: scripts/build_alt.sh <name> <extra nvcc flags...>
set -e
cd "$(dirname "$0")/../rnmf_b200/csrc"
name=$1; shift
mkdir -p _alt/$name
for f in *.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
     -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math "$@" -c $f -o _alt/$name/${f%.cu}.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o _alt/$name/librnmf_b200.so _alt/$name/*.o -ldl
echo built _alt/$name/librnmf_b200.so
