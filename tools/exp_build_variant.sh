#!/bin/bash
# build an experimental variant of the library next to the shipped one: tools/exp_build_variant.sh <suffix> <extra nvcc flags...>
set -e
suf=$1; shift
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/exp_$suf
objs=""
for f in api layout em post comm csv; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden "$@" -c scsplit_b200/csrc/$f.cu -o gpurun_out/exp_$suf/$f.o &
  objs="$objs gpurun_out/exp_$suf/$f.o"
done
wait
nvcc -shared -o scsplit_b200/libscsplit_b200_$suf.so $objs -gencode arch=compute_100a,code=sm_100a -ldl -lpthread
echo built scsplit_b200/libscsplit_b200_$suf.so
