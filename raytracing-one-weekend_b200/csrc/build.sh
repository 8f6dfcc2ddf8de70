#!/bin/bash
# Builds librtw_b200.so (CUDA kernels + C ABI) for sm_100a, in tree.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${RTW_OUT:-../librtw_b200.so}   # RTW_OUT / RTW_BUILD_DIR / RTW_EXTRA: A/B builds next to the product library
B=${RTW_BUILD_DIR:-../build}
FLAGS="${RTW_EXTRA:-} -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -ccbin /usr/bin/g++"
newer=0
for f in rtw_kernels.cu rtw_abi.cu rtw_multi.cu rtw_build.cu rtw_device.cuh rtw_internal.h rtw_host.h rtw_bvh.h ../../include/rtw_b200.h build.sh; do
  if [ ! -e "$OUT" ] || [ "$f" -nt "$OUT" ]; then newer=1; fi
done
if [ "$newer" = 0 ] && [ "${1:-}" != "-f" ]; then echo "librtw_b200.so up to date"; exit 0; fi
mkdir -p $B
rm -f $B/rtw_kernels.o $B/rtw_abi.o $B/rtw_multi.o $B/rtw_build.o
$NVCC $FLAGS ${RTW_PTXAS_V:+-Xptxas -v} -c rtw_kernels.cu -o $B/rtw_kernels.o & p1=$!
$NVCC $FLAGS -c rtw_abi.cu -o $B/rtw_abi.o & p2=$!
$NVCC $FLAGS -c rtw_multi.cu -o $B/rtw_multi.o & p3=$!
$NVCC $FLAGS -c rtw_build.cu -o $B/rtw_build.o & p4=$!
wait $p1; wait $p2; wait $p3; wait $p4   # each wait returns that compiler's status; set -e stops on the first failure
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $B/rtw_kernels.o $B/rtw_abi.o $B/rtw_multi.o $B/rtw_build.o -lcudart_static -ldl -lpthread -lrt
echo "built $OUT"
