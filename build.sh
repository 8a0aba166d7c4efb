#!/bin/bash
# Build the C-ABI shared library for sm_100a (in-tree, travels to the GPU box with the snapshot).
set -e
cd "$(dirname "$0")"
PKG="multitask-hydranet_b200"
OUT="$PKG/libhydranet_b200.so"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v"
mkdir -p build
for f in hn_api hn_conv_gemm hn_direct hn_postproc hn_train hn_wgrad; do
  src="$PKG/csrc/$f.cu"
  obj="build/$f.o"
  if [ ! -f "$obj" ] || [ "$src" -nt "$obj" ] || [ "$PKG/csrc/hn_common.cuh" -nt "$obj" ] || [ "$PKG/csrc/hn_ops.h" -nt "$obj" ] || [ include/hydranet_b200.h -nt "$obj" ]; then
    echo "nvcc $src"
    $NVCC $FLAGS -c "$src" -o "$obj" 2> "build/$f.ptxas.log" || { cat "build/$f.ptxas.log"; exit 1; }
  fi
done
$NVCC -shared -o "$OUT" build/hn_api.o build/hn_conv_gemm.o build/hn_direct.o build/hn_postproc.o build/hn_train.o build/hn_wgrad.o -lcudart_static -ldl -lrt -lpthread
echo "built $OUT"
