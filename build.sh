#!/bin/bash
# Build libxeofs_b200.so in-tree for sm_100a (B200).  Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"; mkdir -p build
SRC=xeofs_b200/csrc
OUT=xeofs_b200/libxeofs_b200.so
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
     -Xcompiler -fPIC -shared -Xptxas -v "$@" \
     $SRC/api.cu $SRC/stats.cu $SRC/project_simt.cu $SRC/project_tc.cu $SRC/smallmat.cu $SRC/dense64.cu $SRC/gram_bf16.cu $SRC/rotation.cu $SRC/rotation_tc.cu $SRC/reconstruct.cu \
     -o $OUT -lcudart 2> build/nvcc.log || { cat build/nvcc.log; exit 1; }
grep -E "error|warning" build/nvcc.log | grep -v "ptxas info" | head -20 || true
echo "built $OUT"
