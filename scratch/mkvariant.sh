#!/bin/bash
# build a library variant: scratch/mkvariant.sh name "-DFLAG=..."   -> scratch/lib_name.so (kernels.cu recompiled, other objects reused)
set -e
name=$1; flags=$2
cd /root/repo/chroma_lite_b200/csrc
make -s -j4 >/dev/null
NV=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NV $ARCH $flags -O3 -std=c++17 -lineinfo --use_fast_math -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v -diag-suppress=177,550 \
    -c -o /tmp/kernels_$name.o kernels.cu 2> /root/repo/scratch/lib_$name.ptxas.log
$NV $ARCH -shared -o /root/repo/scratch/lib_$name.so runtime.o geometry.o bvh_native.o pdf.o hostmesh.o comm.o /tmp/kernels_$name.o -Xcompiler -fPIC -lcudart_static -lpthread -ldl -lrt
echo built scratch/lib_$name.so
