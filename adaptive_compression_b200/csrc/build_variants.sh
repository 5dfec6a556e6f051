#!/bin/bash
# dev tool: build libambc_<tag>.so with extra -D flags for compress.cu only (the other objects are reused)
# usage: ./build_variants.sh tag "-DSF_G=4 ..."
set -e
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
tag=$1; shift
mkdir -p build_var
$NVCC -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC --fmad=true $@ -c compress.cu -o build_var/compress_$tag.o
$NVCC $ARCH -shared -o ../libambc_$tag.so build_var/compress_$tag.o api.o decode.o index.o codec_batch.o marker.o synth.o -lcudart
