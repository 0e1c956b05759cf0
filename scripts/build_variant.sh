#!/bin/bash
# scripts/build_variant.sh NAME [extra nvcc flags...]  -> quantumattention_b200/libqattn_sm100_NAME.so  (dev A/B builds)
set -e
name=$1; shift
cd "$(dirname "$0")/../quantumattention_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC "$@" --shared \
  -o ../libqattn_sm100_$name.so api.cu quantize.cu attn_fwd.cu
echo built ../libqattn_sm100_$name.so
