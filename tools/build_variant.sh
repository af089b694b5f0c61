#!/bin/bash
# Tuning build: tools/build_variant.sh NAME -DGTB_CHAIN_MIN_BLOCKS=12 ...  ->  graphtyper_b200/libgtb200_NAME.so
# (run with GTB_LIB=graphtyper_b200/libgtb200_NAME.so python bench.py ...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared "$@" \
  -o graphtyper_b200/libgtb200_$name.so graphtyper_b200/csrc/gtb_api.cu graphtyper_b200/csrc/gtb_kernels.cu \
  graphtyper_b200/csrc/gtb_sw.cu graphtyper_b200/csrc/gtb_index_dev.cu graphtyper_b200/csrc/gtb_bam.cu -lcudart -ldl 2>&1 | grep -i "error" || true
