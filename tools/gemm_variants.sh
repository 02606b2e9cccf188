#!/bin/bash
# Builds variants of the GEMM kernel for A/B timing on the GPU box (select with VB200_LIB=...):
#   tools/gemm_variants.sh name1:"-DFOO=1" name2:"..."   ->  build/variants/lib<name>.so
set -e
cd "$(dirname "$0")/../tts-with-diffusion-model_b200"
python build.py > /dev/null
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  nvcc $FLAGS $defs -c csrc/gemm_tcgen05.cu -o build/variants/gemm_$name.o
  nvcc -shared -o build/variants/lib$name.so build/obj/api.o build/obj/elementwise.o build/obj/d3pm.o \
       build/obj/attn_tcgen05.o build/obj/head_sample_tcgen05.o build/obj/debug_simt.o build/variants/gemm_$name.o -gencode arch=compute_100a,code=sm_100a
  echo "built build/variants/lib$name.so ($defs)"
done
