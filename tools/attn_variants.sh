#!/bin/bash
# Builds variants of the attention kernel for A/B timing on the GPU box:
#   tools/attn_variants.sh name1:"-DFOO=1 -DBAR=2" name2:"..."   ->  build/variants/lib<name>.so
set -e
cd "$(dirname "$0")/../tts-with-diffusion-model_b200"
python build.py > /dev/null
mkdir -p build/variants
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  nvcc $FLAGS $defs -c csrc/attn_tcgen05.cu -o build/variants/attn_$name.o
  nvcc -shared -o build/variants/lib$name.so build/obj/api.o build/obj/elementwise.o build/obj/d3pm.o \
       build/obj/gemm_tcgen05.o build/obj/head_sample_tcgen05.o build/obj/debug_simt.o build/variants/attn_$name.o -gencode arch=compute_100a,code=sm_100a
  echo "built build/variants/lib$name.so ($defs)"
done

# Timeline build for tools/attn_trace.py:  tools/attn_variants.sh trace:"-DVB200_ATTN_TRACE"  then
#   VB200_LIB=tts-with-diffusion-model_b200/build/variants/libtrace.so python tools/attn_trace.py 64 1024
