#!/bin/bash
# Development build of the library with -DGQ_CYCLES (per-role cycle accounting inside the GQA attention kernels):
#   tools/build_gqa_cycles.sh && AUDIOLLM_B200_LIB=$PWD/build/libaudiollm_cyc.so CYCLES=1 python tools/bench_gqa_attention.py
set -e
cd "$(dirname "$0")/.."
python __graft_entry__.py > /dev/null
mkdir -p build/obj
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DGQ_CYCLES $GQ_EXTRA \
  -c audio_llama_b200/csrc/gqa_attention_sm100.cu -o build/obj/gqa_cyc.o
objs=$(ls build/obj/*.o | grep -v "gqa_attention_sm100\|gqa_cyc")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/libaudiollm_cyc.so $objs build/obj/gqa_cyc.o
echo built build/libaudiollm_cyc.so
