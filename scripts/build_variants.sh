#!/bin/bash
# Ablation builds of the GEMM (TDM_DBG_NOEPI / NOMMA / NOTMA) -> tdm_b200/variants/libtdm_<name>.so (git-ignored, travels with gpurun).
set -e
cd "$(dirname "$0")/.."
mkdir -p tdm_b200/variants
OBJS=$(ls tdm_b200/build/*.o | grep -v gemm.o)
for v in ${VARIANTS:-NOEPI NOMMA NOTMA}; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
       -DTDM_DBG_$v -c tdm_b200/csrc/gemm.cu -o tdm_b200/variants/gemm_$v.o &
done
wait
for v in ${VARIANTS:-NOEPI NOMMA NOTMA}; do
  nvcc -shared -o tdm_b200/variants/libtdm_$v.so $OBJS tdm_b200/variants/gemm_$v.o -ldl
done
ls -la tdm_b200/variants/*.so
