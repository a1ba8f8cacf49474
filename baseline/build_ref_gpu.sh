#!/bin/bash
# GPU baseline: the reference's own kernels compiled UNMODIFIED, from where they lie, for sm_100 (what
# `make GPU_ARCH=sm_100` of /root/reference/Makefile does per kernel: nvcc --std=c++14 -O3 -rdc=true -Iinclude),
# linked with the timing harness baseline/ref_kernels_bench.cu into baseline/_ref/ (git-ignored, travels to the
# GPU box).  No reference source is copied into the repository.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}"
[ -d "$REF/kernels" ] || { echo "no reference tree at $REF: keeping the prebuilt baseline/_ref"; exit 0; }
mkdir -p "$HERE/_ref"
K="$REF/kernels"
nvcc --std=c++14 -O3 -gencode arch=compute_100,code=sm_100 -Xcompiler -fPIC -rdc=true -I"$REF/include" --shared \
  -o "$HERE/_ref/libref_gpu_kernels.so" "$HERE/ref_kernels_bench.cu" \
  "$K/fct_ale_a1.cu" "$K/fct_ale_a2.cu" "$K/fct_ale_a3.cu" "$K/fct_ale_b1_vertical.cu" "$K/fct_ale_b1_horizontal.cu" \
  "$K/fct_ale_b2.cu" "$K/fct_ale_b3_vertical.cu" "$K/fct_ale_b3_horizontal.cu" "$K/fct_ale_c_vertical.cu" "$K/fct_ale_c_horizontal.cu"
echo "built baseline/_ref/libref_gpu_kernels.so from $K (sm_100)"
