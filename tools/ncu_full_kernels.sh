#!/bin/bash
# `ncu --set full` captures of single representative launches of the main kernel classes (run under gpurun; reports -> gpurun_out/)
#   bash tools/ncu_full_kernels.sh
cap() {  # name regex skip
  timeout 500 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/r01_full_$1 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$1.log 2>&1
  ls -la gpurun_out/r01_full_$1.ncu-rep 2>&1 | cut -c20-
}
cap gemm256 'gemm_tma_kernel<\(int\)256, \(int\)0' 40
cap conv1p  'gemm_tma_kernel<\(int\)128, \(int\)1' 2
cap conv0   'gemm_tma_kernel<\(int\)64, \(int\)1' 1
cap attn    'attention_mma_kernel' 30
cap dw3     'dwconv3x3_gelu_kernel' 30
