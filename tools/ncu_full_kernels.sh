#!/bin/bash
# `ncu --set full` captures of single representative launches of the main kernel classes (run under gpurun; reports -> gpurun_out/,
# summaries are made from them with tools/ncu_summary.py --full and committed under profiles/; the .ncu-rep files stay out of git)
#   bash tools/ncu_full_kernels.sh [round-tag]
R=${1:-r02}
cap() {  # name regex skip
  timeout 500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/${R}_full_$1 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-profile-passes > gpurun_out/ncu_full_$1.log 2>&1
  ls -la gpurun_out/${R}_full_$1.ncu-rep 2>&1 | cut -c20-
}
cap halo256 'gemm_tma_kernel<\(int\)256, \(int\)1' 15     # an 80x80 RefineNet conv (the dominant kernel; launches 14-17 of 18 per forward)
cap pair256 'gemm2_tma_kernel<\(int\)256' 20              # stage-3 fc1 on the CTA-pair kernel
cap conv0   'gemm_tma_kernel<\(int\)64, \(int\)1' 0       # conv_fuse_conv0 (dual-N folded)
cap attn_tc 'attention_tc_kernel' 10
cap post    'postprocess_kernel' 0
cap dw3     'dwconv3x3_gelu_kernel' 10
cap dw7     'dwconv7x7_kernel' 0
cap ln      'layernorm_kernel<\(int\)3' 5
