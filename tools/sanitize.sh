#!/bin/bash
# compute-sanitizer passes over the hand-rolled mbarrier / TMEM / TMA pipelines (run on the GPU box: gpurun -- tools/sanitize.sh).
# racecheck + synccheck on the operator tests of the tcgen05 engine and the attention core (small shapes), memcheck on one whole
# forward.  Logs: gpurun_out/sanitize_<tool>.log (copy the summaries into profiles/).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
SEL='test_conv3x3_tma_halo or test_conv_gemm_tma_engine or (test_attention and (mma or tc)) or test_fused_pred_argmax or test_postprocess_op'
for tool in racecheck synccheck; do
  timeout ${SAN_TIMEOUT:-900} $SAN --tool $tool --print-limit 20 --log-file gpurun_out/sanitize_$tool.log \
    python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "$SEL" > gpurun_out/sanitize_${tool}_pytest.log 2>&1
  echo "$tool exit $?" >> gpurun_out/sanitize_$tool.log
  tail -3 gpurun_out/sanitize_${tool}_pytest.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_$tool.log | tail -5
done
timeout ${SAN_TIMEOUT:-900} $SAN --tool memcheck --print-limit 20 --log-file gpurun_out/sanitize_memcheck.log \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_memcheck_py.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitize_memcheck.log
tail -2 gpurun_out/sanitize_memcheck_py.log
grep -E "ERROR SUMMARY" gpurun_out/sanitize_memcheck.log | tail -2
