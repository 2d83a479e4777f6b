# compute-sanitizer over the kernel tests and one parity case (ONE tool per gpurun call: `bash tools/sanitize.sh memcheck`).
# The large-shape tests are left out (the tool slows kernels 10-100x); every kernel family is still launched.
cd $GRAFT_REPO_ROOT
TOOL=${1:-memcheck}
SEL='not large and not 150-1 and not test_tc_cta and not wgrad_split'
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_kernels_gpu.py tests/test_attn_fused_gpu.py tests/test_gemm_gpu.py -m gpu -q -x -k "$SEL" \
  > gpurun_out/r02_sanitizer_$TOOL.log 2>&1
echo "kernel tests rc=$?" >> gpurun_out/r02_sanitizer_$TOOL.log
timeout 900 compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "text_r3_train and bf16 and golden" \
  >> gpurun_out/r02_sanitizer_$TOOL.log 2>&1
echo "parity case rc=$?" >> gpurun_out/r02_sanitizer_$TOOL.log
grep -E "ERROR SUMMARY|passed|failed|rc=|Error|error" gpurun_out/r02_sanitizer_$TOOL.log | tail -20
