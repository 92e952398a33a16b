#!/bin/bash
# compute-sanitizer over the kernel tests at SMALL shapes (memcheck / racecheck / synccheck targets
# for the hand-rolled mbarrier / TMEM / TMA protocols).  ONE tool per invocation and per gpurun call (B200_PROFILING.md:
# several tools in one call have left GPUs unusable on this pool):
#     gpurun --timeout 1500 -- tools/sanitize.sh memcheck
#     gpurun --timeout 1500 -- tools/sanitize.sh racecheck
#     gpurun --timeout 1500 -- tools/sanitize.sh synccheck
# The log lands in gpurun_out/sanitize_<tool>.log; copy the summary into profiles/.
set -u
TOOL=${1:-memcheck}
SEL=${2:-"gemm_forward_kk or gemm_dgrad_k_mn or gemm_wgrad_mn_mn or layernorm_fwd_bwd or convmod_eval_and_train or attention_fwd or attention_dropout_mask or convsub_compose_and_forward or cast_colsum_repack or gemm_silu_backward_row_domain or fused_column_sums"}
mkdir -p gpurun_out
export B2C_SANITIZE=1          # tests skip their largest parametrisations (tests/test_kernels_gpu.py)
# the plain run must pass first (a faulting program under a tool is what wedges the GPU)
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_ctc_gpu.py -m gpu -q -x -k "$SEL or ctc_small" > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 1; }
timeout 1300 compute-sanitizer --tool "$TOOL" --error-exitcode 7 --launch-timeout 120 \
    python -m pytest tests/test_kernels_gpu.py tests/test_ctc_gpu.py -m gpu -q -x -k "$SEL or ctc_small" > gpurun_out/sanitize_${TOOL}.log 2>&1
rc=$?
echo "compute-sanitizer --tool $TOOL rc=$rc"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" gpurun_out/sanitize_${TOOL}.log | tail -20
exit 0
