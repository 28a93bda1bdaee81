#!/usr/bin/env bash
# compute-sanitizer over the hand-rolled mbarrier / TMEM / cluster protocols at the smallest shapes:
#   tools/sanitize.sh [memcheck|racecheck|synccheck|initcheck ...]      (default: memcheck synccheck racecheck)
# Needs a B200: run it through gpurun, e.g.  gpurun --timeout 1500 -- 'bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1'
# Covers tc_gemm_kernel in every CG / EPI combination (plain, LN fold, GEGLU, fp32 out, residual ring, tail split, convs, up2x),
# fa_tc_kernel, xattn_tc_kernel and the elementwise / norm kernels through tests/test_kernels_gpu.py.
set -u
cd "$(dirname "$0")/.."
TOOLS="${*:-memcheck synccheck racecheck}"
SEL='test_gemm_plain or test_gemm_tail_split or test_gemm_every_tile_width or test_gemm_epilogues or test_gemm_fp32_stream or test_gemm_layernorm_fold or test_gemm_bf16_tma_store or test_gemm_geglu or test_groupnorm_statistics or test_conv3x3 or test_conv_up2x or test_flash_self_attn or test_decoupled_cross_attn or test_cfg_ddim or test_groupnorm or test_layernorm'
rc=0
for t in $TOOLS; do
  echo "=== compute-sanitizer --tool $t"
  extra=""
  [ "$t" = memcheck ] && extra="--leak-check no"
  timeout 1500 compute-sanitizer --tool "$t" $extra --error-exitcode 7 --print-limit 20 \
      python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "$SEL" -p no:cacheprovider 2>&1 | grep -vE "^$" | tail -25
  r=${PIPESTATUS[0]}
  echo "=== $t exit code $r"
  [ "$r" -ne 0 ] && rc=$r
done
# the persistent prior trunk (grid barriers: give its bounded waits minutes under the sanitizer) and the flash-attention edge cases
for t in $TOOLS; do
  echo "=== compute-sanitizer --tool $t (persistent prior trunk, flash attention incl. peaked rows)"
  extra=""
  [ "$t" = memcheck ] && extra="--leak-check no"
  IA2P_SPIN_LIMIT_S=600 timeout 1500 compute-sanitizer --tool "$t" $extra --error-exitcode 7 --print-limit 20 \
      python -m pytest tests/test_prior_parity_gpu.py tests/test_kernels_gpu.py -x -q -m gpu -k "(test_fused_trunk and (1-14 or 1-11 or 3-7)) or test_flash_self_attn" -p no:cacheprovider 2>&1 | grep -vE "^$" | tail -12
  r=${PIPESTATUS[0]}
  echo "=== $t exit code $r"
  [ "$r" -ne 0 ] && rc=$r
done
exit $rc
