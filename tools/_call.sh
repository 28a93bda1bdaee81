set -u
mkdir -p gpurun_out
IA2P_SPIN_LIMIT_S=600 timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 12 python -m pytest tests/test_prior_parity_gpu.py -x -q -m gpu -k "test_fused_trunk and (1-14 or 1-11 or 3-7)" -p no:cacheprovider > gpurun_out/race_prior3.log 2>&1
echo "exit $?"
grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/race_prior3.log
