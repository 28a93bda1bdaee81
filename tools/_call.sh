set -u
mkdir -p gpurun_out
IA2P_SPIN_LIMIT_S=600 timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 12 python -m pytest tests/test_prior_parity_gpu.py -x -q -m gpu -k "test_fused_trunk and 1-14" -p no:cacheprovider > gpurun_out/race_prior4.log 2>&1
grep -E "passed|failed|SUMMARY" gpurun_out/race_prior4.log
