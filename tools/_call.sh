set -u
mkdir -p gpurun_out
# launch list of one eager step (serialised, cold cache: shares only)
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02b.csv python tools/profile_step.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02b.csv > gpurun_out/launches_r02b_summary.md 2>&1; head -12 gpurun_out/launches_r02b_summary.md
# --set full of: the 30th tc_gemm launch of the step and one xattn launch
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:xattn_tc --launch-skip 20 --launch-count 1 -o gpurun_out/ncu_xattn_tc_r02 -f python tools/profile_step.py > gpurun_out/ncu_x.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc_gemm --launch-skip 200 --launch-count 6 -o gpurun_out/ncu_tc_gemm_r02 -f python tools/profile_step.py > gpurun_out/ncu_g.log 2>&1
ls -la gpurun_out/*.ncu-rep
# the canonical bench lines
python bench.py --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_c3_r02_final.json
python bench.py --impl reference --steps 1 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference_arm_r02.json
bash tools/run_configs.sh c1 c2 c4 b1 rf
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c3_r02_final.json'))
print('c3', d['value'], d['unet_step_ms'], d['unet_tensor_frac'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], d.get('gpu_eager_baseline',{}).get('unet_step_ms'))
PY
