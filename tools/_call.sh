set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
timeout 900 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_c3_r02_final2.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c3_r02_final2.json'))
print('c3', d['value'], d['unet_step_ms'], d['unet_tensor_frac'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['clocks'], 'cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'])
PY
