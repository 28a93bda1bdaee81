set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_prior_parity_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 60 python tools/trace_prior.py 1 2>&1 | sed -n 1,13p
timeout 120 python bench.py --workload c1 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_c1_r02_v5.json
python -c "
import json; d=json.load(open('gpurun_out/bench_c1_r02_v5.json')); print('c1', d['value'], d['unit'], d['e2e']['value'], d['roofline']['trunk_graph_replay_us'])"
