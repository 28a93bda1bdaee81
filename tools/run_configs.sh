#!/usr/bin/env bash
# One bench line per BASELINE.json configuration (c1 prior alone, c2 512^2 batch 1, c3 1024^2 batch 4 [the driver's default],
# c4 prior + 1024^2 batch 8, b1 one interactive request) -> gpurun_out/bench_<workload>_r02.json.  Needs a B200 (gpurun).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in ${*:-c1 c2 c3 c4 b1}; do
  python bench.py --workload "$w" --steps 3 --warmup 3 2>/dev/null | tail -1 > "gpurun_out/bench_${w}_r02.json"
  python - "$w" <<'PY'
import json, sys
w = sys.argv[1]
d = json.load(open(f"gpurun_out/bench_{w}_r02.json"))
print(w, d["metric"], f"value {d['value']:.4g} {d['unit']}", f"e2e {d['e2e']['value']:.4g}", f"unet_step_ms {d.get('unet_step_ms')}",
      f"frac {d.get('unet_tensor_frac')}", f"roofline {d['roofline']['frac']:.3f}", f"eager {d.get('gpu_eager_baseline', {}).get('unet_step_ms')}")
PY
done
