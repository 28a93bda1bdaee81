set -u
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_c3_${N}gpu_r02.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/bench_c3_1gpu_torchrun_r02_same${N}box.json
python - <<PY
import json
for f in ["gpurun_out/bench_c3_${N}gpu_r02.json", "gpurun_out/bench_c3_1gpu_torchrun_r02_same${N}box.json"]:
    d = json.load(open(f)); print(f, "n_gpus", d["n_gpus"], "value", round(d["value"], 4), "e2e", round(d["e2e"]["value"], 4), "sha", d["e2e"].get("first_batch_sha256"), "requests", d["e2e"].get("requests"), "ms/step", round(d["ms_per_step"], 1), "unet", d.get("unet_step_ms"))
PY
