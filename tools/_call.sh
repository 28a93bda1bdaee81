set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "flash or attn" 2>&1 | tail -5
timeout 120 python tools/bench_fa.py 2>&1 | grep self-attn
