set -u
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "flash" 2>&1 | tail -5
