set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py tests/test_prior_parity_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
