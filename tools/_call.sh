for v in "" fap1 fap2; do
  if [ -n "$v" ]; then export IA2P_LIB_OVERRIDE=$PWD/tools/libia2p_$v.so; fi
  echo "== ${v:-product}"
  timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "flash_self" 2>&1 | tail -1
  python tools/bench_kernels.py attn 2>&1 | grep self-attn
  python tools/ab_step.py 2>&1 | grep ab_step
done
