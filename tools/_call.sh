python tools/ab_step.py 2>&1 | grep ab_step
IA2P_GEMM_CG=2 python tools/ab_step.py 2>&1 | grep ab_step
IA2P_PDL=1 python tools/ab_step.py 2>&1 | grep ab_step
IA2P_WEIGHT_PREFETCH=1 python tools/ab_step.py 2>&1 | grep ab_step
