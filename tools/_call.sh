set -u
for r in 1 2; do
timeout 300 python tools/ab_step.py 1 64 60 2>&1 | grep ab_step
IA2P_LIB_OVERRIDE=tools/libia2p_s9.so timeout 300 python tools/ab_step.py 1 64 60 2>&1 | grep ab_step
done
timeout 300 python tools/ab_step.py 1 128 40 2>&1 | grep ab_step
IA2P_LIB_OVERRIDE=tools/libia2p_s9.so timeout 300 python tools/ab_step.py 1 128 40 2>&1 | grep ab_step
timeout 300 python tools/ab_step.py 4 128 30 2>&1 | grep ab_step
IA2P_LIB_OVERRIDE=tools/libia2p_s9.so timeout 300 python tools/ab_step.py 4 128 30 2>&1 | grep ab_step
