python tools/ab_step.py 2>&1 | grep ab_step
IA2P_GEMM_EPI=1 python tools/ab_step.py 2>&1 | grep ab_step
python tools/trace_step.py > gpurun_out/trace_step_c3_r02d.log 2>&1; head -50 gpurun_out/trace_step_c3_r02d.log | cut -c1-180
