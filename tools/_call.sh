set -u
IA2P_PDL=1 timeout 300 python tools/timeline_step.py 1 64 2>&1 | head -9
IA2P_PDL=0 timeout 300 python tools/timeline_step.py 1 64 2>&1 | head -3
