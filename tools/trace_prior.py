#!/usr/bin/env python
"""Phase anatomy of the persistent prior trunk kernel (DEBUG build tools/libia2p_trace.so): globaltimer stamps of CTA 0 around every
grid barrier -> time spent in each phase body and waiting at each barrier, averaged over the layers.  Usage: python tools/trace_prior.py [bs]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from instructany2pix_b200 import _lib

_lib.LIB_PATH = os.path.join(ROOT, "tools", "libia2p_trace.so")
from instructany2pix_b200 import ops  # noqa: E402
from instructany2pix_b200.prior import B200Prior  # noqa: E402

lib = _lib.load()
lib.ia2p_debug_set_pt_trace.argtypes = [ctypes.c_void_p]
lib.ia2p_debug_set_pt_fine.argtypes = [ctypes.c_void_p]
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.set_grad_enabled(False)
torch.manual_seed(0)
pr = B200Prior(n_layer=24, device="cuda", use_cuda_graph=False)
for q in pr.parameters():
    q.normal_(0, 0.02)
seq = torch.randn(2 * bs, 14, 1024, device="cuda")
for _ in range(3):
    pr._trunk_last(seq)
torch.cuda.synchronize()
buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
lib.ia2p_debug_set_pt_trace(buf.data_ptr())
fine = torch.zeros(8192, dtype=torch.int64, device="cuda")
lib.ia2p_debug_set_pt_fine(fine.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pr._trunk_last(seq)
e1.record()
torch.cuda.synchronize()
t = buf.cpu().tolist()
L = 24
n = 1 + 10 * L + 2
names = ["wait b0 (after P5/prologue)", "P1 LN1+QKV", "wait b1", "P2 attention", "wait b2", "P3 O-proj", "wait b3", "P4 LN2+FC+GELU", "wait b4", "P5 FC-out"]
acc = [0.0] * 10
for l in range(1, L):          # layer 0's first interval is the prologue
    s = 1 + 10 * l
    prev_end = t[s - 1]
    for i in range(10):
        # stamps: s+0 = before b0, s+1 = after b0, s+2 = before b1 (= end of P1), ...
        pass
    acc[0] += t[s + 1] - t[s]
    acc[1] += t[s + 2] - t[s + 1]
    acc[2] += t[s + 3] - t[s + 2]
    acc[3] += t[s + 4] - t[s + 3]
    acc[4] += t[s + 5] - t[s + 4]
    acc[5] += t[s + 6] - t[s + 5]
    acc[6] += t[s + 7] - t[s + 6]
    acc[7] += t[s + 8] - t[s + 7]
    acc[8] += t[s + 9] - t[s + 8]
    acc[9] += t[s + 10] - t[s + 9]
print(f"prior trunk bs={bs}: kernel {e0.elapsed_time(e1) * 1e3:.1f} us (events), CTA 0 first->last stamp {(t[n - 1] - t[0]) / 1e3:.1f} us; per layer (us, CTA 0):")
for nm, a in zip(names, acc):
    print(f"  {nm:32s} {a / (L - 1) / 1e3:6.2f}")
print(f"  sum per layer {sum(acc) / (L - 1) / 1e3:6.2f}")
# inside the GEMM phases (8 stamps per phase call at <= 32 rows): start | loads issued | (LN: rows landed) | statistics done | MMA loop done |
# all warps done | partials in smem ... | epilogue done
f = fine.cpu().tolist()
per = 8
pn = ["P1 LN1+QKV", "P3 O-proj", "P4 LN2+FC", "P5 FC-out"]
seg = ["issue cp.async", "wait rows (LN) / -", "LN statistics / -", "fragments + MMA", "wait other warps", "reduce + epilogue"]
tot = [[0.0] * 7 for _ in range(4)]
for l in range(1, L):
    for ph in range(4):
        s0 = (l * 4 + ph) * per
        st = f[s0:s0 + per]
        for i in range(7):
            tot[ph][i] += st[i + 1] - st[i]
print("inside the GEMM phases (us, CTA 0 thread 0, mean over layers):")
for ph in range(4):
    t7 = [x / (L - 1) / 1e3 for x in tot[ph]]
    print(f"  {pn[ph]:12s} issue {t7[0]:5.2f} | rows landed {t7[1]:5.2f} | statistics {t7[2]:5.2f} | (stamp) {t7[3]:5.2f} | fragments + MMA {t7[4]:5.2f} | wait warps {t7[5]:5.2f} | reduce + epilogue {t7[6]:5.2f}")
