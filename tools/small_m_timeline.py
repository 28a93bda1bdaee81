#!/usr/bin/env python
"""in-graph start/end stamps of back-to-back few-row GEMM launches (debug library): where does the per-launch time go?"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from instructany2pix_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "libia2p_trace.so")
from instructany2pix_b200 import ops
lib = _lib.load()
lib.ia2p_debug_set_timeline.argtypes = [ctypes.c_void_p]
dev, BF = "cuda", torch.bfloat16
M, N, K = 512, 1280, 5120
tl = torch.zeros(64, 2, dtype=torch.int64, device=dev)
a = torch.randn(M, K, device=dev).to(BF)
ws = [(torch.randn(N, K, device=dev) * K ** -0.5).to(BF) for _ in range(12)]
res = torch.randn(M, N, device=dev)
one = lambda w: ops.gemm(a, w, residual=res, out_dtype=torch.float32, want_ln=True)
one(ws[0]); torch.cuda.synchronize()
lib.ia2p_debug_set_timeline(tl.data_ptr())
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for w in ws: one(w)
for _ in range(3): g.replay()
torch.cuda.synchronize()
tl[:, 0] = torch.iinfo(torch.int64).max; tl[:, 1] = 0
g.replay(); torch.cuda.synchronize()
t = tl[:12].cpu().double()
print("duration us:", [round(float(x), 1) for x in ((t[:, 1] - t[:, 0]) / 1e3)])
print("gap to next us:", [round(float(x), 1) for x in ((t[1:, 0] - t[:-1, 1]) / 1e3)])
