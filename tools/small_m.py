#!/usr/bin/env python
"""few-row GEMMs (batch-1 512^2 shapes): per-role trace with the debug library.  IA2P_GEMM_SPLITK=0/1"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from instructany2pix_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "tools", "libia2p_trace.so")
from instructany2pix_b200 import ops
lib = _lib.load()
lib.ia2p_debug_set_trace.argtypes = [ctypes.c_void_p]
dev, BF = "cuda", torch.bfloat16
trace = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
lib.ia2p_debug_set_trace(trace.data_ptr())
M = 512
for name, N, K, mode in [("ff_out", 1280, 5120, "res"), ("out_proj", 1280, 1280, "res"), ("qkv", 3840, 1280, "plain"), ("conv-like K11520", 1280, 11520, "res")]:
    a = torch.randn(M, K, device=dev).to(BF)
    w = (torch.randn(N, K, device=dev) * K ** -0.5).to(BF)
    res = torch.randn(M, N, device=dev)
    fn = (lambda: ops.gemm(a, w, residual=res, out_dtype=torch.float32, want_ln=True)) if mode == "res" else (lambda: ops.gemm(a, w))
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    trace.zero_(); fn(); torch.cuda.synchronize()
    t = trace.view(148, 16).cpu().double()
    act = t[:, 8] > 0
    t0 = t[act, 0].min()
    own, par = t[:, 13] > 0, t[:, 10] > 0
    if own.any():
        print(f"   split-K: partial CTAs hand over at +{(t[par, 10] - t0).mean() / 1e3:5.1f} us; owners reach the fix-up at +{(t[own, 11] - t0).mean() / 1e3:5.1f}, "
              f"see all partials at +{(t[own, 12] - t0).mean() / 1e3:5.1f}, finish the fix-up at +{(t[own, 13] - t0).mean() / 1e3:5.1f}; true CTA end +{(t[act, 14].max() - t0) / 1e3:5.1f} us")
    else:
        print(f"   true CTA end +{(t[act, 14].max() - t0) / 1e3:5.1f} us")
    print(f"{name:18s} M{M} N{N} K{K}: {e0.elapsed_time(e1) / 20 * 1e3:6.1f} us/launch back-to-back | CTAs {int(act.sum())} | body avg {(t[act, 8] - t[act, 1]).mean() / 1e3:5.1f} us, "
          f"last end {(t[act, 8].max() - t0) / 1e3:5.1f} us | MMA loop {t[act, 4].mean() / 1e3:6.1f} kclk wait-data {100 * t[act, 2].mean() / max(t[act, 4].mean(), 1):4.1f}% | "
          f"epi wait-acc {t[act, 6].mean() / 1e3:6.1f} kclk busy {t[act, 7].mean() / 1e3:6.1f} kclk")
