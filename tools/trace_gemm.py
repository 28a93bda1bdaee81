#!/usr/bin/env python
"""Where does a tc_gemm_kernel launch spend its time?  Uses the DEBUG build of the library (csrc: `make trace`, per-role wait
counters written per CTA) -- never the product .so.  Usage: python tools/trace_gemm.py  (needs a B200)"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from instructany2pix_b200 import _lib

_lib.LIB_PATH = os.environ.get("IA2P_TRACE_LIB") or os.path.join(ROOT, "tools", "libia2p_trace.so")
from instructany2pix_b200 import ops  # noqa: E402

lib = _lib.load()
lib.ia2p_debug_set_trace.argtypes = [ctypes.c_void_p]
dev, BF = "cuda", torch.bfloat16
trace = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
lib.ia2p_debug_set_trace(trace.data_ptr())


def run(name, fn, kb=0):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    trace.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    t = trace.view(148, 16).cpu().double()
    act = t[:, 9] > 0
    t0 = t[:, 0][t[:, 0] > 0].min()
    clk = (t[act, 4] / ((t[act, 8] - t[act, 1]).clamp_min(1))).median().item()      # cycles per ns ~ GHz (loop ~ kernel body)
    f = lambda col: t[act, col].mean().item()
    print(f"{name:34s} {e0.elapsed_time(e1) * 1e3:7.1f} us | CTA start spread {(t[:, 0][t[:, 0] > 0].max() - t0) / 1e3:5.1f} us, "
          f"setup {(t[act, 1] - t[act, 0]).mean().item() / 1e3:4.1f} us, body {(t[act, 8] - t[act, 1]).mean().item() / 1e3:6.1f} us, "
          f"last end {(t[:, 8].max() - t0) / 1e3:6.1f} us | tiles/CTA {f(9):4.1f} | MMA warp: loop {f(4) / 1e3:7.1f} kclk, "
          f"wait data {100 * f(2) / f(4):4.1f}%, wait epilogue {100 * f(3) / f(4):4.1f}% | producer wait-empty {f(5) / 1e3:7.1f} kclk | "
          f"epilogue w2: wait acc {f(6) / 1e3:7.1f} kclk, busy {f(7) / 1e3:7.1f} kclk | ~{clk:.2f} GHz"
          + (f" | {f(4) / (f(9) * kb):6.1f} clk per k-block" if kb else ""))


def main():
    torch.manual_seed(0)
    M = 8192
    st = torch.rand(M, 20, 2, device=dev) + 1.0                     # LN partial sums as a 1280-wide producer emits them (4 per N tile)
    for name, N, K, mode in [("geglu 1280", 10240, 1280, "geglu"), ("geglu 1280 ln-fold", 10240, 1280, "geglu_ln"), ("ff_out 1280", 1280, 5120, "res"),
                             ("qkv 1280", 3840, 1280, "plain"), ("qkv 1280 ln-fold", 3840, 1280, "ln"),
                             ("out_proj 1280", 1280, 1280, "res"), ("to_q 1280", 1280, 1280, "plain"), ("to_q 1280 ln-fold", 1280, 1280, "ln"),
                             ("square", 8192, 8192, "plain")]:
        a = torch.randn(M, K, device=dev).to(BF)
        w = (torch.randn(N, K, device=dev) * K ** -0.5).to(BF)
        bias = torch.randn(N, device=dev)
        c1 = torch.randn(N, device=dev)
        if mode == "geglu":
            fn = lambda: ops.gemm(a, w, bias=bias, geglu=True)
        elif mode == "geglu_ln":
            fn = lambda: ops.gemm(a, w, bias=bias, geglu=True, ln=(st, c1, 1e-5))
        elif mode == "ln":
            fn = lambda: ops.gemm(a, w, bias=bias, ln=(st, c1, 1e-5))
        elif mode == "res":
            res = torch.randn(M, N, device=dev)
            fn = lambda: ops.gemm(a, w, bias=bias, residual=res, out_dtype=torch.float32, want_ln=True)
        else:
            fn = lambda: ops.gemm(a, w)
        run(f"{name} M{M} N{N} K{K}", fn, K // 64)


if __name__ == "__main__":
    main()
