#!/usr/bin/env python
"""Per-role wait/busy cycle counters of the tcgen05 flash-attention kernel (DEBUG build, tools/libia2p_trace.so)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from instructany2pix_b200 import _lib

_lib.LIB_PATH = os.path.join(ROOT, "tools", "libia2p_trace.so")
from instructany2pix_b200 import ops  # noqa: E402

lib = _lib.load()
lib.ia2p_debug_set_fa_trace.argtypes = [ctypes.c_void_p]
dev = "cuda"
trace = torch.zeros(64 * 16, dtype=torch.int64, device=dev)
lib.ia2p_debug_set_fa_trace(trace.data_ptr())
for B, N, H in [(8, 4096, 10), (8, 1024, 20)]:
    qkv = torch.randn(B * N, 3 * H * 64, device=dev).to(torch.bfloat16)
    fn = lambda: ops.flash_self_attn(qkv, B, N, H)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    trace.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    t = trace.view(64, 16)[: N // 128].cpu().double()
    f = lambda c: t[:, c].mean().item()
    nb = f(9)
    print(f"B{B} N{N} H{H}: {us:7.1f} us  {4.0 * B * H * N * N * 64 / us / 1e6:6.1f} TFLOP/s | per key block (clk): MMA-warp loop {f(4) / nb:6.0f} "
          f"[wait K {f(0) / nb:5.0f}, wait S-free {f(1) / nb:5.0f}, wait V {f(2) / nb:5.0f}, wait P {f(3) / nb:5.0f}] | softmax warp: wait S {f(5) / nb:5.0f}, "
          f"softmax {f(6) / nb:5.0f}, wait PV(prev) {f(7) / nb:5.0f}, P write {f(8) / nb:5.0f}")
