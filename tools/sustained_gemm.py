#!/usr/bin/env python
"""Sustained (power-capped) throughput of the step's GEMM shapes: this library vs cuBLAS (torch.matmul), same box, 2 s loops."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from instructany2pix_b200 import ops
dev, BF = "cuda", torch.bfloat16
def loop(fn, secs=2.0):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    n, t0 = 0, time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(20): fn()
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, M, N, K in [("geglu-shape", 8192, 10240, 1280), ("qkv", 8192, 3840, 1280), ("ff_out-shape", 8192, 1280, 5120), ("out_proj-shape", 8192, 1280, 1280),
                      ("geglu640-shape", 32768, 5120, 640), ("square", 8192, 8192, 8192)]:
    a = [torch.randn(M, K, device=dev).to(BF) for _ in range(3)]
    w = [(torch.randn(N, K, device=dev) * K ** -0.5).to(BF) for _ in range(3)]
    i = [0]
    def ours():
        i[0] += 1; ops.gemm(a[i[0] % 3], w[i[0] % 3])
    def cublas():
        i[0] += 1; torch.matmul(a[i[0] % 3], w[i[0] % 3].t())
    t1, t2 = loop(ours), loop(cublas)
    fl = 2.0 * M * N * K
    print(f"{name:16s} M{M} N{N} K{K}: ours {t1*1e3:7.1f} us {fl/t1/1e9:7.1f} TF/s | cuBLAS {t2*1e3:7.1f} us {fl/t2/1e9:7.1f} TF/s | ours/cuBLAS {t2/t1:.2f}")
