#!/usr/bin/env python
"""few-row GEMMs with COLD weights inside a CUDA graph (host overhead excluded): IA2P_GEMM_SPLITK=0/1"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from instructany2pix_b200 import ops
dev, BF = "cuda", torch.bfloat16
M = int(sys.argv[1]) if len(sys.argv) > 1 else 512
MODE = sys.argv[2] if len(sys.argv) > 2 else "cold"          # cold | hot | hint (cold + ia2p_tc_prefetch_hint of the next weights)
from instructany2pix_b200 import _lib
lib = _lib.load()
for name, N, K, mode in [("ff_out", 1280, 5120, "res"), ("out_proj", 1280, 1280, "res"), ("qkv", 3840, 1280, "plain"), ("geglu", 10240, 1280, "geglu"), ("conv-like", 1280, 11520, "res")]:
    nw = max(2, int(400e6 // (N * K * 2)))                     # > 3 L2's worth of distinct weights
    a = torch.randn(M, K, device=dev).to(BF)
    ws = [(torch.randn(N, K, device=dev) * K ** -0.5).to(BF) for _ in range(nw)]
    res = torch.randn(M, N, device=dev)
    bias = torch.randn(N, device=dev)
    def one(w):
        if mode == "res": return ops.gemm(a, w, residual=res, out_dtype=torch.float32, want_ln=True)
        if mode == "geglu": return ops.gemm(a, w, bias=bias, geglu=True)
        return ops.gemm(a, w)
    one(ws[0]); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i, w in enumerate(ws):
            if MODE == "hint":
                nx = ws[(i + 1) % nw]
                lib.ia2p_tc_prefetch_hint(nx.data_ptr(), nx.numel() * 2)
            one(ws[0] if MODE == "hot" else w)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 / nw * 1e3
    print(f"{name:10s} M{M} N{N} K{K}: {us:6.1f} us per launch in-graph, {MODE} weights ({2.0 * M * N * K / us / 1e6:6.1f} TFLOP/s, W stream {N * K * 2 / us / 1e3:5.0f} GB/s)")
