#!/usr/bin/env python
"""Flash self-attention alone at the c3 shapes (level 1: 4096 tokens x 10 heads, level 2: 1024 tokens x 20 heads, CFG batch 8), CUDA-event
timing over distinct inputs.  IA2P_LIB_OVERRIDE=tools/libia2p_<variant>.so selects an experiment build.  Usage: python tools/bench_fa.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from instructany2pix_b200 import ops
from tools.bench_kernels import timeit

BF = torch.bfloat16
tag = os.path.basename(os.environ.get("IA2P_LIB_OVERRIDE", "product"))
for name, B, N, heads in [("lvl1", 8, 4096, 10), ("lvl2", 8, 1024, 20), ("c2 lvl1", 2, 1024, 10), ("c2 lvl2", 2, 256, 20)]:
    qkv = [torch.randn(B * N, 3 * heads * 64, device="cuda").to(BF) for _ in range(3)]
    ms = min(timeit(lambda i: ops.flash_self_attn(qkv[i % 3], B, N, heads), iters=30) for _ in range(3))
    print(f"[{tag}] self-attn {name} B={B} N={N} h={heads}: {ms * 1e3:8.1f} us  {4.0 * B * heads * N * N * 64 / ms / 1e9:7.1f} TFLOP/s", flush=True)
