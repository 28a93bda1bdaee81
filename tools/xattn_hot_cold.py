import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from instructany2pix_b200 import ops
dev, BF = "cuda", torch.bfloat16
for name, B, N, heads, nbuf in [("lvl2 cold", 8, 1024, 20, 8), ("lvl2 L2-hot", 8, 1024, 20, 1), ("lvl1 cold", 8, 4096, 10, 8), ("lvl1 L2-hot", 8, 4096, 10, 1)]:
    qs = [torch.randn(B * N, heads * 64, device=dev).to(BF) for _ in range(nbuf)]
    kvt = torch.randn(B * 77, 2 * heads * 64, device=dev).to(BF)
    kvi = torch.randn(B * 4, 2 * heads * 64, device=dev).to(BF)
    outs = [torch.empty_like(t) for t in qs]
    ops.cross_attn(qs[0], kvt, 77, kvi, 4, 1.0, B, N, heads, out=outs[0]); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for r in range(8 // nbuf):
            for t, o in zip(qs, outs):
                ops.cross_attn(t, kvt, 77, kvi, 4, 1.0, B, N, heads, out=o)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10 / 8 * 1e3:6.1f} us")
