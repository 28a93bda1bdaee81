#!/usr/bin/env python
"""In-graph timeline of the tcgen05 GEMM/conv launches of one CUDA-graph UNet step (c3 shapes): the DEBUG build of the library
stamps %globaltimer at the start and end of every tc_gemm_kernel launch (id fixed per graph node), so the true in-graph duration
of each GEMM shape and the time between GEMMs (all other kernels + launch gaps) can be read without a profiler serialising the
stream.  Usage: python tools/timeline_step.py [B] [L]   (needs a B200; uses tools/libia2p_trace.so, never the product .so)"""
import ctypes
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from instructany2pix_b200 import _lib

_lib.LIB_PATH = os.path.join(ROOT, "tools", "libia2p_trace.so")
from instructany2pix_b200 import ops  # noqa: E402
import bench  # noqa: E402

lib = _lib.load()
lib.ia2p_debug_set_timeline.argtypes = [ctypes.c_void_p]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda", 0)
torch.set_grad_enabled(False)
NMAX = 4096
tl = torch.zeros(NMAX, 2, dtype=torch.int64, device=dev)

unet, _ = bench.build_models(dev, False)
host = bench.host_inputs(B, L, 1000)
dev_in = {k: v.to(dev) for k, v in host.items()}
added = dict(text_embeds=dev_in["pooled"], time_ids=dev_in["tid"])
kv = unet.context_kv(torch.cat([dev_in["ctx"], torch.randn(2 * B, 4, 2048, device=dev).to(dev_in["ctx"].dtype)], 1))   # 77 text + 4 IP tokens
rb = unet.time_rowbias_table(torch.tensor([981.0]), added, 2 * B)[0].contiguous()
x = dev_in["lat"].float()
for _ in range(2):
    unet.forward_core(x, rb, kv, 2 * B)
torch.cuda.synchronize()
# capture ONE forward into a graph with launch ids 0..n-1 and the shape tag of every GEMM/conv call recorded host-side
lib.ia2p_debug_set_timeline(tl.data_ptr())
tags = []
orig_run = ops._run


def run_tagged(fn, args, what):
    before, tag = lib.ia2p_debug_next_launch_id(), ops._TAG
    r = orig_run(fn, args, what)
    for _ in range(lib.ia2p_debug_next_launch_id() - before):      # conv_up2x: four launches per call
        tags.append(tag or what)
    return r


ops._run = run_tagged
ops.PROFILE = None
ops.TAG_ALWAYS = True
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = unet.forward_core(x, rb, kv, 2 * B)
ops._run = orig_run
n = lib.ia2p_debug_next_launch_id()
for _ in range(30):                       # reach the sustained (power-capped) regime
    g.replay()
torch.cuda.synchronize()
tl[:, 0] = torch.iinfo(torch.int64).max
tl[:, 1] = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(5):
    g.replay()
tl[:, 0] = torch.iinfo(torch.int64).max
tl[:, 1] = 0
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
t = tl[:n].cpu()
start, end = t[:, 0].double(), t[:, 1].double()
dur = (end - start) / 1e3
gap = torch.zeros(n, dtype=torch.double)
gap[:-1] = (start[1:] - end[:-1]) / 1e3
step_us = e0.elapsed_time(e1) * 1e3
print(f"graph replay {step_us:.0f} us; {n} tc_gemm launches: busy {dur.sum():.0f} us ({100 * dur.sum() / step_us:.1f}%), "
      f"between GEMMs {gap.sum():.0f} us; first start -> last end {(end.max() - start.min()) / 1e3:.0f} us")
agg = defaultdict(lambda: [0, 0.0, 0.0])
for i in range(n):
    a = agg[tags[i] if i < len(tags) else "?"]
    a[0] += 1
    a[1] += dur[i].item()
    a[2] += gap[i].item()
print(f"{'shape':70s} {'n':>4s} {'total us':>9s} {'avg us':>8s} {'avg gap after':>13s}")
for k, (c, d, gp) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} {c:4d} {d:9.0f} {d / c:8.1f} {gp / c:13.1f}")
