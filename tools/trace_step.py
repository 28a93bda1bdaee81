#!/usr/bin/env python
"""Per-launch, per-CTA anatomy of every tc_gemm_kernel launch INSIDE the CUDA-graph UNet step (c3 shapes by default): the debug
build keeps one trace row per (launch, CTA), so each shape's in-graph set-up time, cold-start wait, main-loop cycles, MMA-warp stalls,
epilogue tail and tear-down can be read without a profiler.  Usage: python tools/trace_step.py [B] [L]  (needs a B200; uses
tools/libia2p_trace.so, never the product .so)."""
import ctypes
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from instructany2pix_b200 import _lib

_lib.LIB_PATH = os.environ.get("IA2P_TRACE_LIB") or os.path.join(ROOT, "tools", "libia2p_trace.so")
from instructany2pix_b200 import ops  # noqa: E402
import bench  # noqa: E402

lib = _lib.load()
lib.ia2p_debug_set_timeline.argtypes = [ctypes.c_void_p]
lib.ia2p_debug_set_trace.argtypes = [ctypes.c_void_p]
lib.ia2p_debug_set_trace_stride.argtypes = [ctypes.c_int]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda", 0)
torch.set_grad_enabled(False)
NMAX, ROWS = 1024, 160
tl = torch.zeros(NMAX, 2, dtype=torch.int64, device=dev)
tr = torch.zeros(NMAX * ROWS * 16, dtype=torch.int64, device=dev)

unet, _ = bench.build_models(dev, False)
host = bench.host_inputs(B, L, 1000)
dev_in = {k: v.to(dev) for k, v in host.items()}
added = dict(text_embeds=dev_in["pooled"], time_ids=dev_in["tid"])
kv = unet.context_kv(torch.cat([dev_in["ctx"], torch.randn(2 * B, 4, 2048, device=dev).to(dev_in["ctx"].dtype)], 1))
rb = unet.time_rowbias_table(torch.tensor([981.0]), added, 2 * B)[0].contiguous()
x = dev_in["lat"].float()
for _ in range(2):
    unet.forward_core(x, rb, kv, 2 * B)
torch.cuda.synchronize()
lib.ia2p_debug_set_timeline(tl.data_ptr())          # resets the launch-id counter
tags = []
orig_run = ops._run


def run_tagged(fn, args, what):
    before, tag = lib.ia2p_debug_next_launch_id(), ops._TAG
    r = orig_run(fn, args, what)
    for _ in range(lib.ia2p_debug_next_launch_id() - before):
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
for _ in range(30):
    g.replay()
torch.cuda.synchronize()
lib.ia2p_debug_set_trace(tr.data_ptr())
lib.ia2p_debug_set_trace_stride(ROWS)
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
tr.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
lib.ia2p_debug_set_trace(None)
T = tr.view(NMAX, ROWS, 16)[:n].cpu().double()
print(f"graph replay (traced build) {e0.elapsed_time(e1) * 1e3:.0f} us, {n} tc_gemm launches")
agg = defaultdict(lambda: defaultdict(float))
for i in range(n):
    t = T[i]
    act = t[:, 0] > 0                                  # CTAs of this launch
    lead = act & (t[:, 9] > 0)                         # CTAs whose MMA warp ran (pair leaders / all single CTAs)
    if lead.sum() == 0:
        continue
    t0 = t[act, 0].min()
    a = agg[tags[i] if i < len(tags) else "?"]
    a["n"] += 1
    a["dur"] += (t[act, 14].max() - t0) / 1e3                              # first CTA start -> last CTA end (after tear-down)
    a["setup"] += (t[act, 1] - t[act, 0]).mean() / 1e3                     # barrier init, TMEM alloc, cluster sync
    a["cold"] += (t[lead, 10]).mean()                                      # cycles: first operands landed after the MMA warp began waiting
    a["loop_us"] += (t[lead, 15] - t[lead, 1]).mean() / 1e3                # MMA warp: role start -> last MMA issued
    a["loop_clk"] += t[lead, 4].mean()
    a["wait_data"] += t[lead, 2].mean()
    a["wait_epi"] += t[lead, 3].mean()
    a["tail"] += (t[act, 11].max() - t[lead, 15].max()) / 1e3              # last MMA issued -> last epilogue done
    a["teardown"] += (t[act, 14] - t[act, 11]).mean() / 1e3
    a["tiles"] += t[lead, 9].mean()
    a["spread"] += (t[act, 11].max() - t[act, 11].min()) / 1e3             # imbalance: first CTA done -> last CTA done
print(f"{'shape':64s} {'n':>3s} {'dur':>7s} {'setup':>6s} {'cold':>6s} {'loop':>7s} {'GHz':>5s} {'w.data':>6s} {'w.epi':>6s} {'tail':>6s} {'tear':>5s} {'tiles':>5s} {'spread':>6s}")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["dur"]):
    c = a["n"]
    ghz = a["loop_clk"] / max(a["loop_us"], 1e-9) / 1e3
    print(f"{k:64s} {int(c):3d} {a['dur'] / c:7.1f} {a['setup'] / c:6.2f} {a['cold'] / c / 1e3:5.1f}k {a['loop_us'] / c:7.1f} {ghz:5.2f} "
          f"{100 * a['wait_data'] / max(a['loop_clk'], 1):5.1f}% {100 * a['wait_epi'] / max(a['loop_clk'], 1):5.1f}% {a['tail'] / c:6.1f} "
          f"{a['teardown'] / c:5.2f} {a['tiles'] / c:5.1f} {a['spread'] / c:6.1f}")
print("columns: us unless noted; cold = kclk the MMA warp waited for the first k-block; loop = MMA warp role start -> last MMA issued; "
      "w.data / w.epi = share of the MMA loop spent waiting for operands / for a free accumulator; tail = last MMA issued -> last epilogue done")
