#!/usr/bin/env python
"""Flash self-attention back to back for ~3 s per shape while nvidia-smi samples the SM clock and power: is the kernel clock- (power-) limited?
Usage: [IA2P_LIB_OVERRIDE=...] python tools/fa_clocks.py"""
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from instructany2pix_b200 import ops

tag = os.path.basename(os.environ.get("IA2P_LIB_OVERRIDE", "product"))
for name, B, N, heads in [("lvl1", 8, 4096, 10), ("lvl2", 8, 1024, 20)]:
    qkv = [torch.randn(B * N, 3 * heads * 64, device="cuda").to(torch.bfloat16) for _ in range(3)]
    for i in range(20):
        ops.flash_self_attn(qkv[i % 3], B, N, heads)
    torch.cuda.synchronize()
    samples, stop = [], False

    def poll():
        while not stop:
            o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout
            try:
                c, p = o.strip().split(",")
                samples.append((float(c), float(p)))
            except ValueError:
                pass
            time.sleep(0.05)

    th = threading.Thread(target=poll)
    th.start()
    iters = 6000 if N == 4096 else 30000
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        ops.flash_self_attn(qkv[i % 3], B, N, heads)
    e1.record()
    torch.cuda.synchronize()
    stop = True
    th.join()
    ms = e0.elapsed_time(e1) / iters
    s = samples[len(samples) // 3:]
    clk = sorted(x[0] for x in s)[len(s) // 2] if s else 0
    pw = sorted(x[1] for x in s)[len(s) // 2] if s else 0
    print(f"[{tag}] {name}: {ms * 1e3:7.1f} us  {4.0 * B * heads * N * N * 64 / ms / 1e9:6.1f} TFLOP/s sustained  median SM clock {clk:.0f} MHz, power {pw:.0f} W ({len(s)} samples)", flush=True)
