#!/usr/bin/env python
"""GPU comparator named in SURVEY 8(d): the reference-equivalent PyTorch EAGER bf16 path on the same B200 -- the oracle's restated
SDXL UNet + the decoupled cross-attention processors moved to CUDA in bf16 (cuDNN convs, cuBLAS linears, SDPA flash attention,
~2 000 kernel launches per forward), one CFG UNet step at the c3 shapes (CFG batch 8, 128x128 latent) incl. the CFG + DDIM
arithmetic.  Test infrastructure only (imports oracle/); not part of bench.py's numbers.  Usage: python tools/eager_gpu_baseline.py [B] [L]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from oracle.attention import IPAttnProcessor2_0
from oracle.unet import SDXL_BASE, OracleUNet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev, dt = torch.device("cuda", 0), torch.bfloat16
torch.set_grad_enabled(False)
with torch.device("meta"):
    m = OracleUNet(SDXL_BASE)
m = m.to_empty(device=dev).to(dt)
procs = {}
for name, p in m.attn_processors.items():
    if name.endswith("attn2.processor"):
        hs = dict(m.named_modules())[name[: -len(".processor")]].to_q.weight.shape[0]
        procs[name] = IPAttnProcessor2_0(hs, 2048).to(dev, dt)
    else:
        procs[name] = p
m.set_attn_processor(procs)
g = torch.Generator(device=dev)
g.manual_seed(0)
for n, p in m.named_parameters():
    if p.ndim >= 2:
        p.copy_(((torch.rand(p.shape, generator=g, device=dev) * 2 - 1) * p[0].numel() ** -0.5).to(dt))
    elif n.endswith("weight"):
        p.fill_(1.0)
    else:
        p.zero_()
x = torch.randn(B, 4, L, L, device=dev, dtype=dt)
ctx = torch.randn(2 * B, 81, 2048, device=dev, dtype=dt)
added = dict(text_embeds=torch.randn(2 * B, 1280, device=dev, dtype=dt),
             time_ids=torch.tensor([[L * 8.0, L * 8.0, 0, 0, L * 8.0, L * 8.0]] * (2 * B), device=dev, dtype=dt))


def step():
    eps = m(torch.cat([x, x]), torch.tensor(981, device=dev), ctx, added_cond_kwargs=added)[0]
    eu, ec = eps.chunk(2)
    return 1.0 * x + 0.1 * (eu + 10.0 * (ec - eu))


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
F = {128: 6.7656e12, 64: 1.5919e12}[L]
print(f"PyTorch eager bf16 (oracle modules on cuda:0): {ms:.1f} ms per CFG UNet step at batch {B}, {L}x{L} latent "
      f"= {B / (50 * ms * 1e-3):.3f} images/s for 50 steps, {2 * B * F / (ms * 1e-3) / 1e12:.0f} TFLOP/s")
