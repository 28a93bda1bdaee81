#!/usr/bin/env python
"""One eager prior step (GPT-2-medium trunk, 2 x 14 rows) between cudaProfilerStart/Stop, for an ncu launch list:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file prior_launches.csv python tools/profile_prior.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from instructany2pix_b200.prior import B200Prior

dev = torch.device("cuda", 0)
torch.set_grad_enabled(False)
g = torch.Generator(device=dev).manual_seed(0)
prior = B200Prior(device=dev, use_cuda_graph=False)
for name, p in prior.named_parameters():
    if p.ndim >= 2:
        p.copy_(0.02 * torch.randn(p.shape, generator=g, device=dev))
    elif name.endswith("weight"):
        p.fill_(1.0)
    else:
        p.zero_()
prior.set_clip_hidden(0.5 * torch.randn(1, 2, 1024, generator=g, device=dev))
src = torch.randn(1, 1, 1024, device=dev)
kw = dict(num_inference_steps=1, guidance_scale=10, score=6.5, dtype=torch.float32)
for _ in range(3):
    prior.generate_diffusion(3, 0, src, device=dev, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.start()
prior.generate_diffusion(3, 0, src, device=dev, **kw)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
