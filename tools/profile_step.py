#!/usr/bin/env python
"""One eager UNet step at the c3 shapes (CFG batch 8, 128x128 latent) between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/profile_step.py
  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
      -k regex:tc_gemm --clock-control none --csv --log-file gemm_traffic.csv python tools/profile_step.py

Everything before the profiled step (model build, K/V hoist, two warm forwards) is outside the capture range."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda", 0)
torch.set_grad_enabled(False)
unet, _ = bench.build_models(dev, False)
host = bench.host_inputs(B, L, 1000)
d = {k: v.to(dev) for k, v in host.items()}
added = dict(text_embeds=d["pooled"], time_ids=d["tid"])
kv = unet.context_kv(torch.cat([d["ctx"], torch.randn(2 * B, 4, 2048, device=dev).to(d["ctx"].dtype)], 1))   # 77 text + 4 IP tokens
rb = unet.time_rowbias_table(torch.tensor([981.0]), added, 2 * B)[0].contiguous()
x = d["lat"].float()
from instructany2pix_b200 import ops
from instructany2pix_b200.scheduler import B200DDIMScheduler
s = B200DDIMScheduler()
s.set_timesteps(50)
for _ in range(2):
    eps = unet.forward_core(x, rb, kv, 2 * B)
    s.cfg_step(eps, 981, x.clone(), 10.0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eps = unet.forward_core(x, rb, kv, 2 * B)
s.cfg_step(eps, 981, x.clone(), 10.0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one UNet step: CFG batch", 2 * B, "latent", L)
