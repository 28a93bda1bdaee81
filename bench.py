#!/usr/bin/env python
"""Benchmark of the B200-native InstructAny2Pix denoising hot path.

    python bench.py --gpus N --steps K --warmup W [--workload c3|c2|c4] [--impl reference]

A "step" is ONE complete 50-step DDIM + CFG sampling trajectory for the per-GPU batch (the unit of BASELINE.json's
metric, images/sec at 1024^2): 50 x (CUDA-graph replay of the SDXL-class UNet at CFG batch 2B + fused CFG/DDIM kernel),
preceded by the per-request hoisted work (cross-attention K/V of all 70 layers, time-embedding table).  Workloads
(BASELINE.json configs): c3 = 1024^2 (128x128 latent) batch 4 [default: the config the metric is quoted on that fits one
GPU], c2 = 512^2 batch 1, c4 = prior + 1024^2 batch 8.  Weights are random-init of the named architecture, inputs
synthetic (no network).  Multi-GPU (torchrun, one rank per GPU): every rank samples its own batch -- whole trajectories
shard by prompt/seed, no collective inside the loop -- so scaling is "weak"; time = max over ranks.

JSON keys beyond the base contract: "roofline" (dominant kernel = tcgen05 implicit GEMM, measured with CUDA events around
every launch of one eager UNet forward), "cpu_baseline" (the oracle port of the reference path on the host cores, bounded
sample), "e2e" (same trajectory through the public API with pinned-host inputs and a device->host read of the result).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

F_UNET = {128: 6.7104e12, 64: 1.5367e12}          # FLOPs / sample-forward with K,V hoisted
WORKLOADS = {
    "c2": dict(L=64, B=1, prior=False, name="SDXL-class UNet 512^2 (64x64 latent) 50-step DDIM CFG, batch 1"),
    "c3": dict(L=128, B=4, prior=False, name="SDXL-class UNet 1024^2 (128x128 latent) 50-step DDIM CFG, decoupled image+text cross-attn, batch 4"),
    "c4": dict(L=128, B=8, prior=True, name="instruction-edit: prior + 1024^2 UNet sampling, batch 8 per GPU"),
    "b1": dict(L=128, B=1, prior=False, name="single interactive request: SDXL-class UNet 1024^2 50-step DDIM CFG, batch 1"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(burst=d["bf16_tflops"], sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), hbm=d["hbm_gbs"], src="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, src="fallback")


def ncu_traffic():
    """DRAM bytes (read + write) per tc_gemm_kernel launch, averaged over the 480 launches of one UNet step, from the committed
    ncu capture (profiles/gemm_traffic_r01.json; not measured live -- a run under ncu is never a bench value)."""
    p = os.path.join(ROOT, "profiles", "gemm_traffic_r01.json")
    if os.path.exists(p):
        return json.load(open(p))["dram_bytes_per_launch"]
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ model + inputs
def build_models(dev, want_prior):
    from instructany2pix_b200.attention_processor import B200IPAttnProcessor
    from instructany2pix_b200.unet import B200UNet

    g = torch.Generator(device=dev)
    g.manual_seed(0)
    unet = B200UNet(device=dev)                                  # SDXL-base config: 2 567 463 684 parameters
    procs = {}
    for name, p in unet.attn_processors.items():
        if name.endswith("attn2.processor"):
            hs = dict(unet.named_modules())[name[: -len(".processor")]].to_q.weight.shape[0]
            procs[name] = B200IPAttnProcessor(hs, unet.config.cross_attention_dim, scale=1.0, num_tokens=4, device=dev)
        else:
            procs[name] = p
    unet.set_attn_processor(procs)                               # +340 787 200 IP-adapter parameters
    for name, p in unet.named_parameters():                      # PyTorch-default-scale random init, on the device
        if p.ndim >= 2:
            fan_in = p[0].numel()
            p.copy_(((torch.rand(p.shape, generator=g, device=dev) * 2 - 1) * fan_in ** -0.5).to(p.dtype))
        elif name.endswith("weight"):
            p.fill_(1.0)
        else:
            p.copy_(0.02 * torch.randn(p.shape, generator=g, device=dev))
    unet.invalidate()
    prior = None
    if want_prior:
        from instructany2pix_b200.prior import B200Prior
        prior = B200Prior(device=dev)
        for name, p in prior.named_parameters():
            if p.ndim >= 2:
                p.copy_(0.02 * torch.randn(p.shape, generator=g, device=dev))
            elif name.endswith("weight"):
                p.fill_(1.0)
            else:
                p.zero_()
        prior.set_clip_hidden(0.5 * torch.randn(1, 2, 1024, generator=g, device=dev))
    return unet, prior


def build_vae(dev):
    """SDXL AutoencoderKL decoder (random init of the named architecture) for the end-to-end leg: latents -> images."""
    from instructany2pix_b200.vae import B200VAE
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    vae = B200VAE(device=dev, with_encoder=False)
    for name, p in vae.named_parameters():
        if p.ndim >= 2:
            fan_in = p[0].numel()
            p.copy_(((torch.rand(p.shape, generator=g, device=dev) * 2 - 1) * fan_in ** -0.5).to(p.dtype))
        elif name.endswith("weight"):
            p.fill_(1.0)
        else:
            p.zero_()
    vae.invalidate()
    return vae


def host_inputs(B, L, seed):
    """Synthetic per-request conditioning in PINNED host memory: what a caller of the public API holds."""
    g = torch.Generator()
    g.manual_seed(seed)
    pin = lambda t: t.pin_memory() if torch.cuda.is_available() else t
    ctx = torch.randn(2 * B, 77, 2048, generator=g)              # [neg ; pos] text tokens; the 4 IP tokens are projected per request
    pooled = torch.randn(2 * B, 1280, generator=g)
    H = float(L * 8)
    tid = torch.tensor([[H, H, 0.0, 0.0, H, H]]).repeat(2 * B, 1)
    lat = torch.randn(B, 4, L, L, generator=g)
    e = torch.randn(B, 1, 1024, generator=g)
    e = e / e.norm(dim=-1, keepdim=True) * 100.0
    return dict(ctx=pin(ctx), pooled=pin(pooled), tid=pin(tid), lat=pin(lat), llm=pin(e))


def run_trajectory(hot, dev_in, steps):
    """one request batch: (prior ->) LLM embedding -> ImageProj -> 4 IP tokens appended to the text tokens -> CFG sampling"""
    y = None
    if hot.prior is not None:
        y = hot.prior.generate_diffusion(3, 0, dev_in["llm"], device=dev_in["llm"].device, dtype=torch.float32,
                                         num_inference_steps=25, guidance_scale=10, score=6.5)[0]
    ctx = hot.ip_context(dev_in["ctx"], dev_in["llm"], prior_embed=y)
    return hot.sampler.generate(dev_in["lat"], ctx, dict(text_embeds=dev_in["pooled"], time_ids=dev_in["tid"]),
                                num_inference_steps=steps, guidance_scale=10.0)


# ------------------------------------------------------------------------------------------------ roofline of the dominant kernel
def profile_dominant_kernel(unet, sampler, dev_in, B, ctx81):
    """CUDA events around every C-ABI launch of ONE eager UNet forward (CFG batch 2B) on the launching stream."""
    from instructany2pix_b200 import ops
    added = dict(text_embeds=dev_in["pooled"], time_ids=dev_in["tid"])
    kv = unet.context_kv(ctx81)
    rb = unet.time_rowbias_table(torch.tensor([981.0]), added, 2 * B)[0].contiguous()
    x = dev_in["lat"].float()
    for _ in range(2):
        unet.forward_core(x, rb, kv, 2 * B)
    torch.cuda.synchronize()
    ops.PROFILE = []
    unet.forward_core(x, rb, kv, 2 * B)
    torch.cuda.synchronize()
    rec, ops.PROFILE = ops.PROFILE, None
    by, shapes = {}, {}
    for name, flops, e0, e1, tag, nbytes in rec:
        ms = e0.elapsed_time(e1)
        d = by.setdefault(name, dict(ms=0.0, flops=0.0, n=0, bytes=0.0))
        d["ms"] += ms
        d["flops"] += flops
        d["bytes"] += nbytes
        d["n"] += 1
        if tag:
            t = shapes.setdefault(tag, dict(ms=0.0, flops=0.0, n=0))
            t["ms"] += ms
            t["flops"] += flops
            t["n"] += 1
    if os.environ.get("IA2P_BENCH_SHAPES"):
        for tag, t in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"]):
            print(f"# {t['ms']:8.3f} ms  n={t['n']:3d}  {t['flops'] / max(t['ms'], 1e-9) / 1e9:7.1f} TFLOP/s  {tag}", file=sys.stderr)
    return by


# ------------------------------------------------------------------------------------------------ CPU baseline (oracle port)
_CPU_UNET = None


def _cpu_unet(threads):
    """the fp32 oracle restatement at SDXL-base width with decoupled cross-attention processors, built once per process"""
    global _CPU_UNET
    torch.set_num_threads(threads)
    if _CPU_UNET is not None:
        return _CPU_UNET
    from oracle.attention import IPAttnProcessor2_0
    from oracle.unet import SDXL_BASE, OracleUNet
    with torch.device("meta"):
        m = OracleUNet(SDXL_BASE)
    m = m.to_empty(device="cpu")
    procs = {}
    for name, p in m.attn_processors.items():
        if name.endswith("attn2.processor"):
            hs = dict(m.named_modules())[name[: -len(".processor")]].to_q.weight.shape[0]
            procs[name] = IPAttnProcessor2_0(hs, 2048)
        else:
            procs[name] = p
    m.set_attn_processor(procs)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.ndim >= 2:
                p.uniform_(-0.02, 0.02)
            elif n.endswith("weight"):
                p.fill_(1.0)
            else:
                p.zero_()
    _CPU_UNET = m
    return m


def cpu_unet_step_seconds(L, threads):
    """One CFG UNet step (2 sample-forwards, batch 1) + the CFG/DDIM arithmetic of the fp32 oracle restatement on the host
    cores (model construction excluded)."""
    m = _cpu_unet(threads)
    with torch.no_grad():
        x = torch.randn(2, 4, L, L)
        ctx = torch.randn(2, 81, 2048)
        added = dict(text_embeds=torch.randn(2, 1280), time_ids=torch.tensor([[L * 8.0, L * 8.0, 0, 0, L * 8.0, L * 8.0]] * 2))
        t0 = time.time()
        eps = m(x, torch.tensor(981), ctx, added_cond_kwargs=added)[0]
        eu, ec = eps.chunk(2)
        _ = 1.0 * x[:1] + 0.1 * (eu + 10.0 * (ec - eu))
        return time.time() - t0


def reference_arm(args, wl):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: diffusers is not installable
    here) on all host cores.  Each step = one CFG UNet step at batch 1 of the workload's resolution; images/sec
    extrapolates x50 steps (stated in `sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    L = wl["L"]
    ts = []
    for i in range(args.warmup + args.steps):
        dt = cpu_unet_step_seconds(L, cores)
        if i >= args.warmup:
            ts.append(dt)
    t = sum(ts) / len(ts)
    v = 1.0 / (50.0 * t)
    out = dict(metric="images/sec 1024^2 50-step DDIM CFG" if L == 128 else "images/sec 512^2 50-step DDIM CFG",
               value=v, unit="images/sec", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup, ms_per_step=t * 1e3 * 50,
               higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", impl="reference",
               config=dict(workload=wl["name"], latent=L, num_inference_steps=50, guidance_scale=10.0),
               cpu_baseline=dict(value=v, unit="images/sec", cores=cores, kind="port",
                                 sample=f"{len(ts)} x one CFG UNet step (2 sample-forwards, batch 1, {L}x{L} latent) of the fp32 "
                                        "oracle port, x50 extrapolated"),
               e2e=dict(value=v, unit="images/sec", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--num-inference-steps", type=int, default=50)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return reference_arm(args, wl)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG", "WARN")               # keep NCCL's version banner out of stdout: ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)

    from instructany2pix_b200 import ops
    L, B, NS = wl["L"], wl["B"], args.num_inference_steps
    unet, prior = build_models(dev, wl["prior"])
    vae = build_vae(dev)
    from instructany2pix_b200.hotpath import B200HotPath
    from instructany2pix_b200.image_proj import B200ImageProj
    proj = B200ImageProj(device=dev)                             # Linear(1024 -> 4 x 2048) + LayerNorm(2048), PyTorch-default init
    hot = B200HotPath(unet, vae, prior=prior, use_cuda_graph=not args.no_graph, image_proj=proj)
    sampler = hot.sampler
    host = host_inputs(B, L, seed=1000 + rank)                   # every rank samples different prompts/seeds
    dev_in = {k: v.to(dev) for k, v in host.items()}

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        lat_w = run_trajectory(hot, dev_in, NS)
    vae.decode(lat_w)                                            # warm-up of the decode path (function attributes, allocator)
    # ---- device-resident timing (value)
    sync_all()
    clocks = ClockSampler(local)
    clocks.start()
    l0 = ops.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run_trajectory(hot, dev_in, NS)
    e1.record()
    sync_all()
    clk = clocks.stop()
    ms = e0.elapsed_time(e1)
    launches_eager = ops.LAUNCHES - l0
    # ---- end-to-end through the public API: pinned host inputs -> device, result -> host, every step
    sync_all()
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        cur = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
        out = run_trajectory(hot, cur, NS)
        v0 = torch.cuda.Event(enable_timing=True)
        v0.record()
        img = vae.decode(out)                                    # sdxl_pipeline.py:859-871: latents -> (B,3,8L,8L) images
        res = img.to("cpu")                                      # device -> host read of the decoded images (syncs)
    f1.record()
    sync_all()
    ms_e2e = f0.elapsed_time(f1)
    ms_decode = v0.elapsed_time(f1)                              # last step's decode + image read-back
    d2h = res.numel() * res.element_size()
    if world > 1:
        tt = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e = tt.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    ms_step = ms / args.steps
    value = world * B * args.steps / (ms / 1e3)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    unet_step_ms = ms_step / NS
    step_flops = 2 * B * F_UNET[L]
    whole_frac = step_flops / (unet_step_ms * 1e-3) / (pk["sustained"] * 1e12)
    # launches: graph replays re-issue the captured kernels
    per_forward = getattr(sampler, "launches_per_forward", None)
    by = profile_dominant_kernel(unet, sampler, dev_in, B, hot.ip_context(dev_in["ctx"], dev_in["llm"]))
    n_forward_kernels = sum(d["n"] * ops._KERNELS_PER_CALL.get(k, 1) for k, d in by.items())
    gpu_launches = launches_eager + (0 if args.no_graph else args.steps * NS * n_forward_kernels)
    tc = dict(ms=0.0, flops=0.0, n=0, bytes=0.0)
    for k in ("ia2p_gemm_bf16", "ia2p_gemm_ln_bf16", "ia2p_conv3x3_nhwc_bf16", "ia2p_conv_up2x_nhwc_bf16"):
        if k in by:
            for f in tc:
                tc[f] += by[k][f] * (ops._KERNELS_PER_CALL.get(k, 1) if f == "n" else 1)
    total_ms = sum(d["ms"] for d in by.values())
    ach = tc["flops"] / (tc["ms"] * 1e-3) / 1e12 if tc["ms"] else 0.0
    roof = dict(kernel="tc_gemm_kernel (tcgen05 implicit GEMM: linears + 3x3 convs)", bound="tensor", achieved=ach,
                peak=pk["sustained"], unit="TFLOP/s", frac=ach / pk["sustained"], traffic=ncu_traffic(), peak_source=pk["src"] + " bf16_tflops_sustained",
                algorithmic_bytes_per_launch=tc["bytes"] / max(tc["n"], 1), algorithmic_flops_per_launch=tc["flops"] / max(tc["n"], 1),
                launches_per_forward=tc["n"], avg_launch_ms=tc["ms"] / max(tc["n"], 1), share_of_forward=tc["ms"] / total_ms if total_ms else None,
                forward_breakdown_ms={k: round(d["ms"], 3) for k, d in sorted(by.items(), key=lambda kv: -kv[1]["ms"])})
    out = dict(metric="images/sec 1024^2 50-step DDIM CFG" if L == 128 else "images/sec 512^2 50-step DDIM CFG",
               value=value, unit="images/sec", n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms_step,
               higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
               config=dict(workload=wl["name"], per_gpu_batch=B, latent=L, num_inference_steps=NS, guidance_scale=10.0,
                           parallelism=f"replicas x{world} (whole trajectories per GPU, no in-loop collective)",
                           l2_policy="inputs larger than L2 (5.8 GB weights + activations stream every step); no explicit flush",
                           cuda_graph=not args.no_graph, residual_stream="fp32"),
               unet_step_ms=unet_step_ms, unet_tensor_frac=whole_frac, clocks=clk,
               e2e=dict(value=e2e_value, unit="images/sec", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                        includes="H2D of the text conditioning + LLM embedding + noise, (prior,) image projector -> IP tokens, 50-step trajectory, VAE decode to fp32 images, D2H of the images",
                        vae_decode_and_readback_ms=ms_decode),
               gpu_launches=int(gpu_launches), roofline=roof)
    if not args.no_cpu_baseline and world == 1:                 # contract: rank 0 at N = 1 only
        cores = os.cpu_count() or 1
        t = cpu_unet_step_seconds(L, cores)
        out["cpu_baseline"] = dict(value=1.0 / (50.0 * t), unit="images/sec", cores=cores, kind="port",
                                   sample=f"1 CFG UNet step (2 sample-forwards, batch 1, {L}x{L} latent) of the fp32 oracle port "
                                          f"= {t:.1f} s, x50 extrapolated")
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
