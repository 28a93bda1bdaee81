#!/usr/bin/env python
"""Micro-benchmark of the hot kernels at the c3 (1024^2, batch 4 -> CFG batch 8) shapes.  CUDA-event timing, L2 flushed
between iterations by cycling through enough distinct buffers.  Usage: python tools/bench_kernels.py [gemm|conv|attn|norm|all]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from instructany2pix_b200 import ops
from instructany2pix_b200.packing import pack_conv3x3

dev = "cuda"
BF = torch.bfloat16


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn(0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_gemm():
    shapes = [("geglu 1280", 8192, 10240, 1280, "geglu"), ("ff_out 1280", 8192, 1280, 5120, "res32"), ("qkv 1280", 8192, 3840, 1280, "bf16"),
              ("out_proj 1280", 8192, 1280, 1280, "res32"), ("to_q 1280", 8192, 1280, 1280, "bf16"),
              ("geglu 640", 32768, 5120, 640, "geglu"), ("ff_out 640", 32768, 640, 2560, "res32"), ("qkv 640", 32768, 1920, 640, "bf16"),
              ("out_proj 640", 32768, 640, 640, "res32")]
    for name, M, N, K, mode in shapes:
        nb = max(2, int(3e8 // (M * K * 2 + N * K * 2)) + 1)          # cycle > L2 worth of operands
        a = [torch.randn(M, K, device=dev).to(BF) for _ in range(nb)]
        w = [(torch.randn(N, K, device=dev) * K ** -0.5).to(BF) for _ in range(nb)]
        bias = torch.randn(N, device=dev)
        if mode == "geglu":
            f = lambda i: ops.gemm(a[i % nb], w[i % nb], bias=bias, geglu=True)
        elif mode == "res32":
            res = [torch.randn(M, N, device=dev) for _ in range(nb)]
            out = torch.empty(M, N, device=dev)
            f = lambda i: ops.gemm(a[i % nb], w[i % nb], bias=bias, residual=res[i % nb], out=out)
        else:
            f = lambda i: ops.gemm(a[i % nb], w[i % nb])
        ms = timeit(f)
        print(f"gemm {name:14s} M={M:6d} N={N:6d} K={K:5d} {mode:6s}: {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s")


def bench_mainloop():
    """main-loop rate in isolation: exactly one or two tiles per CTA, K so long that the epilogue is negligible"""
    for name, M, N, K in [("1 tile/CTA", 128 * 148, 256, 16384), ("2 tiles/CTA", 128 * 148, 512, 8192), ("square 8192", 8192, 8192, 8192)]:
        a = torch.randn(M, K, device=dev).to(BF)
        w = (torch.randn(N, K, device=dev) * K ** -0.5).to(BF)
        ms = timeit(lambda i: ops.gemm(a, w), iters=10)
        print(f"mainloop {name:12s} M={M:6d} N={N:6d} K={K:5d}: {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s")
    a = torch.randn(8192, 8192, device=dev).to(BF)
    w = torch.randn(8192, 8192, device=dev).to(BF)
    ms = timeit(lambda i: torch.matmul(a, w.t()), iters=10)
    print(f"cuBLAS square 8192: {ms * 1e3:8.1f} us  {2.0 * 8192 ** 3 / ms / 1e9:7.1f} TFLOP/s")


def bench_vae():
    """SDXL VAE at 1024^2 (random init): decode of B latents, encode of B images; FLOPs counted analytically from the convs/linears"""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    vae = bench.build_vae(torch.device("cuda", 0))
    for B in (1, 4):
        lat = torch.randn(B, 4, 128, 128, device=dev)
        ops.PROFILE = []
        vae.decode(lat)
        torch.cuda.synchronize()
        rec, ops.PROFILE = ops.PROFILE, None
        flops = sum(r[1] for r in rec)
        ms = timeit(lambda i: vae.decode(lat), iters=5, warm=2)
        print(f"vae decode 1024^2 B={B}: {ms:7.2f} ms  {flops / ms / 1e9:7.1f} TFLOP/s (tensor-core FLOPs {flops / 1e12:.2f} T)  "
              f"peak mem {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB")


def bench_lnfold():
    """consumer GEMMs with and without the LayerNorm fold, producer GEMM with and without the LN outputs."""
    M, C = 8192, 1280
    a = [torch.randn(M, C, device=dev).to(BF) for _ in range(6)]
    st = torch.rand(M, 10, 2, device=dev) + 1.0
    for name, N, geglu in [("geglu", 10240, True), ("qkv", 3840, False), ("to_q", 1280, False)]:
        w = [(torch.randn(N, C, device=dev) * C ** -0.5).to(BF) for _ in range(6)]
        bias, c1 = torch.randn(N, device=dev), torch.randn(N, device=dev)
        t0 = timeit(lambda i: ops.gemm(a[i % 6], w[i % 6], bias=bias, geglu=geglu))
        t1 = timeit(lambda i: ops.gemm(a[i % 6], w[i % 6], bias=bias, geglu=geglu, ln=(st, c1, 1e-5)))
        print(f"consumer {name:6s}: plain {t0 * 1e3:7.1f} us   ln-fold {t1 * 1e3:7.1f} us")
    w = [(torch.randn(C, C, device=dev) * C ** -0.5).to(BF) for _ in range(6)]
    res = [torch.randn(M, C, device=dev) for _ in range(6)]
    bias = torch.randn(C, device=dev)
    t0 = timeit(lambda i: ops.gemm(a[i % 6], w[i % 6], bias=bias, residual=res[i % 6], out_dtype=torch.float32))
    t1 = timeit(lambda i: ops.gemm(a[i % 6], w[i % 6], bias=bias, residual=res[i % 6], out_dtype=torch.float32, want_ln=True))
    print(f"producer out_proj: plain {t0 * 1e3:7.1f} us   +bf16 copy +stats {t1 * 1e3:7.1f} us")


def bench_conv():
    for name, B, H, Cin, Cout in [("1280@32", 8, 32, 1280, 1280), ("640@64", 8, 64, 640, 640), ("320@128", 8, 128, 320, 320),
                                  ("2560->1280@32", 8, 32, 2560, 1280), ("960->320@128", 8, 128, 960, 320)]:
        x = [torch.randn(B, H, H, Cin, device=dev).to(BF) for _ in range(3)]
        w = pack_conv3x3(torch.randn(Cout, Cin, 3, 3, device=dev) * (9 * Cin) ** -0.5)
        bias = torch.randn(Cout, device=dev)
        ms = timeit(lambda i: ops.conv3x3(x[i % 3], w, Cout, bias=bias, out_dtype=torch.float32))
        print(f"conv {name:14s} B={B} HxW={H}x{H}: {ms * 1e3:8.1f} us  {2.0 * B * H * H * Cout * 9 * Cin / ms / 1e9:7.1f} TFLOP/s")


def bench_attn():
    for name, B, N, heads in [("lvl1", 8, 4096, 10), ("lvl2", 8, 1024, 20)]:
        qkv = [torch.randn(B * N, 3 * heads * 64, device=dev).to(BF) for _ in range(3)]
        ms = timeit(lambda i: ops.flash_self_attn(qkv[i % 3], B, N, heads))
        print(f"self-attn {name} B={B} N={N} h={heads}: {ms * 1e3:8.1f} us  {4.0 * B * heads * N * N * 64 / ms / 1e9:7.1f} TFLOP/s")
    B, N, heads = 8, 1024, 20
    q = torch.randn(B * N, heads * 64, device=dev).to(BF)
    kvt = torch.randn(B * 77, 2 * heads * 64, device=dev).to(BF)
    kvi = torch.randn(B * 4, 2 * heads * 64, device=dev).to(BF)
    ms = timeit(lambda i: ops.cross_attn(q, kvt, 77, kvi, 4, 1.0, B, N, heads))
    print(f"cross-attn lvl2: {ms * 1e3:8.1f} us  ({(2 * q.numel() * 2) / ms / 1e6:7.1f} GB/s of Q read + O write)  [eager launches: host-bound]")
    for name, B, N, heads in [("lvl2", 8, 1024, 20), ("lvl1", 8, 4096, 10), ("c2 lvl2", 2, 256, 20)]:
        qs = [torch.randn(B * N, heads * 64, device=dev).to(BF) for _ in range(8)]       # > L2 at the c3 shapes
        kvt = torch.randn(B * 77, 2 * heads * 64, device=dev).to(BF)
        kvi = torch.randn(B * 4, 2 * heads * 64, device=dev).to(BF)
        outs = [torch.empty_like(t) for t in qs]
        ops.cross_attn(qs[0], kvt, 77, kvi, 4, 1.0, B, N, heads, out=outs[0])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for t, o in zip(qs, outs):
                ops.cross_attn(t, kvt, 77, kvi, 4, 1.0, B, N, heads, out=o)
        ms = timeit(lambda i: g.replay(), iters=10) / len(qs)
        print(f"cross-attn {name} in a CUDA graph: {ms * 1e3:8.1f} us  ({(2 * qs[0].numel() * 2) / ms / 1e6:7.1f} GB/s of Q read + O write)")


def bench_norm():
    for name, rows, cols in [("ln 1280", 8192, 1280), ("ln 640", 32768, 640)]:
        x = [torch.randn(rows, cols, device=dev) for _ in range(4)]
        g, b = torch.ones(cols, device=dev), torch.zeros(cols, device=dev)
        ms = timeit(lambda i: ops.layernorm(x[i % 4], g, b, 1e-5, out_dtype=BF))
        print(f"{name}: {ms * 1e3:8.1f} us  {rows * cols * 6 / ms / 1e6:7.1f} GB/s")
    for name, B, HW, C in [("gn 320@128", 8, 16384, 320), ("gn 1280@32", 8, 1024, 1280), ("gn 640@64", 8, 4096, 640)]:
        x = [torch.randn(B, HW, C, device=dev) for _ in range(4)]
        g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        ms = timeit(lambda i: ops.groupnorm(x[i % 4], None, g, b, 32, 1e-5, True))
        print(f"{name}: {ms * 1e3:8.1f} us  {B * HW * C * 6 / ms / 1e6:7.1f} GB/s (algorithmic 4 B read + 2 B write)")


def bench_elementwise():
    """the fused CFG + DDIM update and the inverse-DDIM axpby: 128-bit loads / stores.  At the c3 latent size (4 x 4 x 128 x 128
    fp32: 2 eps reads + x read + x write = 1 MiB) a launch is latency-bound, so the kernel's own bandwidth is shown on a size that
    is not (same kernel, 256x the elements, buffers cycled through > L2)."""
    for name, B, L, rep in [("c3 latents", 4, 128, 1), ("256 x c3 latents", 4, 128, 256)]:
        n = B * 4 * L * L * rep
        nb = 1 if rep == 1 else 3
        eps2 = [torch.randn(2 * n, device=dev) for _ in range(nb)]
        x = [torch.randn(n, device=dev) for _ in range(nb)]
        out = torch.empty(n, device=dev)
        f = lambda i: ops.cfg_ddim_step(eps2[i % nb].view(2, -1), x[i % nb].view(1, -1), 10.0, 0.99, -0.1, x_out=out.view(1, -1))
        g = torch.cuda.CUDAGraph()
        f(0)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            for i in range(nb * 4):
                f(i)
        ms = timeit(lambda i: g.replay(), iters=10) / (nb * 4)
        print(f"cfg_ddim_step {name:18s}: {ms * 1e3:8.2f} us  {16.0 * n / ms / 1e6:8.1f} GB/s (2 eps reads + x read + x write, fp32)")
        f2 = lambda i: ops.axpby(eps2[i % nb][:n], x[i % nb], 0.99, -0.1, out=out)
        g2 = torch.cuda.CUDAGraph()
        f2(0)
        torch.cuda.synchronize()
        with torch.cuda.graph(g2):
            for i in range(nb * 4):
                f2(i)
        ms = timeit(lambda i: g2.replay(), iters=10) / (nb * 4)
        print(f"axpby         {name:18s}: {ms * 1e3:8.2f} us  {12.0 * n / ms / 1e6:8.1f} GB/s (eps read + x read + x write, fp32)")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.manual_seed(0)
    for k, fn in (("gemm", bench_gemm), ("conv", bench_conv), ("attn", bench_attn), ("norm", bench_norm), ("elementwise", bench_elementwise), ("lnfold", bench_lnfold), ("mainloop", bench_mainloop), ("vae", bench_vae)):
        if which in (k, "all"):
            fn()
