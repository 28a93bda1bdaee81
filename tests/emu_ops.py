"""TEST DOUBLE for ``instructany2pix_b200.ops`` (never importable from the product package).

Restates each C-ABI op in plain torch on CPU with the kernels' rounding points (bf16 outputs, fp32 accumulation) so the
host-side orchestration -- weight packing, tap/K ordering, K/V hoisting, processor plumbing, sampler bookkeeping -- can
be checked against the oracle in the GPU-less ``-m "not gpu"`` suite.  Installed by the ``emu`` fixture via monkeypatch.
"""
import math

import torch
import torch.nn.functional as F

ACT_NONE, ACT_GELU_NEW, ACT_SILU = 0, 1, 2
BF = torch.bfloat16


class IA2PError(RuntimeError):
    pass


def require_cuda(t, what):
    return None


class pdl:
    def __init__(self, on=True):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _act(h, act):
    if act == ACT_GELU_NEW:
        return 0.5 * h * (1 + torch.tanh(math.sqrt(2 / math.pi) * (h + 0.044715 * h ** 3)))
    if act == ACT_SILU:
        return F.silu(h)
    return h


def cfg_ddim_step(eps2, x, g, c_x, c_e, x_out=None, x_in_next2=None):
    eu, ec = eps2.float().chunk(2)
    v = (c_x * x.float() + c_e * (eu + g * (ec - eu))).to(x.dtype)
    if x_in_next2 is not None:
        x_in_next2.copy_(torch.cat([v, v]).to(x_in_next2.dtype))
    if x_out is not None:
        x_out.copy_(v)
        return x_out
    return v


def axpby(eps, x, c_x, c_e, out=None):
    v = (c_x * x.float() + c_e * eps.float()).to(x.dtype)
    if out is not None:
        out.copy_(v)
        return out
    return v


def polar_interpolate(x, y, alpha, out=None):
    """pipeline.py:295-300 over the WHOLE tensor it is given"""
    x, y = x.float(), y.float()
    ll = x * alpha + y * (1 - alpha)
    v = ll / ll.norm() * (x.norm() * alpha + y.norm() * (1 - alpha))
    if out is not None:
        out.copy_(v)
        return out
    return v


def inpaint_blend(latents, orig, noise, mask, c_x, c_e, out=None):
    ref = c_x * orig + (c_e * noise if noise is not None else 0.0)
    v = (1 - mask) * ref + mask * latents
    if out is not None:
        out.copy_(v)
        return out
    return v


def prior_cfg_ddpm_step(x0_pair, x, noise, sqrt_a, sqrt_1ma, g, c_x0, c_x, sigma, out=None):
    x0_pair = x0_pair.reshape(2, -1)
    xf = x.reshape(-1)
    ec, eu = (xf - sqrt_a * x0_pair[0]) / sqrt_1ma, (xf - sqrt_a * x0_pair[1]) / sqrt_1ma
    e = eu + g * (ec - eu)
    v = c_x0 * (xf - sqrt_1ma * e) / sqrt_a + c_x * xf
    if noise is not None:
        v = v + sigma * noise.reshape(-1)
    v = v.reshape(x.shape)
    if out is not None:
        out.copy_(v)
        return out
    return v


def timestep_embedding(t, dim, flip_sin_to_cos=True, shift=0.0, dtype=torch.float32):
    half = dim // 2
    f = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / (half - shift))
    a = t.reshape(-1).float()[:, None] * f[None]
    e = torch.cat([a.sin(), a.cos()], -1)
    if flip_sin_to_cos:
        e = torch.cat([e[:, half:], e[:, :half]], -1)
    return e.to(dtype)


def upsample2x(x):
    return x.repeat_interleave(2, 1).repeat_interleave(2, 2).to(BF)


def to_bf16(x):
    return x.to(BF)


def groupnorm(xa, xb, gamma, beta, groups, eps, silu, want_raw=False):
    x = xa if xb is None else torch.cat([xa, xb], -1)
    shp = x.shape
    v = x.float().reshape(shp[0], -1, shp[-1]).permute(0, 2, 1)
    y = F.group_norm(v, groups, gamma, beta, eps).permute(0, 2, 1).reshape(shp)
    if silu:
        y = F.silu(y)
    if want_raw:
        return y.to(BF), x.to(BF)
    return y.to(BF)


def layernorm(x, gamma, beta, eps, out_dtype=None):
    return F.layer_norm(x.float(), (x.shape[-1],), gamma, beta, eps).to(out_dtype or x.dtype)


def gemm(a, w, bias=None, a2=None, rowbias=None, rows_per_batch=0, residual=None, geglu=False, out=None, out_dtype=BF,
         want_ln=False, ln=None, want_colstats=False):
    assert a.dtype == BF and w.dtype == BF
    A = a.float() if a2 is None else torch.cat([a, a2], 1).float()
    h = A @ w.float().t()
    if ln is not None:                                  # consumer side of the LayerNorm fold
        stats, c1, eps = ln
        s = stats.sum(1)
        mean = s[:, 0:1] / A.shape[1]
        rstd = torch.rsqrt((s[:, 1:2] / A.shape[1] - mean * mean).clamp_min(0) + eps)
        h = rstd * (h - mean * c1[None])
    if bias is not None:
        h = h + bias
    if rowbias is not None:
        h = h + rowbias.repeat_interleave(rows_per_batch, 0)[: h.shape[0]]
    if geglu:
        n = h.shape[1]
        hv = h.reshape(-1, n // 64, 2, 32)
        h = (hv[:, :, 0] * F.gelu(hv[:, :, 1])).reshape(-1, n // 2)
    if residual is not None:
        h = h + residual.float()
    if want_ln:                                         # producer side: bf16 copy + per-row partial sums (2 parts here)
        half = h.shape[1] // 2
        stats = torch.stack([torch.stack([h[:, :half].sum(1), (h[:, :half] ** 2).sum(1)], -1),
                             torch.stack([h[:, half:].sum(1), (h[:, half:] ** 2).sum(1)], -1)], 1)
        return h.to(out_dtype), h.to(BF), stats
    h = h.to(out_dtype)
    if out is not None:
        out.copy_(h)
        return out
    return h


def conv3x3(x, w, cout, stride=1, sc_a=None, sc_b=None, bias=None, rowbias=None, residual=None, out_dtype=BF, want_colstats=False):
    assert x.dtype == BF and (sc_a is None or sc_a.dtype == BF) and (sc_b is None or sc_b.dtype == BF)
    B, H, W, Cin = x.shape
    w3 = w[:, : 9 * Cin].float().reshape(cout, 3, 3, Cin).permute(0, 3, 1, 2)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w3, stride=stride, padding=1).permute(0, 2, 3, 1)
    if sc_a is not None:
        sc = sc_a if sc_b is None else torch.cat([sc_a, sc_b], -1)
        y = y + sc.float() @ w[:, 9 * Cin:].float().t()
    if bias is not None:
        y = y + bias
    if rowbias is not None:
        y = y + rowbias[:, None, None, :]
    if residual is not None:
        y = y + residual.float()
    return y.to(out_dtype)


def conv_up2x(x, w4, cout, bias=None, want_colstats=False):
    """four parity 2x2 convs over the low-res map with the pre-summed bf16 weights, exactly as the kernel evaluates them"""
    assert x.dtype == BF and w4.dtype == BF
    B, H, W, Cin = x.shape
    xp = F.pad(x.float().permute(0, 3, 1, 2), (1, 1, 1, 1))                      # index (y + 1, x + 1)
    out = torch.zeros(B, 2 * H, 2 * W, cout)
    for py in range(2):
        for px in range(2):
            wk = w4[py * 2 + px].float().reshape(cout, 2, 2, Cin).permute(0, 3, 1, 2)      # taps (i, j)
            y0, x0 = (0 if py == 0 else 1), (0 if px == 0 else 1)                          # first tap offset (-1 or 0) + pad 1
            win = xp[:, :, y0:y0 + H + 1, x0:x0 + W + 1]
            out[:, py::2, px::2, :] = F.conv2d(win, wk).permute(0, 2, 3, 1)
    if bias is not None:
        out = out + bias
    return out


def conv_out_tc(x, w32, b32, cout):
    y = conv3x3(x, w32, 32, bias=b32, out_dtype=torch.float32)
    return y[..., :cout].permute(0, 3, 1, 2).contiguous()


def conv_in(x_nchw, w, bias, out_batch=None, out_dtype=BF):
    B = out_batch or x_nchw.shape[0]
    x = x_nchw.float().repeat(B // x_nchw.shape[0], 1, 1, 1)
    return F.conv2d(x, w, bias, padding=1).permute(0, 2, 3, 1).contiguous().to(out_dtype)


def conv_out(x, w, bias, out_dtype=torch.float32):
    return F.conv2d(x.float().permute(0, 3, 1, 2), w.permute(0, 3, 1, 2), bias, padding=1).to(out_dtype)


def _heads(t, b, n, h):
    return t.float().reshape(b, n, h, 64).transpose(1, 2)


def flash_self_attn(qkv, batch, n_tokens, heads, out=None):
    C = heads * 64
    q, k, v = qkv.split(C, dim=1)
    o = F.scaled_dot_product_attention(_heads(q, batch, n_tokens, heads), _heads(k, batch, n_tokens, heads),
                                       _heads(v, batch, n_tokens, heads))
    return o.transpose(1, 2).reshape(batch * n_tokens, C).to(BF)


def cross_attn(q, kv_text, n_text, kv_ip, n_ip, ip_scale, batch, n_q, heads, out=None):
    C = heads * 64
    qh = _heads(q, batch, n_q, heads)
    o = F.scaled_dot_product_attention(qh, _heads(kv_text[:, :C], batch, n_text, heads), _heads(kv_text[:, C:], batch, n_text, heads))
    if n_ip:
        o = o + ip_scale * F.scaled_dot_product_attention(qh, _heads(kv_ip[:, :C], batch, n_ip, heads),
                                                          _heads(kv_ip[:, C:], batch, n_ip, heads))
    return o.transpose(1, 2).reshape(batch * n_q, C).to(BF)


def gemm_smallm(a, w, bias=None, residual=None, act_in=ACT_NONE, act=ACT_NONE, out=None):
    h = _act(a.float(), act_in) @ w.float().t()
    if bias is not None:
        h = h + bias
    h = _act(h, act)
    if residual is not None:
        h = h + residual
    return h


def causal_attn_small(qkv, batch, T, heads, out=None):
    E = heads * 64
    q, k, v = [t.reshape(batch, T, heads, 64).transpose(1, 2) for t in qkv.reshape(batch, T, 3 * E).split(E, dim=2)]
    return F.scaled_dot_product_attention(q, k, v, is_causal=True).transpose(1, 2).reshape(batch, T, E)


def prior_trunk_workspace(rows, device):
    return torch.zeros(256, dtype=torch.uint8, device=device)


def prior_trunk(seq, wpe, layers, lnf_g, lnf_b, heads, workspace, out=None, cache=None):
    """the persistent GPT-2 trunk kernel, restated op by op (fp32)"""
    B2, T, E = seq.shape
    h = (seq.float() + wpe[:T].float()[None]).reshape(B2 * T, E)
    for wqkv, wo, wfc, wpr, bqkv, bo, bfc, bpr, g1, b1, g2, b2 in layers:
        qkv = gemm_smallm(layernorm(h, g1, b1, 1e-5), wqkv, bias=bqkv)
        att = causal_attn_small(qkv, B2, T, heads).reshape(B2 * T, E)
        h = gemm_smallm(att, wo, bias=bo, residual=h)
        f = gemm_smallm(layernorm(h, g2, b2, 1e-5), wfc, bias=bfc, act=ACT_GELU_NEW)
        h = gemm_smallm(f, wpr, bias=bpr, residual=h)
    return layernorm(h.reshape(B2, T, E)[:, -1].contiguous(), lnf_g, lnf_b, 1e-5)
