"""Oracle (test infrastructure): restated attention processors + projector.

Restates, for the only configuration the hot path reaches (no spatial_norm, no
group_norm, no norm_cross, attention_mask None, residual_connection False,
rescale_output_factor 1):
  * AttnProcessor2_0.__call__     diffusion/ip_adapter/attention_processor.py:205-279
  * IPAttnProcessor2_0.__call__   diffusion/ip_adapter/attention_processor.py:310-412
  * ImageProjModel.forward        diffusion/ip_adapter/ip_adapter.py:42-67
Pinned against the reference code itself by tests/golden/attn_*.npz and
tests/golden/image_proj.npz (made by oracle/gen_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _heads(x, b, h):
    return x.view(b, -1, h, x.shape[-1] // h).transpose(1, 2)


class AttnProcessor2_0(nn.Module):
    def __init__(self, hidden_size=None, cross_attention_dim=None):
        super().__init__()

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, *a, **k):
        assert attention_mask is None
        b = hidden_states.shape[0]
        q = attn.to_q(hidden_states)
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k_ = attn.to_k(ctx)
        v_ = attn.to_v(ctx)
        o = F.scaled_dot_product_attention(_heads(q, b, attn.heads), _heads(k_, b, attn.heads), _heads(v_, b, attn.heads))
        o = o.transpose(1, 2).reshape(b, -1, q.shape[-1]).to(q.dtype)
        o = attn.to_out[0](o)
        o = attn.to_out[1](o)
        return o / attn.rescale_output_factor


class IPAttnProcessor2_0(nn.Module):
    def __init__(self, hidden_size, cross_attention_dim=None, scale=1.0, num_tokens=4):
        super().__init__()
        self.hidden_size = hidden_size
        self.cross_attention_dim = cross_attention_dim
        self.scale = scale
        self.num_tokens = num_tokens
        self.to_k_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)
        self.to_v_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False)

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None, *a, **k):
        assert attention_mask is None and encoder_hidden_states is not None
        b = hidden_states.shape[0]
        q = attn.to_q(hidden_states)
        end_pos = encoder_hidden_states.shape[1] - self.num_tokens       # :350
        text, ip = encoder_hidden_states[:, :end_pos], encoder_hidden_states[:, end_pos:]
        h = attn.heads
        qh = _heads(q, b, h)
        o = F.scaled_dot_product_attention(qh, _heads(attn.to_k(text), b, h), _heads(attn.to_v(text), b, h))
        o = o.transpose(1, 2).reshape(b, -1, q.shape[-1]).to(q.dtype)
        ip_k = _heads(self.to_k_ip(ip), b, h)
        ip_v = _heads(self.to_v_ip(ip), b, h)
        o_ip = F.scaled_dot_product_attention(qh, ip_k, ip_v)
        o_ip = o_ip.transpose(1, 2).reshape(b, -1, q.shape[-1]).to(q.dtype)
        o = o + self.scale * o_ip                                        # :397
        o = attn.to_out[0](o)
        o = attn.to_out[1](o)
        return o / attn.rescale_output_factor


class Attention(nn.Module):
    """Duck-type of diffusers ``Attention`` carrying exactly what the processors read
    (attention_processor.py:322-410); bias layout per SURVEY.md A.3."""

    def __init__(self, query_dim, heads, dim_head=64, cross_attention_dim=None):
        super().__init__()
        inner = heads * dim_head
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_attention_dim or query_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim, bias=True), nn.Dropout(0.0)])
        self.spatial_norm = None
        self.group_norm = None
        self.norm_cross = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        self.processor = AttnProcessor2_0()

    def prepare_attention_mask(self, *a, **k):  # never reached: attention_mask is None on this path
        raise NotImplementedError

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states,
                              attention_mask=attention_mask, **kw)


class ImageProjModel(nn.Module):
    """Restatement of ImageProjModel (ip_adapter.py:28-67) without in-place slicing."""

    def __init__(self, cross_attention_dim=2048, clip_embeddings_dim=1024, clip_extra_context_tokens=4, num_crops=2):
        super().__init__()
        self.cross_attention_dim = cross_attention_dim
        self.clip_extra_context_tokens = clip_extra_context_tokens
        self.proj = nn.Linear(clip_embeddings_dim, clip_extra_context_tokens * cross_attention_dim)
        self.norm = nn.LayerNorm(cross_attention_dim)
        self.raw_embed = nn.Parameter(torch.zeros(2, cross_attention_dim))
        self.num_crops = num_crops

    def forward(self, image_embeds, mode, scales=(1.0, 1.0)):
        bs = image_embeds.shape[0]
        t = self.proj(image_embeds).reshape(bs, self.num_crops, self.clip_extra_context_tokens, self.cross_attention_dim)
        g = t[:, 0:1]
        l = g * (1 - scales[1]) + t[:, 1:] * scales[1]
        g = g + self.raw_embed[0][None, None]
        l = l + self.raw_embed[1][None, None]
        if mode == "global":
            t = g
        elif mode == "local":
            t = l
        else:
            assert mode == "both", f"Invalid Mode {mode}"
            t = torch.cat([g, l], dim=1)
        t = t.reshape(bs, -1, self.cross_attention_dim)
        return self.norm(t)


def get_image_embeds(image_proj_model, clip_image_embeds=None, clip_image_embeds_local=None, mode="global", scale_g=1.0,
                     scale_l=1.0):
    """IPAdapter.get_image_embeds with the CLIP embeddings given (ip_adapter.py:171-209): a missing crop := zeros,
    uncond := projector(zeros) with the default scales.  Pinned by tests/golden/image_proj.npz (k64/gie_*)."""
    if clip_image_embeds is None:
        clip_image_embeds = torch.zeros_like(clip_image_embeds_local)
    elif clip_image_embeds_local is None:
        clip_image_embeds_local = torch.zeros_like(clip_image_embeds)
    e = torch.stack([clip_image_embeds, clip_image_embeds_local], dim=1)
    return image_proj_model(e, mode=mode, scales=[scale_g, scale_l]), image_proj_model(torch.zeros_like(e), mode=mode)
