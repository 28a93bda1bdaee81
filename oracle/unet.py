"""Oracle (test infrastructure): restated diffusers==0.26.3 ``UNet2DConditionModel``
for the SDXL-base configuration family.

Third-party algorithm (not under /root/reference; pinned at requirements.txt:3),
restated from its published semantics.  Reference call sites:
ddim/pnp_pipeline.py:253-260, :465-483; diffusion/ip_adapter/custom_pipelines.py:
338-345; ddim/sdxl_pipeline.py:832-839.  Attention math is delegated to the
(restated, golden-pinned) in-tree processors in :mod:`oracle.attention`.
Structural known answers: parameter count 2 567 463 684 for ``SDXL_BASE`` and
140 attention processors (tests/test_oracle_structure.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .attention import Attention, AttnProcessor2_0
from .schedulers import get_timestep_embedding


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    sample_size: int = 128
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 2, 10)   # level 0 has no attention (DownBlock2D/UpBlock2D)
    attention_head_dim: Tuple[int, ...] = (5, 10, 20)           # = head COUNTS
    cross_attention_dim: int = 2048
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    time_cond_proj_dim: object = None
    addition_embed_type: str = "text_time"
    down_block_types: Tuple[str, ...] = ("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D")
    up_block_types: Tuple[str, ...] = ("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D")

    @property
    def time_embed_dim(self):
        return self.block_out_channels[0] * 4

    def get(self, k, default=None):
        return getattr(self, k, default)


SDXL_BASE = UNetConfig()

# small same-topology config for quick CPU/GPU parity runs (channels stay multiples of 64: kernel K-chunk)
TINY = UNetConfig(sample_size=32, block_out_channels=(64, 128, 256), transformer_layers_per_block=(1, 1, 2),
                  attention_head_dim=(1, 2, 4), cross_attention_dim=256, addition_time_embed_dim=32,
                  projection_class_embeddings_input_dim=6 * 32 + 128)


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_dim, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, emb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(emb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class GEGLU(nn.Module):
    def __init__(self, dim, inner):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, ctx_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads, dim // heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, dim // heads, cross_attention_dim=ctx_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, ctx, cross_attention_kwargs=None):
        kw = cross_attention_kwargs or {}
        x = x + self.attn1(self.norm1(x), **kw)
        x = x + self.attn2(self.norm2(x), encoder_hidden_states=ctx, **kw)
        x = x + self.ff(self.norm3(x))
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, dim, heads, depth, ctx_dim, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(dim, heads, ctx_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(dim, dim)

    def forward(self, x, ctx, cross_attention_kwargs=None):
        b, c, h, w = x.shape
        res = x
        t = self.norm(x).permute(0, 2, 3, 1).reshape(b, h * w, c)
        t = self.proj_in(t)
        for blk in self.transformer_blocks:
            t = blk(t, ctx, cross_attention_kwargs)
        t = self.proj_out(t)
        return t.reshape(b, h, w, c).permute(0, 3, 1, 2) + res


class Downsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cfg, cin, cout, depth, heads, add_down):
        super().__init__()
        n = cfg.layers_per_block
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, cfg.time_embed_dim,
                                                    cfg.norm_num_groups, cfg.norm_eps) for i in range(n)])
        if depth:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, depth, cfg.cross_attention_dim,
                                                                cfg.norm_num_groups) for _ in range(n)])
        self.has_attn = bool(depth)
        if add_down:
            self.downsamplers = nn.ModuleList([Downsample2D(cout)])
        self.add_down = add_down

    def forward(self, x, emb, ctx, kw):
        outs = []
        for i, r in enumerate(self.resnets):
            x = r(x, emb)
            if self.has_attn:
                x = self.attentions[i](x, ctx, kw)
            outs.append(x)
        if self.add_down:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, cfg, c, depth, heads):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps)
                                      for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(c, heads, depth, cfg.cross_attention_dim, cfg.norm_num_groups)])

    def forward(self, x, emb, ctx, kw):
        x = self.resnets[0](x, emb)
        x = self.attentions[0](x, ctx, kw)
        return self.resnets[1](x, emb)


class UpBlock(nn.Module):
    def __init__(self, cfg, prev_out, cout, cin_skip_last, depth, heads, add_up):
        super().__init__()
        n = cfg.layers_per_block + 1
        res = []
        for i in range(n):
            skip = cin_skip_last if i == n - 1 else cout
            rin = prev_out if i == 0 else cout
            res.append(ResnetBlock2D(rin + skip, cout, cfg.time_embed_dim, cfg.norm_num_groups, cfg.norm_eps))
        self.resnets = nn.ModuleList(res)
        if depth:
            self.attentions = nn.ModuleList([Transformer2DModel(cout, heads, depth, cfg.cross_attention_dim,
                                                                cfg.norm_num_groups) for _ in range(n)])
        self.has_attn = bool(depth)
        if add_up:
            self.upsamplers = nn.ModuleList([Upsample2D(cout)])
        self.add_up = add_up

    def forward(self, x, skips, emb, ctx, kw):
        for i, r in enumerate(self.resnets):
            x = torch.cat([x, skips.pop()], dim=1)
            x = r(x, emb)
            if self.has_attn:
                x = self.attentions[i](x, ctx, kw)
        if self.add_up:
            x = self.upsamplers[0](x)
        return x


class OracleUNet(nn.Module):
    """SDXL-family UNet2DConditionModel; state-dict keys follow diffusers."""

    def __init__(self, cfg: UNetConfig = SDXL_BASE):
        super().__init__()
        self.config = cfg
        ch = cfg.block_out_channels
        ted = cfg.time_embed_dim
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], ted)
        self.add_embedding = TimestepEmbedding(cfg.projection_class_embeddings_input_dim, ted)
        # diffusers registers down_blocks, up_blocks (as empty lists) BEFORE mid_block: attn_processors /
        # IP-adapter checkpoint indices follow that order (down -> up -> mid; ip_adapter.py:165-169)
        self.down_blocks = nn.ModuleList()
        self.up_blocks = nn.ModuleList()
        downs = []
        out = ch[0]
        for i, c in enumerate(ch):
            cin, out = out, c
            depth = cfg.transformer_layers_per_block[i] if cfg.down_block_types[i].startswith("CrossAttn") else 0
            downs.append(DownBlock(cfg, cin, out, depth, cfg.attention_head_dim[i], add_down=i < len(ch) - 1))
        self.down_blocks.extend(downs)
        self.mid_block = MidBlock(cfg, ch[-1], cfg.transformer_layers_per_block[-1], cfg.attention_head_dim[-1])
        ups = []
        rch = list(reversed(ch))
        rdepth = list(reversed(cfg.transformer_layers_per_block))
        rheads = list(reversed(cfg.attention_head_dim))
        out = rch[0]
        for i, c in enumerate(rch):
            prev, out = out, c
            skip_last = rch[min(i + 1, len(ch) - 1)]
            depth = rdepth[i] if cfg.up_block_types[i].startswith("CrossAttn") else 0
            ups.append(UpBlock(cfg, prev, out, skip_last, depth, rheads[i], add_up=i < len(ch) - 1))
        self.up_blocks.extend(ups)
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=cfg.norm_eps)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    # ---- diffusers attention-processor plugin API (ip_adapter.py:120-154) ----
    def _attn_modules(self):
        for name, m in self.named_modules():
            if isinstance(m, Attention):
                yield name, m

    @property
    def attn_processors(self) -> Dict[str, object]:
        return {f"{n}.processor": m.processor for n, m in self._attn_modules()}

    def set_attn_processor(self, processor):
        mods = list(self._attn_modules())
        if isinstance(processor, dict):
            assert len(processor) == len(mods)
            for n, m in mods:
                m.processor = processor[f"{n}.processor"]
        else:
            for _, m in mods:
                m.processor = processor

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def embed(self, timestep, added_cond_kwargs, batch, dtype):
        cfg = self.config
        t = torch.as_tensor(timestep).reshape(-1).to(torch.float32).expand(batch) if torch.as_tensor(timestep).ndim == 0 \
            else torch.as_tensor(timestep).to(torch.float32)
        t_emb = get_timestep_embedding(t, cfg.block_out_channels[0], flip_sin_to_cos=True, downscale_freq_shift=0).to(dtype)
        emb = self.time_embedding(t_emb)
        text_embeds = added_cond_kwargs["text_embeds"]
        time_ids = added_cond_kwargs["time_ids"]
        te = get_timestep_embedding(time_ids.flatten(), cfg.addition_time_embed_dim, flip_sin_to_cos=True,
                                    downscale_freq_shift=0).reshape(text_embeds.shape[0], -1)
        add = torch.cat([text_embeds, te], dim=-1).to(dtype)
        return emb + self.add_embedding(add)

    def forward(self, sample, timestep, encoder_hidden_states, cross_attention_kwargs=None,
                added_cond_kwargs=None, return_dict=False, **unused):
        emb = self.embed(timestep, added_cond_kwargs, sample.shape[0], sample.dtype)
        x = self.conv_in(sample)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, encoder_hidden_states, cross_attention_kwargs)
            skips += outs
        x = self.mid_block(x, emb, encoder_hidden_states, cross_attention_kwargs)
        for blk in self.up_blocks:
            x = blk(x, skips, emb, encoder_hidden_states, cross_attention_kwargs)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        return (x,)
