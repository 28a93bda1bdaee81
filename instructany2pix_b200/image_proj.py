"""``B200ImageProj``: drop-in for ``ImageProjModel`` (diffusion/ip_adapter/ip_adapter.py:28-67) and for the projector half of
``IPAdapter.get_image_embeds`` (ip_adapter.py:171-209) -- SURVEY 8a rows a12/a13: the LLM / prior embedding
``(B, 1024)`` becomes the 4 (``mode='both'``: 8) image tokens ``(B, 4, 2048)`` the decoupled cross-attention consumes, plus
the unconditional tokens projected from zeros.

Same constructor arguments, state-dict keys (``proj.{weight,bias}``, ``norm.{weight,bias}``, ``raw_embed`` = the
``"image_proj"`` part of the IP-adapter checkpoint, ip_adapter.py:165-166) and call signature as the reference module.
Arithmetic on the sm_100a kernels: the local/global blend is linear, so it is applied to the 1024-wide INPUT rows
(``ia2p_axpby``) instead of the 8192-wide projections; one small-M weight-streaming GEMM per crop (``ia2p_gemm_smallm``,
bias pre-summed with the crop's ``raw_embed`` row) and one fp32 LayerNorm over the token rows (``ia2p_layernorm``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class B200ImageProj(nn.Module):
    def __init__(self, cross_attention_dim=2048, clip_embeddings_dim=1024, clip_extra_context_tokens=4, num_crops=2,
                 device="cuda"):
        super().__init__()
        assert num_crops == 2, "the reference's forward hard-codes a global and a local crop (ip_adapter.py:50-53)"
        self.cross_attention_dim = cross_attention_dim
        self.clip_extra_context_tokens = clip_extra_context_tokens
        self.num_crops = num_crops
        self.proj = nn.Linear(clip_embeddings_dim, clip_extra_context_tokens * cross_attention_dim, device=device)
        self.norm = nn.LayerNorm(cross_attention_dim, device=device)
        self.raw_embed = nn.Parameter(torch.zeros(2, cross_attention_dim, device=device))
        self.requires_grad_(False)
        self._packed = None

    @classmethod
    def from_module(cls, m, device="cuda"):
        """Adopt a reference ``ImageProjModel`` (same keys)."""
        out_f, in_f = m.proj.weight.shape
        new = cls(m.cross_attention_dim, in_f, m.clip_extra_context_tokens, getattr(m, "num_crops", 2), device=device)
        new.load_state_dict(m.state_dict())
        return new

    def load_state_dict(self, *a, **k):
        self._packed = None
        return super().load_state_dict(*a, **k)

    @property
    def device(self):
        return self.proj.weight.device

    @property
    def dtype(self):
        return self.proj.weight.dtype

    def prepare(self):
        """bf16 weight + per-crop bias (``proj.bias`` + ``raw_embed[crop]`` tiled over the tokens); once per load."""
        if self._packed is None:
            T = self.clip_extra_context_tokens
            w = self.proj.weight.detach().to(torch.bfloat16).contiguous()
            b = self.proj.bias.detach().float()
            bias = [(b + self.raw_embed[c].detach().float().repeat(T)).contiguous() for c in range(2)]
            self._packed = (w, bias, self.norm.weight.detach().float().contiguous(), self.norm.bias.detach().float().contiguous())
        return self._packed

    @torch.no_grad()
    def forward(self, image_embeds, mode, scales=(1.0, 1.0)):
        """image_embeds (B, 2, D_clip): [:, 0] global crop, [:, 1] local crop -> (B, T * (1 | 2), cross_attention_dim) in the
        dtype of ``image_embeds`` (the reference feeds fp16, ip_adapter.py:182; computed in fp32 either way)."""
        ops.require_cuda(image_embeds, "B200ImageProj")
        assert mode in ("global", "local", "both"), f"Invalid Mode {mode}"
        w, bias, gamma, beta = self.prepare()
        e = image_embeds.float().contiguous()
        B, T, D = e.shape[0], self.clip_extra_context_tokens, self.cross_attention_dim
        eg = e[:, 0].contiguous()
        parts = []
        if mode in ("global", "both"):
            parts.append(ops.gemm_smallm(eg, w, bias=bias[0]))
        if mode in ("local", "both"):
            # proj(g) (1 - s) + proj(l) s  ==  proj((1 - s) g + s l): blend the 1024-wide inputs, not the 8192-wide outputs
            s = float(scales[1])
            el = ops.axpby(e[:, 1].contiguous(), eg, 1.0 - s, s)
            parts.append(ops.gemm_smallm(el, w, bias=bias[1]))
        t = parts[0] if len(parts) == 1 else torch.stack(parts, dim=1)           # (B, [crop,] T * D)
        out = ops.layernorm(t.reshape(-1, D), gamma, beta, self.norm.eps, out_dtype=torch.float32).reshape(B, -1, D)
        return out if image_embeds.dtype == torch.float32 else out.to(image_embeds.dtype)

    @torch.no_grad()
    def get_image_embeds(self, clip_image_embeds=None, clip_image_embeds_local=None, mode="global", scale_g=1.0, scale_l=1.0):
        """The tensor half of ``IPAdapter.get_image_embeds`` (ip_adapter.py:171-209; the PIL / CLIP-vision branch stays on
        the reference): a missing crop is zeros, the unconditional tokens are the projector applied to zeros with the
        DEFAULT scales -> (image_prompt_embeds, uncond_image_prompt_embeds)."""
        if clip_image_embeds is None:
            assert clip_image_embeds_local is not None
            clip_image_embeds = torch.zeros_like(clip_image_embeds_local)
        elif clip_image_embeds_local is None:
            clip_image_embeds_local = torch.zeros_like(clip_image_embeds)
        e = torch.stack([clip_image_embeds, clip_image_embeds_local], dim=1).to(self.device)
        return self.forward(e, mode=mode, scales=[scale_g, scale_l]), self.forward(torch.zeros_like(e), mode=mode)
