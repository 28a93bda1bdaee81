"""Attention processors for ``B200UNet`` -- the reference's plugin API.

The reference installs processors with ``unet.set_attn_processor({...})`` (ip_adapter.py:120-148), mutates
``processor.scale`` through ``isinstance(p, IPAttnProcessor)`` checks (ip_adapter.py:211-214, custom_pipelines.py:17-20)
and loads ``to_k_ip/to_v_ip`` through ``ModuleList(unet.attn_processors.values())`` (ip_adapter.py:165-169).
These classes keep all of that working; the arithmetic itself is NOT here: ``B200UNet`` reads ``scale``,
``num_tokens``, ``to_k_ip``, ``to_v_ip`` at forward time and runs the fused decoupled cross-attention kernel
(replacing attention_processor.py:310-412).  ``B200IPAttnProcessor`` subclasses the reference's
``IPAttnProcessor2_0`` whenever the reference package is importable so those ``isinstance`` checks hold.
"""
from __future__ import annotations

import torch
import torch.nn as nn

try:  # the drop-in case: running inside the reference's environment
    from instructany2pix.diffusion.ip_adapter.attention_processor import IPAttnProcessor2_0 as _RefIP  # type: ignore
    from instructany2pix.diffusion.ip_adapter.attention_processor import AttnProcessor2_0 as _RefPlain  # type: ignore
except Exception:  # standalone (bench / tests / GPU box)
    _RefIP = None
    _RefPlain = None

_PLAIN_NAMES = {"AttnProcessor", "AttnProcessor2_0", "B200AttnProcessor"}


def is_ip_processor(p) -> bool:
    return hasattr(p, "to_k_ip") and hasattr(p, "to_v_ip")


def is_plain_processor(p) -> bool:
    return type(p).__name__ in _PLAIN_NAMES and not is_ip_processor(p)


class B200AttnProcessor(nn.Module if _RefPlain is None else _RefPlain):
    """Self-attention (attn1) / text-only cross-attention marker (AttnProcessor2_0, attention_processor.py:191-279)."""

    def __init__(self, hidden_size=None, cross_attention_dim=None):
        nn.Module.__init__(self)

    def __call__(self, *a, **k):
        raise RuntimeError("B200 processors are parameter holders; B200UNet.forward runs the fused CUDA kernels")


class B200IPAttnProcessor(nn.Module if _RefIP is None else _RefIP):
    """Decoupled cross-attention parameters: scale, num_tokens, to_k_ip, to_v_ip (IPAttnProcessor2_0, :282-308)."""

    def __init__(self, hidden_size, cross_attention_dim=None, scale=1.0, num_tokens=4, device=None, dtype=torch.bfloat16):
        nn.Module.__init__(self)
        self.hidden_size = hidden_size
        self.cross_attention_dim = cross_attention_dim
        self.scale = scale
        self.num_tokens = num_tokens
        self.to_k_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False, device=device, dtype=dtype)
        self.to_v_ip = nn.Linear(cross_attention_dim or hidden_size, hidden_size, bias=False, device=device, dtype=dtype)
        for p in self.parameters():
            p.requires_grad_(False)

    @classmethod
    def from_reference(cls, proc, device=None):
        w = proc.to_k_ip.weight
        new = cls(w.shape[0], w.shape[1], scale=proc.scale, num_tokens=proc.num_tokens, device=device or w.device)
        new.to_k_ip.weight.data.copy_(proc.to_k_ip.weight.data)
        new.to_v_ip.weight.data.copy_(proc.to_v_ip.weight.data)
        return new

    def __call__(self, *a, **k):
        raise RuntimeError("B200 processors are parameter holders; B200UNet.forward runs the fused CUDA kernels")
