"""Oracle (test infrastructure): deterministic synthetic weights and inputs.

There is no network, hence no checkpoints: every parity run uses random weights of
the named architecture.  To make the SAME weights available to the reference code
(run in the build container), to the oracle and to the B200 modules without
shipping gigabytes, every tensor is generated from a seed derived from its
state-dict *name* -- independent of module construction order and of the
process that generates it (CPU ``torch.Generator``, bit-reproducible for a fixed
torch version).

Scales follow PyTorch's default initialisers: Linear/Conv matrices
~ U(+-1/sqrt(fan_in)); HF ``Conv1D`` ([in,out]) ~ N(0, 0.02) like GPT-2's own init; norm
scales 1 + 0.1 N(0,1); biases / embeddings small.  Every tensor with >= 2 dims is rounded
to a bf16-representable value: the "checkpoint" is a bf16 checkpoint, so the fp32 oracle
and the bf16 tensor-core path consume IDENTICAL weights and a parity gap measures the
arithmetic, not the storage format of the synthetic weights.
"""
from __future__ import annotations

import zlib

import torch

_CONV1D_TAGS = (".c_attn.", ".c_fc.", ".c_proj.")


def _gen(name: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def synth_tensor(name: str, shape, seed: int = 0, dtype=torch.float32) -> torch.Tensor:
    shape = tuple(shape)
    g = _gen(name, seed)
    if name.endswith("wte.weight"):                       # unused by the prior (inputs_embeds path)
        return torch.zeros(shape, dtype=dtype)
    if name.endswith("raw_embed"):
        return (0.02 * torch.randn(shape, generator=g)).to(dtype)
    if len(shape) <= 1:
        if name.endswith(".weight"):                       # norm scale
            return (1.0 + 0.1 * torch.randn(shape, generator=g)).to(dtype)
        return (0.05 * torch.randn(shape, generator=g)).to(dtype)
    q = lambda t: t.to(torch.bfloat16).to(dtype)
    if "embedding" in name or "_tokens" in name or name.endswith("wpe.weight"):
        return q(0.02 * torch.randn(shape, generator=g))
    if any(t in name for t in _CONV1D_TAGS) and "text_model" not in name:
        return q(0.02 * torch.randn(shape, generator=g))
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    bound = (1.0 / fan_in) ** 0.5
    return q((torch.rand(shape, generator=g) * 2 - 1) * bound)


def synth_state_dict(module_or_shapes, seed: int = 0, dtype=torch.float32, prefix_filter=None):
    """state dict for ``module`` (or a ``{name: shape}`` map) with name-seeded tensors."""
    if hasattr(module_or_shapes, "state_dict"):
        shapes = {k: tuple(v.shape) for k, v in module_or_shapes.state_dict().items()}
    else:
        shapes = dict(module_or_shapes)
    out = {}
    for k, shp in shapes.items():
        if prefix_filter is not None and not prefix_filter(k):
            continue
        out[k] = synth_tensor(k, shp, seed, dtype)
    return out


def synth_input(tag: str, shape, seed: int = 0, scale: float = 1.0) -> torch.Tensor:
    """seeded N(0, scale^2) input tensor keyed by a tag (e.g. 'ctx/3')."""
    return scale * torch.randn(tuple(shape), generator=_gen("input/" + tag, seed))


def synth_request(i: int, cfg, L: int, seed: int = 1000):
    """Synthetic conditioning for image ``i``: text ctx (77,D), pooled, negatives,
    LLM embedding e (norm 20), time_ids, initial latent (4,L,L)."""
    s = seed + i
    D = cfg.cross_attention_dim
    pooled_dim = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
    e = synth_input("llm", (1024,), s)
    e = e / e.norm() * 20.0
    H = float(L * 8)
    return dict(
        ctx=synth_input("ctx", (77, D), s), neg_ctx=synth_input("neg_ctx", (77, D), s),
        pooled=synth_input("pooled", (pooled_dim,), s), neg_pooled=synth_input("neg_pooled", (pooled_dim,), s),
        llm_embed=e, time_ids=torch.tensor([H, H, 0.0, 0.0, H, H]),
        latent=synth_input("latent", (cfg.in_channels, L, L), s),
    )
