"""Oracle (test infrastructure): restated embedding-prior denoiser.

Follows, line by line where it matters (quirks included -- SURVEY.md section 0.5):
  * InstructAny2PixPrior.generate_diffusion     prior/model.py:527-658
  * get_input_sequence_and_mask / add_sos_eos   prior/model.py:299-381, :272-287
  * get_eps                                     prior/model.py:208-239
  * prior_config (fused key "noise_leveltgt_type")   prior/__init__.py:2-41
  * [3P] transformers GPT2Model (gpt2-medium)   called at prior/model.py:624-626
State-dict key names equal the reference's (``prior/model.bin`` layout)
minus the CLIP text tower, whose output for the constant prompt "" is an INPUT here
(``clip_hidden``): the tower runs once per request and stays on PyTorch.
Pinned against the reference code executed under shims: tests/golden/prior_*.npz.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .schedulers import DDPMSchedulerOracle, get_timestep_embedding

SEQUENCE_INPUT_KEY = ["src_type", "imagebind", "crossattn_clip", "score", "noisy_inputs", "noise_leveltgt_type"]
SEQUENCE_INPUT_EMBED_DIM = [0, 1024, 1024, 512, 0, 0, 0]


def gelu_new(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


class _Conv1D(nn.Module):
    """HF Conv1D: weight stored [in, out]."""

    def __init__(self, nin, nout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(nin, nout).normal_(std=0.02))
        self.bias = nn.Parameter(torch.zeros(nout))

    def forward(self, x):
        return x @ self.weight + self.bias


class _Attn(nn.Module):
    def __init__(self, d, h):
        super().__init__()
        self.c_attn = _Conv1D(d, 3 * d)
        self.c_proj = _Conv1D(d, d)
        self.h = h

    def forward(self, x):
        b, t, d = x.shape
        q, k, v = self.c_attn(x).split(d, dim=2)
        sh = lambda z: z.view(b, t, self.h, d // self.h).transpose(1, 2)
        q, k, v = sh(q), sh(k), sh(v)
        w = (q @ k.transpose(-1, -2)) / math.sqrt(d // self.h)
        mask = torch.tril(torch.ones(t, t, dtype=torch.bool, device=x.device))
        w = w.masked_fill(~mask, torch.finfo(w.dtype).min).softmax(dim=-1)
        o = (w @ v).transpose(1, 2).reshape(b, t, d)
        return self.c_proj(o)


class _MLP(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.c_fc = _Conv1D(d, 4 * d)
        self.c_proj = _Conv1D(4 * d, d)

    def forward(self, x):
        return self.c_proj(gelu_new(self.c_fc(x)))


class _Block(nn.Module):
    def __init__(self, d, h, eps):
        super().__init__()
        self.ln_1 = nn.LayerNorm(d, eps=eps)
        self.attn = _Attn(d, h)
        self.ln_2 = nn.LayerNorm(d, eps=eps)
        self.mlp = _MLP(d)

    def forward(self, x):
        x = x + self.attn(self.ln_1(x))
        return x + self.mlp(self.ln_2(x))


class OracleGPT2(nn.Module):
    def __init__(self, n_embd=1024, n_layer=24, n_head=16, n_positions=1024, vocab_size=50257, eps=1e-5):
        super().__init__()
        self.wte = nn.Embedding(vocab_size, n_embd)    # unused (inputs_embeds path) but present in model.bin
        self.wpe = nn.Embedding(n_positions, n_embd)
        self.h = nn.ModuleList([_Block(n_embd, n_head, eps) for _ in range(n_layer)])
        self.ln_f = nn.LayerNorm(n_embd, eps=eps)

    def forward(self, inputs_embeds, attention_mask=None):
        t = inputs_embeds.shape[1]
        x = inputs_embeds + self.wpe.weight[:t][None]
        for blk in self.h:
            x = blk(x)
        return {"last_hidden_state": self.ln_f(x)}


class OraclePrior(nn.Module):
    def __init__(self, n_layer=24, embed_dim=1024, n_head=16):
        super().__init__()
        self.embed_dim = embed_dim
        self.mae_token_num = 1
        self.start_of_sequence_tokens = nn.Embedding(32, embed_dim)
        self.end_of_sequence_tokens = nn.Embedding(32, embed_dim)
        self.input_sequence_embed_linear = nn.ModuleList(
            [nn.Identity() if d == 0 else nn.Linear(d, embed_dim) for d in SEQUENCE_INPUT_EMBED_DIM])
        self.modality_embedding = nn.Embedding(10, embed_dim)
        self.model = OracleGPT2(n_embd=embed_dim, n_layer=n_layer, n_head=n_head)
        self.noise_scheduler = DDPMSchedulerOracle()

    def _wrap(self, _id, seq):
        b = seq.shape[0]
        kid = torch.tensor([_id])
        sos = self.start_of_sequence_tokens(kid).expand(b, 1, -1)
        eos = self.end_of_sequence_tokens(kid).expand(b, 1, -1)
        return torch.cat([sos, seq, eos], dim=1)

    def input_sequence(self, cond):
        parts = []
        for _id, key in enumerate(SEQUENCE_INPUT_KEY):
            if key not in cond:
                continue
            v = cond[key]
            if key in ("src_type", "tgt_type"):
                parts.append(self.modality_embedding(v))
            elif isinstance(v, list):
                parts.append(self._wrap(_id, self.input_sequence_embed_linear[_id](v[0])))
            else:
                parts.append(self._wrap(_id, self.input_sequence_embed_linear[_id](v)))
        x = torch.cat(parts, dim=1)
        return x[:, : 1024 - self.mae_token_num]

    def get_eps(self, t, sample, x0):
        a = self.noise_scheduler.alphas_cumprod[int(t)]
        return (sample - a ** 0.5 * x0) / (1 - a) ** 0.5

    @torch.no_grad()
    def generate_diffusion(self, src_type, tgt_type, src, clip_hidden, num_inference_steps=25, generator=None,
                           guidance_scale=5, score=6.8, negative_score=2.0, no_diffusion=False, trace=None):
        """One sample (the reference cannot batch: prior/model.py:569,580).  ``clip_hidden`` (1,2,E) is the
        CLIP-H text hidden state of "".  Draws from the GLOBAL torch RNG in the reference's order:
        randn(1,1,E) then one randn per scheduler step with t>0."""
        if no_diffusion:
            num_inference_steps = 1
        E = self.embed_dim
        src = src.reshape(1, 1, E).float()
        score_emb = get_timestep_embedding(torch.tensor([score]).float(), 512, flip_sin_to_cos=True,
                                           downscale_freq_shift=0).view(1, 1, -1)
        cond = dict(
            src_type=torch.tensor(src_type).view(1, 1).repeat(2, 1),
            imagebind=torch.cat([src, src * 0.0], dim=0),
            crossattn_clip=[clip_hidden.expand(2, -1, -1).float(), torch.ones(2, clip_hidden.shape[1])],
            score=torch.cat([score_emb, score_emb * 0.0 + negative_score], dim=0),
        )
        self.noise_scheduler.set_timesteps(num_inference_steps)
        key = "noisy_input" if no_diffusion else "noisy_inputs"
        x = torch.randn(1, 1, E).to(torch.int64)                 # prior/model.py:597 (int64 truncation)
        cond[key] = x.repeat(2, 1, 1)
        for t in self.noise_scheduler.timesteps:
            seq = self.input_sequence(cond)
            out = self.model(inputs_embeds=seq)["last_hidden_state"][:, -1:, :]
            eps = self.get_eps(t, cond[key], out)
            e_c, e_u = eps.chunk(2)                               # first half = cond (:643)
            eps = e_u + guidance_scale * (e_c - e_u)
            lat = self.noise_scheduler.step(eps, t, cond[key][:1], generator=generator)[0]
            if trace is not None:
                trace.append(dict(t=int(t), seq=seq.clone(), x0=out.clone(), eps=eps.clone(), x=lat.clone()))
            cond[key] = lat.repeat(2, 1, 1)
        return cond[key][:1], cond
