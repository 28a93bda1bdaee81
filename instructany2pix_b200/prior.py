"""``B200Prior``: drop-in for ``InstructAny2PixPrior.generate_diffusion`` (prior/model.py:527-658).

Same signature, return contract ``(Tensor[raw_bs,1,1024], cond_dict)`` and state-dict key names as the reference
(``prior/model.bin`` layout); the reference's quirks are reproduced on purpose:
the fused config key means neither the timestep embedding nor ``tgt_type`` is consumed (prior/__init__.py:19-20), the
initial noise is truncated to integers (prior/model.py:597), CFG takes the FIRST half as conditional (:643-644), and
with ``no_diffusion=True`` the noisy sample is not part of the sequence (11 tokens instead of 14, :593-596).
Unlike the reference (which only works for one sample, :569,580) the "" text conditioning is broadcast, so
``raw_bs`` samples run as one batch of ``2*raw_bs`` rows.

The GPT-2-medium trunk ([3P] transformers ``GPT2Model``) runs on the small-M weight-streaming kernels
(``ia2p_gemm_smallm`` with fp32 activations split hi/lo onto bf16 tensor-core MMAs, ``ia2p_layernorm`` fp32,
``ia2p_causal_attn_small_f32``) and one fused x0->eps + CFG + DDPM kernel per step; the step-invariant prefix of the
token sequence is built once per call.  The CLIP-H text tower for the constant prompt "" stays a PyTorch module
(``cond_stage_models[0]``; once per request) and its output is cached.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .packing import conv1d_to_linear
from .scheduler import B200DDPMScheduler

SEQUENCE_INPUT_KEY = ["src_type", "imagebind", "crossattn_clip", "score", "noisy_inputs", "noise_leveltgt_type"]
SEQUENCE_INPUT_EMBED_DIM = [0, 1024, 1024, 512, 0, 0, 0]


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder: arithmetic runs in libia2p_sm100a.so via B200Prior.generate_diffusion")


class _Conv1D(_Holder):
    def __init__(self, nin, nout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(nin, nout), requires_grad=False)   # HF layout [in, out]
        self.bias = nn.Parameter(torch.empty(nout), requires_grad=False)


class _GPTAttn(_Holder):
    def __init__(self, d):
        super().__init__()
        self.c_attn = _Conv1D(d, 3 * d)
        self.c_proj = _Conv1D(d, d)


class _GPTMLP(_Holder):
    def __init__(self, d):
        super().__init__()
        self.c_fc = _Conv1D(d, 4 * d)
        self.c_proj = _Conv1D(4 * d, d)


class _GPTBlock(_Holder):
    def __init__(self, d):
        super().__init__()
        self.ln_1 = nn.LayerNorm(d)
        self.attn = _GPTAttn(d)
        self.ln_2 = nn.LayerNorm(d)
        self.mlp = _GPTMLP(d)


class _GPT2(_Holder):
    def __init__(self, d, n_layer, n_positions, vocab):
        super().__init__()
        self.wte = nn.Embedding(vocab, d)        # present in model.bin, unused (inputs_embeds path)
        self.wpe = nn.Embedding(n_positions, d)
        self.h = nn.ModuleList([_GPTBlock(d) for _ in range(n_layer)])
        self.ln_f = nn.LayerNorm(d)


class B200Prior(nn.Module):
    def __init__(self, n_layer=24, embed_dim=1024, n_head=16, n_positions=1024, vocab_size=50257, device="cuda",
                 cond_stage_model=None, use_cuda_graph=True):
        super().__init__()
        assert embed_dim == n_head * 64, "head_dim must be 64"
        self.embed_dim, self.n_head, self.n_layer = embed_dim, n_head, n_layer
        self.mae_token_num = 1
        with torch.device(device):
            self.start_of_sequence_tokens = nn.Embedding(32, embed_dim)
            self.end_of_sequence_tokens = nn.Embedding(32, embed_dim)
            self.input_sequence_embed_linear = nn.ModuleList(
                [nn.Identity() if d == 0 else nn.Linear(d, embed_dim) for d in SEQUENCE_INPUT_EMBED_DIM])
            self.modality_embedding = nn.Embedding(10, embed_dim)
            self.model = _GPT2(embed_dim, n_layer, n_positions, vocab_size)
        self.cond_stage_models = nn.ModuleList([cond_stage_model] if cond_stage_model is not None else [])
        for p in self.parameters():
            p.requires_grad_(False)
        self.noise_scheduler = B200DDPMScheduler()
        self.use_cuda_graph = use_cuda_graph
        self.fused_trunk = True          # False: the per-op kernels (7 launches per layer), kept for A/B and for other widths
        self.fused_trunk_max_rows = 32
        self._trunk_ws = {}
        self._packed = None
        self._clip_hidden = None
        self._graphs = {}
        self._seq_static = {}

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_module(cls, prior, device="cuda"):
        """Build from the reference ``InstructAny2PixPrior`` (or the oracle): same config, weights, CLIP tower."""
        gpt = prior.model
        n_layer = len(gpt.h)
        embed = prior.embed_dim
        cond = prior.cond_stage_models[0] if len(getattr(prior, "cond_stage_models", [])) else None
        new = cls(n_layer=n_layer, embed_dim=embed, device=device, cond_stage_model=cond)
        sd = {k: v for k, v in prior.state_dict().items() if not k.startswith("cond_stage_models.")}
        new.load_state_dict(sd, strict=False)
        return new

    def load_state_dict(self, state_dict, strict=True, **kw):
        # accept (and ignore) buffers written by older transformers: attn.bias / masked_bias / position_ids
        sd = {k: v for k, v in state_dict.items()
              if not (k.endswith(".attn.bias") or k.endswith(".attn.masked_bias") or k.endswith("position_ids"))}
        own_cond = any(k.startswith("cond_stage_models.") for k in self.state_dict())
        if not own_cond:
            sd = {k: v for k, v in sd.items() if not k.startswith("cond_stage_models.")}
        r = super().load_state_dict(sd, strict=strict, **kw)
        self._packed = None
        self._graphs.clear()
        self._seq_static.clear()
        self._trunk_ws.clear()
        return r

    @property
    def device(self):
        return self.modality_embedding.weight.device

    def set_clip_hidden(self, hidden):
        """Hidden state (1,T,E) of the CLIP-H text tower for the constant prompt "" (skips running the tower)."""
        self._clip_hidden = hidden.detach().to(self.device, torch.float32)

    def _clip(self):
        if self._clip_hidden is None:
            if not len(self.cond_stage_models):
                raise ops.IA2PError("B200Prior: no cond_stage_model and no clip hidden state set (set_clip_hidden)")
            hidden, _mask = self.cond_stage_models[0]([""])
            self._clip_hidden = hidden.detach().to(self.device, torch.float32)
        return self._clip_hidden

    # ------------------------------------------------------------------ weight packing
    def prepare(self):
        if self._packed is not None:
            return self._packed
        f32 = lambda t: t.detach().float().contiguous()
        lin = lambda t: t.detach().to(torch.bfloat16).contiguous()              # nn.Linear [out,in]
        c1d = lambda t: conv1d_to_linear(t.detach()).to(torch.bfloat16).contiguous()   # Conv1D [in,out] -> [out,in]
        P = dict(layers=[])
        for blk in self.model.h:
            P["layers"].append(dict(
                ln1=(f32(blk.ln_1.weight), f32(blk.ln_1.bias)), ln2=(f32(blk.ln_2.weight), f32(blk.ln_2.bias)),
                wqkv=c1d(blk.attn.c_attn.weight), bqkv=f32(blk.attn.c_attn.bias),
                wo=c1d(blk.attn.c_proj.weight), bo=f32(blk.attn.c_proj.bias),
                wfc=c1d(blk.mlp.c_fc.weight), bfc=f32(blk.mlp.c_fc.bias),
                wpr=c1d(blk.mlp.c_proj.weight), bpr=f32(blk.mlp.c_proj.bias)))
        P["lnf"] = (f32(self.model.ln_f.weight), f32(self.model.ln_f.bias))
        P["wpe"] = f32(self.model.wpe.weight)
        P["emb"] = {i: (lin(m.weight), f32(m.bias)) for i, m in enumerate(self.input_sequence_embed_linear)
                    if isinstance(m, nn.Linear)}
        self._packed = P
        return P

    # ------------------------------------------------------------------ step-invariant sequence prefix
    def _prefix(self, src_type, src, score, negative_score, bs, with_noisy_slot):
        """[2*bs, T, E] fp32: rows [0,bs) conditional, [bs,2bs) unconditional (prior/model.py:562-584, 299-381)."""
        P = self.prepare()
        dev, E = self.device, self.embed_dim
        sos, eos = self.start_of_sequence_tokens.weight.float(), self.end_of_sequence_tokens.weight.float()

        def wrap(_id, seq):              # add_sos_eos_tokens (:272-287)
            n = seq.shape[0]
            return torch.cat([sos[_id].expand(n, 1, E), seq, eos[_id].expand(n, 1, E)], dim=1)

        def embed(_id, x):               # input_sequence_embed_linear[_id] on the small-M kernel
            w, b = P["emb"][_id]
            shp = x.shape
            return ops.gemm_smallm(x.reshape(-1, shp[-1]).contiguous(), w, bias=b).reshape(*shp[:-1], E)

        parts = [self.modality_embedding.weight.float()[torch.full((2 * bs, 1), int(src_type), device=dev)]]
        ib = torch.cat([src.reshape(bs, 1, E), torch.zeros(bs, 1, E, device=dev)], 0)          # uncond: imagebind * 0
        parts.append(wrap(1, embed(1, ib)))
        clip = self._clip()
        parts.append(wrap(2, embed(2, clip)).expand(2 * bs, -1, -1))
        sc = ops.timestep_embedding(torch.tensor([float(score)], device=dev), 512, True, 0.0).view(1, 1, 512)
        sc = torch.cat([sc.expand(bs, 1, 512), torch.full((bs, 1, 512), float(negative_score), device=dev)], 0)  # :583
        parts.append(wrap(3, embed(3, sc.contiguous())))
        if with_noisy_slot:
            parts.append(wrap(4, torch.zeros(2 * bs, 1, E, device=dev)))     # x_t is written into the middle token per step
        seq = torch.cat(parts, dim=1).contiguous()
        return seq[:, : 1024 - self.mae_token_num]

    # ------------------------------------------------------------------ GPT-2 trunk on the small-M kernels
    def _trunk_last(self, seq):
        """GPT2Model(inputs_embeds=seq)["last_hidden_state"][:, -1] -> [2*bs, E] fp32."""
        P = self.prepare()
        B2, T, E = seq.shape
        # the persistent kernel keeps the weights resident and walks 32-row chunks one after the other: it wins for one request
        # (2 x 14 rows); batched requests spread their row chunks over more CTAs on the per-op kernels
        if self.fused_trunk and E == 1024 and self.n_head == 16 and B2 * T <= self.fused_trunk_max_rows and len(P["layers"]) <= 30:
            return self._trunk_last_fused(P, seq, B2, T)
        with ops.pdl():      # ~175 tiny dependent kernels: each one's prologue / weight prefetch overlaps its predecessor's tail
            return self._trunk_last_impl(P, seq, B2, T, E)

    def _trunk_last_fused(self, P, seq, B2, T):
        """ONE persistent cooperative kernel per step (csrc/smallm.cu prior_trunk_kernel) instead of 7 launches per layer."""
        layers = P.get("_flat")
        if layers is None:
            layers = P["_flat"] = [(L["wqkv"], L["wo"], L["wfc"], L["wpr"], L["bqkv"], L["bo"], L["bfc"], L["bpr"],
                                    L["ln1"][0], L["ln1"][1], L["ln2"][0], L["ln2"][1]) for L in P["layers"]]
            P["_cache"] = {}
        ws = self._trunk_ws.get(B2 * T)
        if ws is None:
            ws = self._trunk_ws[B2 * T] = ops.prior_trunk_workspace(B2 * T, seq.device)
        return ops.prior_trunk(seq, P["wpe"], layers, P["lnf"][0], P["lnf"][1], self.n_head, ws, cache=P["_cache"])

    def _trunk_last_impl(self, P, seq, B2, T, E):
        h = ops.axpby(P["wpe"][:T].unsqueeze(0).expand(B2, T, E).contiguous(), seq, 1.0, 1.0).reshape(B2 * T, E)
        for L in P["layers"]:
            a = ops.layernorm(h, L["ln1"][0], L["ln1"][1], 1e-5)
            qkv = ops.gemm_smallm(a, L["wqkv"], bias=L["bqkv"])
            att = ops.causal_attn_small(qkv, B2, T, self.n_head).reshape(B2 * T, E)
            h = ops.gemm_smallm(att, L["wo"], bias=L["bo"], residual=h)
            m = ops.layernorm(h, L["ln2"][0], L["ln2"][1], 1e-5)
            f = ops.gemm_smallm(m, L["wfc"], bias=L["bfc"], act=ops.ACT_GELU_NEW)
            h = ops.gemm_smallm(f, L["wpr"], bias=L["bpr"], residual=h)
        last = h.reshape(B2, T, E)[:, -1].contiguous()
        return ops.layernorm(last, P["lnf"][0], P["lnf"][1], 1e-5)

    def _static_seq(self, seq):
        """The sequence buffer the captured trunk reads: ONE per shape, owned by the module and reused by every call, so the
        CUDA graph of a shape is captured once per weight load -- not once per request (a capture costs a warm-up pass, the
        capture itself and torch's gc / synchronise on graph entry: ~200 ms, 8x the whole 25-step sampling it would serve)."""
        key = tuple(seq.shape)
        st = self._seq_static.get(key)
        if st is None:
            if len(self._seq_static) >= 4:
                self._seq_static.clear()
                self._graphs.clear()
            st = self._seq_static[key] = torch.empty_like(seq)
        st.copy_(seq)
        return st

    def _step_graph(self, seq):
        """x0 = trunk(seq): eager, or one CUDA-graph replay over the static sequence buffer (updated in place by the caller)."""
        if not self.use_cuda_graph:
            return self._trunk_last(seq)
        key = (tuple(seq.shape), seq.data_ptr(), id(self._packed))
        ent = self._graphs.get(key)
        if ent is None:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._trunk_last(seq)
            torch.cuda.current_stream().wait_stream(s)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._trunk_last(seq)
            ent = self._graphs[key] = (g, out)
        ent[0].replay()
        return ent[1]

    # ------------------------------------------------------------------ the reference entry point
    @torch.no_grad()
    def generate_diffusion(self, src_type, tgt_type, src, no_grad=False, num_inference_steps=25, eta: float = 0.0,
                           generator=None, image_bind_overwrite=None, guidance_scale=5, score=6.8, negative_score=2.0,
                           do_classifier_free_guidance=True, device="cuda", dtype=torch.float16, no_diffusion=False,
                           force_guidence_t0=False, trace=None):
        # Arguments of the reference signature (prior/model.py:527-541) and what they mean here:
        #   do_classifier_free_guidance=False -> ``noise_pred = output`` of the conditional rows (:645-646).  The trunk is bound by
        #       streaming its weights, rows are free: the unconditional rows are still evaluated and combined with weight 0
        #       (guidance 1: eps_u + 1 * (eps_c - eps_u)), one code path, equal to the reference to an ulp.
        #   eta -> reaches ``prepare_extra_step_kwargs`` only; [3P] ``DDPMScheduler.step`` takes no eta, so it is ignored there too.
        #   image_bind_overwrite -> only read when the source is text (:558-566); for every other source it is overwritten by ``src``.
        #   no_grad -> unused in the reference body as well.
        if not do_classifier_free_guidance:
            guidance_scale = 1.0
        if src_type == 2:
            raise NotImplementedError("text-source prior (MODALITY.TEXT, prior/model.py:554-556) needs the CLIP text tower on every call and is "
                                      "never used by pipeline.py (it passes MODALITY.VIDEO at :313): not on the hot path")
        if no_diffusion:
            num_inference_steps = 1
        dev, E = self.device, self.embed_dim
        bs = len(src)
        out_dev = torch.device(device)
        sched = self.noise_scheduler
        sched.set_timesteps(num_inference_steps)
        src_d = src.reshape(bs, 1, -1).to(dev, torch.float32)
        seq = self._prefix(src_type, src_d, score, negative_score, bs, with_noisy_slot=not no_diffusion)
        if self.use_cuda_graph:
            seq = self._static_seq(seq)
        slot = seq.shape[1] - 2          # [.. SOS4 x_t EOS4]
        # initial noise: drawn where the reference draws it (``device``), truncated like `.to(src_type)` (int64) does
        x = torch.randn(bs, 1, E, device=out_dev).to(torch.int64).to(torch.float32).to(dev).contiguous()
        for t in sched.timesteps.tolist():
            if not no_diffusion:
                seq[:, slot] = torch.cat([x, x], 0)[:, 0]
            x0 = self._step_graph(seq)                                   # [2*bs, E]: cond rows then uncond rows
            c = sched.coefficients(t)
            noise = None
            if c["sigma"] > 0.0:
                # DDPMScheduler.step draws randn_tensor(shape, generator, device=model_output.device) (global RNG if None)
                noise = torch.randn(bs, 1, E, generator=generator, device=out_dev).to(dev)
            x = ops.prior_cfg_ddpm_step(x0, x, noise, c["sqrt_a"], c["sqrt_1ma"], guidance_scale, c["c_x0"], c["c_x"],
                                        c["sigma"]).reshape(bs, 1, E)
            if trace is not None:
                trace.append(dict(t=t, x0=x0.clone(), x=x.clone()))
        out = x.to(out_dev)
        key = "noisy_input" if no_diffusion else "noisy_inputs"
        cond_dict = {key: torch.cat([out, out], 0), "src_type": torch.full((2 * bs, 1), int(src_type)),
                     "tgt_type": torch.full((2 * bs, 1), int(tgt_type))}
        return out, cond_dict
