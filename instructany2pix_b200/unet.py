"""``B200UNet``: drop-in for the diffusers ``UNet2DConditionModel`` the reference drives (SDXL-base family).

Boundary kept intact: ``unet(sample, timestep, encoder_hidden_states=, cross_attention_kwargs=,
added_cond_kwargs={"text_embeds","time_ids"}, return_dict=False) -> (eps,)`` as called at
ddim/pnp_pipeline.py:253-260 and diffusion/ip_adapter/custom_pipelines.py:338-345; ``.config`` attributes read at
ddim/sdxl_pipeline.py:160,770 / pnp_pipeline.py:45 / ip_adapter.py:124-132; ``.add_embedding.linear_1.in_features``
(pnp_pipeline.py:47); the attention-processor plugin API ``attn_processors`` / ``set_attn_processor`` used by
ip_adapter.py:120-154,165-169,211-214; diffusers state-dict key names.

Inside, nothing is diffusers: activations are NHWC bf16 and every op is one of the hand-written sm_100a kernels in
``libia2p_sm100a.so`` (tcgen05 implicit-GEMM convs and linears with fused bias/temb/residual/GEGLU epilogues, fused
GroupNorm+SiLU over un-materialised skip concats, flash self-attention, single-kernel decoupled cross-attention).
Step-invariant work is hoisted: text/IP K,V projections of all 70 cross-attention layers are two GEMMs per request,
time-embedding projections are cached per timestep.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn as nn

from . import ops
from .attention_processor import B200AttnProcessor, B200IPAttnProcessor, is_ip_processor, is_plain_processor
from .packing import interleave_geglu, pack_conv3x3, pack_conv3x3_up2x, pack_conv_out


@dataclass
class B200UNetConfig:
    """Attribute-style config mirroring the diffusers fields callers read."""
    in_channels: int = 4
    out_channels: int = 4
    sample_size: int = 128
    block_out_channels: Tuple[int, ...] = (320, 640, 1280)
    layers_per_block: int = 2
    transformer_layers_per_block: Tuple[int, ...] = (1, 2, 10)
    attention_head_dim: Tuple[int, ...] = (5, 10, 20)          # head COUNTS, head_dim = 64
    cross_attention_dim: int = 2048
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    time_cond_proj_dim: object = None
    addition_embed_type: str = "text_time"
    down_block_types: Tuple[str, ...] = ("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D")
    up_block_types: Tuple[str, ...] = ("CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D")
    use_linear_projection: bool = True

    @property
    def time_embed_dim(self):
        return self.block_out_channels[0] * 4

    def get(self, k, default=None):
        return getattr(self, k, default)


_CONFIG_FIELDS = set(B200UNetConfig.__dataclass_fields__)

TINY_CONFIG = dict(sample_size=32, block_out_channels=(64, 128, 256), transformer_layers_per_block=(1, 1, 2),
                   attention_head_dim=(1, 2, 4), cross_attention_dim=256, addition_time_embed_dim=32,
                   projection_class_embeddings_input_dim=6 * 32 + 128)


# SDXL-refiner UNet (stabilityai/stable-diffusion-xl-refiner-1.0, loaded at pipeline.py:128-131 as ``piperf``): four
# levels, attention on the two middle ones, 4 transformer layers per block, text context 1280, five micro-conditioning ids.
REFINER_CONFIG = dict(block_out_channels=(384, 768, 1536, 1536), transformer_layers_per_block=(4, 4, 4, 4),
                      attention_head_dim=(6, 12, 24, 24), cross_attention_dim=1280, addition_time_embed_dim=256,
                      projection_class_embeddings_input_dim=2560,
                      down_block_types=("DownBlock2D", "CrossAttnDownBlock2D", "CrossAttnDownBlock2D", "DownBlock2D"),
                      up_block_types=("UpBlock2D", "CrossAttnUpBlock2D", "CrossAttnUpBlock2D", "UpBlock2D"))


# ------------------------------------------------------------------ parameter holders (names == diffusers names)
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - never called: arithmetic lives in the CUDA library
        raise RuntimeError("parameter holder: arithmetic runs in libia2p_sm100a.so via B200UNet.forward")


class _TimestepEmbedding(_Holder):
    def __init__(self, cin, dim):
        super().__init__()
        self.linear_1 = nn.Linear(cin, dim)
        self.linear_2 = nn.Linear(dim, dim)


class _Resnet(_Holder):
    def __init__(self, cin, cout, temb_dim, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_dim, cout)
        self.norm2 = nn.GroupNorm(groups, cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.conv_shortcut = nn.Conv2d(cin, cout, 1)
        self.cin, self.cout = cin, cout


class _Attention(_Holder):
    def __init__(self, dim, heads, ctx_dim=None):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim or dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim or dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
        self.is_cross = ctx_dim is not None
        self.processor = B200AttnProcessor()


class _GEGLU(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Linear(dim, dim * 8)


class _FeedForward(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([_GEGLU(dim), nn.Dropout(0.0), nn.Linear(dim * 4, dim)])


class _TBlock(_Holder):
    def __init__(self, dim, heads, ctx_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = _Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = _Attention(dim, heads, ctx_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = _FeedForward(dim)


class _Transformer2D(_Holder):
    def __init__(self, dim, heads, depth, ctx_dim, groups):
        super().__init__()
        self.norm = nn.GroupNorm(groups, dim, eps=1e-6)
        self.proj_in = nn.Linear(dim, dim)
        self.transformer_blocks = nn.ModuleList([_TBlock(dim, heads, ctx_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(dim, dim)
        self.dim, self.heads = dim, heads


class _Sampler(_Holder):
    def __init__(self, c, stride=1):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=stride, padding=1)


class _Block(_Holder):
    pass


def _is_matrix(name, p):
    return p.ndim >= 2


def _fold_ln(w, bias, norm):
    """LayerNorm folded into the consuming Linear (ia2p_gemm_ln_bf16): y = LN(x) W^T + b
    = rstd * (x W'^T - mean * c1) + c2 with W' = bf16(W * gamma), c1 = W' 1, c2 = W beta + b."""
    w32 = w.detach().float()
    wp = (w32 * norm.weight.detach().float()[None]).to(torch.bfloat16).contiguous()
    c1 = wp.float().sum(1).contiguous()
    c2 = w32 @ norm.bias.detach().float()
    if bias is not None:
        c2 = c2 + bias.detach().float()
    return wp, c1, c2.contiguous(), float(norm.eps)


class B200UNet(nn.Module):
    def __init__(self, config=None, device="cuda", stream_dtype=torch.float32, **config_overrides):
        """``stream_dtype``: storage type of the residual stream (block inputs/outputs, skip tensors, conv1 output).
        fp32 (default) keeps only tensor-core OPERANDS in bf16 -- needed to hold the 1e-2 per-step eps tolerance over
        70 transformer blocks (numerics); bf16 trades ~40%% more rounding error for less HBM traffic."""
        super().__init__()
        self.stream_dtype = stream_dtype
        # LayerNorm folded into the consuming GEMM's epilogue (no separate normalisation pass); False restores the
        # standalone ia2p_layernorm kernels (A/B measurements, or nets whose token means dwarf their spread)
        self.fuse_ln = True
        self.fold_upsample = True     # Upsample2D as four parity 2x2 convs over the low-res map (4/9 of the MACs)
        if config is None:
            config = B200UNetConfig(**config_overrides)
        elif not isinstance(config, B200UNetConfig):
            src = config if isinstance(config, dict) else {k: getattr(config, k) for k in _CONFIG_FIELDS if hasattr(config, k)}
            kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in src.items() if k in _CONFIG_FIELDS}
            config = B200UNetConfig(**kw)
        cfg = self.config = config
        if isinstance(cfg.transformer_layers_per_block, int):
            cfg.transformer_layers_per_block = (cfg.transformer_layers_per_block,) * len(cfg.block_out_channels)
        if isinstance(cfg.attention_head_dim, int):
            cfg.attention_head_dim = (cfg.attention_head_dim,) * len(cfg.block_out_channels)
        ch, ted, G = cfg.block_out_channels, cfg.time_embed_dim, cfg.norm_num_groups
        assert all(c % 64 == 0 for c in ch), "channel counts must be multiples of 64 (kernel K chunk)"
        with torch.device(device):
            self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
            self.time_embedding = _TimestepEmbedding(ch[0], ted)
            self.add_embedding = _TimestepEmbedding(cfg.projection_class_embeddings_input_dim, ted)
            # registration order down_blocks, up_blocks, mid_block == diffusers (IP checkpoint indices, ip_adapter.py:165-169)
            self.down_blocks = nn.ModuleList()
            self.up_blocks = nn.ModuleList()
            n = cfg.layers_per_block
            out = ch[0]
            for i, c in enumerate(ch):
                cin, out = out, c
                depth = cfg.transformer_layers_per_block[i] if cfg.down_block_types[i].startswith("CrossAttn") else 0
                blk = _Block()
                blk.resnets = nn.ModuleList([_Resnet(cin if j == 0 else out, out, ted, G) for j in range(n)])
                if depth:
                    blk.attentions = nn.ModuleList([_Transformer2D(out, cfg.attention_head_dim[i], depth,
                                                                   cfg.cross_attention_dim, G) for _ in range(n)])
                if i < len(ch) - 1:
                    blk.downsamplers = nn.ModuleList([_Sampler(out, 2)])
                self.down_blocks.append(blk)
            self.mid_block = _Block()
            self.mid_block.attentions = nn.ModuleList([_Transformer2D(ch[-1], cfg.attention_head_dim[-1],
                                                                      cfg.transformer_layers_per_block[-1],
                                                                      cfg.cross_attention_dim, G)])
            self.mid_block.resnets = nn.ModuleList([_Resnet(ch[-1], ch[-1], ted, G) for _ in range(2)])
            rch = list(reversed(ch))
            rdepth = list(reversed(cfg.transformer_layers_per_block))
            rheads = list(reversed(cfg.attention_head_dim))
            out = rch[0]
            for i, c in enumerate(rch):
                prev, out = out, c
                skip_last = rch[min(i + 1, len(ch) - 1)]
                depth = rdepth[i] if cfg.up_block_types[i].startswith("CrossAttn") else 0
                blk = _Block()
                blk.resnets = nn.ModuleList([
                    _Resnet((prev if j == 0 else out) + (skip_last if j == n else out), out, ted, G) for j in range(n + 1)])
                blk.skip_widths = [(skip_last if j == n else out) for j in range(n + 1)]
                if depth:
                    blk.attentions = nn.ModuleList([_Transformer2D(out, rheads[i], depth, cfg.cross_attention_dim, G)
                                                    for _ in range(n + 1)])
                if i < len(ch) - 1:
                    blk.upsamplers = nn.ModuleList([_Sampler(out)])
                self.up_blocks.append(blk)
            self.conv_norm_out = nn.GroupNorm(G, ch[0])
            self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)
        # storage dtypes: matrices bf16 (tensor-core operands), vectors (norm scales, biases) fp32
        for name, p in self.named_parameters():
            p.requires_grad_(False)
            if _is_matrix(name, p):
                p.data = p.data.to(torch.bfloat16)
        self._packed = None
        self._kv_cache = {}
        self._temb_cache = {}
        self._proc_version = 0
        self._wplans = {}

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_module(cls, unet, device=None):
        """Build from a diffusers ``UNet2DConditionModel`` (or the oracle): copies config, weights and processors."""
        device = device or getattr(unet, "device", "cuda")
        # Build where the source lives and move ONCE: for a CPU source (a checkpoint just loaded, the oracle) construction,
        # parameter initialisation and the state-dict copy stay on the host and the device only sees memcpys -- no stream of
        # initialisation kernels in front of the first real launch (the driver's launch record of smoke() starts there).
        src_dev = next(iter(unet.parameters())).device if hasattr(unet, "parameters") else torch.device(device)
        new = cls(unet.config, device=src_dev)
        # processors first: nn.Module processors are registered sub-modules, so their weights appear in state_dict()
        # under "...attn2.processor.to_k_ip.weight" on both sides (same as diffusers)
        new.set_attn_processor({k: new._adopt_processor(v, k) for k, v in unet.attn_processors.items()})
        new.load_state_dict(unet.state_dict())
        if torch.device(device) != src_dev:
            new.to(device)
            new.invalidate()
        return new

    def _adopt_processor(self, proc, name):
        if isinstance(proc, (B200AttnProcessor, B200IPAttnProcessor)):
            return proc
        if is_ip_processor(proc):
            return B200IPAttnProcessor.from_reference(proc, device=self.device)
        if is_plain_processor(proc):
            return B200AttnProcessor()
        raise NotImplementedError(f"attention processor {type(proc).__name__} at {name} is outside the B200 hot path "
                                  "(supported: AttnProcessor[2_0], IPAttnProcessor[2_0])")

    def load_state_dict(self, state_dict, strict=True, **kw):
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        self.invalidate()
        return r

    def invalidate(self):
        self._packed = None
        self._kv_cache.clear()
        self._temb_cache.clear()
        self._wplans.clear()

    @property
    def dtype(self):
        return torch.bfloat16

    @property
    def device(self):
        return self.conv_in.weight.device

    # ------------------------------------------------------------------ attention-processor plugin API
    def _attn_modules(self):
        for name, m in self.named_modules():
            if isinstance(m, _Attention):
                yield name, m

    @property
    def attn_processors(self) -> Dict[str, object]:
        return {f"{n}.processor": m.processor for n, m in self._attn_modules()}

    def set_attn_processor(self, processor):
        mods = list(self._attn_modules())
        if isinstance(processor, dict):
            if len(processor) != len(mods):
                raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} does not "
                                 f"match the number of attention layers: {len(mods)}.")
            for n, m in mods:
                m.processor = self._adopt_processor(processor[f"{n}.processor"], n)
        else:
            for n, m in mods:
                m.processor = self._adopt_processor(processor, n)
        self._proc_version += 1
        self._packed = None
        self._kv_cache.clear()

    # ------------------------------------------------------------------ weight packing (once per load)
    def _iter_tblocks(self):
        for name, m in self.named_modules():
            if isinstance(m, _TBlock):
                yield name, m

    def _ip_weight_signature(self):
        sig = 0
        for _, m in self._iter_tblocks():
            p = m.attn2.processor
            if is_ip_processor(p):
                sig += p.to_k_ip.weight._version + p.to_v_ip.weight._version + 7 * p.to_k_ip.weight.data_ptr() % 1000003
        return sig

    def prepare(self):
        if self._packed is not None and self._packed.get("ip_sig") != self._ip_weight_signature():
            self.invalidate()          # load_ip_adapter() rewrote to_k_ip / to_v_ip in place (ip_adapter.py:165-169)
        if self._packed is not None:
            return self._packed
        P = {}
        f32 = lambda t: t.detach().float().contiguous()
        bf = lambda t: t.detach().to(torch.bfloat16).contiguous()
        for name, m in self.named_modules():
            if isinstance(m, _Resnet):
                sc = getattr(m, "conv_shortcut", None)
                P[name] = dict(
                    g1=f32(m.norm1.weight), b1=f32(m.norm1.bias), w1=pack_conv3x3(m.conv1.weight), cb1=f32(m.conv1.bias),
                    wt=bf(m.time_emb_proj.weight), bt=f32(m.time_emb_proj.bias),
                    g2=f32(m.norm2.weight), b2=f32(m.norm2.bias),
                    w2=pack_conv3x3(m.conv2.weight, None if sc is None else sc.weight),
                    cb2=f32(m.conv2.bias if sc is None else m.conv2.bias.float() + sc.bias.float()), has_sc=sc is not None)
            elif isinstance(m, _Transformer2D):
                P[name] = dict(g=f32(m.norm.weight), b=f32(m.norm.bias), wi=bf(m.proj_in.weight), bi=f32(m.proj_in.bias),
                               wo=bf(m.proj_out.weight), bo=f32(m.proj_out.bias))
            elif isinstance(m, _TBlock):
                wg, bg = interleave_geglu(m.ff.net[0].proj.weight.detach(), m.ff.net[0].proj.bias.detach().float())
                wqkv_raw = torch.cat([m.attn1.to_q.weight, m.attn1.to_k.weight, m.attn1.to_v.weight], 0)
                ln_qkv = _fold_ln(wqkv_raw, None, m.norm1)
                ln_q2 = _fold_ln(m.attn2.to_q.weight, None, m.norm2)
                ln_g = _fold_ln(wg, bg, m.norm3)
                P[name] = dict(
                    ln_qkv=ln_qkv, ln_q2=ln_q2, ln_g=ln_g,
                    ln1=(f32(m.norm1.weight), f32(m.norm1.bias)), ln2=(f32(m.norm2.weight), f32(m.norm2.bias)),
                    ln3=(f32(m.norm3.weight), f32(m.norm3.bias)),
                    wqkv=bf(torch.cat([m.attn1.to_q.weight, m.attn1.to_k.weight, m.attn1.to_v.weight], 0)),
                    wo1=bf(m.attn1.to_out[0].weight), bo1=f32(m.attn1.to_out[0].bias),
                    wq2=bf(m.attn2.to_q.weight), wo2=bf(m.attn2.to_out[0].weight), bo2=f32(m.attn2.to_out[0].bias),
                    wg=bf(wg), bg=f32(bg), wf=bf(m.ff.net[2].weight), bf=f32(m.ff.net[2].bias))
            elif isinstance(m, _Sampler):
                P[name] = dict(w=pack_conv3x3(m.conv.weight), b=f32(m.conv.bias))
                if ".upsamplers." in name:       # nearest-2x upsample folded into the conv: four parity 2x2 convs
                    P[name]["w4"] = pack_conv3x3_up2x(m.conv.weight)
        P["conv_in"] = (f32(self.conv_in.weight), f32(self.conv_in.bias))
        P["conv_out"] = (f32(self.conv_out.weight.permute(0, 2, 3, 1)), f32(self.conv_out.bias))
        P["conv_out_tc"] = pack_conv_out(self.conv_out.weight, self.conv_out.bias)       # tensor-core form (output channels padded to 32)
        P["norm_out"] = (f32(self.conv_norm_out.weight), f32(self.conv_norm_out.bias))
        P["time"] = (bf(self.time_embedding.linear_1.weight), f32(self.time_embedding.linear_1.bias),
                     bf(self.time_embedding.linear_2.weight), f32(self.time_embedding.linear_2.bias))
        P["add"] = (bf(self.add_embedding.linear_1.weight), f32(self.add_embedding.linear_1.bias),
                    bf(self.add_embedding.linear_2.weight), f32(self.add_embedding.linear_2.bias))
        # all time_emb_proj stacked: one small-M GEMM per step gives every resnet's per-image channel bias
        rn = [n for n, m in self.named_modules() if isinstance(m, _Resnet)]
        P["temb_w"] = torch.cat([P[n]["wt"] for n in rn], 0).contiguous()
        P["temb_b"] = torch.cat([P[n]["bt"] + P[n]["cb1"] for n in rn], 0).contiguous()   # conv1 bias folded in
        off = 0
        for n in rn:
            P[n]["temb_slice"] = (off, off + P[n]["wt"].shape[0])
            P.setdefault("temb_slices", []).append(P[n]["temb_slice"])
            off += P[n]["wt"].shape[0]
        # cross-attention K,V projections of ALL layers as one stacked weight per branch (step-invariant: SURVEY 7.1)
        tb = list(self._iter_tblocks())
        wkv_t, wkv_i, col = [], [], 0
        mode = None
        for n, m in tb:
            proc = m.attn2.processor
            C = m.attn2.to_k.weight.shape[0]
            wkv_t.append(torch.cat([m.attn2.to_k.weight, m.attn2.to_v.weight], 0))
            if is_ip_processor(proc):
                wkv_i.append(torch.cat([proc.to_k_ip.weight, proc.to_v_ip.weight], 0).to(self.device))
            P[n]["kv_cols"] = (col, col + 2 * C)
            col += 2 * C
        if tb:
            assert len(wkv_i) in (0, len(tb)), "mixed IP / plain processors on attn2 layers are not supported"
            P["wkv_text"] = bf(torch.cat(wkv_t, 0))
            P["wkv_ip"] = bf(torch.cat(wkv_i, 0)) if wkv_i else None
        P["ip_sig"] = self._ip_weight_signature()
        self._packed = P
        return P

    # ------------------------------------------------------------------ hoisted, step-invariant pieces
    def _ip_state(self):
        for _, m in self._iter_tblocks():
            p = m.attn2.processor
            if is_ip_processor(p):
                return int(p.num_tokens), float(p.scale)
            return 0, 0.0
        return 0, 0.0

    def context_kv(self, ctx):
        """K,V of every cross-attention layer for this context: [B*n_text, sum 2C], [B*n_ip, sum 2C].

        One entry is cached for the drop-in ``forward()`` path (the reference passes the SAME ``prompt_embeds`` tensor object on
        every step of a loop, custom_pipelines.py:338-345).  The entry holds a reference to that tensor and hits only on object
        identity + unchanged version counter: an address can never be recycled to a different request while it is cached."""
        P = self.prepare()
        n_ip, _ = self._ip_state()
        hit = self._kv_cache.get("entry")
        if (hit is not None and hit["ctx"] is ctx and hit["version"] == ctx._version and hit["proc"] == self._proc_version
                and hit["n_ip"] == n_ip):
            return hit["val"]
        self._kv_cache.clear()
        B, S, D = ctx.shape
        n_text = S - n_ip                                    # attention_processor.py:350-354 (also the 77-token quirk)
        c = ctx.to(torch.bfloat16)
        text = c[:, :n_text].reshape(B * n_text, D).contiguous()
        kv_t = ops.gemm(text, P["wkv_text"])
        kv_i = None
        if n_ip:
            ip = c[:, n_text:].reshape(B * n_ip, D).contiguous()
            kv_i = ops.gemm(ip, P["wkv_ip"])
        val = (kv_t, kv_i, n_text, n_ip)
        self._kv_cache["entry"] = dict(ctx=ctx, version=ctx._version, proc=self._proc_version, n_ip=n_ip, val=val)
        return val

    def time_rowbias(self, timestep, added_cond_kwargs, batch):
        """emb = time_embedding(sinus(t)) + add_embedding([text_embeds, sinus(time_ids)]) and, stacked for all resnets,
        time_emb_proj(SiLU(emb)) + conv1.bias -> fp32 [batch, sum Cout]."""
        P = self.prepare()
        cfg = self.config
        dev = self.device
        te, tid = added_cond_kwargs["text_embeds"], added_cond_kwargs["time_ids"]
        key = None
        if not (torch.is_tensor(timestep) and timestep.is_cuda):
            # per-timestep cache for the drop-in loop, valid only for the SAME conditioning tensor objects (held here, so their
            # addresses cannot be recycled) at unchanged version counters
            c = self._temb_cache
            if not (c.get("te") is te and c.get("tid") is tid and c.get("ver") == (te._version, tid._version)):
                c.clear()
                c.update(te=te, tid=tid, ver=(te._version, tid._version), rows={})
            key = (float(timestep), batch)
            hit = c["rows"].get(key)
            if hit is not None:
                return hit
            t = torch.full((batch,), float(timestep), device=dev, dtype=torch.float32)
        else:
            t = timestep.reshape(-1).to(torch.float32).expand(batch).contiguous()
        w1, b1, w2, b2 = P["time"]
        e = ops.timestep_embedding(t, cfg.block_out_channels[0], True, 0.0)
        e = ops.gemm_smallm(e, w1, bias=b1, act=ops.ACT_SILU)
        emb = ops.gemm_smallm(e, w2, bias=b2)
        a1, ab1, a2, ab2 = P["add"]
        ids = ops.timestep_embedding(tid.to(dev, torch.float32).reshape(-1), cfg.addition_time_embed_dim, True, 0.0)
        add_in = torch.cat([te.to(dev, torch.float32), ids.reshape(te.shape[0], -1)], dim=-1).contiguous()
        a = ops.gemm_smallm(add_in, a1, bias=ab1, act=ops.ACT_SILU)
        emb = ops.gemm_smallm(a, a2, bias=ab2, residual=emb)
        rb = ops.gemm_smallm(emb, P["temb_w"], bias=P["temb_b"], act_in=ops.ACT_SILU)
        if key is not None:
            rows = self._temb_cache["rows"]
            if len(rows) > 256:
                rows.clear()
            rows[key] = rb
        return rb

    def time_rowbias_table(self, timesteps, added_cond_kwargs, batch):
        """``time_rowbias`` for ALL timesteps of a trajectory in five small GEMMs -> fp32 [steps, batch * sum Cout] in the BLOCKED
        layout ``forward_core`` takes as a 1-D ``rowbias`` (block r = the [batch, Cout_r] bias of ResnetBlock r)."""
        P = self.prepare()
        cfg = self.config
        dev = self.device
        ts = torch.as_tensor(timesteps, dtype=torch.float32).reshape(-1)
        steps = ts.numel()
        te, tid = added_cond_kwargs["text_embeds"], added_cond_kwargs["time_ids"]
        t = ts.to(dev).repeat_interleave(batch)
        w1, b1, w2, b2 = P["time"]
        e = ops.timestep_embedding(t, cfg.block_out_channels[0], True, 0.0)
        e = ops.gemm_smallm(e, w1, bias=b1, act=ops.ACT_SILU)
        a1, ab1, a2, ab2 = P["add"]
        ids = ops.timestep_embedding(tid.to(dev, torch.float32).reshape(-1), cfg.addition_time_embed_dim, True, 0.0)
        add_in = torch.cat([te.to(dev, torch.float32), ids.reshape(te.shape[0], -1)], dim=-1).contiguous()
        a = ops.gemm_smallm(add_in, a1, bias=ab1, act=ops.ACT_SILU)
        aug = ops.gemm_smallm(a, a2, bias=ab2)                              # [batch, ted]
        emb = ops.gemm_smallm(e, w2, bias=b2, residual=aug.repeat(steps, 1))
        rb = ops.gemm_smallm(emb, P["temb_w"], bias=P["temb_b"], act_in=ops.ACT_SILU).reshape(steps, batch, -1)
        # blocked layout: per step the 17 ResnetBlocks' [batch, Cout] biases one after the other, each block contiguous, so that
        # the step itself slices views (no copy kernels inside the CUDA graph); one re-layout per request
        return torch.cat([rb[:, :, lo:hi].reshape(steps, -1) for lo, hi in P["temb_slices"]], dim=1).contiguous()

    # ------------------------------------------------------------------ forward
    def forward(self, sample, timestep, encoder_hidden_states, class_labels=None, timestep_cond=None,
                attention_mask=None, cross_attention_kwargs=None, added_cond_kwargs=None, return_dict=False, **unused):
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is always None on the reference hot path (SURVEY 8b-4)")
        ops.require_cuda(sample, "B200UNet.forward(sample)")
        B = sample.shape[0]
        ctx = encoder_hidden_states
        if ctx.shape[0] != B:
            raise ValueError(f"encoder_hidden_states batch {ctx.shape[0]} != sample batch {B}")
        rowbias = self.time_rowbias(timestep, added_cond_kwargs, B)
        kv = self.context_kv(ctx)
        eps = self.forward_core(sample, rowbias, kv, B, out_dtype=sample.dtype)
        if return_dict:
            from types import SimpleNamespace
            return SimpleNamespace(sample=eps)
        return (eps,)

    def forward_core(self, sample, rowbias, kv, batch, out_dtype=torch.float32):
        """All per-step device work (capturable in a CUDA graph: no host sync, no data-dependent control flow).
        ``sample`` may hold ``batch`` or ``batch // 2`` images (CFG duplication is folded into conv_in)."""
        P = self.prepare()
        cfg = self.config
        G = cfg.norm_num_groups
        kv_t, kv_i, n_text, n_ip = kv
        _, ip_scale = self._ip_state()
        # weight prefetch plan: keyed by everything that fixes the launch order (packed weights, batch, resolution, processors)
        pkey = (id(P), batch, tuple(sample.shape[-2:]), n_text, n_ip, self._proc_version)
        plan = self._wplans.setdefault(pkey, {"seq": [], "ready": False})
        ops.plan_begin(plan)
        try:
            return self._forward_core(sample, rowbias, kv, batch, out_dtype, P)
        finally:
            ops.plan_end()

    def _forward_core(self, sample, rowbias, kv, batch, out_dtype, P):
        cfg = self.config
        G = cfg.norm_num_groups
        kv_t, kv_i, n_text, n_ip = kv
        _, ip_scale = self._ip_state()

        SD = self.stream_dtype

        def resnet(name, x, skip=None):
            p = P[name]
            raw = None
            if p["has_sc"] and (SD != torch.bfloat16):
                # the 1x1 shortcut is a tensor-core operand: GroupNorm also emits the raw bf16 concat it just read
                h, raw = ops.groupnorm(x, skip, p["g1"], p["b1"], G, cfg.norm_eps, True, want_raw=True)
            else:
                h = ops.groupnorm(x, skip, p["g1"], p["b1"], G, cfg.norm_eps, True)
            lo, hi = p["temb_slice"]
            # 1-D rowbias = blocked table row (time_rowbias_table): a view; 2-D [batch, sum Cout] (generic forward): a small copy
            rbias = rowbias[batch * lo:batch * hi].view(batch, hi - lo) if rowbias.ndim == 1 else rowbias[:, lo:hi].contiguous()
            h = ops.conv3x3(h, p["w1"], p["w1"].shape[0], rowbias=rbias, out_dtype=SD, want_colstats=True)
            h = ops.groupnorm(h, None, p["g2"], p["b2"], G, cfg.norm_eps, True)
            if p["has_sc"]:
                if raw is not None:
                    return ops.conv3x3(h, p["w2"], p["w2"].shape[0], sc_a=raw, bias=p["cb2"], out_dtype=SD, want_colstats=True)
                return ops.conv3x3(h, p["w2"], p["w2"].shape[0], sc_a=x, sc_b=skip, bias=p["cb2"], out_dtype=SD, want_colstats=True)
            assert skip is None
            return ops.conv3x3(h, p["w2"], p["w2"].shape[0], bias=p["cb2"], residual=x, out_dtype=SD, want_colstats=True)

        def transformer(name, mod, x):
            p = P[name]
            Bx, H, W, C = x.shape
            M, N, heads = Bx * H * W, H * W, mod.heads
            t = ops.groupnorm(x, None, p["g"], p["b"], G, 1e-6, False).reshape(M, C)
            BF = torch.bfloat16
            nblk = len(mod.transformer_blocks)
            if self.fuse_ln and SD == torch.float32:
                # LayerNorm folded into the consumer GEMMs: producers of the fp32 stream also emit raw bf16 rows + row statistics
                t, tb, st = ops.gemm(t, p["wi"], bias=p["bi"], out_dtype=SD, want_ln=True)
                for k in range(nblk):
                    q = P[f"{name}.transformer_blocks.{k}"]
                    w_, c1, c2, eps = q["ln_qkv"]
                    qkv = ops.gemm(tb, w_, bias=c2, ln=(st, c1, eps))
                    a = ops.flash_self_attn(qkv, Bx, N, heads)
                    t, tb, st = ops.gemm(a, q["wo1"], bias=q["bo1"], residual=t, out_dtype=SD, want_ln=True)
                    w_, c1, c2, eps = q["ln_q2"]
                    qq = ops.gemm(tb, w_, bias=c2, ln=(st, c1, eps))
                    c0, c1k = q["kv_cols"]
                    a = ops.cross_attn(qq, kv_t[:, c0:c1k], n_text, None if kv_i is None else kv_i[:, c0:c1k], n_ip, ip_scale,
                                       Bx, N, heads)
                    t, tb, st = ops.gemm(a, q["wo2"], bias=q["bo2"], residual=t, out_dtype=SD, want_ln=True)
                    w_, c1, c2, eps = q["ln_g"]
                    h = ops.gemm(tb, w_, bias=c2, ln=(st, c1, eps), geglu=True)
                    if k == nblk - 1:   # consumed only by proj_out as a tensor-core operand: write bf16
                        t = ops.gemm(h, q["wf"], bias=q["bf"], residual=t, out_dtype=BF)
                    else:
                        t, tb, st = ops.gemm(h, q["wf"], bias=q["bf"], residual=t, out_dtype=SD, want_ln=True)
            else:
                t = ops.gemm(t, p["wi"], bias=p["bi"], out_dtype=SD)
                for k in range(nblk):
                    q = P[f"{name}.transformer_blocks.{k}"]
                    h = ops.layernorm(t, q["ln1"][0], q["ln1"][1], 1e-5, out_dtype=BF)
                    qkv = ops.gemm(h, q["wqkv"])
                    a = ops.flash_self_attn(qkv, Bx, N, heads)
                    t = ops.gemm(a, q["wo1"], bias=q["bo1"], residual=t, out_dtype=SD)
                    h = ops.layernorm(t, q["ln2"][0], q["ln2"][1], 1e-5, out_dtype=BF)
                    qq = ops.gemm(h, q["wq2"])
                    c0, c1 = q["kv_cols"]
                    a = ops.cross_attn(qq, kv_t[:, c0:c1], n_text, None if kv_i is None else kv_i[:, c0:c1], n_ip, ip_scale,
                                       Bx, N, heads)
                    t = ops.gemm(a, q["wo2"], bias=q["bo2"], residual=t, out_dtype=SD)
                    h = ops.layernorm(t, q["ln3"][0], q["ln3"][1], 1e-5, out_dtype=BF)
                    h = ops.gemm(h, q["wg"], bias=q["bg"], geglu=True)
                    # the last block's output is consumed only by proj_out as a tensor-core operand: write it as bf16
                    last = k == nblk - 1
                    t = ops.gemm(h, q["wf"], bias=q["bf"], residual=t, out_dtype=BF if last else SD)
            out = ops.gemm(t, p["wo"], bias=p["bo"], residual=x.reshape(M, C), out_dtype=SD, want_colstats=True)
            cs = getattr(out, "_ia2p_cs", None)
            out = out.reshape(Bx, H, W, C)
            if cs is not None:
                out._ia2p_cs = cs              # GroupNorm statistics of the next block come from this epilogue
            return out

        w_in, b_in = P["conv_in"]
        x = ops.conv_in(sample, w_in, b_in, out_batch=batch, out_dtype=SD)
        skips = [x]
        for i, blk in enumerate(self.down_blocks):
            for j in range(len(blk.resnets)):
                x = resnet(f"down_blocks.{i}.resnets.{j}", x)
                if hasattr(blk, "attentions"):
                    x = transformer(f"down_blocks.{i}.attentions.{j}", blk.attentions[j], x)
                skips.append(x)
            if hasattr(blk, "downsamplers"):
                p = P[f"down_blocks.{i}.downsamplers.0"]
                x = ops.conv3x3(ops.to_bf16(x), p["w"], p["w"].shape[0], stride=2, bias=p["b"], out_dtype=SD, want_colstats=True)
                skips.append(x)
        x = resnet("mid_block.resnets.0", x)
        x = transformer("mid_block.attentions.0", self.mid_block.attentions[0], x)
        x = resnet("mid_block.resnets.1", x)
        for i, blk in enumerate(self.up_blocks):
            for j in range(len(blk.resnets)):
                x = resnet(f"up_blocks.{i}.resnets.{j}", x, skips.pop())
                if hasattr(blk, "attentions"):
                    x = transformer(f"up_blocks.{i}.attentions.{j}", blk.attentions[j], x)
            if hasattr(blk, "upsamplers"):
                p = P[f"up_blocks.{i}.upsamplers.0"]
                if SD == torch.float32 and self.fold_upsample:
                    x = ops.conv_up2x(ops.to_bf16(x), p["w4"], p["w"].shape[0], bias=p["b"], want_colstats=True)
                else:
                    x = ops.conv3x3(ops.upsample2x(x), p["w"], p["w"].shape[0], bias=p["b"], out_dtype=SD)
        g, b = P["norm_out"]
        x = ops.groupnorm(x, None, g, b, G, cfg.norm_eps, True)
        if out_dtype == torch.float32 and x.shape[-1] % 64 == 0:
            return ops.conv_out_tc(x, *P["conv_out_tc"], cfg.out_channels)
        w_out, b_out = P["conv_out"]
        return ops.conv_out(x, w_out, b_out, out_dtype=out_dtype)
