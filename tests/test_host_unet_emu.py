"""Host-side orchestration of B200UNet / sampler checked against the oracle with the kernel test double (CPU)."""
import torch

from instructany2pix_b200.attention_processor import B200IPAttnProcessor
from instructany2pix_b200.sampler import B200Sampler
from instructany2pix_b200.unet import TINY_CONFIG, B200UNet
from oracle import sampler as osampler
from oracle.attention import ImageProjModel, IPAttnProcessor2_0, get_image_embeds
from oracle.synth import synth_request, synth_state_dict
from oracle.unet import TINY, OracleUNet

torch.set_grad_enabled(False)


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def build_pair(with_ip=True, device="cpu"):
    o = OracleUNet(TINY).eval()
    o.load_state_dict(synth_state_dict(o, 0))
    if with_ip:
        procs = {}
        for name, p in o.attn_processors.items():
            if name.endswith("attn2.processor"):
                C = dict(o.named_modules())[name[: -len(".processor")]].to_q.weight.shape[0]
                ip = IPAttnProcessor2_0(C, TINY.cross_attention_dim, scale=0.8, num_tokens=4)
                ip.load_state_dict(synth_state_dict({k: v.shape for k, v in ip.state_dict().items()}, 5))
                # per-layer distinct weights
                for k, v in ip.state_dict().items():
                    v.copy_(synth_state_dict({name + k: v.shape}, 5)[name + k])
                procs[name] = ip
            else:
                procs[name] = p
        o.set_attn_processor(procs)
    b = B200UNet.from_module(o, device=device)
    return o, b


def make_inputs(cfg, B=1, L=16):
    reqs = [synth_request(i, cfg, L) for i in range(B)]
    proj = ImageProjModel(cfg.cross_attention_dim, 1024, 4)
    proj.load_state_dict(synth_state_dict(proj, 9))
    ip, ip_u = get_image_embeds(proj, torch.stack([r["llm_embed"] for r in reqs]))
    ctx, added = osampler.cfg_inputs(reqs, ip, ip_u)
    lat = torch.stack([r["latent"] for r in reqs])
    return lat, ctx, added


def test_structure_and_plugin_api():
    b = B200UNet(device="meta", **TINY_CONFIG)
    o = OracleUNet(TINY)
    assert {k: tuple(v.shape) for k, v in b.state_dict().items()} == {k: tuple(v.shape) for k, v in o.state_dict().items()}
    assert list(b.attn_processors) == list(o.attn_processors)
    assert b.add_embedding.linear_1.in_features == TINY.projection_class_embeddings_input_dim
    assert b.config.time_cond_proj_dim is None and b.config.in_channels == 4


def test_forward_matches_oracle_with_ip(emu):
    o, b = build_pair(True)
    lat, ctx, added = make_inputs(TINY)
    x = torch.cat([lat, lat])
    ref = o(x, torch.tensor(981), ctx, added_cond_kwargs=added)[0]
    out = b(x, torch.tensor(981), ctx, added_cond_kwargs=added)[0]
    assert out.shape == ref.shape
    assert rel(out, ref) < 1.5e-2        # bf16 activations end to end vs fp32 oracle
    # set_scale semantics (ip_adapter.py:211-214): mutate .scale on the processors, next forward sees it
    for p in b.attn_processors.values():
        if isinstance(p, B200IPAttnProcessor):
            p.scale = 0.0
    for p in o.attn_processors.values():
        if isinstance(p, IPAttnProcessor2_0):
            p.scale = 0.0
    assert rel(b(x, 981, ctx, added_cond_kwargs=added)[0], o(x, torch.tensor(981), ctx, added_cond_kwargs=added)[0]) < 1.5e-2


def test_forward_plain_processors_and_quirk(emu):
    """disable(): plain processors on attn2 -> text-only attention over ALL tokens (ip_adapter.py:153-154);
    inversion quirk: 77-token context with IP processors -> last 4 text tokens act as image tokens."""
    o, b = build_pair(False)
    lat, ctx, added = make_inputs(TINY)
    x = torch.cat([lat, lat])
    assert rel(b(x, 501, ctx, added_cond_kwargs=added)[0], o(x, torch.tensor(501), ctx, added_cond_kwargs=added)[0]) < 1.5e-2
    o2, b2 = build_pair(True)
    c77 = ctx[:1, :77]
    add1 = {k: v[:1] for k, v in added.items()}
    assert rel(b2(lat, 21, c77, added_cond_kwargs=add1)[0], o2(lat, torch.tensor(21), c77, added_cond_kwargs=add1)[0]) < 1.5e-2


def test_sampler_generate_and_invert(emu):
    o, b = build_pair(True)
    lat, ctx, added = make_inputs(TINY, B=1, L=8)
    s = B200Sampler(b, use_cuda_graph=False)
    tr_o, tr_b = [], []
    ref = osampler.generate(o, lat, ctx, added, num_inference_steps=4, guidance_scale=7.5, trace=tr_o)
    out = s.generate(lat, ctx, added, num_inference_steps=4, guidance_scale=7.5, trace=tr_b,
                     teacher=[t["x"] for t in tr_o])
    for a, r in zip(tr_b, tr_o):
        assert a["t"] == r["t"] and rel(a["eps2"], r["eps2"]) < 1.5e-2
    assert rel(out, ref) < 1.5e-2
    c1 = ctx[1:, :77]
    add1 = {k: v[1:] for k, v in added.items()}
    ref_i = osampler.invert(o, lat, c1, add1, num_inference_steps=3)
    out_i = s.invert(lat, c1, add1, num_inference_steps=3)
    assert rel(out_i, ref_i) < 1.5e-2


def test_kv_and_temb_caches_never_serve_a_recycled_address(emu):
    """the K/V and time-embedding caches were keyed on data_ptr()/_version; a freed context's address is handed to
    the next request's tensor (same shape, version 0) -> stale conditioning.  Now the entry holds the tensor and hits on identity."""
    o, b = build_pair(True)
    lat, ctx, added = make_inputs(TINY)
    seen = set()
    prev_k = None
    stale = 0
    for i in range(10):
        c = torch.randn(2, 81, TINY.cross_attention_dim)            # a fresh request: same shape, new contents
        seen.add(c.data_ptr())
        k = b.context_kv(c)[0].clone()
        assert b.context_kv(c)[0] is b._kv_cache["entry"]["val"][0]  # same object again -> hit
        if prev_k is not None and torch.equal(k, prev_k):
            stale += 1
        prev_k = k
        del c
    assert stale == 0
    # in-place modification of the SAME tensor bumps the version counter -> recomputed
    c = torch.randn(2, 81, TINY.cross_attention_dim)
    k0 = b.context_kv(c)[0].clone()
    c.mul_(2.0)
    assert not torch.equal(b.context_kv(c)[0], k0)
    # time-embedding bias: fresh pooled embeddings at a (possibly recycled) address must not hit the previous prompt's rows
    prev = None
    for i in range(10):
        a = dict(text_embeds=torch.randn_like(added["text_embeds"]), time_ids=added["time_ids"].clone())
        rb = b.time_rowbias(981, a, 2).clone()
        assert b.time_rowbias(981, a, 2) is b._temb_cache["rows"][(981.0, 2)]
        assert prev is None or not torch.equal(rb, prev)
        prev = rb
        del a


def test_sampler_keeps_inversion_and_cfg_graph_entries(emu):
    """`_static` cleared every entry on a new key, so an edit request (invert -> generate) recaptured both graphs."""
    o, b = build_pair(True)
    lat, ctx, added = make_inputs(TINY, B=1, L=8)
    s = B200Sampler(b, use_cuda_graph=False)
    c1, add1 = ctx[1:, :77], {k: v[1:] for k, v in added.items()}
    s.invert(lat, c1, add1, num_inference_steps=2)
    s.generate(lat, ctx, added, num_inference_steps=2)
    ents = {k[0]: id(v) for k, v in s._graphs.items()}
    assert set(ents) == {"inv", "gen"}
    s.invert(lat, c1, add1, num_inference_steps=2)
    s.generate(lat, ctx, added, num_inference_steps=2)
    assert {k[0]: id(v) for k, v in s._graphs.items()} == ents


def test_inpaint_mask_needs_init_latents(emu):
    import pytest
    o, b = build_pair(True)
    lat, ctx, added = make_inputs(TINY, B=1, L=8)
    with pytest.raises(ValueError, match="init_latents"):
        B200Sampler(b, use_cuda_graph=False).generate(lat, ctx, added, num_inference_steps=2, inpaint_mask=torch.ones(1, 1, 8, 8))


def test_start_latent_is_per_sample(emu):
    """a batch is a batch of independent requests: sample b of a batched call == the single-sample call (reference: B = 1 only)"""
    o, b = build_pair(True)
    s = B200Sampler(b, use_cuda_graph=False)
    x, n = torch.randn(3, 4, 8, 8), torch.randn(3, 4, 8, 8) * 2.0
    out = s.start_latent(x, 0.7, noise=n)
    for i in range(3):
        assert torch.allclose(out[i:i + 1], s.start_latent(x[i:i + 1], 0.7, noise=n[i:i + 1]), atol=1e-6)


def nine_channel_case(device="cpu"):
    """TINY topology with a 9-channel conv_in: latents + mask + masked-image latents (the released SDXL-inpainting layout)"""
    import dataclasses
    from oracle.synth import synth_input
    cfg = dataclasses.replace(TINY, in_channels=9)
    o = OracleUNet(cfg).eval()
    o.load_state_dict(synth_state_dict(o, 31))
    b = B200UNet.from_module(o, device=device)
    lat, ctx, added = make_inputs(TINY, B=2, L=16)
    init = synth_input("ip9/init", (2, 4, 16, 16), seed=4) * 0.8
    mask = (synth_input("ip9/mask", (2, 1, 16, 16), seed=5) > 0).float()
    return o, b, lat, ctx, added, init, mask, init * (1 - mask)


def test_nine_channel_inpainting_unet(emu):
    o, b, lat, ctx, added, init, mask, masked = nine_channel_case()
    assert b.conv_in.weight.shape[1] == 9 and b.config.in_channels == 9
    kw = dict(num_inference_steps=4, guidance_scale=7.5, init_latents=init, strength=0.75, inpaint_mask=mask, masked_image_latents=masked)
    ref = osampler.generate(o, lat, ctx, added, **kw)
    out = B200Sampler(b, use_cuda_graph=False).generate(lat, ctx, added, **kw)
    assert rel(out, ref) < 2e-2
    import pytest
    with pytest.raises(ValueError, match="9-channel"):
        B200Sampler(b, use_cuda_graph=False).generate(lat, ctx, added, num_inference_steps=2, init_latents=init, inpaint_mask=mask)

def test_lcm_scheduler_tables_and_sampler(emu):
    """B200LCMScheduler (the reference's 4-step ``ipa_lcm`` mode) against the restated [3P] LCMScheduler: the published
    4-step timesteps of a 50-point training schedule, the three coefficients of every step against the oracle's step on random
    tensors, and a free-running 4-step CFG trajectory through B200Sampler with identical noise draws."""
    from instructany2pix_b200.scheduler import B200LCMScheduler
    from oracle.schedulers import LCMSchedulerOracle
    bs, os_ = B200LCMScheduler(), LCMSchedulerOracle()
    for n in (1, 2, 4, 8):
        bs.set_timesteps(n)
        os_.set_timesteps(n)
        assert bs.timesteps.tolist() == os_.timesteps.tolist()
    bs.set_timesteps(4)
    assert bs.timesteps.tolist() == [999, 759, 499, 259]
    os_.set_timesteps(4)
    g = torch.Generator().manual_seed(3)
    x, e = torch.randn(1, 4, 8, 8, generator=g), torch.randn(1, 4, 8, 8, generator=g)
    for t in bs.timesteps.tolist():
        c_x, c_e, c_n = bs.coefficients3(t)
        ref = os_.step(e, t, x, generator=torch.Generator().manual_seed(11))[0]
        noise = torch.randn(x.shape, generator=torch.Generator().manual_seed(11))
        assert rel(c_x * x + c_e * e + c_n * noise, ref) < 1e-5
    assert bs.coefficients3(259)[2] == 0.0                       # the last step is not re-noised
    o, b = build_pair(True)
    lat, ctx, added = make_inputs(TINY, B=1, L=8)
    os2, bs2 = LCMSchedulerOracle(), B200LCMScheduler()
    os2.generator, bs2.generator = torch.Generator().manual_seed(5), torch.Generator().manual_seed(5)
    ref = osampler.generate(o, lat, ctx, added, num_inference_steps=4, guidance_scale=1.5, scheduler=os2)
    out = B200Sampler(b, scheduler=bs2, use_cuda_graph=False).generate(lat, ctx, added, num_inference_steps=4, guidance_scale=1.5)
    assert rel(out, ref) < 2e-2
