"""Pin the oracle restatements against outputs of the REFERENCE code (tests/golden, made by oracle/gen_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import gen_golden as G
from oracle import ref_shims
from oracle.attention import AttnProcessor2_0, IPAttnProcessor2_0, ImageProjModel
from oracle.prior import OraclePrior
from oracle.schedulers import DDIMSchedulerOracle, backward_ddim, polar_interpolate
from oracle.synth import synth_input, synth_state_dict

GOLD = os.path.join(os.path.dirname(__file__), "golden")
torch.set_grad_enabled(False)


def _load(name):
    return np.load(os.path.join(GOLD, name))


@pytest.mark.parametrize("case", G.ATTN_CASES, ids=[c[0] for c in G.ATTN_CASES])
def test_attention_processors_match_reference(case):
    name, C, heads, ctx_dim, N, T, scale, B = case
    gold = _load("attn.npz")[name]
    attn = G.make_attn(C, heads, ctx_dim)
    x = synth_input(f"attn/{name}/x", (B, N, C))
    if ctx_dim is None:
        y = AttnProcessor2_0()(attn, x)
    else:
        ctx = synth_input(f"attn/{name}/ctx", (B, T, ctx_dim))
        p = IPAttnProcessor2_0(C, ctx_dim, scale=scale, num_tokens=4)
        p.load_state_dict(synth_state_dict(p, 1))
        y = p(attn, x, encoder_hidden_states=ctx)
    np.testing.assert_allclose(y.numpy(), gold, rtol=1e-5, atol=1e-5)


def test_image_proj_matches_reference():
    gold = _load("image_proj.npz")
    m = ImageProjModel(cross_attention_dim=64, clip_embeddings_dim=48, clip_extra_context_tokens=4)
    m.load_state_dict(synth_state_dict(m, 2))
    e = synth_input("image_proj/e", (3, 2, 48))
    for mode, scales in [("global", [1.0, 1.0]), ("local", [1.0, 0.5]), ("both", [0.7, 0.25])]:
        np.testing.assert_allclose(m(e, mode, scales=scales).numpy(), gold[mode], rtol=1e-5, atol=1e-6)


def test_backward_ddim_and_polar_match_reference():
    gold = _load("scalar_fns.npz")
    s = DDIMSchedulerOracle()
    s.set_timesteps(50)
    x = synth_input("bd/x", (1, 4, 8, 8))
    eps = synth_input("bd/eps", (1, 4, 8, 8))
    prev = None
    for i, t in enumerate(gold["ts"].tolist()):
        a_p = s.alphas_cumprod[prev] if prev is not None else s.final_alpha_cumprod
        np.testing.assert_allclose(backward_ddim(x, s.alphas_cumprod[t], a_p, eps).numpy(), gold["backward_ddim"][i],
                                   rtol=1e-6, atol=1e-6)
        prev = t
    y = synth_input("polar/y", (1, 4, 8, 8))
    np.testing.assert_allclose(polar_interpolate(x, y, 0.7).numpy(), gold["polar"], rtol=1e-6, atol=1e-6)


def test_ddim_step_inverts_reference_backward_ddim():
    """x_t = _backward_ddim(x_{t-1}, a_t, a_prev, eps) computed BY THE REFERENCE (golden) -> DDIMScheduler.step(eps, t, x_t)
    must give x_{t-1} back: pins the restated step formula, its prev-timestep rule and final_alpha_cumprod on real DDIM
    pairs of the 50-step schedule (t = 1 -> final alpha, 21 -> 1, 41 -> 21)."""
    gold = _load("scalar_fns.npz")
    s = DDIMSchedulerOracle()
    s.set_timesteps(50)
    x = synth_input("bd/x", (1, 4, 8, 8))
    eps = synth_input("bd/eps", (1, 4, 8, 8))
    for i, t in enumerate(gold["ts"].tolist()[:3]):
        back = s.step(eps, t, torch.from_numpy(gold["backward_ddim"][i]))[0]
        np.testing.assert_allclose(back.numpy(), x.numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("case", G.PRIOR_CASES, ids=[c[0] for c in G.PRIOR_CASES])
def test_prior_matches_reference(case):
    name, n_layer, kw = case
    gold = _load("prior.npz")
    p = OraclePrior(n_layer=n_layer).eval()
    p.load_state_dict(synth_state_dict(p, seed=3))
    src, clip_hidden = G.prior_inputs(name)
    trace = []
    torch.manual_seed(1234)
    y, _ = p.generate_diffusion(3, 0, src, clip_hidden, num_inference_steps=kw["num_inference_steps"],
                                guidance_scale=kw["guidance_scale"], score=kw["score"],
                                no_diffusion=kw["no_diffusion"], trace=trace)
    assert len(trace) == int(gold[name + "/nfwd"])
    np.testing.assert_allclose(trace[0]["seq"].numpy(), gold[name + "/seq0"], rtol=1e-5, atol=1e-5)
    ref = torch.from_numpy(gold[name])
    cos = torch.nn.functional.cosine_similarity(y.flatten(), ref.flatten(), dim=0).item()
    rel = ((y - ref).norm() / ref.norm()).item()
    assert cos > 0.99999 and rel < 2e-3, (cos, rel)


@pytest.mark.skipif(not ref_shims.available(), reason="reference tree not mounted (GPU box)")
def test_live_reference_processor_agrees():
    """When /root/reference is present, also run the reference file live (guards stale fixtures)."""
    ref = ref_shims.load_attention_processors()
    name, C, heads, ctx_dim, N, T, scale, B = G.ATTN_CASES[1]
    attn = G.make_attn(C, heads, ctx_dim)
    x = synth_input("live/x", (B, N, C))
    ctx = synth_input("live/ctx", (B, T, ctx_dim))
    pr = ref.IPAttnProcessor2_0(hidden_size=C, cross_attention_dim=ctx_dim, scale=0.8, num_tokens=4)
    pr.load_state_dict(synth_state_dict(pr, 1))
    po = IPAttnProcessor2_0(C, ctx_dim, scale=0.8, num_tokens=4)
    po.load_state_dict(synth_state_dict(po, 1))
    torch.testing.assert_close(po(attn, x, encoder_hidden_states=ctx), pr(attn, x, encoder_hidden_states=ctx),
                               rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------ in-tree LDM blocks
# tests/golden/ldm_blocks.npz = outputs of the reference's own llm/model/vae/modules/{blocks,attention,util}.py, the code the
# (absent) diffusers blocks are ports of; the oracle restatements must reproduce them from the same name-seeded weights.
@pytest.fixture(scope="module")
def ldm():
    return _load("ldm_blocks.npz"), G.ldm_oracle_modules(), G.ldm_inputs()


@pytest.mark.parametrize("case", G.LDM_RES_CASES, ids=[c[0] for c in G.LDM_RES_CASES])
def test_resnet_block_matches_reference_ldm(ldm, case):
    gold, mods, x = ldm
    np.testing.assert_allclose(mods[case[0]](x["res_x"], x["res_temb"]).numpy(), gold[case[0]], rtol=1e-5, atol=2e-5)


def test_transformer_blocks_match_reference_ldm(ldm):
    """GEGLU feed-forward, BasicTransformerBlock (LN -> self-attn -> LN -> cross-attn -> LN -> FF, residuals) and the
    Transformer2D wrapper (GroupNorm eps 1e-6 -> proj_in -> blocks -> proj_out -> + input)"""
    gold, mods, x = ldm
    np.testing.assert_allclose(mods["ff"](x["tok"]).numpy(), gold["ff"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(mods["tblock"](x["tok"], x["ctx"]).numpy(), gold["tblock"], rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(mods["t2d"](x["map"], x["ctx"]).numpy(), gold["t2d"], rtol=1e-5, atol=2e-5)


def test_vae_trunks_match_reference_ldm(ldm):
    """AutoencoderKL decoder / encoder trunks == the in-tree LDM Decoder / Encoder (ragged 6x4 latent, 32x48 image; the
    encoder's pad-end stride-2 convs and the single-head mid attention included)"""
    gold, mods, x = ldm
    np.testing.assert_allclose(mods["vae_dec"].trunk(x["z"]).numpy(), gold["vae_dec"], rtol=1e-4, atol=5e-5)
    np.testing.assert_allclose(mods["vae_enc"].trunk(x["img"]).numpy(), gold["vae_enc"], rtol=1e-4, atol=5e-5)


def test_sinusoid_and_schedule_tables_match_reference_ldm(ldm):
    from oracle.schedulers import get_timestep_embedding
    gold, _, x = ldm
    for dim in (320, 256):          # UNet time_proj (320) and add_time_proj (256): flip_sin_to_cos=True, freq_shift 0
        np.testing.assert_allclose(get_timestep_embedding(x["t"], dim, flip_sin_to_cos=True, downscale_freq_shift=0).numpy(),
                                   gold[f"sinusoid_{dim}"], rtol=1e-6, atol=1e-6)
    s = DDIMSchedulerOracle()
    # the reference builds the schedule in fp64, diffusers (and the oracle) in fp32: agreement to fp32 round-off
    np.testing.assert_allclose(s.alphas_cumprod.numpy(), gold["alphas_cumprod"], rtol=2e-5)
    for n in (50, 25):
        s.set_timesteps(n)
        assert s.timesteps.tolist() == gold[f"ddim_timesteps_{n}"][::-1].tolist()
    s.set_timesteps(50)
    ts = s.timesteps.tolist()[::-1]
    np.testing.assert_allclose(s.alphas_cumprod[ts].numpy(), gold["ddim_alphas_50"], rtol=2e-5)
    # a_prev of the first (lowest) step: alphas_cumprod[0] = diffusers' final_alpha_cumprod with set_alpha_to_one=False
    assert abs(float(s.final_alpha_cumprod) - gold["ddim_alphas_prev_50"][0]) < 1e-6


@pytest.mark.skipif(not ref_shims.available(), reason="reference tree not mounted (GPU box)")
def test_live_reference_ldm_resnet_agrees():
    from oracle import ldm_map as M
    B, _, _ = ref_shims.load_ldm_modules()
    name, cin, cout, tw = G.LDM_RES_CASES[0]
    m = G.ldm_oracle_modules()[name]
    ref = B.ResnetBlock(in_channels=cin, out_channels=cout, dropout=0.0, temb_channels=tw).eval()
    ref.load_state_dict(M.resnet_to_ldm(m.state_dict()))
    x, temb = synth_input("live/res/x", (1, cin, 7, 9)), synth_input("live/res/temb", (1, tw))
    torch.testing.assert_close(m(x, temb), ref(x, temb), rtol=1e-5, atol=2e-5)
