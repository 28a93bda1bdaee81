"""Oracle (test infrastructure): generate tests/golden/*.npz by RUNNING THE REFERENCE CODE.

Run in the build container only (needs /root/reference):  ``python -m oracle.gen_golden``.
Every fixture stores the seeds/sizes that regenerate its inputs and weights (oracle/synth.py:
name-seeded tensors) plus the reference's outputs, so the fixtures stay a few KB each and the
GPU box -- which has no /root/reference -- can still check the oracle (and, through the oracle,
the CUDA path) against numbers produced by the reference itself.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ref_shims
from .attention import Attention
from .synth import synth_input, synth_state_dict

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

ATTN_CASES = [  # name, C, heads, ctx_dim, N, n_ctx_tokens, scale, batch
    ("self", 128, 2, None, 40, 0, 1.0, 2),
    ("ip81", 128, 2, 96, 40, 81, 1.0, 2),
    ("ip77_quirk", 128, 2, 96, 24, 77, 0.6, 1),     # inversion: last 4 TEXT tokens become "image" tokens
    ("ip81_scale0", 192, 3, 96, 17, 81, 0.0, 2),
]


def make_attn(C, heads, ctx_dim, seed=0):
    a = Attention(C, heads, C // heads, cross_attention_dim=ctx_dim)
    a.load_state_dict(synth_state_dict(a, seed))
    return a.eval()


def gen_attention():
    ref = ref_shims.load_attention_processors()
    out = {}
    for name, C, heads, ctx_dim, N, T, scale, B in ATTN_CASES:
        attn = make_attn(C, heads, ctx_dim)
        x = synth_input(f"attn/{name}/x", (B, N, C))
        if ctx_dim is None:
            proc = ref.AttnProcessor2_0()
            y = proc(attn, x)
        else:
            ctx = synth_input(f"attn/{name}/ctx", (B, T, ctx_dim))
            proc = ref.IPAttnProcessor2_0(hidden_size=C, cross_attention_dim=ctx_dim, scale=scale, num_tokens=4)
            proc.load_state_dict(synth_state_dict(proc, 1))
            with torch.no_grad():
                y = proc(attn, x, encoder_hidden_states=ctx)
        out[name] = y.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "attn.npz"), **out)


IMAGE_PROJ_K64 = dict(cross=128, clip=64)
IMAGE_PROJ_MODES = [("global", [1.0, 1.0]), ("local", [1.0, 0.5]), ("both", [0.7, 0.25])]


def image_proj_k64_inputs():
    return synth_input("image_proj/k64/e", (3, IMAGE_PROJ_K64["clip"])), synth_input("image_proj/k64/el", (3, IMAGE_PROJ_K64["clip"]))


def gen_image_proj():
    Ref = ref_shims.extract_source("instructany2pix/diffusion/ip_adapter/ip_adapter.py", "ImageProjModel")
    m = Ref(cross_attention_dim=64, clip_embeddings_dim=48, clip_extra_context_tokens=4)
    m.load_state_dict(synth_state_dict(m, 2))
    e = synth_input("image_proj/e", (3, 2, 48))
    out = {}
    with torch.no_grad():
        for mode, scales in [("global", [1.0, 1.0]), ("local", [1.0, 0.5]), ("both", [0.7, 0.25])]:
            out[mode] = m(e.clone(), mode, scales=scales).numpy()
    # second configuration with kernel-friendly widths (K multiple of 32) for the GPU drop-in (B200ImageProj), non-zero
    # raw_embed, plus IPAdapter.get_image_embeds itself (ip_adapter.py:171-209, a method: extracted and run with a stand-in
    # ``self``; the reference casts the embeddings to fp16 there, so the module is run in fp16 like the reference does)
    m2 = Ref(cross_attention_dim=IMAGE_PROJ_K64["cross"], clip_embeddings_dim=IMAGE_PROJ_K64["clip"], clip_extra_context_tokens=4)
    m2.load_state_dict(synth_state_dict(m2, 2))
    e2, e2l = image_proj_k64_inputs()
    with torch.no_grad():
        for mode, scales in IMAGE_PROJ_MODES:
            out["k64/" + mode] = m2(torch.stack([e2, e2l], 1), mode, scales=scales).numpy()
        import types
        from PIL import Image
        gie = ref_shims.extract_source("instructany2pix/diffusion/ip_adapter/ip_adapter.py", "get_image_embeds", {"Image": Image})
        import contextlib
        import io
        fake = types.SimpleNamespace(image_proj_model=m2.half(), device="cpu")
        with contextlib.redirect_stdout(io.StringIO()):          # the reference prints a comparison tensor
            c, u = gie(fake, clip_image_embeds=e2, mode="global", scale_g=1.0, scale_l=1.0)
            c2, u2 = gie(fake, clip_image_embeds=e2, clip_image_embeds_local=e2l, mode="both", scale_g=1.0, scale_l=0.4)
        out["k64/gie_cond"], out["k64/gie_uncond"] = c.float().numpy(), u.float().numpy()
        out["k64/gie_both_cond"], out["k64/gie_both_uncond"] = c2.float().numpy(), u2.float().numpy()
    np.savez_compressed(os.path.join(OUT, "image_proj.npz"), **out)


def gen_scalar_fns():
    bd = ref_shims.extract_source("instructany2pix/ddim/pnp_pipeline.py", "_backward_ddim")
    polar = ref_shims.extract_source("instructany2pix/pipeline.py", "polar_intrtpolate")
    from .schedulers import DDIMSchedulerOracle
    s = DDIMSchedulerOracle()
    s.set_timesteps(50)
    x = synth_input("bd/x", (1, 4, 8, 8))
    eps = synth_input("bd/eps", (1, 4, 8, 8))
    outs = []
    ts = [1, 21, 41, 981]
    prev = None
    for t in ts:
        a_t = s.alphas_cumprod[t]
        a_p = s.alphas_cumprod[prev] if prev is not None else s.final_alpha_cumprod
        outs.append(bd(x, a_t, a_p, eps).numpy())
        prev = t
    y = synth_input("polar/y", (1, 4, 8, 8))
    np.savez_compressed(os.path.join(OUT, "scalar_fns.npz"), backward_ddim=np.stack(outs), ts=np.array(ts),
                        polar=polar(None, x, y, 0.7).numpy())


PRIOR_CASES = [  # name, n_layer, kwargs
    ("l2_nodiff", 2, dict(no_diffusion=True, num_inference_steps=25, guidance_scale=10, score=6.5, force_guidence_t0=True)),
    ("l2_diff25", 2, dict(no_diffusion=False, num_inference_steps=25, guidance_scale=10, score=6.5)),
    ("l24_nodiff", 24, dict(no_diffusion=True, num_inference_steps=25, guidance_scale=10, score=6.5, force_guidence_t0=True)),
    ("l24_diff4", 24, dict(no_diffusion=False, num_inference_steps=4, guidance_scale=5, score=6.8)),
]


def prior_inputs(name):
    e = synth_input(f"prior/{name}/src", (1, 1, 1024))
    src = e / e.norm() * 100.0                                 # pipeline.py:313
    clip_hidden = synth_input(f"prior/{name}/clip", (1, 2, 1024), scale=0.5)
    return src, clip_hidden


def gen_prior():
    out = {}
    for name, n_layer, kw in PRIOR_CASES:
        prior, mod, fake = ref_shims.load_prior(n_layer=n_layer)
        sd = synth_state_dict({k: v.shape for k, v in prior.state_dict().items()}, seed=3)
        missing, unexpected = prior.load_state_dict(sd, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        src, clip_hidden = prior_inputs(name)
        fake.hidden = clip_hidden
        captured = []
        orig_fwd = prior.model.forward

        def spy(*a, **k):
            captured.append(k["inputs_embeds"].detach().clone())
            return orig_fwd(*a, **k)

        prior.model.forward = spy
        torch.manual_seed(1234)
        y, _ = prior.generate_diffusion(mod.MODALITY.VIDEO, mod.MODALITY.IMAGE, src, device="cpu",
                                        dtype=torch.float32, image_bind_overwrite=None,
                                        do_classifier_free_guidance=True, **kw)
        out[name] = y.detach().numpy()
        out[name + "/seq0"] = captured[0].numpy()
        out[name + "/nfwd"] = np.array(len(captured))
        print(name, tuple(y.shape), captured[0].shape, len(captured), float(y.norm()))
    np.savez_compressed(os.path.join(OUT, "prior.npz"), **out)


# ---------------------------------------------------------------------------------------------- in-tree LDM blocks
# The diffusers blocks of the hot path are absent from /root/reference, but the LDM modules they are ports of are in-tree
# (llm/model/vae/modules).  Each case below builds the ORACLE module, draws its name-seeded weights, renames them
# (oracle/ldm_map.py) into the reference module and records what the reference computes.
LDM_VAE_CFG = dict(latent_channels=4, out_channels=3, block_out_channels=(64, 64, 128, 128), layers_per_block=1,
                   norm_num_groups=32, scaling_factor=0.13025)
LDM_RES_CASES = [("res_64_128", 64, 128, 96), ("res_64_64", 64, 64, 96)]       # name, cin, cout, temb width
LDM_TBLOCK = dict(dim=128, heads=2, ctx_dim=96, depth=2, tokens_hw=(6, 5), ctx_tokens=11, batch=2)


def ldm_oracle_modules():
    """the oracle modules of every LDM case with their name-seeded weights loaded: {case: module}"""
    from . import unet as U
    from .vae import OracleVAEDecoder, OracleVAEEncoder
    mods = {}
    for name, cin, cout, tw in LDM_RES_CASES:
        mods[name] = U.ResnetBlock2D(cin, cout, tw, 32, 1e-6)        # LDM Normalize: eps 1e-6
    mods["ff"] = U.FeedForward(LDM_TBLOCK["dim"])
    mods["tblock"] = U.BasicTransformerBlock(LDM_TBLOCK["dim"], LDM_TBLOCK["heads"], LDM_TBLOCK["ctx_dim"])
    mods["t2d"] = U.Transformer2DModel(LDM_TBLOCK["dim"], LDM_TBLOCK["heads"], LDM_TBLOCK["depth"], LDM_TBLOCK["ctx_dim"], 32)
    mods["vae_dec"] = OracleVAEDecoder(LDM_VAE_CFG)
    mods["vae_enc"] = OracleVAEEncoder(LDM_VAE_CFG)
    for i, (k, m) in enumerate(sorted(mods.items())):
        m.load_state_dict(synth_state_dict(m, 40 + i))
        m.eval()
    return mods


def ldm_inputs():
    T = LDM_TBLOCK
    h, w = T["tokens_hw"]
    return dict(
        res_x=synth_input("ldm/res/x", (2, 64, 12, 10)), res_temb=synth_input("ldm/res/temb", (2, 96)),
        tok=synth_input("ldm/tok", (T["batch"], h * w, T["dim"])), ctx=synth_input("ldm/ctx", (T["batch"], T["ctx_tokens"], T["ctx_dim"])),
        map=synth_input("ldm/map", (T["batch"], T["dim"], h, w)),
        z=synth_input("ldm/z", (2, 4, 6, 4)), img=synth_input("ldm/img", (2, 3, 32, 48)),
        t=torch.tensor([0.0, 1.0, 21.0, 981.0, 999.0, 500.5]),
    )


def gen_ldm():
    from . import ldm_map as M
    B, A, Ut = ref_shims.load_ldm_modules()
    mods, x = ldm_oracle_modules(), ldm_inputs()
    out = {}

    def load(ref, sd):
        missing, unexpected = ref.load_state_dict(sd, strict=True)
        assert not missing and not unexpected
        return ref.eval()

    for name, cin, cout, tw in LDM_RES_CASES:
        ref = load(B.ResnetBlock(in_channels=cin, out_channels=cout, dropout=0.0, temb_channels=tw), M.resnet_to_ldm(mods[name].state_dict()))
        out[name] = ref(x["res_x"], x["res_temb"]).numpy()
    T = LDM_TBLOCK
    dh = T["dim"] // T["heads"]
    out["ff"] = load(A.FeedForward(T["dim"], glu=True), mods["ff"].state_dict())(x["tok"]).numpy()
    ref = load(A.BasicTransformerBlock(T["dim"], T["heads"], dh, context_dim=T["ctx_dim"], checkpoint=False), mods["tblock"].state_dict())
    out["tblock"] = ref(x["tok"], context=x["ctx"]).numpy()
    ref = load(A.SpatialTransformer(T["dim"], T["heads"], dh, depth=T["depth"], context_dim=T["ctx_dim"]), M.transformer2d_to_ldm(mods["t2d"].state_dict()))
    for blk in ref.transformer_blocks:
        blk.checkpoint = False
    out["t2d"] = ref(x["map"], context=x["ctx"]).numpy()
    ch = LDM_VAE_CFG["block_out_channels"]
    kw = dict(ch=ch[0], out_ch=3, ch_mult=tuple(c // ch[0] for c in ch), num_res_blocks=LDM_VAE_CFG["layers_per_block"],
              attn_resolutions=[], in_channels=3, resolution=256, z_channels=4)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):       # the constructors print
        dec, enc = B.Decoder(**kw), B.Encoder(**kw, double_z=True)
    sd = {k: v for k, v in mods["vae_dec"].state_dict().items() if not k.startswith("post_quant_conv")}
    out["vae_dec"] = load(dec, M.vae_decoder_to_ldm(sd, len(ch)))(x["z"]).numpy()
    sd = {k: v for k, v in mods["vae_enc"].state_dict().items() if not k.startswith("quant_conv")}
    out["vae_enc"] = load(enc, M.vae_encoder_to_ldm(sd, len(ch)))(x["img"]).numpy()
    # scalar tables: sinusoid (util.py:271-292 = diffusers flip_sin_to_cos=True, shift 0), SDXL "scaled_linear" betas
    # (util.py:141-145 "linear", fp64), DDIM "uniform" timesteps + 1 (util.py:166-180 = leading spacing, steps_offset 1)
    out["sinusoid_320"] = Ut.timestep_embedding(x["t"], 320).numpy()
    out["sinusoid_256"] = Ut.timestep_embedding(x["t"], 256).numpy()
    betas = Ut.make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
    out["alphas_cumprod"] = np.cumprod(1.0 - betas, axis=0)
    out["ddim_timesteps_50"] = Ut.make_ddim_timesteps("uniform", 50, 1000, verbose=False)
    out["ddim_timesteps_25"] = Ut.make_ddim_timesteps("uniform", 25, 1000, verbose=False)
    _, a, a_prev = Ut.make_ddim_sampling_parameters(out["alphas_cumprod"], out["ddim_timesteps_50"], 0.0, verbose=False)
    out["ddim_alphas_50"], out["ddim_alphas_prev_50"] = a, a_prev
    np.savez_compressed(os.path.join(OUT, "ldm_blocks.npz"), **out)
    print("ldm:", {k: tuple(v.shape) for k, v in out.items()})


def main():
    assert ref_shims.available(), "needs /root/reference (build container only)"
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    gen_attention()
    gen_image_proj()
    gen_scalar_fns()
    gen_prior()
    gen_ldm()
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
