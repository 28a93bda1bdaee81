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


def gen_image_proj():
    Ref = ref_shims.extract_source("instructany2pix/diffusion/ip_adapter/ip_adapter.py", "ImageProjModel")
    m = Ref(cross_attention_dim=64, clip_embeddings_dim=48, clip_extra_context_tokens=4)
    m.load_state_dict(synth_state_dict(m, 2))
    e = synth_input("image_proj/e", (3, 2, 48))
    out = {}
    with torch.no_grad():
        for mode, scales in [("global", [1.0, 1.0]), ("local", [1.0, 0.5]), ("both", [0.7, 0.25])]:
            out[mode] = m(e.clone(), mode, scales=scales).numpy()
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


def main():
    assert ref_shims.available(), "needs /root/reference (build container only)"
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    gen_attention()
    gen_image_proj()
    gen_scalar_fns()
    gen_prior()
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
