"""Row a13: ``B200ImageProj`` = ImageProjModel + the tensor half of IPAdapter.get_image_embeds, against outputs of the
reference's own code (tests/golden/image_proj.npz, k64/*: made by oracle/gen_golden.py::gen_image_proj)."""
import os

import numpy as np
import pytest
import torch

from instructany2pix_b200.image_proj import B200ImageProj
from oracle import gen_golden as G
from oracle.attention import ImageProjModel, get_image_embeds
from oracle.synth import synth_state_dict

torch.set_grad_enabled(False)
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "image_proj.npz"))
CFG = G.IMAGE_PROJ_K64


def _oracle():
    m = ImageProjModel(cross_attention_dim=CFG["cross"], clip_embeddings_dim=CFG["clip"], clip_extra_context_tokens=4)
    m.load_state_dict(synth_state_dict(m, 2))
    return m.eval()


def _check(m, dev, tol, tol16):
    e, el = (t.to(dev) for t in G.image_proj_k64_inputs())
    for mode, scales in G.IMAGE_PROJ_MODES:
        y = m(torch.stack([e, el], 1), mode, scales=scales)
        assert y.dtype == torch.float32
        np.testing.assert_allclose(y.cpu().numpy(), GOLD["k64/" + mode], rtol=tol, atol=tol)
    # get_image_embeds: the reference ran this leg in fp16 (ip_adapter.py:182) -> fp16-level tolerance
    c, u = m.get_image_embeds(clip_image_embeds=e, mode="global") if hasattr(m, "get_image_embeds") else get_image_embeds(m, e)
    np.testing.assert_allclose(c.cpu().numpy(), GOLD["k64/gie_cond"], rtol=tol16, atol=tol16)
    np.testing.assert_allclose(u.cpu().numpy(), GOLD["k64/gie_uncond"], rtol=tol16, atol=tol16)
    kw = dict(clip_image_embeds=e, clip_image_embeds_local=el, mode="both", scale_g=1.0, scale_l=0.4)
    c, u = m.get_image_embeds(**kw) if hasattr(m, "get_image_embeds") else get_image_embeds(m, **kw)
    np.testing.assert_allclose(c.cpu().numpy(), GOLD["k64/gie_both_cond"], rtol=tol16, atol=tol16)
    np.testing.assert_allclose(u.cpu().numpy(), GOLD["k64/gie_both_uncond"], rtol=tol16, atol=tol16)


def test_oracle_matches_reference():
    _check(_oracle(), "cpu", 1e-5, 1e-2)


def test_state_dict_keys_match_reference_layout():
    b = B200ImageProj(CFG["cross"], CFG["clip"], 4, device="meta")
    assert {k: tuple(v.shape) for k, v in b.state_dict().items()} == {k: tuple(v.shape) for k, v in _oracle().state_dict().items()}
    assert set(b.state_dict()) == {"proj.weight", "proj.bias", "norm.weight", "norm.bias", "raw_embed"}


def test_host_logic_matches_reference_with_kernel_double(emu):
    """input-side blend, per-crop bias fold, crop selection / ordering, zeros for a missing crop, default scales for uncond"""
    _check(B200ImageProj.from_module(_oracle(), device="cpu"), "cpu", 2e-5, 1e-2)


@pytest.mark.gpu
def test_gpu_matches_reference():
    _check(B200ImageProj.from_module(_oracle(), device="cuda"), "cuda", 2e-4, 1e-2)


@pytest.mark.gpu
def test_gpu_production_width_matches_oracle():
    """SDXL IP-adapter widths: Linear(1024 -> 4 x 2048), LayerNorm(2048), batch 8"""
    o = ImageProjModel()
    o.load_state_dict(synth_state_dict(o, 5))
    b = B200ImageProj.from_module(o.eval(), device="cuda")
    e = torch.randn(8, 1024, generator=torch.Generator().manual_seed(3))
    e = e / e.norm(dim=1, keepdim=True) * 20.0                                    # pipeline.py:168,324
    c, u = b.get_image_embeds(clip_image_embeds=e.cuda())
    co, uo = get_image_embeds(o, e)
    assert c.shape == (8, 4, 2048)
    assert ((c.cpu() - co).norm() / co.norm()).item() < 1e-4 and ((u.cpu() - uo).norm() / uo.norm()).item() < 1e-4
