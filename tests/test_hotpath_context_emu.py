"""``B200HotPath.ip_context``: the conditioning hand-over of an edit request (pipeline.py:324-326 mix, ip_adapter.py:171-209 projector,
ip_adapter.py:336-342 concatenation) against a plain restatement built from the oracle projector, with the kernel test double."""
import torch

from instructany2pix_b200.hotpath import B200HotPath
from instructany2pix_b200.image_proj import B200ImageProj
from oracle.attention import ImageProjModel, get_image_embeds
from oracle.synth import synth_input, synth_state_dict

torch.set_grad_enabled(False)


class _NoUNet:                      # ip_context touches neither the UNet nor the VAE
    pass


def _hot(monkeypatch):
    o = ImageProjModel(cross_attention_dim=128, clip_embeddings_dim=64, clip_extra_context_tokens=4)
    o.load_state_dict(synth_state_dict(o, 7))
    import instructany2pix_b200.hotpath as hp
    monkeypatch.setattr(hp, "B200Sampler", lambda *a, **k: None)
    return o.eval(), B200HotPath(_NoUNet(), None, image_proj=B200ImageProj.from_module(o, device="cpu"))


def test_ip_context_matches_the_reference_formula(emu, monkeypatch):
    o, hot = _hot(monkeypatch)
    B = 3
    text = synth_input("hot/text", (2 * B, 77, 128))
    llm, y, base = synth_input("hot/llm", (B, 64)), synth_input("hot/y", (B, 1, 64)), synth_input("hot/base", (B, 64))
    h, norm = (0.2, 0.4, 1.0), 20.0
    ctx = hot.ip_context(text, llm, prior_embed=y, base_embed=base, h=h, norm=norm)
    # pipeline.py:324-326, per request
    la = base * h[0] + llm * h[1] + y[:, 0] / y[:, 0].norm(dim=-1, keepdim=True) * 20.0 * h[2]
    la = la / la.norm(dim=-1, keepdim=True) * norm
    cond, uncond = get_image_embeds(o, la)
    assert ctx.shape == (2 * B, 81, 128)
    assert torch.equal(ctx[:, :77], text)                                   # text tokens untouched, [negative ; positive] order kept
    torch.testing.assert_close(ctx[B:, 77:], cond, rtol=2e-5, atol=2e-5)   # positive rows carry the projected embedding
    torch.testing.assert_close(ctx[:B, 77:], uncond, rtol=2e-5, atol=2e-5)  # negative rows: projector(zeros)


def test_ip_context_without_prior_uses_the_llm_embedding(emu, monkeypatch):
    o, hot = _hot(monkeypatch)
    text, llm = synth_input("hot/text2", (2, 77, 128)), synth_input("hot/llm2", (1, 64))
    ctx = hot.ip_context(text, llm)
    cond, _ = get_image_embeds(o, llm / llm.norm(dim=-1, keepdim=True) * 20.0)
    torch.testing.assert_close(ctx[1:, 77:], cond, rtol=2e-5, atol=2e-5)
