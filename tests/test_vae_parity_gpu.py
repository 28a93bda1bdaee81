"""GPU parity of ``B200VAE`` against the fp32 CPU oracle restatement of the SDXL AutoencoderKL, on
identical name-seeded synthetic weights, plus unit tests of the kernels only the VAE uses.

Tolerances: decoded image PSNR >= 35 dB (north-star figure) and rel-L2 <= 1e-2; encoder latents rel-L2 <= 1e-2.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from instructany2pix_b200 import ops  # noqa: E402
from instructany2pix_b200.vae import B200VAE  # noqa: E402
from oracle.synth import synth_input, synth_state_dict  # noqa: E402
from oracle.vae import SDXL_VAE, OracleVAEDecoder, OracleVAEEncoder, psnr, to_diffusers_keys  # noqa: E402

torch.set_grad_enabled(False)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
# SDXL's block structure (4 levels, 2 layers per block, 3 resnets per decoder level, one mid attention) at a quarter of the width
SMALL = dict(latent_channels=4, out_channels=3, block_out_channels=(64, 64, 128, 128), layers_per_block=2, norm_num_groups=32,
             scaling_factor=0.13025)


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def build(cfg):
    dec, enc = OracleVAEDecoder(cfg).eval(), OracleVAEEncoder(cfg).eval()
    dec.load_state_dict(synth_state_dict(dec, seed=11))
    enc.load_state_dict(synth_state_dict(enc, seed=12))
    vae = B200VAE({k: v for k, v in cfg.items()})
    sd = to_diffusers_keys(dec.state_dict(), "decoder")
    sd.update(to_diffusers_keys(enc.state_dict(), "encoder"))
    vae.load_state_dict(sd)
    return dec, enc, vae


@pytest.mark.parametrize("cfg,L,B", [(SMALL, 16, 2), (SMALL, 24, 1), (SDXL_VAE, 32, 1), (SDXL_VAE, 128, 1)],
                         ids=["small-16", "small-24", "sdxl-width-32", "sdxl-width-128-full-1024px"])
def test_decode_parity(cfg, L, B):
    """the last case is BASELINE.json's full size: a 128x128 latent decoded to 1024^2 at the real SDXL VAE width (16 384-token
    single-head mid attention, 10.5 TFLOP); the fp32 CPU oracle takes some tens of seconds there"""
    dec, _, vae = build(cfg)
    lat = synth_input("vae/lat", (B, 4, L, L), seed=L) * 0.9            # sampler-scale latents (std ~ 0.9)
    ref = dec.decode(lat)
    out = vae.decode(lat.cuda())
    assert out.shape == ref.shape == (B, 3, 8 * L, 8 * L) and out.dtype == torch.float32
    e, db = rel(out, ref), psnr(out.cpu(), ref, data_range=float(ref.max() - ref.min()))
    print(f"VAE decode {L}x{L} latent: rel-L2 {e:.2e}, PSNR {db:.1f} dB")
    assert e < 1e-2 and db >= 35.0


@pytest.mark.parametrize("cfg,S,B", [(SMALL, 128, 2), (SDXL_VAE, 256, 1)], ids=["small-128", "sdxl-width-256"])
def test_encode_parity(cfg, S, B):
    _, enc, vae = build(cfg)
    img = synth_input("vae/img", (B, 3, S, S), seed=S).clamp(-1, 1)
    noise = synth_input("vae/noise", (B, 4, S // 8, S // 8), seed=S)
    ref_mode, ref_sample = enc.encode(img), enc.encode(img, noise)
    out_mode, out_sample = vae.encode(img.cuda(), sample=False), vae.encode(img.cuda(), noise=noise.cuda())
    assert out_mode.shape == ref_mode.shape == (B, 4, S // 8, S // 8)
    print(f"VAE encode {S}x{S}: mode rel-L2 {rel(out_mode, ref_mode):.2e}, sample rel-L2 {rel(out_sample, ref_sample):.2e}")
    assert rel(out_mode, ref_mode) < 1e-2 and rel(out_sample, ref_sample) < 1e-2


def test_round_trip_is_consistent_with_the_oracle_round_trip():
    """size-independent property: decode(encode(x)) through the B200 path tracks the oracle's own round trip."""
    dec, enc, vae = build(SMALL)
    img = synth_input("vae/rt", (1, 3, 192, 192), seed=3).clamp(-1, 1)
    ref = dec.decode(enc.encode(img))
    out = vae.decode(vae.encode(img.cuda(), sample=False))
    e = rel(out, ref)
    print(f"VAE round trip: rel-L2 {e:.2e}")
    assert e < 5e-2            # two random-weight networks back to back: the encoder's 8.5e-3 is amplified by the decoder


# ------------------------------------------------------------------------------------------------ kernels only the VAE uses
@pytest.mark.parametrize("B,H,W,C,Co", [(2, 16, 16, 64, 64), (1, 32, 64, 128, 128), (1, 128, 128, 64, 64)])
def test_conv3x3_down_padend(B, H, W, C, Co):
    from instructany2pix_b200.packing import pack_conv3x3
    g = torch.Generator().manual_seed(H * W + C)
    x = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Co, C, 3, 3, generator=g) * (9 * C) ** -0.5).to(torch.bfloat16).cuda()
    b = torch.randn(Co, generator=g).cuda()
    out = ops.conv3x3_down_padend(x, pack_conv3x3(w), Co, bias=b, out_dtype=torch.float32)
    ref = F.conv2d(F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1)), w.float(), b, stride=2).permute(0, 2, 3, 1)
    assert out.shape == ref.shape and rel(out, ref) < 1e-5


def test_softmax_rows_conv1x1_gaussian_sample():
    g = torch.Generator().manual_seed(7)
    s = (torch.randn(300, 1024, generator=g) * 30).cuda()
    p = ops.softmax_rows(s, 0.05)
    assert rel(p, torch.softmax(s.double() * 0.05, -1)) < 4e-3 and (p.float().sum(-1) - 1).abs().max() < 2e-2
    x = torch.randn(2, 4, 24, 24, generator=g).cuda()
    w, b = torch.randn(4, 4, 1, 1, generator=g).cuda(), torch.randn(4, generator=g).cuda()
    assert rel(ops.conv1x1_nchw_small(x, w, b, 1 / 0.13025), F.conv2d(x / 0.13025, w, b)) < 1e-6
    m = torch.randn(2, 8, 12, 12, generator=g).cuda() * 3
    nz = torch.randn(2, 4, 12, 12, generator=g).cuda()
    mean, logvar = m.chunk(2, 1)
    assert rel(ops.gaussian_sample(m, nz, 0.13025), (mean + torch.exp(0.5 * logvar.clamp(-30, 20)) * nz) * 0.13025) < 1e-6
    assert torch.equal(ops.gaussian_sample(m, None, 1.0), mean.contiguous())


def test_edit_chain_matches_the_oracle_chain():
    """encode -> DDIM inversion -> polar start latent -> CFG sampling -> decode (pipeline.py:303-361 after the encoders), B200
    path vs the same chain on the fp32 oracles with identical weights and noise draws: image PSNR >= 35 dB."""
    from instructany2pix_b200.hotpath import B200HotPath
    from oracle import sampler as osampler
    from oracle.schedulers import polar_interpolate
    from oracle.unet import TINY
    from tests.test_host_unet_emu import build_pair, make_inputs
    dec, enc, vae = build(SMALL)
    o, b = build_pair(True, device="cuda")
    lat, ctx, added = make_inputs(TINY, B=1, L=16)
    img = synth_input("chain/img", (1, 3, 128, 128), seed=1).clamp(-1, 1)
    n_enc, n_pol = synth_input("chain/enc", (1, 4, 16, 16), seed=2), synth_input("chain/pol", (1, 4, 16, 16), seed=3)
    ctx_inv, added_inv = ctx[1:, :77], {k: v[1:] for k, v in added.items()}
    # oracle chain
    z0 = enc.encode(img, n_enc)
    z_inv = osampler.invert(o, z0, ctx_inv, added_inv, num_inference_steps=4)
    z_t = polar_interpolate(z_inv, n_pol, 0.7)
    z = osampler.generate(o, z_t, ctx, added, num_inference_steps=4, guidance_scale=7.5)
    ref = dec.decode(z)
    # B200 chain
    cu = lambda t: {k: v.cuda() for k, v in t.items()} if isinstance(t, dict) else t.cuda()
    hp = B200HotPath(b, vae)
    out, zb = hp.edit(img.cuda(), cu(ctx_inv), cu(added_inv), cu(ctx), cu(added), alpha=0.7, num_inference_steps=4, guidance_scale=7.5,
                      noise=n_pol.cuda(), encode_noise=n_enc.cuda(), return_latents=True)
    s = ref.abs().max().clamp_min(1e-6)
    db = psnr(out.cpu() / s, ref / s)
    print(f"edit chain: final latent rel-L2 {rel(zb, z):.2e}, image PSNR {db:.1f} dB")
    assert db >= 35.0
