"""Oracle (test infrastructure): restated diffusers==0.26.3 ``AutoencoderKL`` DECODER (SDXL VAE config), used only to turn
latents into images for the north-star PSNR gate ("decoded image PSNR >= 35 dB, same decoder on both latents").

Third-party algorithm (requirements.txt:3), reference call site ddim/sdxl_pipeline.py:859-871 (``vae.decode(latents /
scaling_factor)``, fp32 upcast).  The diffusers module itself is absent, but it is a key-renamed port of the LDM
``Encoder`` / ``Decoder`` the reference carries in-tree (llm/model/vae/modules/blocks.py:369-570): both trunks are PINNED
against that code run in the build container (oracle/gen_golden.py::gen_ldm -> tests/golden/ldm_blocks.npz,
tests/test_oracle_golden.py::test_vae_trunks_match_reference_ldm); only the two 1x1 (post_)quant convs and the SDXL channel
widths are taken from the published config.  Structure follows the published SDXL VAE config:
block_out_channels (128,256,512,512), layers_per_block 2 (decoder uses 3 resnets per up block), one single-head mid attention,
GroupNorm(32, eps 1e-6), SiLU, scaling_factor 0.13025.  ``TINY_VAE`` shrinks the widths for fast tests.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

SDXL_VAE = dict(latent_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                norm_num_groups=32, scaling_factor=0.13025)
TINY_VAE = dict(latent_channels=4, out_channels=3, block_out_channels=(32, 32, 64, 64), layers_per_block=1,
                norm_num_groups=32, scaling_factor=0.13025)


class _Res(nn.Module):
    def __init__(self, cin, cout, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class _Attn(nn.Module):
    def __init__(self, c, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, c, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x).view(b, c, h * w).transpose(1, 2)
        o = F.scaled_dot_product_attention(self.to_q(t)[:, None], self.to_k(t)[:, None], self.to_v(t)[:, None])[:, 0]
        return x + self.to_out[0](o).transpose(1, 2).reshape(b, c, h, w)


class _Up(nn.Module):
    def __init__(self, cin, cout, n, groups, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([_Res(cin if i == 0 else cout, cout, groups) for i in range(n)])
        if add_up:
            self.upsamplers = nn.ModuleList([nn.ModuleDict(dict(conv=nn.Conv2d(cout, cout, 3, padding=1)))])
        self.add_up = add_up

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.add_up:
            x = self.upsamplers[0]["conv"](F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class OracleVAEDecoder(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        cfg = dict(SDXL_VAE if cfg is None else cfg)
        self.cfg = cfg
        ch, G = cfg["block_out_channels"], cfg["norm_num_groups"]
        self.post_quant_conv = nn.Conv2d(cfg["latent_channels"], cfg["latent_channels"], 1)
        self.conv_in = nn.Conv2d(cfg["latent_channels"], ch[-1], 3, padding=1)
        self.mid_res0, self.mid_attn, self.mid_res1 = _Res(ch[-1], ch[-1], G), _Attn(ch[-1], G), _Res(ch[-1], ch[-1], G)
        rch = list(reversed(ch))
        ups, prev = [], rch[0]
        for i, c in enumerate(rch):
            ups.append(_Up(prev, c, cfg["layers_per_block"] + 1, G, add_up=i < len(rch) - 1))
            prev = c
        self.up_blocks = nn.ModuleList(ups)
        self.conv_norm_out = nn.GroupNorm(G, ch[0], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[0], cfg["out_channels"], 3, padding=1)

    @torch.no_grad()
    def decode(self, latents):
        """latents (B,4,L,L) as produced by the sampler -> images (B,3,8L,8L) in [-1, 1] nominal range."""
        return self.trunk(self.post_quant_conv(latents.float() / self.cfg["scaling_factor"]))

    @torch.no_grad()
    def trunk(self, z):
        """``Decoder`` proper (conv_in .. conv_out).  Pinned against the reference's in-tree LDM ``Decoder``
        (llm/model/vae/modules/blocks.py:463-570, the module diffusers' AutoencoderKL decoder is a port of) by
        tests/golden/ldm_blocks.npz."""
        x = self.conv_in(z)
        x = self.mid_res1(self.mid_attn(self.mid_res0(x)))
        for u in self.up_blocks:
            x = u(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class _Down(nn.Module):
    def __init__(self, cin, cout, n, groups, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([_Res(cin if i == 0 else cout, cout, groups) for i in range(n)])
        if add_down:
            # [3P] Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) then a stride-2 3x3 conv without padding
            self.downsamplers = nn.ModuleList([nn.ModuleDict(dict(conv=nn.Conv2d(cout, cout, 3, stride=2, padding=0)))])
        self.add_down = add_down

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        if self.add_down:
            x = self.downsamplers[0]["conv"](F.pad(x, (0, 1, 0, 1)))
        return x


class OracleVAEEncoder(nn.Module):
    """Restated diffusers==0.26.3 ``AutoencoderKL`` ENCODER + ``quant_conv`` + ``DiagonalGaussianDistribution`` (SDXL VAE
    config): the path ``prepare_latents`` of the inversion pipeline takes (ddim/pnp_pipeline.py:195-204:
    ``vae.encode(image).latent_dist.sample(generator) * vae.config.scaling_factor``).  The trunk is pinned against the in-tree LDM
    ``Encoder`` (see the module header); ``quant_conv`` + the gaussian sample are restated only."""

    def __init__(self, cfg=None, in_channels=3):
        super().__init__()
        cfg = dict(SDXL_VAE if cfg is None else cfg)
        self.cfg = cfg
        ch, G = cfg["block_out_channels"], cfg["norm_num_groups"]
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        downs, prev = [], ch[0]
        for i, c in enumerate(ch):
            downs.append(_Down(prev, c, cfg["layers_per_block"], G, add_down=i < len(ch) - 1))
            prev = c
        self.down_blocks = nn.ModuleList(downs)
        self.mid_res0, self.mid_attn, self.mid_res1 = _Res(ch[-1], ch[-1], G), _Attn(ch[-1], G), _Res(ch[-1], ch[-1], G)
        self.conv_norm_out = nn.GroupNorm(G, ch[-1], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[-1], 2 * cfg["latent_channels"], 3, padding=1)
        self.quant_conv = nn.Conv2d(2 * cfg["latent_channels"], 2 * cfg["latent_channels"], 1)

    @torch.no_grad()
    def moments(self, images):
        return self.quant_conv(self.trunk(images))

    @torch.no_grad()
    def trunk(self, images):
        """``Encoder`` proper (conv_in .. conv_out); pinned against the in-tree LDM ``Encoder`` (blocks.py:369-460)."""
        x = self.conv_in(images.float())
        for d in self.down_blocks:
            x = d(x)
        x = self.mid_res1(self.mid_attn(self.mid_res0(x)))
        return self.conv_out(F.silu(self.conv_norm_out(x)))

    @torch.no_grad()
    def encode(self, images, noise=None):
        """images (B,3,H,W) in [-1,1] -> latents (B,4,H/8,W/8) = sample * scaling_factor (noise None: the mode)."""
        mean, logvar = self.moments(images).chunk(2, dim=1)
        z = mean if noise is None else mean + torch.exp(0.5 * logvar.clamp(-30.0, 20.0)) * noise
        return z * self.cfg["scaling_factor"]


def to_diffusers_keys(sd, side):
    """state dict of an Oracle VAE half -> the diffusers AutoencoderKL key names ``B200VAE`` uses."""
    out = {}
    for k, v in sd.items():
        if k.startswith(("post_quant_conv", "quant_conv")):
            out[k] = v
            continue
        k = k.replace("mid_res0", "mid_block.resnets.0").replace("mid_res1", "mid_block.resnets.1").replace("mid_attn", "mid_block.attentions.0")
        out[f"{side}.{k}"] = v
    return out


def psnr(a, b, data_range=2.0):
    """PSNR in dB between two image batches in [-1, 1] (data_range 2), computed over the whole batch."""
    mse = ((a.float() - b.float()) ** 2).mean().clamp_min(1e-20)
    return float(10.0 * torch.log10(torch.tensor(data_range ** 2) / mse))
