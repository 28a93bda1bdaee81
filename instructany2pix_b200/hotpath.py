"""``B200HotPath``: the denoising sequence of ``InstructAny2PixPipeline.__call__`` (pipeline.py:303-361) in one object.

What the reference does per edit request once the LLM / prompt encoders have produced their embeddings (those stay on the
reference code, SURVEY 8 "out of scope"), and which call below replaces it:

  pipeline.py:313-316   y = self.model.generate_diffusion(...)                      -> ``prior.generate_diffusion`` (B200Prior)
  pnp_pipeline.py:195   latents = vae.encode(image).latent_dist.sample() * sf       -> ``vae.encode``
  pnp_pipeline.py:251   DDIM inversion, N UNet forwards, no CFG                      -> ``sampler.invert``
  pipeline.py:332-336   polar_intrtpolate(latent_inv, randn_like(latent_inv), alpha) -> ``sampler.start_latent``
  ip_adapter.py:341-354 50-step (here N) CFG sampling with text + IP tokens          -> ``sampler.generate``
  sdxl_pipeline.py:859  vae.decode(latents / sf)                                     -> ``vae.decode``

Everything runs on the sm_100a kernels; tensors stay on the device between the stages.
"""
from __future__ import annotations

import torch

from .sampler import B200Sampler


class B200HotPath:
    def __init__(self, unet, vae, scheduler=None, prior=None, use_cuda_graph=True):
        self.unet, self.vae, self.prior = unet, vae, prior
        self.sampler = B200Sampler(unet, scheduler=scheduler, use_cuda_graph=use_cuda_graph)

    @torch.no_grad()
    def edit(self, image, ctx_inv, added_inv, ctx_cfg, added_cfg, alpha=0.7, num_inference_steps=25, guidance_scale=10.0,
             noise=None, encode_noise=None, generator=None, return_latents=False):
        """image (B,3,H,W) in [-1,1]; ctx_inv (B,S,D) / added_inv: conditioning of the inversion pass (prompt '' in the
        reference, pipeline.py:330); ctx_cfg (2B,S',D) / added_cfg: [negative; positive] conditioning incl. the IP tokens.
        ``noise`` / ``encode_noise``: optional fixed draws (parity tests); otherwise drawn like the reference does.
        Returns images (B,3,H,W) fp32 (and the final latents if asked)."""
        z0 = self.vae.encode(image, noise=encode_noise, generator=generator)
        z_inv = self.sampler.invert(z0, ctx_inv, added_inv, num_inference_steps=num_inference_steps)
        z_t = self.sampler.start_latent(z_inv, alpha=alpha, noise=noise, generator=generator)
        z = self.sampler.generate(z_t, ctx_cfg, added_cfg, num_inference_steps=num_inference_steps, guidance_scale=guidance_scale)
        img = self.vae.decode(z)
        return (img, z) if return_latents else img

    @torch.no_grad()
    def generate(self, latents, ctx_cfg, added_cfg, num_inference_steps=50, guidance_scale=10.0):
        """text/IP-conditioned generation from given start latents (ip_adapter.py:341-354 + sdxl_pipeline.py:859-871)."""
        return self.vae.decode(self.sampler.generate(latents, ctx_cfg, added_cfg, num_inference_steps=num_inference_steps,
                                                     guidance_scale=guidance_scale))
