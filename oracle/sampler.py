"""Oracle (test infrastructure): restated sampler loops.

  * generate(): the SDXL denoising loop as the reference runs it through IPAdapterXL.generate
    (diffusion/ip_adapter/ip_adapter.py:341-354 -> [3P] StableDiffusionXLPipeline.__call__; in-tree text of
    the loop: diffusion/ip_adapter/custom_pipelines.py:324-363, ddim/sdxl_pipeline.py:823-857):
    x_in = cat([x]*2) -> unet -> eps_u + g (eps_c - eps_u) (uncond FIRST) -> scheduler.step.
  * invert(): DDIM inversion, ddim/pnp_pipeline.py:251-275 (batch 1, no CFG, ascending t).
"""
from __future__ import annotations

import torch

from .schedulers import DDIMSchedulerOracle, backward_ddim


def cfg_inputs(reqs, ip_tokens, ip_tokens_uncond):
    """Build the CFG-doubled conditioning exactly as ip_adapter.py:341-342 + custom_pipelines.py:296-302:
    encoder_hidden_states = cat([neg ⊕ ip_uncond, pos ⊕ ip]) (2B,81,D); text_embeds; time_ids."""
    pos = torch.stack([torch.cat([r["ctx"], ip_tokens[i]], 0) for i, r in enumerate(reqs)])
    neg = torch.stack([torch.cat([r["neg_ctx"], ip_tokens_uncond[i]], 0) for i, r in enumerate(reqs)])
    ctx = torch.cat([neg, pos], 0)
    pooled = torch.cat([torch.stack([r["neg_pooled"] for r in reqs]), torch.stack([r["pooled"] for r in reqs])], 0)
    tid = torch.stack([r["time_ids"] for r in reqs]).repeat(2, 1)
    return ctx, dict(text_embeds=pooled, time_ids=tid)


@torch.no_grad()
def generate(unet, latents, ctx, added, num_inference_steps=50, guidance_scale=10.0, scheduler=None,
             trace=None, teacher=None, init_latents=None, strength=1.0, inpaint_mask=None, masked_image_latents=None):
    """-> final latents (B,4,L,L).  ``trace``: list collecting (t, x_in, eps2B, x_next) per step.
    ``teacher``: optional list of per-step input latents (teacher forcing for per-step parity).
    ``init_latents`` + ``strength``: [3P] StableDiffusionXLImg2ImgPipeline semantics (get_timesteps + add_noise; the refiner call
    at pipeline.py:358-361).  ``inpaint_mask`` (B,1,L,L, 1 = repaint): [3P] StableDiffusionXLInpaintPipeline with a 4-channel UNet
    (gdino/lib.py:85-102): after every step ``latents = (1 - m) * add_noise(init, noise, t_next) + m * latents``.
    With a 9-channel inpainting UNet ([3P] same pipeline, ``num_channels_unet == 9``; not what pipeline.py:132-139 builds -- it
    passes the base UNet -- but the configuration the released SDXL-inpainting checkpoints use) the model input is
    ``cat([scaled latents, mask, masked_image_latents], 1)`` and NO blending happens."""
    nine = getattr(unet.config, "in_channels", 4) == 9
    s = scheduler or DDIMSchedulerOracle()
    s.set_timesteps(num_inference_steps)
    timesteps = s.timesteps
    noise = latents
    if init_latents is None:
        x = latents * s.init_noise_sigma
    else:
        n_run = min(int(num_inference_steps * strength), num_inference_steps)
        timesteps = timesteps[max(num_inference_steps - n_run, 0):]
        if strength >= 1.0 and inpaint_mask is not None:
            x = noise * s.init_noise_sigma
        else:
            x = s.add_noise(init_latents, noise, timesteps[0])
        if hasattr(s, "_step_index"):
            s._step_index = None
    for i, t in enumerate(timesteps):
        if teacher is not None:
            x = teacher[i]
        x_in = s.scale_model_input(torch.cat([x] * 2), t)
        if nine:
            x_in = torch.cat([x_in, torch.cat([inpaint_mask] * 2), torch.cat([masked_image_latents] * 2)], dim=1)
        eps2 = unet(x_in, t, encoder_hidden_states=ctx, added_cond_kwargs=added, return_dict=False)[0]
        e_u, e_c = eps2.chunk(2)
        eps = e_u + guidance_scale * (e_c - e_u)
        x_next = s.step(eps, t, x, eta=0.0)[0]
        if inpaint_mask is not None and not nine:
            keep = init_latents if i + 1 >= len(timesteps) else s.add_noise(init_latents, noise, timesteps[i + 1])
            x_next = (1 - inpaint_mask) * keep + inpaint_mask * x_next
        if trace is not None:
            trace.append(dict(t=int(t), x=x.clone(), eps2=eps2.clone(), x_next=x_next.clone()))
        x = x_next
    return x


@torch.no_grad()
def invert(unet, latents, ctx, added, num_inference_steps=50, scheduler=None, trace=None):
    s = scheduler or DDIMSchedulerOracle()
    s.set_timesteps(num_inference_steps)
    x = latents
    prev = None
    for t in reversed(s.timesteps):
        eps = unet(x, t, encoder_hidden_states=ctx, added_cond_kwargs=added, return_dict=False)[0]
        a_t = s.alphas_cumprod[int(t)]
        a_p = s.alphas_cumprod[int(prev)] if prev is not None else s.final_alpha_cumprod
        prev = t
        x = backward_ddim(x, a_t, a_p, eps)
        if trace is not None:
            trace.append(dict(t=int(t), eps=eps.clone(), x=x.clone()))
    return x
