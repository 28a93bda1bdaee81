"""Schedulers behind the diffusers interface the reference uses.

``B200DDIMScheduler`` replaces the ``DDIMScheduler`` the app swaps in (serve.py:9; pipeline.py:105,307;
ddim/pnp_pipeline.py:133): ``set_timesteps / timesteps / scale_model_input / step / init_noise_sigma /
alphas_cumprod / final_alpha_cumprod / order / config / from_config``.  ``step`` is ONE kernel launch
(x_prev = c_x x + c_e eps with host-precomputed fp64 coefficients; no device->host sync, unlike the ~8 elementwise
launches + CPU indexing of the original), and ``cfg_step`` additionally folds the classifier-free-guidance combine and
the duplication of the next UNet input into the same pass (custom_pipelines.py:332-357).
``B200EulerDiscreteScheduler`` is the SDXL pipelines' default (pipeline.py:101, refiner :128-131).
``B200DDPMScheduler`` is the prior's ancestral sampler (prior/model.py:134,585,648).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

from . import ops

_DEFAULTS = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                 prediction_type="epsilon", steps_offset=1, timestep_spacing="leading", clip_sample=False,
                 set_alpha_to_one=False)


class _SchedulerBase:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, **kw):
        cfg = dict(_DEFAULTS)
        cfg.update({k: v for k, v in kw.items() if k in _DEFAULTS})
        if cfg["beta_schedule"] != "scaled_linear" or cfg["prediction_type"] != "epsilon" or cfg["timestep_spacing"] != "leading":
            raise NotImplementedError(f"scheduler config outside the reference hot path: {cfg}")
        if cfg["clip_sample"]:
            raise NotImplementedError("clip_sample=True is not used by the reference (SDXL scheduler config)")
        self.config = SimpleNamespace(**cfg)
        self._cfg = cfg
        # fp32 table exactly as diffusers builds it (kept on CPU: indexing never touches the device)
        betas = torch.linspace(cfg["beta_start"] ** 0.5, cfg["beta_end"] ** 0.5, cfg["num_train_timesteps"], dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.one = torch.tensor(1.0)
        self.final_alpha_cumprod = torch.tensor(1.0) if cfg["set_alpha_to_one"] else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, cfg["num_train_timesteps"])[::-1].copy().astype(np.int64))

    @classmethod
    def from_config(cls, config, **kw):
        src = config if isinstance(config, dict) else {k: getattr(config, k) for k in _DEFAULTS if hasattr(config, k)}
        src = dict(src)
        src.update(kw)
        return cls(**src)

    def set_timesteps(self, num_inference_steps, device=None):
        """``device`` is accepted for interface parity; timesteps stay on the host so the loop never syncs."""
        self.num_inference_steps = int(num_inference_steps)
        ratio = self._cfg["num_train_timesteps"] // self.num_inference_steps
        ts = (np.arange(0, self.num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64) + self._cfg["steps_offset"]
        self.timesteps = torch.from_numpy(ts)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _alphas(self, timestep):
        t = int(timestep)
        prev = t - self._cfg["num_train_timesteps"] // self.num_inference_steps
        return t, prev, float(self.alphas_cumprod[t])


class B200DDIMScheduler(_SchedulerBase):
    def coefficients(self, timestep):
        """x_prev = c_x * x + c_e * eps (eta = 0; SURVEY A.5 fused form), computed in fp64 on the host."""
        t, prev, a_t = self._alphas(timestep)
        a_p = float(self.alphas_cumprod[prev]) if prev >= 0 else float(self.final_alpha_cumprod)
        c_x = math.sqrt(a_p / a_t)
        c_e = math.sqrt(1.0 - a_p) - math.sqrt(a_p * (1.0 - a_t) / a_t)
        return c_x, c_e

    def add_noise_coefficients(self, timestep):
        """``add_noise(x0, noise, t) = c_x x0 + c_e noise`` ([3P] DDIMScheduler.add_noise; img2img / inpainting start latents)."""
        a = float(self.alphas_cumprod[int(timestep)])
        return math.sqrt(a), math.sqrt(1.0 - a)

    def inverse_coefficients(self, timestep, prev_timestep):
        """Inverse DDIM step x_t = c_x x + c_e eps (``_backward_ddim``, pnp_pipeline.py:73-85, :262-275)."""
        a = float(self.alphas_cumprod[int(timestep)])
        b = float(self.alphas_cumprod[int(prev_timestep)]) if prev_timestep is not None else float(self.final_alpha_cumprod)
        return math.sqrt(a / b), math.sqrt(a) * (math.sqrt(1.0 / a - 1.0) - math.sqrt(1.0 / b - 1.0))

    def step(self, model_output, timestep, sample, eta=0.0, use_clipped_model_output=False, generator=None,
             variance_noise=None, return_dict=False):
        if eta != 0.0:
            raise NotImplementedError("eta != 0 is not used by the reference hot path (custom_pipelines.py:357 passes eta=0)")
        c_x, c_e = self.coefficients(timestep)
        prev = ops.axpby(model_output, sample, c_x, c_e)
        return SimpleNamespace(prev_sample=prev) if return_dict else (prev,)

    def cfg_step(self, eps2, timestep, sample, guidance_scale, x_in_next2=None, out=None):
        """Fused: eps_u + g (eps_c - eps_u) -> DDIM update -> (optionally) duplicated next UNet input."""
        c_x, c_e = self.coefficients(timestep)
        return ops.cfg_ddim_step(eps2, sample, guidance_scale, c_x, c_e, x_out=out, x_in_next2=x_in_next2)

    def inverse_step(self, model_output, timestep, prev_timestep, sample):
        c_x, c_e = self.inverse_coefficients(timestep, prev_timestep)
        return ops.axpby(model_output, sample, c_x, c_e)


class B200EulerDiscreteScheduler(_SchedulerBase):
    """``EulerDiscreteScheduler`` of the SDXL pipelines (base pipeline default at pipeline.py:101, refiner at :128-131; SURVEY
    8f-4): leading spacing, linear sigma interpolation, no churn.  Like DDIM it is linear in (x, eps) --
    x_next = x + (sigma_next - sigma) eps, model input x / sqrt(sigma^2 + 1) -- so ``step`` / ``cfg_step`` reuse the fused
    kernels with host-side fp64 coefficients; ``input_scale`` is applied by the sampler when it refreshes the UNet input."""

    def __init__(self, **kw):
        super().__init__(**kw)
        ac = self.alphas_cumprod.double()
        self.all_sigmas = ((1 - ac) / ac) ** 0.5
        self.sigmas = torch.cat([self.all_sigmas.flip(0), torch.zeros(1, dtype=torch.float64)]).float()
        self.timesteps = torch.arange(self._cfg["num_train_timesteps"] - 1, -1, -1, dtype=torch.float32)

    @property
    def init_noise_sigma(self):
        return float((float(self.sigmas.max()) ** 2 + 1) ** 0.5)

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = int(num_inference_steps)
        ratio = self._cfg["num_train_timesteps"] // self.num_inference_steps
        ts = (np.arange(0, self.num_inference_steps) * ratio).round()[::-1].copy().astype(np.float32) + self._cfg["steps_offset"]
        sig = np.interp(ts, np.arange(0, len(self.all_sigmas)), self.all_sigmas.float().numpy())
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)

    def _sigmas_at(self, timestep):
        i = int((self.timesteps == float(timestep)).nonzero()[0])
        return float(self.sigmas[i]), float(self.sigmas[i + 1])

    def add_noise_coefficients(self, timestep):
        """[3P] EulerDiscreteScheduler.add_noise: x0 + sigma_t noise"""
        sigma, _ = self._sigmas_at(timestep)
        return 1.0, sigma

    def input_scale(self, timestep):
        sigma, _ = self._sigmas_at(timestep)
        return 1.0 / math.sqrt(sigma * sigma + 1.0)

    def scale_model_input(self, sample, timestep=None):
        return ops.axpby(sample, sample, self.input_scale(timestep), 0.0)

    def coefficients(self, timestep):
        sigma, nxt = self._sigmas_at(timestep)
        return 1.0, nxt - sigma

    def step(self, model_output, timestep, sample, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0, generator=None,
             return_dict=False):
        if s_churn != 0.0:
            raise NotImplementedError("s_churn != 0 is not used by the reference pipelines")
        c_x, c_e = self.coefficients(timestep)
        prev = ops.axpby(model_output, sample.float(), c_x, c_e)
        return SimpleNamespace(prev_sample=prev) if return_dict else (prev,)

    def cfg_step(self, eps2, timestep, sample, guidance_scale, x_in_next2=None, out=None):
        c_x, c_e = self.coefficients(timestep)
        return ops.cfg_ddim_step(eps2, sample, guidance_scale, c_x, c_e, x_out=out, x_in_next2=x_in_next2)


class B200DDPMScheduler(_SchedulerBase):
    """DDPM ancestral step, variance_type 'fixed_small' (SURVEY A.5); used by the prior with injectable noise."""

    def coefficients(self, timestep):
        t, prev, a_t = self._alphas(timestep)
        a_p = float(self.alphas_cumprod[prev]) if prev >= 0 else 1.0
        cur_a = a_t / a_p
        cur_b = 1.0 - cur_a
        c_x0 = math.sqrt(a_p) * cur_b / (1.0 - a_t)
        c_x = math.sqrt(cur_a) * (1.0 - a_p) / (1.0 - a_t)
        sigma = math.sqrt(max((1.0 - a_p) / (1.0 - a_t) * cur_b, 1e-20)) if t > 0 else 0.0
        return dict(sqrt_a=math.sqrt(a_t), sqrt_1ma=math.sqrt(1.0 - a_t), c_x0=c_x0, c_x=c_x, sigma=sigma)

class B200LCMScheduler(B200DDIMScheduler):
    """[3P] ``LCMScheduler`` (latent consistency models): the reference's 4-step ``ipa_lcm`` mode (``serve.py:90``,
    ``sdxl_img2img_pipeline.py:91-104``: ``LCMScheduler.from_config(pipeline.scheduler.config)`` + the LCM LoRA; commented out in
    the shipped ``pipeline.py:102-103``).  Restated from the published diffusers 0.26 algorithm (parity unpinned:
    diffusers is not in this image): timesteps = every ``len / n``-th of the ``original_inference_steps`` training-schedule
    points, x0 prediction, consistency boundary scalings c_skip / c_out with sigma_data 0.5, and -- on every step but the last --
    re-noising to the next timestep.  Like DDIM the deterministic part is linear in (x, eps), so ``cfg_step`` is the same fused
    CFG + update kernel with other host-side fp64 coefficients, followed by one axpby with fresh noise (drawn on the latents'
    device from ``self.generator`` / the global RNG, like ``randn_tensor`` in the original)."""

    def __init__(self, original_inference_steps=50, timestep_scaling=10.0, **kw):
        super().__init__(**kw)
        self.config.original_inference_steps = int(original_inference_steps)
        self.config.timestep_scaling = float(timestep_scaling)
        self.generator = None

    def set_timesteps(self, num_inference_steps, device=None, original_inference_steps=None, strength=1.0):
        self.num_inference_steps = int(num_inference_steps)
        orig = int(original_inference_steps or self.config.original_inference_steps)
        k = self._cfg["num_train_timesteps"] // orig
        origin = (np.arange(1, int(orig * strength) + 1) * k - 1)[::-1].copy()
        if self.num_inference_steps > len(origin):
            raise ValueError(f"num_inference_steps {num_inference_steps} exceeds the {len(origin)} points of the LCM training schedule")
        idx = np.floor(np.linspace(0, len(origin), num=self.num_inference_steps, endpoint=False)).astype(np.int64)
        self.timesteps = torch.from_numpy(origin[idx].astype(np.int64))

    def _index(self, timestep):
        return int((self.timesteps == int(timestep)).nonzero()[0])

    def coefficients3(self, timestep):
        """x_next = c_x x + c_e eps + c_n noise (c_n = 0 on the last step), fp64 on the host."""
        i = self._index(timestep)
        last = i == len(self.timesteps) - 1
        t = int(timestep)
        a_t = float(self.alphas_cumprod[t])
        st = t * self.config.timestep_scaling
        c_skip = 0.25 / (st * st + 0.25)
        c_out = st / math.sqrt(st * st + 0.25)
        d_x = c_out / math.sqrt(a_t) + c_skip                         # denoised = c_out (x - sqrt(1 - a_t) eps) / sqrt(a_t) + c_skip x
        d_e = -c_out * math.sqrt(1.0 - a_t) / math.sqrt(a_t)
        if last:
            return d_x, d_e, 0.0
        a_p = float(self.alphas_cumprod[int(self.timesteps[i + 1])])
        return math.sqrt(a_p) * d_x, math.sqrt(a_p) * d_e, math.sqrt(1.0 - a_p)

    def coefficients(self, timestep):
        return self.coefficients3(timestep)[:2]

    def _renoise(self, x, c_n, generator):
        if c_n != 0.0:
            gen = generator if generator is not None else self.generator
            # [3P] randn_tensor: a CPU generator draws on the CPU and the noise is moved to the latents' device
            where = gen.device if gen is not None else x.device
            noise = torch.randn(x.shape, generator=gen, device=where, dtype=torch.float32).to(x.device)
            ops.axpby(noise, x, 1.0, c_n, out=x)
        return x

    def step(self, model_output, timestep, sample, generator=None, return_dict=False, **_):
        c_x, c_e, c_n = self.coefficients3(timestep)
        prev = self._renoise(ops.axpby(model_output, sample, c_x, c_e), c_n, generator)
        return SimpleNamespace(prev_sample=prev) if return_dict else (prev,)

    def cfg_step(self, eps2, timestep, sample, guidance_scale, x_in_next2=None, out=None):
        if x_in_next2 is not None:
            raise NotImplementedError("LCM re-noises after the update: the duplicated next input cannot be written by the same kernel")
        c_x, c_e, c_n = self.coefficients3(timestep)
        return self._renoise(ops.cfg_ddim_step(eps2, sample, guidance_scale, c_x, c_e, x_out=out), c_n, None)

    def inverse_coefficients(self, timestep, prev_timestep):
        raise NotImplementedError("DDIM inversion runs on B200DDIMScheduler (pnp_pipeline.py:133)")
