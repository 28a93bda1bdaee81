"""Oracle (test infrastructure): restated diffusers==0.26.3 schedulers + sinusoid.

Third-party algorithm restated from its published semantics; the reference
pins the version at requirements.txt:3 and calls it at
  * DDIMScheduler: pipeline.py:105,307; ddim/pnp_pipeline.py:133,192,262-267;
    diffusion/ip_adapter/custom_pipelines.py:250,334,357
  * DDPMScheduler: prior/model.py:134,585,648
  * get_timestep_embedding: prior/model.py:565-568,613-614
Scheduler config = the SDXL-base hub ``scheduler_config.json``.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

SDXL_SCHEDULER_CONFIG = dict(
    num_train_timesteps=1000,
    beta_start=0.00085,
    beta_end=0.012,
    beta_schedule="scaled_linear",
    prediction_type="epsilon",
    steps_offset=1,
    timestep_spacing="leading",
    clip_sample=False,
    set_alpha_to_one=False,
)


def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False,
                           downscale_freq_shift=1, scale=1, max_period=10000):
    """diffusers.models.embeddings.get_timestep_embedding."""
    assert timesteps.ndim == 1
    half = embedding_dim // 2
    exponent = -math.log(max_period) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if embedding_dim % 2 == 1:
        emb = torch.nn.functional.pad(emb, (0, 1, 0, 0))
    return emb


def _alphas_cumprod(cfg):
    assert cfg["beta_schedule"] == "scaled_linear"
    betas = torch.linspace(cfg["beta_start"] ** 0.5, cfg["beta_end"] ** 0.5,
                           cfg["num_train_timesteps"], dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def _leading_timesteps(cfg, n):
    ratio = cfg["num_train_timesteps"] // n
    ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64)
    return ts + cfg["steps_offset"]


class _Base:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, **overrides):
        cfg = dict(SDXL_SCHEDULER_CONFIG)
        cfg.update(overrides)
        self.config = SimpleNamespace(**cfg)
        self._cfg = cfg
        self.alphas_cumprod = _alphas_cumprod(cfg)
        self.one = torch.tensor(1.0)
        self.final_alpha_cumprod = torch.tensor(1.0) if cfg["set_alpha_to_one"] else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, cfg["num_train_timesteps"])[::-1].copy().astype(np.int64))

    @classmethod
    def from_config(cls, config):
        if not isinstance(config, dict):
            config = {k: v for k, v in vars(config).items() if k in SDXL_SCHEDULER_CONFIG}
        return cls(**{k: v for k, v in config.items() if k in SDXL_SCHEDULER_CONFIG})

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        self.timesteps = torch.from_numpy(_leading_timesteps(self._cfg, num_inference_steps)).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample


class DDIMSchedulerOracle(_Base):
    """DDIMScheduler.step, eta / variance-noise supported.  Pinned (eta = 0) as the exact inverse of the
    reference's own ``_backward_ddim`` (ddim/pnp_pipeline.py:73-85) on its golden outputs:
    tests/test_oracle_golden.py::test_ddim_step_inverts_reference_backward_ddim."""

    def step(self, model_output, timestep, sample, eta=0.0, generator=None,
             variance_noise=None, return_dict=False):
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        variance = ((1 - a_p) / (1 - a_t)) * (1 - a_t / a_p)
        std = eta * variance ** 0.5
        direction = (1 - a_p - std ** 2) ** 0.5 * model_output
        prev = a_p ** 0.5 * x0 + direction
        if eta > 0:
            if variance_noise is None:
                variance_noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            prev = prev + std * variance_noise
        return (prev,)


def _ddim_add_noise(self, original, noise, timestep):
    """[3P] DDIMScheduler.add_noise: sqrt(a_t) x0 + sqrt(1 - a_t) noise (img2img / inpainting start and re-noising)."""
    a = self.alphas_cumprod[int(timestep)]
    return a ** 0.5 * original + (1 - a) ** 0.5 * noise


DDIMSchedulerOracle.add_noise = _ddim_add_noise


class DDPMSchedulerOracle(_Base):
    """DDPMScheduler.step with variance_type="fixed_small" (SURVEY.md A.5)."""

    def step(self, model_output, timestep, sample, generator=None, return_dict=False):
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t = 1 - a_t
        b_p = 1 - a_p
        cur_a = a_t / a_p
        cur_b = 1 - cur_a
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        c0 = (a_p ** 0.5 * cur_b) / b_t
        cx = cur_a ** 0.5 * b_p / b_t
        prev = c0 * x0 + cx * sample
        if t > 0:
            # randn_tensor(shape, generator=generator, device, dtype): global RNG when generator is None
            noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            var = torch.clamp((1 - a_p) / (1 - a_t) * cur_b, min=1e-20)
            prev = prev + (var ** 0.5) * noise
        return (prev,)


def backward_ddim(x_tm1, alpha_t, alpha_tm1, eps_xt):
    """Restatement of ``_backward_ddim`` (ddim/pnp_pipeline.py:73-85)."""
    a, b = alpha_t, alpha_tm1
    sa = a ** 0.5
    sb = b ** 0.5
    return sa * ((1 / sb) * x_tm1 + ((1 / a - 1) ** 0.5 - (1 / b - 1) ** 0.5) * eps_xt)


def polar_interpolate(x, y, alpha):
    """Restatement of ``polar_intrtpolate`` (pipeline.py:295-300)."""
    n0 = x.norm()
    n1 = y.norm()
    ll = x * alpha + y * (1 - alpha)
    n = n0 * alpha + n1 * (1 - alpha)
    return ll / ll.norm() * n


class EulerDiscreteSchedulerOracle:
    """Restated diffusers==0.26.3 ``EulerDiscreteScheduler`` with the SDXL scheduler config (the base pipeline's default at
    pipeline.py:101 before serve.py:9 swaps in DDIM, and the refiner's scheduler at pipeline.py:128-131 / :358-361):
    timestep_spacing "leading" + steps_offset 1, interpolation_type "linear", use_karras_sigmas False, s_churn 0 (the pipelines
    pass no churn, so ``step`` is deterministic).  Parity unpinned (third-party); known answers pinned in
    tests/test_oracle_structure.py: sigma_max = 14.6146, sigma_min = 0.0292 of the SDXL noise schedule."""
    order = 1

    def __init__(self, **kw):
        cfg = dict(SDXL_SCHEDULER_CONFIG)
        cfg.update(kw)
        self.config = SimpleNamespace(**cfg)
        self._cfg = cfg
        self.alphas_cumprod = _alphas_cumprod(cfg)
        self.all_sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
        self.sigmas = torch.cat([self.all_sigmas.flip(0), torch.zeros(1)])
        self.timesteps = torch.arange(cfg["num_train_timesteps"] - 1, -1, -1, dtype=torch.float32)
        self._step_index = None

    @property
    def init_noise_sigma(self):
        return float((self.sigmas.max() ** 2 + 1) ** 0.5)          # "leading" spacing

    def set_timesteps(self, n, device=None):
        ts = _leading_timesteps(self._cfg, n).astype(np.float32)
        sig = self.all_sigmas.numpy()
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self._step_index = None

    def _index(self, t):
        if self._step_index is None:
            self._step_index = int((self.timesteps == float(t)).nonzero()[0])
        return self._step_index

    def add_noise(self, original, noise, timestep):
        """[3P] EulerDiscreteScheduler.add_noise: x0 + sigma_t noise"""
        i = int((self.timesteps == float(timestep)).nonzero()[0])
        return original + self.sigmas[i] * noise

    def scale_model_input(self, sample, t):
        sigma = self.sigmas[self._index(t)]
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output, t, sample, return_dict=False, **_):
        i = self._index(t)
        sigma = self.sigmas[i]
        sample = sample.float()
        pred_original = sample - sigma * model_output.float()
        derivative = (sample - pred_original) / sigma
        prev = sample + derivative * (self.sigmas[i + 1] - sigma)
        self._step_index = i + 1
        return (prev,)

class LCMSchedulerOracle(_Base):
    """[3P] diffusers ``LCMScheduler`` (0.26: ``set_timesteps`` by ``np.linspace`` over the reversed training-schedule points,
    ``get_scalings_for_boundary_condition_discrete`` with sigma_data 0.5 and timestep_scaling 10, multistep re-noising),
    epsilon prediction, no clipping / thresholding -- the scheduler of the reference's commented-out ``ipa_lcm`` mode
    (sdxl_img2img_pipeline.py:91-104, serve.py:90).  Restated from the published algorithm: PARITY UNPINNED (no diffusers here,
    and the reference never instantiates it)."""

    def __init__(self, original_inference_steps=50, timestep_scaling=10.0, **kw):
        super().__init__(**kw)
        self.original_inference_steps, self.timestep_scaling, self.generator = original_inference_steps, timestep_scaling, None

    def set_timesteps(self, num_inference_steps, device=None, original_inference_steps=None, strength=1.0):
        self.num_inference_steps = num_inference_steps
        orig = original_inference_steps or self.original_inference_steps
        k = self._cfg["num_train_timesteps"] // orig
        origin = np.asarray(list(range(1, int(orig * strength) + 1))) * k - 1
        origin = origin[::-1].copy()
        idx = np.floor(np.linspace(0, len(origin), num=num_inference_steps, endpoint=False)).astype(np.int64)
        self.timesteps = torch.from_numpy(origin[idx]).to(dtype=torch.long)

    def step(self, model_output, timestep, sample, generator=None, return_dict=False, **_):
        i = int((self.timesteps == int(timestep)).nonzero()[0])
        prev_t = self.timesteps[i + 1] if i + 1 < len(self.timesteps) else timestep
        a_t = self.alphas_cumprod[int(timestep)]
        a_p = self.alphas_cumprod[int(prev_t)] if int(prev_t) >= 0 else self.final_alpha_cumprod
        scaled = int(timestep) * self.timestep_scaling
        c_skip = 0.5 ** 2 / (scaled ** 2 + 0.5 ** 2)
        c_out = scaled / (scaled ** 2 + 0.5 ** 2) ** 0.5
        x0 = (sample - (1 - a_t) ** 0.5 * model_output) / a_t ** 0.5
        denoised = c_out * x0 + c_skip * sample
        if i != self.num_inference_steps - 1:
            noise = torch.randn(model_output.shape, generator=generator if generator is not None else self.generator, dtype=denoised.dtype)
            prev = a_p ** 0.5 * denoised + (1 - a_p) ** 0.5 * noise
        else:
            prev = denoised
        return (prev,)


LCMSchedulerOracle.add_noise = _ddim_add_noise
