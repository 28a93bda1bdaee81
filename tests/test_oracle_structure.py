"""Structural known answers for the third-party restatements (the only pins the reference offers)."""
import torch

from oracle.attention import IPAttnProcessor2_0
from oracle.schedulers import DDIMSchedulerOracle, DDPMSchedulerOracle, get_timestep_embedding
from oracle.unet import SDXL_BASE, OracleUNet


def test_sdxl_param_count_and_processors():
    with torch.device("meta"):
        m = OracleUNet(SDXL_BASE)
    assert sum(p.numel() for p in m.parameters()) == 2_567_463_684
    procs = m.attn_processors
    assert len(procs) == 140
    names = list(procs)
    assert names[0].startswith("down_blocks.1") and names[-1].startswith("mid_block")   # down -> up -> mid
    assert names[0].endswith("attn1.processor") and names[1].endswith("attn2.processor")
    ip = 0
    with torch.device("meta"):
        for n in names:
            if n.endswith("attn2.processor"):
                hs = 1280 if ("mid_block" in n or "down_blocks.2" in n or "up_blocks.0" in n) else 640
                ip += sum(p.numel() for p in IPAttnProcessor2_0(hs, 2048).parameters())
    assert ip == 340_787_200
    assert m.add_embedding.linear_1.in_features == 2816 == 256 * 6 + 1280      # pnp_pipeline.py:44-47


def test_timestep_tables():
    s = DDIMSchedulerOracle(); s.set_timesteps(50)
    assert s.timesteps.tolist() == list(range(981, 0, -20))
    p = DDPMSchedulerOracle(); p.set_timesteps(25)
    assert p.timesteps.tolist() == list(range(961, 0, -40))
    p.set_timesteps(1)
    assert p.timesteps.tolist() == [1]
    assert float(s.init_noise_sigma) == 1.0 and abs(float(s.final_alpha_cumprod) - (1 - 0.00085)) < 1e-6


def test_ddim_step_is_linear_form():
    """x_prev = c_x x + c_e eps (SURVEY A.5 fused form) -- the identity the fused CUDA epilogue relies on."""
    s = DDIMSchedulerOracle(); s.set_timesteps(50)
    x, e = torch.randn(2, 4, 8, 8), torch.randn(2, 4, 8, 8)
    for t in (981, 501, 1):
        a_t = s.alphas_cumprod[t]
        prev = t - 20
        a_p = s.alphas_cumprod[prev] if prev >= 0 else s.final_alpha_cumprod
        cx = (a_p / a_t) ** 0.5
        ce = (1 - a_p) ** 0.5 - (a_p * (1 - a_t) / a_t) ** 0.5
        torch.testing.assert_close(s.step(e, t, x)[0], cx * x + ce * e, rtol=1e-5, atol=1e-5)


def test_sinusoid_layout():
    e = get_timestep_embedding(torch.tensor([3.0]), 8, flip_sin_to_cos=True, downscale_freq_shift=0)
    f = torch.exp(-torch.log(torch.tensor(10000.0)) * torch.arange(4) / 4)
    torch.testing.assert_close(e[0], torch.cat([torch.cos(3 * f), torch.sin(3 * f)]))


def test_refiner_topology_structure():
    """SDXL-refiner UNet config (pipeline.py:128-131): the oracle built from that config and the B200 holder
    modules agree key for key; 2 259 526 660 parameters (the published model card says "2.3 B")."""
    from instructany2pix_b200.unet import REFINER_CONFIG, B200UNet
    from oracle.unet import OracleUNet, UNetConfig
    with torch.device("meta"):
        o = OracleUNet(UNetConfig(**REFINER_CONFIG))
    b = B200UNet(device="meta", **REFINER_CONFIG)
    so, sb = {k: tuple(v.shape) for k, v in o.state_dict().items()}, {k: tuple(v.shape) for k, v in b.state_dict().items()}
    assert so == sb
    n = sum(torch.Size(s).numel() for s in so.values())
    print("refiner parameters:", n)
    assert n == 2_259_526_660
    assert len(b.attn_processors) == 2 * (4 * (2 + 2) + 4 * (3 + 3) + 4)      # 4 layers x (down 2+2, up 3+3, mid 1) blocks x 2


def test_euler_schedule_known_answers():
    """SDXL noise schedule as the Euler scheduler sees it: sigma_max 14.6146, sigma_min 0.0292 (the published constants of the
    SDXL/SD k-diffusion schedule), leading timesteps 961..1 for N = 25, sigmas decreasing to the appended 0, and the B200
    scheduler's host-side tables equal to the oracle's."""
    from instructany2pix_b200.scheduler import B200EulerDiscreteScheduler
    from oracle.schedulers import EulerDiscreteSchedulerOracle
    o = EulerDiscreteSchedulerOracle()
    assert abs(float(o.all_sigmas.max()) - 14.6146) < 1e-3 and abs(float(o.all_sigmas.min()) - 0.0292) < 1e-4
    o.set_timesteps(25)
    assert o.timesteps.tolist() == [float(t) for t in range(961, 0, -40)]
    assert float(o.sigmas[-1]) == 0.0 and bool((o.sigmas[:-1] > o.sigmas[1:]).all())
    assert abs(o.init_noise_sigma - (float(o.sigmas.max()) ** 2 + 1) ** 0.5) < 1e-6
    b = B200EulerDiscreteScheduler()
    b.set_timesteps(25)
    assert torch.equal(b.timesteps, o.timesteps) and torch.allclose(b.sigmas, o.sigmas, rtol=1e-6, atol=0)
    c_x, c_e = b.coefficients(b.timesteps[3])
    assert c_x == 1.0 and abs(c_e - float(o.sigmas[4] - o.sigmas[3])) < 1e-6
    assert abs(b.input_scale(b.timesteps[3]) - 1.0 / (float(o.sigmas[3]) ** 2 + 1) ** 0.5) < 1e-6
