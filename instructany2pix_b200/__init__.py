"""B200-native (sm_100a) implementation of the InstructAny2Pix denoising hot path.

Drop-in modules behind the reference's own interfaces (see DESIGN.md / INTEGRATION.md):
``B200UNet`` (diffusers ``UNet2DConditionModel`` call surface + attention-processor plugin API),
``B200DDIMScheduler`` / ``B200DDPMScheduler``, ``B200Prior`` (``generate_diffusion``), plus the fused sampler.
All arithmetic on the hot path runs in hand-written CUDA kernels from ``libia2p_sm100a.so`` (C ABI:
``include/ia2p.h``); there is no PyTorch / CPU fallback.
"""
from ._lib import IA2PError, LIB_PATH  # noqa: F401

__all__ = ["IA2PError", "LIB_PATH"]
