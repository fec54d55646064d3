"""diffusionmodel_b200: the DiffusionModel hot path (ContextUnet denoiser fwd/bwd inside the DDPM train
step, CFG reverse sampling) on hand-written sm_100a kernels, behind the reference's module API.

    from diffusionmodel_b200 import ContextUnet, MnistContextUnet, DDPM, FusedAdamW

The CUDA library (libdm_b200.so, C-ABI in include/dm_b200.h) is loaded lazily on the first compute
call; it is required -- there is no CPU / PyTorch fallback for the hot path.
"""
from .data import CachedCrackBatches         # noqa: F401
from .ddpm import DDPM, ddpm_schedules          # noqa: F401
from .optim import FusedAdamW                   # noqa: F401
from .unet import ContextUnet, MnistContextUnet  # noqa: F401

__all__ = ["ContextUnet", "MnistContextUnet", "DDPM", "ddpm_schedules", "FusedAdamW", "CachedCrackBatches"]
