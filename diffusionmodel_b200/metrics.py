"""Sample-quality metrics on the device: the SSIM / PSNR half of the reference's ``ImageMetrics``
(new_scripy.py:1111-1290).  ``evaluate_batch`` mirrors the reference's method (mean SSIM and PSNR over one-to-one
pairs) without its per-image ``.cpu().numpy()`` round trips; FID needs pretrained Inception-v3 weights (downloaded by
``models.inception_v3(pretrained=True)``, :1123) and is not provided."""
from __future__ import annotations

import ctypes

import torch

from . import ops


def ssim_psnr(real_images: torch.Tensor, gen_images: torch.Tensor) -> torch.Tensor:
    """[N, 2] = (ssim, psnr) per pair for two CUDA fp32 tensors [N, C, H, W] (calc_ssim / calc_psnr, :1189-1251)."""
    if real_images.shape != gen_images.shape:
        raise ValueError("one-to-one comparison needs equal shapes")
    if real_images.device.type != "cuda":
        raise ops._lib.DmB200Error("ssim_psnr runs on CUDA only; there is no CPU fallback")
    a = real_images.detach().to(torch.float32).contiguous()
    b = gen_images.detach().to(torch.float32).contiguous()
    n = a.shape[0]
    out = torch.empty((n, 2), device=a.device, dtype=torch.float32)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    ops.call("dm_image_metrics", p(a), p(b), p(out), n, a[0].numel() if n else 0, ops._stream())
    return out


def evaluate_batch(real_images, gen_images):
    """``ImageMetrics.evaluate_batch`` (:1253-1288) minus FID: {'ssim': mean, 'psnr': mean} when the batches pair up."""
    metrics = {}
    if len(real_images) == len(gen_images) and len(real_images) > 0:
        sp = ssim_psnr(real_images, gen_images).double().mean(0)
        metrics["ssim"], metrics["psnr"] = float(sp[0]), float(sp[1])
    return metrics
