"""Sample-quality metrics on the device: the reference's ``ImageMetrics`` (new_scripy.py:1111-1290).
``evaluate_batch`` mirrors the reference's method: mean SSIM and PSNR over one-to-one pairs through one kernel
(``dm_image_metrics``) instead of per-image ``.cpu().numpy()`` round trips, and FID.  FID's feature network is the
pretrained Inception-v3 that ``models.inception_v3(pretrained=True)`` downloads (:1123); those weights are not available
offline, so the extractor is pluggable (``feature_fn``: any ``[B, 3, 299, 299] in [0,1] -> [B, D]`` callable, e.g. a
torchvision Inception with ``fc = Identity`` where the weights exist) and everything around it -- batching, the batch-level
[-1,1] -> [0,1] decision, the resize, the Frechet distance -- is implemented here.  Without an extractor FID is NaN, which
is what the reference reports when its download fails (:1266-1270)."""
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


def inception_input(images: torch.Tensor) -> torch.Tensor:
    """``_extract_features`` up to the network call (:1133-1142): the batch goes to [0,1] when its minimum is negative and
    is resized to 299 x 299 (bilinear, align_corners=False)."""
    images = images.to(torch.float32)
    if images.min() < 0:
        images = (images + 1) / 2
    if images.shape[2] != 299 or images.shape[3] != 299:
        images = torch.nn.functional.interpolate(images, size=(299, 299), mode="bilinear", align_corners=False)
    return images


def frechet_distance(real_feats: torch.Tensor, gen_feats: torch.Tensor) -> float:
    """|mu_r - mu_g|^2 + Tr(S_r + S_g - 2 (S_r S_g)^(1/2)) of two [N, D] feature sets (:1168-1187), in float64 on the
    features' device.  Tr((S_r S_g)^(1/2)) is the sum of the square roots of the eigenvalues of S_r S_g (the reference takes
    the real part of scipy's sqrtm: negative round-off eigenvalues contribute nothing there either)."""
    r, g = real_feats.to(torch.float64), gen_feats.to(torch.float64)
    mu_r, mu_g = r.mean(0), g.mean(0)
    s_r, s_g = torch.cov(r.T), torch.cov(g.T)
    ev = torch.linalg.eigvals(s_r @ s_g)
    tr_sqrt = torch.sqrt(ev).real.sum()
    diff = mu_r - mu_g
    return float(diff.dot(diff) + torch.trace(s_r) + torch.trace(s_g) - 2 * tr_sqrt)


@torch.no_grad()
def calc_fid(real_images, gen_images, feature_fn, batch_size=8, device=None):
    """``ImageMetrics.calc_fid`` (:1146-1187) with a caller-supplied feature network."""
    device = device or real_images.device
    feats = []
    for imgs in (real_images, gen_images):
        part = [feature_fn(inception_input(imgs[i:i + batch_size].to(device))) for i in range(0, len(imgs), batch_size)]
        feats.append(torch.cat(part, 0))
    return frechet_distance(feats[0], feats[1])


def evaluate_batch(real_images, gen_images, feature_fn=None):
    """``ImageMetrics.evaluate_batch`` (:1253-1288): {'fid' (>= 10 samples each), 'ssim', 'psnr' (batches pair up)}."""
    metrics = {}
    if len(real_images) >= 10 and len(gen_images) >= 10:
        if feature_fn is None:
            print("FID calculation failed: no feature network (the pretrained Inception-v3 weights are not available offline)")
            metrics["fid"] = float("nan")
        else:
            metrics["fid"] = calc_fid(real_images, gen_images, feature_fn)
    if len(real_images) == len(gen_images) and len(real_images) > 0:
        sp = ssim_psnr(real_images, gen_images).double().mean(0)
        metrics["ssim"], metrics["psnr"] = float(sp[0]), float(sp[1])
    return metrics
