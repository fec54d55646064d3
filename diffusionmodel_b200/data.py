"""Device-side input pipeline for the training loop.

The reference's ``CrackDataset`` (new_scripy.py:479-551) re-opens a JPEG and re-parses a VOC XML file for every sample of
every epoch and pushes the result through PIL transforms in five worker processes (:53,683-688); at ~180 img/s per
GPU that loader is the bottleneck.  Here the decode + resize happen ONCE (``from_voc_dir``, same PIL calls as the
reference), the uint8 images, labels and boxes are cached on the device, and every batch is produced by one kernel
(``dm_prep_batch``): random horizontal flip, ToTensor, Normalize and the attention-mask rasterisation.

Reference quirks kept: the mask is NOT flipped with the image (:535-549 build it before the transform runs); box
coordinates are scaled with Python ``round`` (banker's rounding) and clamped to [0, IMG_SIZE-1]; the box is the
half-open slice ``[ymin:ymax, xmin:xmax]``.
"""
from __future__ import annotations

import ctypes
import os

import torch

from . import ops

LOW_WEIGHT, MID_WEIGHT, HIGH_WEIGHT = 0.5, 1.0, 3.0          # new_scripy.py:33-35
NORM_MEAN, NORM_STD = 0.5, 0.5                                # new_scripy.py:66-67 (same for the three channels)


def scale_box(xmin, ymin, xmax, ymax, orig_w, orig_h, img_size):
    """new_scripy.py:542-545, host integers (Python round = round-half-to-even, like the reference)."""
    cl = lambda v: max(0, min(img_size - 1, v))
    return (cl(round(xmin * img_size / orig_w)), cl(round(ymin * img_size / orig_h)),
            cl(round(xmax * img_size / orig_w)), cl(round(ymax * img_size / orig_h)))


class CachedCrackBatches:
    """uint8 image cache on the device + one-kernel batch preparation.

    ``images_u8`` [M, H, W, 3] uint8 (decoded, resized to the training resolution), ``labels`` [M] int64,
    ``boxes`` [M, 4] int32 scaled boxes (xmin, ymin, xmax, ymax: ``scale_box``)."""

    def __init__(self, images_u8, labels, boxes, device, classes=None):
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[3] != 3:
            raise ValueError("images_u8 must be uint8 [M, H, W, 3]")
        self.images = images_u8.to(device).contiguous()
        self.labels = labels.to(device=device, dtype=torch.int64)
        self.boxes = boxes.to(device=device, dtype=torch.int32).contiguous()
        self.device = torch.device(device)
        self.classes = classes
        self.p_flip = 0.5                                       # transforms.RandomHorizontalFlip(0.5), :685

    def __len__(self):
        return self.images.shape[0]

    def batch(self, indices, flips=None, generator=None):
        """(x fp32 [B,3,H,W] in [-1,1], c int64 [B], attn_mask fp32 [B,H,W]) for the given sample indices.
        ``flips``: optional 0/1 tensor [B]; default: ``torch.rand(B) < 0.5`` from ``generator`` (host RNG, as torchvision)."""
        idx = torch.as_tensor(indices, dtype=torch.int64)
        b = idx.numel()
        if flips is None:
            flips = torch.rand(b, generator=generator) < self.p_flip
        flips = torch.as_tensor(flips).to(device=self.device, dtype=torch.int32)
        idx = idx.to(self.device)
        img = self.images.index_select(0, idx)                  # [B,H,W,3] uint8 gather (the only other launch)
        box = self.boxes.index_select(0, idx)
        _, h, w, _ = img.shape
        x = torch.empty((b, 3, h, w), device=self.device, dtype=torch.float32)
        mask = torch.empty((b, h, w), device=self.device, dtype=torch.float32)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        ops.call("dm_prep_batch", p(img), p(flips), p(box), p(x), p(mask), b, h, w, NORM_MEAN, NORM_STD,
                 LOW_WEIGHT, MID_WEIGHT, HIGH_WEIGHT, ops._stream())
        return x, self.labels.index_select(0, idx), mask

    @classmethod
    def from_voc_dir(cls, root_dir, img_size, device):
        """Build the cache from the reference's directory layout (``images/<class>/*.jpg`` + ``annotations/*.xml``,
        new_scripy.py:496-511) with the reference's own decode + ``transforms.Resize`` (PIL bilinear), once."""
        import xml.etree.ElementTree as ET

        import numpy as np
        from PIL import Image
        classes = sorted(d for d in os.listdir(os.path.join(root_dir, "images"))
                         if os.path.isdir(os.path.join(root_dir, "images", d)))
        imgs, labels, boxes = [], [], []
        for ci, cname in enumerate(classes):
            cdir = os.path.join(root_dir, "images", cname)
            for name in os.listdir(cdir):
                if not name.endswith((".png", ".jpg", ".jpeg")):
                    continue
                xml = os.path.join(root_dir, "annotations", name.rsplit(".", 1)[0] + ".xml")
                if not os.path.exists(xml):
                    continue
                root = ET.parse(xml).getroot()
                bb = root.find(".//bndbox")
                vals = [int(bb.find(k).text) for k in ("xmin", "ymin", "xmax", "ymax")]
                ow, oh = int(root.find(".//width").text), int(root.find(".//height").text)
                im = Image.open(os.path.join(cdir, name)).convert("RGB").resize((img_size, img_size), Image.BILINEAR)
                imgs.append(torch.from_numpy(np.asarray(im, dtype=np.uint8).copy()))
                labels.append(ci)
                boxes.append(scale_box(*vals, ow, oh, img_size))
        if not imgs:
            raise FileNotFoundError(f"no annotated images under {root_dir}")
        return cls(torch.stack(imgs), torch.tensor(labels), torch.tensor(boxes, dtype=torch.int32), device, classes)
