"""Deterministic synthetic weights and inputs shared by the golden generator and the tests.

TEST INFRASTRUCTURE ONLY (see oracle/ref_port.py).  Weights are filled from per-key seeded
CPU generators so a fixture only has to store (seed, shapes); every tensor -- BatchNorm
running statistics and the CoordAttn gate scalars included -- gets a non-trivial value so
eval-mode parity is not the "BN is the identity at init" easy case of SURVEY.md Appendix D.
"""
from __future__ import annotations

import zlib

import torch

from . import ref_port as P


def _gen(seed, key):
    g = torch.Generator()
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31 - 1))
    return g


def fill_state_dict_(sd, seed):
    """In-place deterministic fill of a reference-layout state dict (any variant)."""
    sched = {"alpha_t", "oneover_sqrta", "sqrt_beta_t", "alphabar_t", "sqrtab", "sqrtmab", "mab_over_sqrtmab"}
    for k in sorted(sd):
        v = sd[k]
        if k in sched or not v.is_floating_point():
            continue
        g = _gen(seed, k)
        leaf = k.rsplit(".", 1)[-1]
        with torch.no_grad():
            if leaf == "running_mean":
                v.copy_((torch.rand(v.shape, generator=g) - 0.5) * 0.2)
            elif leaf == "running_var":
                v.copy_(0.5 + torch.rand(v.shape, generator=g))
            elif leaf in ("gamma_h", "gamma_w", "alpha", "beta"):
                v.copy_(torch.rand(v.shape, generator=g) * 2 - 1)
            elif v.dim() == 1 and leaf == "weight":          # norm scale
                v.copy_(0.75 + 0.5 * torch.rand(v.shape, generator=g))
            elif v.dim() == 1:                                # biases
                v.copy_((torch.rand(v.shape, generator=g) - 0.5) * 0.2)
            else:
                if k.endswith("up0.0.weight") or ".model.0.weight" in k and v.dim() == 4 and "up" in k:
                    fan_in = v.shape[0]                       # ConvTranspose2d: [Cin, Cout, k, k]
                else:
                    fan_in = v[0].numel()
                bound = (3.0 / fan_in) ** 0.5
                v.copy_((torch.rand(v.shape, generator=g) * 2 - 1) * bound)
    return sd


def make_inputs(variant, batch, in_ch, size, n_classes, n_T, seed):
    g = torch.Generator().manual_seed(seed + 101)
    if variant == "rdd":
        x = (torch.rand(batch, in_ch, size, size, generator=g) * 2 - 1)
        ctx = torch.bernoulli(torch.full((batch,), 0.9), generator=g)
    else:
        x = torch.rand(batch, in_ch, size, size, generator=g)
        ctx = torch.bernoulli(torch.full((batch,), 0.1), generator=g)
    c = torch.randint(0, n_classes, (batch,), generator=g)
    ts = torch.randint(1, n_T + 1, (batch,), generator=g)
    noise = torch.randn(batch, in_ch, size, size, generator=g)
    attn = P.synth_attn_mask(batch, size, g)
    # exercise the exact-threshold edge cases of the mask compares (1.2 / 0.8 are NOT > themselves)
    attn[0, 0, 0] = P.HIGH_THRESH
    attn[0, 0, 1] = P.MID_THRESH
    attn[0, 0, 2] = float("nan") if False else 1.2000001
    return {"x": x, "c": c, "attn_mask": attn, "ts": ts, "noise": noise, "ctx_mask": ctx}
