"""CPU: the oracle (oracle/ref_port.py) against the committed golden fixtures, which
oracle/make_golden.py generated from the imported reference (asserting bit-equality there).  On the
same torch build the oracle reproduces the stored reference outputs bit-exactly; across builds a
1e-5 rel-L2 slack covers library-level summation-order changes."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_port as P
from oracle.synth import fill_state_dict_, make_inputs

GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-5


def _ref_shapes(variant, n_feat, n_classes, n_T):
    """State-dict layout comes from the product modules (same names/shapes as the reference)."""
    import diffusionmodel_b200 as D
    net = D.ContextUnet(3, n_feat, n_classes) if variant == "rdd" else D.MnistContextUnet(1, n_feat, n_classes)
    return {k: v.clone() for k, v in D.DDPM(net, (1e-4, 0.02), n_T, "cpu").state_dict().items()}


def test_schedule_known_answers():
    g = np.load(os.path.join(GOLD, "schedules_T700.npz"))
    s = P.ddpm_schedules(1e-4, 0.02, 700)
    for k in s:
        assert np.array_equal(s[k].numpy(), g[k]), k
    # SURVEY.md 8(c) known answers
    assert abs(float(s["alpha_t"][0]) - 0.999899983) < 1e-8
    assert abs(float(s["sqrt_beta_t"][0]) - 0.01) < 1e-8
    assert abs(float(s["oneover_sqrta"][700]) - 1.010152459) < 1e-7
    assert abs(float(s["alphabar_t"][700]) - 0.000831294) < 1e-8
    assert abs(float(s["sqrtab"][350]) - 0.409153134) < 1e-7
    assert abs(float(s["mab_over_sqrtmab"][1]) - 0.008498023) < 1e-8
    s4 = P.ddpm_schedules(1e-4, 0.02, 400)
    assert abs(float(s4["alphabar_t"][400]) - 0.017296989) < 1e-8
    assert abs(float(s4["sqrtab"][200]) - 0.599440336) < 1e-7
    import diffusionmodel_b200 as D
    mine = D.ddpm_schedules(1e-4, 0.02, 700)
    for k in s:
        assert torch.equal(mine[k], s[k]), k


@pytest.mark.parametrize("tag", ["mnist_f16_b8", "rdd_f16_s128_b2", "rdd_f32_s128_b1_nomap"])
def test_oracle_reproduces_golden(tag):
    g = np.load(os.path.join(GOLD, tag + ".npz"))
    n_feat, size, batch, n_classes, seed, n_T, steps, use_map = (int(v) for v in g["meta"])
    variant = "mnist" if tag.startswith("mnist") else "rdd"
    in_ch = 1 if variant == "mnist" else 3
    sd0 = _ref_shapes(variant, n_feat, n_classes, n_T)
    fill_state_dict_(sd0, seed)
    inp = make_inputs(variant, batch, in_ch, size, n_classes, n_T, seed)
    sched = P.ddpm_schedules(1e-4, 0.02, n_T)
    amap = inp["attn_mask"] if (variant == "rdd" and use_map) else None
    for mode in ("eval", "train"):
        sd = {k: v.clone() for k, v in sd0.items()}
        for k, v in sd.items():
            if v.is_floating_point() and k.startswith("nn_model.") and "running" not in k:
                v.requires_grad_(True)
        loss = P.ddpm_loss(sd, sched, inp["x"], inp["c"], inp["attn_mask"], inp["ts"], inp["noise"], inp["ctx_mask"],
                           variant=variant, n_T=n_T, training=(mode == "train"), attn_map=amap)
        loss.backward()
        assert abs(float(loss) - float(g[f"loss_{mode}"])) <= TOL * abs(float(g[f"loss_{mode}"]))
        names = [str(s) for s in g["grad_names"]]
        gn = np.array([float(sd[k].grad.double().norm()) for k in names])
        ref = g[f"gradnorm_{mode}"]
        assert np.all(np.abs(gn - ref) <= 1e-4 * np.abs(ref) + 1e-9)
        x_t = P.q_sample(sched, inp["x"], inp["ts"], inp["noise"])
        with torch.no_grad():
            pred = P.unet_forward({k: v.detach().clone() for k, v in sd0.items()}, x_t, inp["c"], inp["ts"] / n_T,
                                  inp["ctx_mask"], variant=variant, training=(mode == "train"), prefix="nn_model.",
                                  attn_map=amap)
        assert P.rel_l2(pred, torch.from_numpy(g[f"pred_{mode}"])) <= TOL
        if mode == "train":
            bn = [str(s) for s in g["bn_names"]]
            got = torch.cat([sd[k].detach().flatten() for k in bn])
            assert P.rel_l2(got, torch.from_numpy(g["bn_after_train"])) <= TOL


def test_mask_threshold_edges():
    """(mask > 1.2) / (mask > 0.8) index sets: the thresholds themselves are NOT included, NaN is low."""
    up = float(torch.nextafter(torch.tensor(1.2), torch.tensor(2.0)))
    m = torch.tensor([[[1.2, up, 0.8, 0.80000001, float("nan"), 3.0, 0.5, float("inf")]]])
    noise = torch.zeros(1, 3, 1, 8)
    pred = torch.ones(1, 3, 1, 8)
    w = torch.tensor([1.0, 3.0, 0.5, 0.5, 0.5, 3.0, 0.5, 3.0])      # 1.2 is mid, 0.8 and fp32(0.80000001)==0.8 low, NaN low
    h = torch.tensor([0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 1.0])
    want = (w.mean() * 1.0) + 2.0 * h.mean()
    assert abs(float(P.weighted_loss(noise, pred, m)) - float(want)) < 1e-6
