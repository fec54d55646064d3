"""Checkpoint I/O (new_scripy.py:730-744): the optimizer state in torch.optim.AdamW layout, the
reference's checkpoint dict, and resuming."""
import os

import pytest
import torch

from oracle.synth import make_inputs
from tests.test_gpu_model import build

pytestmark = pytest.mark.gpu


def _steps(ddpm, opt, inp, dev, n):
    x, c, attn, ts, noise, ctx = (inp[k].to(dev) for k in ("x", "c", "attn_mask", "ts", "noise", "ctx_mask"))
    losses = []
    for _ in range(n):
        lo = ddpm(x, c, attn, randoms=(ts, noise, ctx))
        lo.backward()
        opt.step()
        opt.zero_grad()
        losses.append(float(lo))
    return losses


def test_fused_adamw_state_dict_is_torch_adamw_layout_and_resumes(dev):
    """state_dict() has torch.optim.AdamW's layout (per-parameter step / exp_avg / exp_avg_sq in the parameter's own
    NCHW shape, GEMM-native moments permuted back); loading it into a FRESH FusedAdamW continues bit-identically to the
    uninterrupted run, and torch.optim.AdamW accepts it.  Eval-mode norms: the forward is deterministic."""
    import diffusionmodel_b200 as D
    inp = make_inputs("rdd", 1, 3, 128, 5, 700, 5)

    def fresh():
        ddpm, _ = build("rdd", 64, 5, 700, 5, dev, enhance_with_attn_map=True)
        ddpm.eval()
        return ddpm, D.FusedAdamW(ddpm.parameters(), lr=1e-3, weight_decay=1e-2, max_grad_norm=1.0)
    ddpm, opt = fresh()
    assert opt.state_dict()["state"] == {}                       # nothing to save before the first step
    _steps(ddpm, opt, inp, dev, 2)
    osd = opt.state_dict()
    msd = {k: v.detach().cpu().clone() for k, v in ddpm.state_dict().items()}
    n_params = len(list(ddpm.parameters()))
    assert sorted(osd["state"]) == list(range(n_params)) and osd["param_groups"][0]["params"] == list(range(n_params))
    assert sum(opt._native) > 20
    for i, p in enumerate(ddpm.parameters()):
        st = osd["state"][i]
        assert st["exp_avg"].shape == p.shape and st["exp_avg"].is_contiguous() and float(st["step"]) == 2.0
    w = ddpm.nn_model.down1.down[0].weight                       # a GEMM-native weight: its moment comes back in NCHW order
    idx = [i for i, q in enumerate(ddpm.parameters()) if q is w][0]
    o = opt._offsets[idx]
    cout, cin, kh, kw = w.shape
    flat = opt.exp_avg[o:o + w.numel()].view(cout, kh, kw, cin).permute(0, 3, 1, 2)
    assert torch.equal(osd["state"][idx]["exp_avg"], flat.contiguous())
    # the uninterrupted run
    cont = _steps(ddpm, opt, inp, dev, 2)
    want = {k: v.detach().cpu().clone() for k, v in ddpm.state_dict().items()}
    # resume in fresh objects through a file, as the CLI does
    path = os.path.join(os.environ.get("TMPDIR", "/tmp"), "dm_ckpt_test.pt")
    torch.save({"model_state_dict": msd, "optimizer_state_dict": osd}, path)
    ck = torch.load(path, map_location="cpu")
    os.remove(path)
    ddpm2, opt2 = fresh()
    ddpm2.load_state_dict(ck["model_state_dict"])
    opt2.load_state_dict(ck["optimizer_state_dict"])
    assert opt2._step == 2 and opt2.param_groups[0]["lr"] == 1e-3
    resumed = _steps(ddpm2, opt2, inp, dev, 2)
    got = {k: v.detach().cpu() for k, v in ddpm2.state_dict().items()}
    print(f"uninterrupted {cont}  resumed {resumed}")
    assert resumed[0] == pytest.approx(cont[0], rel=1e-6)
    # the second resumed step follows an update computed from a weight gradient whose fp32 red.add order is not fixed
    assert resumed[1] == pytest.approx(cont[1], rel=2e-3)
    for k in want:
        if want[k].is_floating_point():
            assert torch.allclose(got[k], want[k], rtol=0, atol=4.1e-3), k        # <= two lr-sized Adam steps apart
    # and torch's own AdamW takes the same dict
    ref_params = [torch.nn.Parameter(p.detach().cpu().clone().contiguous()) for p in ddpm.parameters()]
    topt = torch.optim.AdamW(ref_params, lr=1e-3, weight_decay=1e-2)
    topt.load_state_dict({"state": osd["state"], "param_groups": osd["param_groups"]})
    assert float(topt.state[ref_params[idx]]["step"]) == 2.0
    assert torch.equal(topt.state[ref_params[idx]]["exp_avg"], osd["state"][idx]["exp_avg"].cpu())
    # a torch.optim.AdamW checkpoint loads the other way too
    opt3 = fresh()[1]
    opt3.load_state_dict(topt.state_dict())
    assert opt3._step == 2
    flat3 = opt3.exp_avg[o:o + w.numel()].view(cout, kh, kw, cin).permute(0, 3, 1, 2)
    assert torch.equal(flat3.contiguous().cpu(), osd["state"][idx]["exp_avg"].cpu())


def test_cli_checkpoint_has_the_reference_keys_and_resumes(dev, tmp_path, monkeypatch):
    from diffusionmodel_b200 import cli
    monkeypatch.setattr(cli.Cfg, "SAVE_DIR", str(tmp_path / "ckpt") + "/")
    monkeypatch.setattr(cli.Cfg, "N_T", 20)
    base = ["--mode", "train", "--steps_per_epoch", "4", "--n_feat", "16", "--img", "128"]
    cli.main(base + ["--epochs", "1"])
    path = os.path.join(cli.Cfg.SAVE_DIR, "best_model.pt")
    ck = torch.load(path)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "scheduler_state_dict", "loss", "metrics"}   # new_scripy.py:736-743
    assert ck["epoch"] == 0 and len(ck["optimizer_state_dict"]["state"]) > 100
    assert ck["scheduler_state_dict"]["last_epoch"] == 1
    cli.main(base + ["--epochs", "2", "--resume", path])
    ck2 = torch.load(path)
    assert ck2["epoch"] == 1 and ck2["scheduler_state_dict"]["last_epoch"] == 2
    assert float(ck2["optimizer_state_dict"]["state"][0]["step"]) == 2 * float(ck["optimizer_state_dict"]["state"][0]["step"])
