"""Parity at the configurations that are benchmarked (GPU): the kernel selections only the full sizes reach.

  * cfg2 (BASELINE.json configs[1]): new_scripy ContextUnet, n_feat 192, 3x256x256, batch 4 -- train-mode loss, per-block
    activations and per-block gradients against the KERNEL-MATCHED oracle (ref_port with operand_dtype = store_dtype =
    bf16: GEMM operands rounded to bf16, fp32 accumulate, and a rounding wherever the kernels store bf16 between
    launches; everything else fp32), and the eval-mode forward against the fp32 oracle at the north-star bar (1e-2);
  * cfg1 (configs[0]): MNIST ContextUnet, n_feat 128, batch 128 -- loss and gradients;
  * cfg5 (configs[4], n_feat 384, 1.41 B parameters): the eval-mode step per block against the oracle executed on the GPU;
  * one convolution per distinct Appendix-A shape class at its true size, each asserting which kernel the dispatcher
    launched (dm_kernel_count / dm_last_kernel);
  * a full 60-step CFG trajectory with injected noise against the fp32 oracle, with the drift curve printed.

The oracle runs on the host cores (about 2.5 s per 256x256 image forward+backward at n_feat 192 on the GPU box).
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import ref_port as P
from oracle.synth import fill_state_dict_, make_inputs
from tests.test_gpu_kernels import BF16_TOL, F32_TOL, bf, nchw, nhwc, rel
from tests.test_gpu_model import build, grads_of

pytestmark = pytest.mark.gpu

RDD_BLOCKS = ["init_conv", "down1", "down2", "down3", "down4", "ca1", "ca2", "ca3", "ca4", "time_emb1", "time_emb2",
              "ctx_emb1", "ctx_emb2", "up0", "up1", "up2", "up3", "up4", "local_enhance", "out"]


def _hook_blocks(net, names, c_of):
    """Forward hooks on the named children: block name -> NCHW fp32 CPU copy of its (bf16 NHWC) output."""
    got, handles = {}, []
    for name in names:
        mod = getattr(net, name)
        handles.append(mod.register_forward_hook(
            lambda m, i, o, name=name: got.__setitem__(name, nchw(o.detach(), c_of[name]))))
    return got, handles


def _block_grad_errors(mine, ref, blocks, prefix="nn_model."):
    out = {}
    for b in blocks:
        num = den = 0.0
        for k, g in mine.items():
            if k.startswith(prefix + b + ".") and k in ref:
                num += float((g.double() - ref[k].double()).pow(2).sum())
                den += float(ref[k].double().pow(2).sum())
        if den > 0:
            out[b] = (num / den) ** 0.5
    return out


def _oracle_grads(sd, sched, inp, variant, n_T, training, use_map, **kw):
    sd_o = {k: v.clone() for k, v in sd.items()}
    for k, v in sd_o.items():
        if v.is_floating_point() and k.startswith("nn_model.") and "running" not in k:
            v.requires_grad_(True)
    tap = {}
    x_t = P.q_sample(sched, inp["x"], inp["ts"], inp["noise"])
    pred = P.unet_forward(sd_o, x_t, inp["c"], inp["ts"] / n_T, inp["ctx_mask"], variant=variant, training=training,
                          attn_map=inp["attn_mask"] if use_map else None, prefix="nn_model.", tap=tap, **kw)
    loss = P.weighted_loss(inp["noise"], pred, inp["attn_mask"]) if variant == "rdd" else F.mse_loss(inp["noise"], pred)
    loss.backward()
    grads = {k: v.grad for k, v in sd_o.items() if v.requires_grad and v.grad is not None}
    return float(loss), pred.detach(), {k: v.detach() for k, v in tap.items()}, grads, sd_o


def _cfg2(dev, training, seed, n_feat=192, batch=4):
    """One cfg2 micro-step (F=192, 3x256x256, B=4, LocalEnhancer fed the attention map) on the GPU with block hooks."""
    from diffusionmodel_b200 import _lib
    size, n_classes, n_T = 256, 5, 700
    ddpm, sd = build("rdd", n_feat, n_classes, n_T, seed, dev, enhance_with_attn_map=True)
    ddpm.train(training)
    inp = make_inputs("rdd", batch, 3, size, n_classes, n_T, seed)
    x, c, attn, ts, noise, ctx = (inp[k].to(dev) for k in ("x", "c", "attn_mask", "ts", "noise", "ctx_mask"))
    f = n_feat
    c_of = {"init_conv": f, "ca1": f, "ca2": 2 * f, "ca3": 4 * f, "ca4": 8 * f, "up1": 4 * f, "up2": 2 * f, "up3": f,
            "up4": f, "local_enhance": f}
    got, handles = _hook_blocks(ddpm.nn_model, list(c_of), c_of)
    with _lib.kernel_counts() as k:
        loss = ddpm(x, c, attn, randoms=(ts, noise, ctx))
        loss.backward()
        torch.cuda.synchronize()
    for h in handles:
        h.remove()
    # the kernel selections the benchmark figure runs on
    print(f"kernel launches in one micro-step (n_feat {n_feat}, batch {batch}):", k.delta)
    if n_feat == 192:
        assert k["conv3x3_halo2"] > 0 and k["conv3x3_halo"] > 0 and k["conv_gemm"] > 0 and k["conv_gemm2"] > 0
        assert k["wgrad_gemm"] > 0 and k["wgrad2_gemm"] > 0 and k["wgrad3_pair"] > 0 and k["skinny_gemm"] == 1
    tap_name = {"ca1": "down1", "ca2": "down2", "ca3": "down3", "ca4": "down4"}
    return ddpm, sd, inp, float(loss.detach()), {tap_name.get(n_, n_): v for n_, v in got.items()}, grads_of(ddpm), n_T


def _fmt(d):
    return {k: f"{v:.2e}" for k, v in d.items()}


def test_cfg2_eval_mode_step_per_block_vs_fp32_oracle(dev):
    """The benchmarked shapes with running-statistics BatchNorm (no batch-statistic amplification): loss and every block's
    output against the FP32 oracle at the north-star bf16 bar (1e-2); every block's parameter gradients at 1e-2 or, where
    bf16 itself does not allow that after ~100 stored tensors of forward + backward, no farther from fp32 than 1.25x the
    kernel-matched oracle (the same algorithm rounding to bf16 where the kernels do) -- and never beyond 2e-2.  Every
    conv / data-gradient / weight-gradient kernel selection of the benchmark runs here at its true size."""
    ddpm, sd, inp, loss, got, mine, n_T = _cfg2(dev, False, 11)
    sched = P.ddpm_schedules(1e-4, 0.02, n_T)
    l_o, _, tap, g_o, _ = _oracle_grads(sd, sched, inp, "rdd", n_T, False, True)
    _, _, _, g_m, _ = _oracle_grads(sd, sched, inp, "rdd", n_T, False, True, operand_dtype=torch.bfloat16,
                                    store_dtype=torch.bfloat16)
    errs = {n_: P.rel_l2(v, tap[n_]) for n_, v in got.items()}
    gerrs = _block_grad_errors(mine, g_o, RDD_BLOCKS)
    gmat = _block_grad_errors(g_m, g_o, RDD_BLOCKS)
    print(f"cfg2 eval-mode loss: ours {loss:.6f}  fp32 oracle {l_o:.6f}")
    print("cfg2 eval-mode per-block output rel-L2 vs fp32 oracle:", _fmt(errs))
    print("cfg2 eval-mode per-block gradient rel-L2 vs fp32 oracle, ours:          ", _fmt(gerrs))
    print("cfg2 eval-mode per-block gradient rel-L2 vs fp32 oracle, kernel-matched:", _fmt(gmat))
    assert abs(loss - l_o) < 1e-2 * abs(l_o)
    assert max(errs.values()) < 1e-2, errs
    for n_, e in gerrs.items():
        assert e < max(1e-2, 1.25 * gmat[n_]) and e < 2e-2, (n_, e, gmat[n_])


def test_cfg2_train_step_vs_kernel_matched_oracle(dev):
    """The benchmarked micro-step itself: train-mode BatchNorm.

    bf16 rounding decorrelates two correct implementations: a relative difference d << 2^-8 in front of a rounding comes
    out as about sqrt(d * 2^-8), so after a handful of stored tensors ANY two bf16 implementations differ by independent
    rounding noise, which batch-statistics BatchNorm then amplifies with depth (PyTorch's own
    autocast(bf16) run of the reference is 1.2e-1 from fp32 in train mode).  So:
      * the first block (two conv-BN-GELU units + SE, 256x256x192) is compared TIGHTLY with the kernel-matched oracle
        (rounds where the kernels store bf16): 1e-3;
      * the loss within 1e-3 of that oracle;
      * at depth, our distance to the fp32 oracle must not exceed the kernel-matched oracle's own distance to fp32 by more
        than 1.5x, block by block, for activations and for gradients (we are indistinguishable from a correct bf16 run)."""
    ddpm, sd, inp, loss, got, mine, n_T = _cfg2(dev, True, 11)
    sched = P.ddpm_schedules(1e-4, 0.02, n_T)
    l_m, _, tap_m, g_m, sd_after = _oracle_grads(sd, sched, inp, "rdd", n_T, True, True, operand_dtype=torch.bfloat16,
                                                 store_dtype=torch.bfloat16)
    l_f, _, tap_f, g_f, _ = _oracle_grads(sd, sched, inp, "rdd", n_T, True, True)
    print(f"cfg2 train loss: ours {loss:.6f}  kernel-matched oracle {l_m:.6f}  fp32 oracle {l_f:.6f}")
    assert abs(loss - l_m) < 1e-3 * abs(l_m) and abs(loss - l_f) < 1e-2 * abs(l_f)
    e_first = P.rel_l2(got["init_conv"], tap_m["init_conv"])
    print(f"cfg2 init_conv output vs kernel-matched oracle: {e_first:.3e}")
    assert e_first < 1e-3
    e_ours = {n_: P.rel_l2(v, tap_f[n_]) for n_, v in got.items()}
    e_mat = {n_: P.rel_l2(tap_m[n_], tap_f[n_]) for n_ in got}
    print("cfg2 train per-block output rel-L2 vs fp32 oracle, ours:          ", _fmt(e_ours))
    print("cfg2 train per-block output rel-L2 vs fp32 oracle, kernel-matched:", _fmt(e_mat))
    for n_ in got:
        assert e_ours[n_] < 1.5 * e_mat[n_] + 1e-3, (n_, e_ours[n_], e_mat[n_])
    ge_ours = _block_grad_errors(mine, g_f, RDD_BLOCKS)
    ge_mat = _block_grad_errors(g_m, g_f, RDD_BLOCKS)
    print("cfg2 train per-block gradient rel-L2 vs fp32 oracle, ours:          ", _fmt(ge_ours))
    print("cfg2 train per-block gradient rel-L2 vs fp32 oracle, kernel-matched:", _fmt(ge_mat))
    for n_ in ge_ours:
        assert ge_ours[n_] < 1.5 * ge_mat[n_] + 1e-3, (n_, ge_ours[n_], ge_mat[n_])
    # the last blocks of the backward pass see few roundings: tight against the kernel-matched oracle
    ge_tail = _block_grad_errors(mine, g_m, ["out", "local_enhance"])
    print("cfg2 train gradient of the head / LocalEnhancer vs kernel-matched oracle:", _fmt(ge_tail))
    assert max(ge_tail.values()) < 1e-2
    # BatchNorm running statistics after the step
    bn = [k_ for k_ in sd if "running_" in k_]
    got_bn = torch.cat([ddpm.state_dict()[k_].flatten().cpu() for k_ in bn])
    ref_bn = torch.cat([sd_after[k_].flatten() for k_ in bn])
    assert P.rel_l2(got_bn, ref_bn) < 1e-2


def test_cfg5_width_step_vs_oracle_run_on_the_gpu(dev):
    """cfg5 (BASELINE.json configs[4], the stress shape: n_feat 384, 1.41 B parameters, 3x256x256), batch 2, running-statistics
    BatchNorm: loss, every block's output and every block's parameter gradients against the fp32 oracle, with the bars of
    the cfg2 test.  The oracle is the same `ref_port` code, executed here on the GPU in fp32 (TF32 off) through aten / cuDNN --
    seconds instead of minutes on the host cores; layers up to 3072 -> 3072 and the 6144-channel dual source run only here."""
    tf32 = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ddpm, sd, inp, loss, got, mine, n_T = _cfg2(dev, False, 21, n_feat=384, batch=2)
        del ddpm
        torch.cuda.empty_cache()
        sched = {k_: v.to(dev) for k_, v in P.ddpm_schedules(1e-4, 0.02, n_T).items()}
        sd_d = {k_: v.to(dev) for k_, v in sd.items()}
        inp_d = {k_: v.to(dev) for k_, v in inp.items()}
        l_o, _, tap, g_o, _ = _oracle_grads(sd_d, sched, inp_d, "rdd", n_T, False, True)
        tap = {k_: v.cpu() for k_, v in tap.items()}
        g_o = {k_: v.cpu() for k_, v in g_o.items()}
        _, _, _, g_m, _ = _oracle_grads(sd_d, sched, inp_d, "rdd", n_T, False, True, operand_dtype=torch.bfloat16,
                                        store_dtype=torch.bfloat16)
        g_m = {k_: v.cpu() for k_, v in g_m.items()}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    errs = {n_: P.rel_l2(v, tap[n_]) for n_, v in got.items()}
    gerrs = _block_grad_errors(mine, g_o, RDD_BLOCKS)
    gmat = _block_grad_errors(g_m, g_o, RDD_BLOCKS)
    print(f"cfg5 eval-mode loss: ours {loss:.6f}  fp32 oracle (on the GPU) {l_o:.6f}")
    print("cfg5 eval-mode per-block output rel-L2 vs fp32 oracle:", _fmt(errs))
    print("cfg5 eval-mode per-block gradient rel-L2 vs fp32 oracle, ours:          ", _fmt(gerrs))
    print("cfg5 eval-mode per-block gradient rel-L2 vs fp32 oracle, kernel-matched:", _fmt(gmat))
    assert abs(loss - l_o) < 1e-2 * abs(l_o)
    assert max(errs.values()) < 1e-2, errs
    for n_, e in gerrs.items():
        assert e < max(1e-2, 1.25 * gmat[n_]) and e < 2e-2, (n_, e, gmat[n_])


def test_cfg2_eval_forward_vs_fp32_oracle(dev):
    """Eval-mode denoiser at F=192, 256x256 against the fp32 oracle: the north-star 1e-2 bar for bf16."""
    n_feat, size, n_classes, n_T, seed = 192, 256, 5, 700, 12
    ddpm, sd = build("rdd", n_feat, n_classes, n_T, seed, dev)
    ddpm.eval()
    inp = make_inputs("rdd", 1, 3, size, n_classes, n_T, seed)
    sched = P.ddpm_schedules(1e-4, 0.02, n_T)
    x_t = P.q_sample(sched, inp["x"], inp["ts"], inp["noise"])
    with torch.no_grad():
        pred = ddpm.nn_model(x_t.to(dev), inp["c"].to(dev), (inp["ts"] / n_T).to(dev), inp["ctx_mask"].to(dev))
        ref = P.unet_forward(sd, x_t, inp["c"], inp["ts"] / n_T, inp["ctx_mask"], variant="rdd", training=False,
                             prefix="nn_model.")
    e = P.rel_l2(pred.cpu(), ref)
    print(f"cfg2 eval forward rel-L2 vs fp32 oracle = {e:.3e}")
    assert e < 1e-2


def test_cfg1_mnist_train_step_at_size(dev):
    """cfg1: MNIST ContextUnet F=128, batch 128 (MNIST_script.py:303-349): loss + gradients, train mode."""
    n_feat, size, batch, n_classes, n_T, seed = 128, 28, 128, 10, 400, 13
    ddpm, sd = build("mnist", n_feat, n_classes, n_T, seed, dev)
    ddpm.train()
    inp = make_inputs("mnist", batch, 1, size, n_classes, n_T, seed)
    x, c, ts, noise, ctx = (inp[k].to(dev) for k in ("x", "c", "ts", "noise", "ctx_mask"))
    loss = ddpm(x, c, randoms=(ts, noise, ctx))
    loss.backward()
    torch.cuda.synchronize()
    sched = P.ddpm_schedules(1e-4, 0.02, n_T)
    l_o, _, _, g_o, _ = _oracle_grads(sd, sched, inp, "mnist", n_T, True, False, operand_dtype=torch.bfloat16,
                                      store_dtype=torch.bfloat16)
    l_32, _, _, g_32, _ = _oracle_grads(sd, sched, inp, "mnist", n_T, True, False)
    mine = grads_of(ddpm)

    def agg(ref):
        num = sum(float((mine[k].double() - ref[k].double()).pow(2).sum()) for k in ref if k in mine)
        return (num / sum(float(ref[k].double().pow(2).sum()) for k in ref)) ** 0.5
    e_m, e_32 = agg(g_o), agg(g_32)
    print(f"cfg1 loss ours {float(loss):.6f} matched {l_o:.6f} fp32 {l_32:.6f}; grad rel-L2 vs matched {e_m:.3e}, vs fp32 {e_32:.3e}")
    assert abs(float(loss) - l_o) < 1e-2 * abs(l_o) and abs(float(loss) - l_32) < 2e-2 * abs(l_32)
    assert e_m < 1e-2


# N, H, W, Cin (c0 [+ c1]), Cout, k, stride, pad, expected forward kernel, expected weight-gradient kernel
SHAPE_CLASSES = [
    ("down4.res 1536->1536 @32^2", 4, 32, 32, (1536,), 1536, 3, 1, 1, "conv3x3_halo2", "wgrad3_pair"),
    ("up1.model.0 3072(dual)->768 @32^2", 4, 32, 32, (1536, 1536), 768, 3, 1, 1, "conv3x3_halo2", "wgrad3_pair"),
    ("down4.down.4 1536 4x4 s2 32^2->16^2", 4, 32, 32, (1536,), 1536, 4, 2, 1, "conv_gemm2", None),
    ("up3.model.0 768(dual)->192 @128^2", 4, 128, 128, (384, 384), 192, 3, 1, 1, "conv3x3_halo2", "wgrad2_gemm"),
    ("down1 192->192 @256^2", 4, 256, 256, (192,), 192, 3, 1, 1, "conv3x3_halo2", "wgrad2_gemm"),
    ("channel_compress 1x1 768->192 @64^2", 4, 64, 64, (768,), 192, 1, 1, 0, "conv_gemm2", "wgrad_gemm"),
]


@pytest.mark.parametrize("case", SHAPE_CLASSES, ids=lambda c: c[0].split()[0])
def test_conv_shape_classes_at_benchmark_size(dev, case):
    """One convolution per SURVEY Appendix A.1 shape class at its cfg2 size: forward, data gradient (both sources),
    weight gradient -- against fp32 CPU convolutions of the bf16-rounded operands -- and which kernels ran."""
    from diffusionmodel_b200 import _lib, ops
    name, n, h, w, cins, cout, k, stride, pad, fwd_kernel, wgrad_kernel = case
    cin = sum(cins)
    g = torch.Generator().manual_seed(len(name))
    xs = [bf(torch.randn(n, ci, h, w, generator=g)) for ci in cins]
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    b = torch.randn(cout, generator=g) * 0.1
    xr = [t.clone().requires_grad_(True) for t in xs]
    wr, br = bf(wt).requires_grad_(True), b.clone().requires_grad_(True)
    y_ref = F.conv2d(torch.cat(xr, 1), wr, br, stride, pad)
    dy = bf(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)
    xd = [nhwc(t, dev).requires_grad_(True) for t in xs]
    wd, bd = torch.nn.Parameter(wt.to(dev)), torch.nn.Parameter(b.to(dev))
    with _lib.kernel_counts() as kc:
        y, _ = ops.conv2d(xd[0], wd, bd, ops.WeightPack(), x1=xd[1] if len(xd) > 1 else None,
                          c1=cins[1] if len(cins) > 1 else 0, stride=stride, pad=pad)
        fwd = _lib.last_kernel()
    assert fwd[0] == fwd_kernel, (fwd, fwd_kernel)
    assert rel(nchw(y, cout), y_ref) < BF16_TOL
    with _lib.kernel_counts() as kb:
        y.backward(nhwc(dy, dev))
        torch.cuda.synchronize()
    print(f"{name}: forward {fwd}, backward launches {kb.delta}")
    if wgrad_kernel is not None:
        assert kb[wgrad_kernel] == 1, kb.delta
    for t, tr, ci in zip(xd, xr, cins):
        assert rel(nchw(t.grad, ci), tr.grad) < BF16_TOL
    assert rel(wd.grad.cpu(), wr.grad) < F32_TOL
    assert rel(bd.grad.cpu(), br.grad) < F32_TOL


def test_up0_conv_transpose_at_benchmark_size(dev):
    """up0 = ConvTranspose2d(1536, 1536, 8, 8) on the 4 x 2 x 2 bottleneck (new_scripy.py:297-301): forward through the
    scatter-epilogue GEMM, data gradient through the skinny split-K kernel (16 x 98304 x 1536), weight gradient."""
    from diffusionmodel_b200 import _lib, ops
    n, cin, cout, k = 4, 1536, 1536, 8
    g = torch.Generator().manual_seed(3)
    x = bf(torch.randn(n, cin, 2, 2, generator=g))
    wt = torch.randn(cin, cout, k, k, generator=g) / math.sqrt(cin)
    b = torch.randn(cout, generator=g) * 0.1
    xr, wr, br = x.clone().requires_grad_(True), bf(wt).requires_grad_(True), b.clone().requires_grad_(True)
    y_ref = F.conv_transpose2d(xr, wr, br, k)
    dy = bf(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)
    xd = nhwc(x, dev).requires_grad_(True)
    wd, bd = torch.nn.Parameter(wt.to(dev)), torch.nn.Parameter(b.to(dev))
    y = ops.conv_transpose(xd, wd, bd, ops.WeightPack(), k)
    assert rel(nchw(y, cout), y_ref) < BF16_TOL
    with _lib.kernel_counts() as kb:
        y.backward(nhwc(dy, dev))
        torch.cuda.synchronize()
    assert kb["skinny_gemm"] == 1, kb.delta
    assert rel(nchw(xd.grad, cin), xr.grad) < BF16_TOL
    assert rel(wd.grad.cpu(), wr.grad) < F32_TOL
    assert rel(bd.grad.cpu(), br.grad) < F32_TOL


def test_cfg_trajectory_drift_60_steps(dev):
    """A complete CFG reverse trajectory (n_T = 60: every step including i == 1, new_scripy.py:441-477) at F=32, 128x128,
    5 trajectories, guide_w 2, injected noise, against the fp32 oracle; the bf16 drift is printed every 10 steps."""
    n_feat, size, n_classes, n_T, seed, w = 32, 128, 5, 60, 17, 2.0
    ddpm, sd = build("rdd", n_feat, n_classes, n_T, seed, dev)
    ddpm.eval()
    gg = torch.Generator().manual_seed(seed)
    x_T = torch.randn(n_classes, 3, size, size, generator=gg)
    zs = {i: torch.randn(n_classes, 3, size, size, generator=gg) for i in range(n_T, 1, -1)}
    sched = P.ddpm_schedules(1e-4, 0.02, n_T)
    trace = []
    with torch.no_grad():
        ref = P.ddpm_sample(sd, sched, x_T, zs, w, variant="rdd", n_T=n_T, n_classes=n_classes, trace=trace)
    drift = {}
    for steps in (1, 10, 20, 30, 40, 50, 60):
        out = ddpm.sample(n_classes, (3, size, size), dev, guide_w=w, steps=steps, noise=(x_T, zs))
        drift[steps] = P.rel_l2(out.cpu(), trace[steps - 1])
    print("CFG trajectory rel-L2 vs fp32 oracle by step:", {k: f"{v:.2e}" for k, v in drift.items()})
    assert torch.equal(trace[-1], ref)
    assert drift[1] < 1e-2 and drift[60] < 3e-2
    assert all(v < 3e-2 for v in drift.values())


def test_sample_after_load_state_dict_uses_the_new_weights(dev):
    """A cached sampling graph must see weights changed behind its back (load_state_dict of another checkpoint, the
    reference's load-best-then-sample flow new_scripy.py:935): sample, load different weights, sample again == a fresh
    DDPM with those weights."""
    n_feat, size, n_classes, n_T = 64, 128, 5, 20
    ddpm, _ = build("rdd", n_feat, n_classes, n_T, 1, dev)
    ddpm.eval()
    gg = torch.Generator().manual_seed(2)
    x_T = torch.randn(n_classes, 3, size, size, generator=gg)
    zs = {i: torch.randn(n_classes, 3, size, size, generator=gg) for i in range(n_T, n_T - 3, -1)}
    first = ddpm.sample(n_classes, (3, size, size), dev, guide_w=2.0, steps=3, noise=(x_T, zs)).cpu()
    fresh, sd2 = build("rdd", n_feat, n_classes, n_T, 2, dev)
    fresh.eval()
    want = fresh.sample(n_classes, (3, size, size), dev, guide_w=2.0, steps=3, noise=(x_T, zs)).cpu()
    ddpm.load_state_dict(sd2)
    again = ddpm.sample(n_classes, (3, size, size), dev, guide_w=2.0, steps=3, noise=(x_T, zs)).cpu()
    assert P.rel_l2(first, want) > 1e-2            # the two checkpoints really differ
    assert P.rel_l2(again, want) < 1e-5


@pytest.mark.parametrize("variant", ["rdd", "mnist"])
def test_batched_guidance_scales_equal_sequential_calls(dev, variant):
    """sample(guide_w=[2, 4, 6]) runs the three scales as one trajectory batch (cfg3); each scale's result must be the
    sequential call's (new_scripy.py:1036-1041 loops) given the same noise: every per-trajectory computation is identical,
    only per-sample reductions may be partitioned differently over the grid for n and 3n samples."""
    n_feat, size, in_ch, ncls, n_T = (32, 128, 3, 5, 30) if variant == "rdd" else (32, 28, 1, 10, 30)
    ddpm, _ = build(variant, n_feat, ncls, n_T, 4, dev)
    ddpm.eval()
    ws, steps = [2.0, 4.0, 6.0], 4
    gg = torch.Generator().manual_seed(6)
    x_T = torch.randn(3 * ncls, in_ch, size, size, generator=gg)
    zs = {i: torch.randn(3 * ncls, in_ch, size, size, generator=gg) for i in range(n_T, n_T - steps, -1)}
    out = ddpm.sample(ncls, (in_ch, size, size), dev, guide_w=ws, steps=steps, noise=(x_T, zs))
    batched = out[0] if variant == "mnist" else out
    assert isinstance(batched, list) and len(batched) == 3
    worst = 0.0
    for s_, w in enumerate(ws):
        sl = slice(s_ * ncls, (s_ + 1) * ncls)
        o = ddpm.sample(ncls, (in_ch, size, size), dev, guide_w=w, steps=steps,
                        noise=(x_T[sl], {i: z[sl] for i, z in zs.items()}))
        o = o[0] if variant == "mnist" else o
        assert o.shape == batched[s_].shape
        worst = max(worst, P.rel_l2(batched[s_].cpu(), o.cpu()))
    different = P.rel_l2(batched[0].cpu(), batched[2].cpu())
    print(f"{variant}: batched vs sequential guidance scales, worst rel-L2 {worst:.3e} (scale 2 vs 6: {different:.3e})")
    assert worst < 1e-5 and different > 1e-3
