"""Model-level parity (GPU): blocks, full denoisers, DDPM loss/gradients and the CFG loop against the
oracle (oracle/ref_port.py, pinned bit-exact to the reference) and the committed golden fixtures.

Bars (north_star): rel-L2 <= 1e-2 for the bf16 path.  Per SURVEY.md Appendix D the bar is reachable
per block and end-to-end in eval mode against the fp32 reference; end-to-end train mode is checked
against the precision-matched oracle (same bf16 operand rounding), with the fp32 distance reported.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ref_port as P
from oracle.synth import fill_state_dict_, make_inputs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
BAR = 1e-2


def build(variant, n_feat, n_classes, n_T, seed, dev, **kw):
    import diffusionmodel_b200 as D
    net = D.ContextUnet(3, n_feat, n_classes) if variant == "rdd" else D.MnistContextUnet(1, n_feat, n_classes)
    ddpm = D.DDPM(net, (1e-4, 0.02), n_T, "cpu", 0.1, **kw)
    sd = {k: v.clone() for k, v in ddpm.state_dict().items()}
    fill_state_dict_(sd, seed)
    ddpm.load_state_dict(sd)
    ddpm.device = dev
    return ddpm.to(dev), sd


def grads_of(ddpm):
    return {k: p.grad.detach().cpu().clone() for k, p in ddpm.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("tag", ["mnist_f16_b8", "rdd_f16_s128_b2", "rdd_f32_s128_b1_nomap"])
def test_against_golden(dev, tag):
    g = np.load(os.path.join(GOLD, tag + ".npz"), allow_pickle=False)
    n_feat, size, batch, n_classes, seed, n_T, steps, use_map = (int(v) for v in g["meta"])
    variant = "mnist" if tag.startswith("mnist") else "rdd"
    in_ch = 1 if variant == "mnist" else 3
    ddpm, sd = build(variant, n_feat, n_classes, n_T, seed, dev, enhance_with_attn_map=bool(use_map))
    inp = make_inputs(variant, batch, in_ch, size, n_classes, n_T, seed)
    x, c, attn, ts, noise, ctx = (inp[k].to(dev) for k in ("x", "c", "attn_mask", "ts", "noise", "ctx_mask"))
    sched = P.ddpm_schedules(1e-4, 0.02, n_T)
    x_t = P.q_sample(sched, inp["x"], inp["ts"], inp["noise"]).to(dev)

    # ---- eval mode vs the fp32 reference (golden): the 1e-2 bar holds end to end
    ddpm.eval()
    with torch.no_grad():
        kw = {"attn_map": attn} if (variant == "rdd" and use_map) else {}
        pred = ddpm.nn_model(x_t, c, ts / n_T, ctx, **kw)
    e = P.rel_l2(pred.cpu(), torch.from_numpy(g["pred_eval"]))
    print(f"{tag}: eval pred rel-L2 vs fp32 reference = {e:.3e}")
    assert e < BAR
    loss = ddpm(x, c, attn if variant == "rdd" else None, randoms=(ts, noise, ctx))
    assert abs(float(loss) - float(g["loss_eval"])) < BAR * abs(float(g["loss_eval"]))

    # ---- eval-mode gradients (no batch-statistic amplification) vs the fp32 oracle: every parameter
    ddpm.zero_grad()
    loss.backward()
    torch.cuda.synchronize()

    def oracle_grads(training, **kw):
        sd_o = {k: v.clone() for k, v in sd.items()}
        for k, v in sd_o.items():
            if v.is_floating_point() and k.startswith("nn_model.") and "running" not in k:
                v.requires_grad_(True)
        lo = P.ddpm_loss(sd_o, sched, inp["x"], inp["c"], inp["attn_mask"], inp["ts"], inp["noise"], inp["ctx_mask"],
                         variant=variant, n_T=n_T, training=training,
                         attn_map=inp["attn_mask"] if (variant == "rdd" and use_map) else None, **kw)
        lo.backward()
        return float(lo), {k: v.grad for k, v in sd_o.items() if v.requires_grad and v.grad is not None}

    def agg_err(mine, ref):
        num = den = 0.0
        worst = (0.0, "")
        for k, gr in mine.items():
            go = ref.get(k)
            if go is None:
                continue
            num += float((gr.double() - go.double()).pow(2).sum()); den += float(go.double().pow(2).sum())
        for k, gr in mine.items():
            go = ref.get(k)
            if go is not None and float(go.double().pow(2).sum()) > 1e-4 * den:      # tensors that carry real signal
                r = P.rel_l2(gr, go)
                if r > worst[0]:
                    worst = (r, k)
        return (num / den) ** 0.5, worst

    # bar: 1e-2 where bf16 itself allows it; on configurations where PyTorch's own autocast(bfloat16) run
    # of the reference is already beyond 0.8e-2 from fp32 (the floor from operand rounding alone is
    # 0.86e-2 on rdd_f16_s128_b2), no farther than 1.25x that distance.
    _, g_eval = oracle_grads(False)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        _, g_eval_ac = oracle_grads(False)
    e_g, worst = agg_err(grads_of(ddpm), g_eval)
    e_ac_eval, _ = agg_err({k: v.float() for k, v in g_eval_ac.items()}, g_eval)
    print(f"{tag}: eval-mode grad rel-L2 vs fp32 oracle = {e_g:.3e} (torch autocast(bf16): {e_ac_eval:.3e}); "
          f"worst significant tensor {worst[1]} {worst[0]:.3e}")
    assert e_g < max(BAR, 1.25 * e_ac_eval) and worst[0] < 3e-2

    # ---- train mode.  Batch-statistic BatchNorm amplifies bf16 rounding chaotically (SURVEY.md App. D:
    # PyTorch's own autocast(bf16) run of the reference is ~1e-1 from fp32), so the end-to-end train-mode
    # criterion is: loss within 2% of the fp32 reference, and our gradient no farther from the fp32
    # reference than 1.5x the distance of the reference under torch.autocast(bfloat16).
    ddpm.train()
    ddpm.zero_grad()
    loss = ddpm(x, c, attn if variant == "rdd" else None, randoms=(ts, noise, ctx))
    loss.backward()
    torch.cuda.synchronize()
    l_ref, g_ref = oracle_grads(True)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        l_ac, g_ac = oracle_grads(True)
    e_ours, _ = agg_err(grads_of(ddpm), g_ref)
    e_ac, _ = agg_err({k: v.float() for k, v in g_ac.items()}, g_ref)
    # the kernel-matched oracle: the reference algorithm rounding to bf16 where the kernels do (ref_port store_dtype)
    l_km, g_km = oracle_grads(True, operand_dtype=torch.bfloat16, store_dtype=torch.bfloat16)
    e_km, _ = agg_err(g_km, g_ref)
    print(f"{tag}: train loss ours {float(loss):.6f} fp32-ref {l_ref:.6f} kernel-matched {l_km:.6f} autocast-bf16 {l_ac:.6f}")
    print(f"{tag}: train grad rel-L2 vs fp32 reference: ours {e_ours:.3e}, kernel-matched oracle {e_km:.3e}, "
          f"torch autocast(bf16) {e_ac:.3e}")
    assert abs(float(loss) - l_ref) < 2e-2 * abs(l_ref) and abs(float(loss) - l_km) < 5e-3 * abs(l_km)
    # no farther from fp32 than a correct bf16 run of the same algorithm (1.25x), nor than torch's own autocast
    assert e_ours < 1.25 * e_km + 2e-3 and e_ours < 1.5 * e_ac + 1e-2
    # BatchNorm running buffers after one train step
    bn_names = [str(s) for s in g["bn_names"]]
    got = torch.cat([ddpm.state_dict()[k].flatten().cpu() for k in bn_names])
    assert P.rel_l2(got, torch.from_numpy(g["bn_after_train"])) < BAR

    # ---- CFG reverse loop, `steps` iterations with injected noise, vs golden
    ddpm.eval()
    ncls = 10 if variant == "mnist" else n_classes
    gg = torch.Generator().manual_seed(seed + 7)
    x_T = torch.randn(ncls, in_ch, size, size, generator=gg)
    zs = {i: torch.randn(ncls, in_ch, size, size, generator=gg) for i in range(n_T, n_T - steps, -1)}
    out = ddpm.sample(ncls, (in_ch, size, size), dev, guide_w=2.0, steps=steps, noise=(x_T, zs))
    xs = out[0] if variant == "mnist" else out
    e = P.rel_l2(xs.cpu(), torch.from_numpy(g["sample_x"]))
    print(f"{tag}: {steps}-step CFG sample rel-L2 vs fp32 reference = {e:.3e}")
    assert e < BAR


@pytest.mark.parametrize("block", ["res_se", "unet_down", "coord_attn", "unet_up", "local_enhancer"])
def test_blocks_vs_oracle(dev, block):
    """Per-block fwd/bwd, train mode, against the fp32 oracle block with shared inputs (bar 1e-2)."""
    from diffusionmodel_b200 import unet as U
    from tests.test_gpu_kernels import bf, nchw, nhwc
    g = torch.Generator().manual_seed(31)
    n, f, s = 4, 64, 32
    if block == "res_se":
        mod, cin, cout = U.ResConvBlock(f, f, is_res=True), f, f
        fn = lambda cx, x: P.res_conv_block(cx, "m", x, True, True)
    elif block == "unet_down":
        mod, cin, cout = U.UnetDown(f, 2 * f), f, 2 * f
        fn = lambda cx, x: P.unet_down_rdd(cx, "m", x)
    elif block == "coord_attn":
        mod, cin, cout = U.CoordAttn(f), f, f
        fn = lambda cx, x: P.coord_attn(cx, "m", x)
    elif block == "unet_up":
        mod, cin, cout = U.UnetUp(2 * f, f // 2), f, f // 2
        s = 16
    else:
        mod, cin, cout = U.LocalEnhancer(f), f, f
    sd = {"m." + k: v.clone() for k, v in mod.state_dict().items()}
    fill_state_dict_(sd, 77)
    mod.load_state_dict({k[2:]: v for k, v in sd.items()})
    mod = mod.to(dev).train()
    x = bf(torch.randn(n, cin, s, s, generator=g))
    skip = bf(torch.randn(n, cin, s, s, generator=g))
    mask = P.synth_attn_mask(n, s, g)
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    xr, sr = x.clone().requires_grad_(True), skip.clone().requires_grad_(True)
    cx = P._Ctx(sd, True)
    if block == "unet_up":
        y_ref = P.unet_up_rdd(cx, "m", xr, sr)
    elif block == "local_enhancer":
        y_ref = P.local_enhancer(cx, "m", xr, mask)
    else:
        y_ref = fn(cx, xr)
    dy = bf(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)
    xd, sdv = nhwc(x, dev).requires_grad_(True), nhwc(skip, dev).requires_grad_(True)
    if block == "unet_up":
        y = mod(xd, sdv, cin, cin)
    elif block == "local_enhancer":
        y = mod(xd, mask.to(dev))
    else:
        y = mod(xd)
    e_out = P.rel_l2(nchw(y, cout), y_ref)
    y.backward(nhwc(dy, dev))
    e_dx = P.rel_l2(nchw(xd.grad, cin), xr.grad)
    num = den = 0.0
    for k, p in mod.named_parameters():
        go = sd["m." + k].grad
        if go is None or p.grad is None:
            continue
        num += float((p.grad.cpu().double() - go.double()).pow(2).sum()); den += float(go.double().pow(2).sum())
    e_dw = (num / max(den, 1e-30)) ** 0.5
    print(f"{block}: out {e_out:.3e} dx {e_dx:.3e} dparams {e_dw:.3e}")
    assert e_out < BAR and e_dx < 1.5e-2 and e_dw < 1.5e-2


def test_shared_encoder_cfg_matches_reference_schedule(dev):
    """Eval mode: running the encoder once for the n trajectories and the decoder on the doubled batch is the
    same computation as the reference's schedule (encoder on the doubled batch).  Exact in exact arithmetic;
    the per-sample reductions (GroupNorm / SE / CoordAttn pooling) split differently for n and 2n samples, so
    the two differ by fp32 summation order only."""
    for variant, n_feat, size, in_ch in (("rdd", 16, 128, 3), ("mnist", 16, 28, 1)):
        ddpm, _ = build(variant, n_feat, 5 if variant == "rdd" else 10, 400, 3, dev)
        ddpm.eval()
        ncls = 10 if variant == "mnist" else 5
        gg = torch.Generator().manual_seed(5)
        x_T = torch.randn(ncls, in_ch, size, size, generator=gg)
        zs = {i: torch.randn(ncls, in_ch, size, size, generator=gg) for i in range(400, 397, -1)}
        outs = []
        for shared in (True, False):
            ddpm.shared_encoder_cfg = shared
            o = ddpm.sample(ncls, (in_ch, size, size), dev, guide_w=2.0, steps=3, noise=(x_T, zs))
            outs.append((o[0] if variant == "mnist" else o).cpu())
        e = P.rel_l2(outs[0], outs[1])
        print(f"{variant}: shared-encoder vs doubled-batch sampling rel-L2 = {e:.3e}")
        assert e < 2e-3


@pytest.mark.parametrize("training", [False, True])
def test_graphed_train_step_matches_eager(dev, training):
    """DDPM.capture_train_step (CUDA graph of forward+backward) against the eager path over two accumulated
    micro-steps, packed weight gradients flushed by the optimizer.  With running-statistics BatchNorm the
    two agree to reduction-order noise; with batch statistics the forward is chaotic in bf16 (an ulp-level
    change of a batch mean re-rolls the rounding of every later activation: two EAGER runs differ by ~1e-1
    in the gradient, see tools/nondet_probe.py), so there only the losses are compared."""
    import diffusionmodel_b200 as D
    inp = make_inputs("rdd", 2, 3, 128, 5, 700, 3)
    x, c, attn, ts, noise, ctx = (inp[k].to(dev) for k in ("x", "c", "attn_mask", "ts", "noise", "ctx_mask"))
    results = []
    for graphed in (False, True):
        ddpm, _ = build("rdd", 16, 5, 700, 3, dev, enhance_with_attn_map=True)
        ddpm.train(training)
        opt = D.FusedAdamW(ddpm.parameters(), lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0)
        if graphed:
            step = ddpm.capture_train_step(x, c, attn, loss_scale=0.5)
            opt.zero_grad()
        losses = []
        for _ in range(2):
            if graphed:
                losses.append(float(step(x, c, attn, randoms=(ts, noise, ctx))))
            else:
                lo = ddpm(x, c, attn, randoms=(ts, noise, ctx)) * 0.5
                lo.backward()
                losses.append(float(lo))
        opt.flush()
        torch.cuda.synchronize()
        results.append((losses, opt.flat_grad.detach().cpu().clone(), float(opt.grad_norm())))
        opt.step()
        torch.cuda.synchronize()
    (l0, g0, n0), (l1, g1, n1) = results
    e = P.rel_l2(g1, g0)
    print(f"training={training}: eager losses {l0} graphed {l1}; grad norms {n0:.5f} {n1:.5f}; rel-L2 {e:.3e}")
    assert abs(l0[0] - l1[0]) < 1e-3 * abs(l0[0]) and abs(l0[1] - l1[1]) < 1e-3 * abs(l0[1])
    assert n1 > 0 and abs(n1 - n0) < 0.2 * n0
    if not training:
        assert e < 1e-3


@pytest.mark.parametrize("variant", ["rdd", "mnist"])
def test_split_backward_graphs_match_single_graph(dev, variant):
    """capture_train_step(split_backward=True) -- forward + decoder-side backward as one graph, the trunk's backward as a
    second (under an SM limit), a callback in between -- accumulates the same gradients as the single graph, and when the
    callback runs every gradient from ``first_decoder_param()`` on is already final."""
    import diffusionmodel_b200 as D
    from diffusionmodel_b200 import parallel
    n_feat, size, in_ch, ncls, n_T = (64, 128, 3, 5, 700) if variant == "rdd" else (32, 28, 1, 10, 400)
    inp = make_inputs(variant, 2, in_ch, size, ncls, n_T, 3)
    x, c, attn, ts, noise, ctx = (inp[k].to(dev) for k in ("x", "c", "attn_mask", "ts", "noise", "ctx_mask"))
    attn = attn if variant == "rdd" else None
    grads, tails = [], []
    for split in (False, True):
        ddpm, _ = build(variant, n_feat, ncls, n_T, 3, dev, **({"enhance_with_attn_map": True} if variant == "rdd" else {}))
        ddpm.eval()
        opt = D.FusedAdamW(ddpm.parameters(), lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0)
        step = ddpm.capture_train_step(x, c, attn, loss_scale=0.5, split_backward=split, trunk_sm_limit=132 if split else 0)
        opt.zero_grad()
        red = parallel.OverlappedGradReduce(opt, ddpm.nn_model.grad_ready_regions())
        seen = {}

        def between(k):
            red.reduce_ready(k)                # world size 1: flushes the packed / queued gradients, no collective
            lo, hi = red.spans[k]
            seen[k] = opt.flat_grad[lo:hi].clone()
        for _ in range(2):
            step(x, c, attn, randoms=(ts, noise, ctx), **({"between": between} if split else {}))
        red.finish()
        torch.cuda.synchronize()
        grads.append(opt.flat_grad.detach().cpu().clone())
        if split:
            assert len(seen) == (2 if variant == "rdd" else 1)
            for k, g in seen.items():                                              # final when the callback ran
                lo, hi = red.spans[k]
                assert torch.equal(g.cpu(), grads[-1][lo:hi]) and 0 < lo < hi <= opt.flat_grad.numel()
    e = P.rel_l2(grads[1], grads[0])
    print(f"{variant}: split-backward vs single-graph gradient rel-L2 {e:.3e}")
    assert e < 1e-3 and float(grads[0].norm()) > 0


def test_state_dict_roundtrip_and_fail_loudly(dev):
    import diffusionmodel_b200 as D
    from diffusionmodel_b200._lib import DmB200Error
    net = D.ContextUnet(3, 16, 5)
    ddpm = D.DDPM(net, (1e-4, 0.02), 700, dev)
    sd = ddpm.state_dict()
    assert len(sd) == 415 and "nn_model.up0.0.weight" in sd and "mab_over_sqrtmab" in sd
    ddpm2 = D.DDPM(D.ContextUnet(3, 16, 5), (1e-4, 0.02), 700, dev)
    ddpm2.load_state_dict(sd)
    with pytest.raises(DmB200Error):
        net(torch.zeros(1, 3, 128, 128), torch.zeros(1, dtype=torch.long), torch.ones(1), torch.ones(1))
    with pytest.raises(RuntimeError):          # the class list would not cover the batch (new_scripy.py:447-448)
        ddpm.to(dev).eval().sample(7, (3, 128, 128), dev, guide_w=1.0, steps=1)


def test_gemm_native_weight_storage(dev):
    """FusedAdamW stores Conv2d weights with Cin % 64 == 0 in the GEMM's [Cout][kh][kw][Cin] order (permuted views:
    same names, shapes and values), so wgrad accumulates straight into ``.grad`` and the optimizer's bf16 shadow is
    the forward pack.  Two optimizer steps (eval-mode norms: deterministic forward) must match the plain layout:
    gradients, updated parameters, the packs used by the next forward, and a state_dict round trip."""
    import diffusionmodel_b200 as D
    from diffusionmodel_b200 import ops
    inp = make_inputs("rdd", 1, 3, 128, 5, 700, 5)
    x, c, attn, ts, noise, ctx = (inp[k].to(dev) for k in ("x", "c", "attn_mask", "ts", "noise", "ctx_mask"))
    outs = []
    for native in (False, True):
        ddpm, _ = build("rdd", 64, 5, 700, 5, dev, enhance_with_attn_map=True)
        ddpm.eval()
        opt = D.FusedAdamW(ddpm.parameters(), lr=1e-3, weight_decay=1e-2, max_grad_norm=1.0, gemm_native_weights=native)
        n_native = sum(opt._native)
        assert (n_native > 20) == native
        losses, g_first = [], None
        for _ in range(2):
            lo = ddpm(x, c, attn, randoms=(ts, noise, ctx))
            lo.backward()
            losses.append(float(lo))
            opt.flush()
            if g_first is None:
                g_first = grads_of(ddpm)
            opt.step()
            opt.zero_grad()
        sd = {k: v.detach().cpu().clone() for k, v in ddpm.state_dict().items()}
        outs.append((losses, g_first, sd))
        if native:
            w = ddpm.nn_model.down1.down[0].weight
            assert not w.is_contiguous() and tuple(w.stride()) == ops.native_strides(w.shape)
            idx = [i for i, q in enumerate(opt._params) if q is w][0]
            assert w.grad.data_ptr() == opt.flat_grad.data_ptr() + 4 * opt._offsets[idx]
            # a reference-layout checkpoint loads into the permuted storage and comes back unchanged
            with torch.no_grad():
                l_before = float(ddpm(x, c, attn, randoms=(ts, noise, ctx)))
                ddpm.load_state_dict({k: v.clone() for k, v in outs[0][2].items()})      # the plain run's weights
                l_other = float(ddpm(x, c, attn, randoms=(ts, noise, ctx)))
                ddpm.load_state_dict({k: v.clone() for k, v in sd.items()})
                back = ddpm.state_dict()
                assert all(torch.equal(back[k].cpu(), sd[k]) for k in sd)
                l_after = float(ddpm(x, c, attn, randoms=(ts, noise, ctx)))
            assert l_after == pytest.approx(l_before, rel=1e-5) and l_other == pytest.approx(l_before, rel=5e-3)
    (l0, g0, s0), (l1, g1, s1) = outs
    print(f"plain losses {l0}  native losses {l1}")
    assert abs(l0[0] - l1[0]) < 1e-5 * abs(l0[0]) and abs(l0[1] - l1[1]) < 2e-3 * abs(l0[1])
    for k in g0:
        assert g0[k].shape == g1[k].shape
        # two runs differ by the order of the wgrad kernels' fp32 red.adds (and of every reduction upstream): ~1e-4..1e-3 on
        # the smallest gradients; a layout mistake would be O(1)
        assert P.rel_l2(g1[k], g0[k]) < 3e-3 or float(g0[k].abs().max()) < 1e-7, k
    # parameters after two Adam steps: the first update is lr*sign(g), so reduction-order noise on near-zero gradient
    # entries flips individual updates (2*lr each); the gradients above are the tight check
    for k in s0:
        if s0[k].is_floating_point():
            assert P.rel_l2(s1[k], s0[k]) < 1e-2, k


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("shape", [(4, 192, 16), (2, 64, 12), (3, 32, 7), (2, 3072, 8), (4, 1536, 16)],
                         ids=lambda s: "x".join(map(str, s)))
def test_coordattn_gate_kernels_vs_oracle(dev, shape, training):
    """CoordAttn (new_scripy.py:97-140) with the gate network on dm_ca_gates_fwd/bwd against the fp32 oracle block on
    the CPU (same bf16-rounded input; everything inside the block is fp32 on both sides): output, dx, every parameter
    gradient, BatchNorm running statistics + counters.  (4, 1536, 16) is ca4 at the benchmarked configuration."""
    from diffusionmodel_b200 import unet as U
    from tests.test_gpu_kernels import bf, nchw, nhwc
    n, c, s = shape
    g = torch.Generator().manual_seed(5)
    mod = U.CoordAttn(c)
    sd = {"m." + k: v.clone() for k, v in mod.state_dict().items()}
    fill_state_dict_(sd, 9)
    for k in ("gamma_h", "gamma_w", "alpha", "beta"):
        sd["m." + k] = torch.randn(1, generator=g) * 0.7
    mod.load_state_dict({k[2:]: v for k, v in sd.items()})
    mod = mod.to(dev).train(training)
    x = bf(torch.randn(n, c, s, s, generator=g))
    dy = bf(torch.randn(n, c, s, s, generator=g))
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    y_ref = P.coord_attn(P._Ctx(sd, training), "m", xr)
    y_ref.backward(dy)
    xd = nhwc(x, dev).requires_grad_(True)
    y = mod(xd)
    y.backward(nhwc(dy, dev))
    torch.cuda.synchronize()
    assert P.rel_l2(nchw(y, c), y_ref) < 4e-3 and P.rel_l2(nchw(xd.grad, c), xr.grad) < 4e-3      # one bf16 rounding
    g_ref = {k[2:]: v.grad for k, v in sd.items() if v.requires_grad and v.grad is not None}
    scale = max(float(v.abs().max()) for v in g_ref.values())
    for k, p in mod.named_parameters():      # (a bias in front of a batch-statistics BatchNorm has zero gradient: noise on both sides)
        go = g_ref[k]
        noise = float(go.abs().max()) < 1e-4 * scale and float(p.grad.abs().max()) < 1e-4 * scale
        assert noise or P.rel_l2(p.grad.cpu(), go) < 1e-3, (k, go.flatten()[:4], p.grad.flatten()[:4])
    for k, v in mod.state_dict().items():
        if "tracked" in k:
            assert int(v) == int(sd["m." + k])
        elif "running" in k:
            assert P.rel_l2(v.cpu(), sd["m." + k]) < 1e-5, k


@pytest.mark.parametrize("variant", ["mnist", "rdd"])
def test_loss_curve_tracks_fp32_reference_training(dev, variant):
    """A dozen optimizer steps (train-mode BatchNorm, clip 1.0, AdamW) from identical weights, data and per-step
    random draws: the bf16 B200 path's loss curve against the fp32 oracle trained with torch.optim.AdamW on the CPU
    (the reference loop, new_scripy.py:784-803 / MNIST_script.py:339-349)."""
    import diffusionmodel_b200 as D
    n_feat, size, batch, n_classes, n_T, in_ch = (16, 28, 16, 10, 400, 1) if variant == "mnist" else (16, 128, 2, 5, 700, 3)
    steps, lr, wd = 12, 1e-3, 1e-5
    ddpm, sd = build(variant, n_feat, n_classes, n_T, 21, dev, **({"enhance_with_attn_map": True} if variant == "rdd" else {}))
    ddpm.train()
    inp = make_inputs(variant, batch, in_ch, size, n_classes, n_T, 21)
    g = torch.Generator().manual_seed(77)
    draws = [(torch.randint(1, n_T + 1, (batch,), generator=g), torch.randn(batch, in_ch, size, size, generator=g),
              torch.bernoulli(torch.full((batch,), 0.9 if variant == "rdd" else 0.1), generator=g)) for _ in range(steps)]
    # ---- fp32 oracle + torch AdamW
    sd_o = {k: v.clone() for k, v in sd.items()}
    params = [v.requires_grad_(True) for k, v in sd_o.items()
              if v.is_floating_point() and k.startswith("nn_model.") and "running" not in k]
    opt_o = torch.optim.AdamW(params, lr=lr, weight_decay=wd)
    sched = P.ddpm_schedules(1e-4, 0.02, n_T)
    ref = []
    for ts, noise, ctx in draws:
        opt_o.zero_grad()
        lo = P.ddpm_loss(sd_o, sched, inp["x"], inp["c"], inp["attn_mask"], ts, noise, ctx, variant=variant, n_T=n_T,
                         training=True, attn_map=inp["attn_mask"] if variant == "rdd" else None)
        lo.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt_o.step()
        ref.append(float(lo))
    # ---- ours
    opt = D.FusedAdamW(ddpm.parameters(), lr=lr, weight_decay=wd, max_grad_norm=1.0)
    x, c, attn = inp["x"].to(dev), inp["c"].to(dev), inp["attn_mask"].to(dev)
    ours = []
    for ts, noise, ctx in draws:
        lo = ddpm(x, c, attn if variant == "rdd" else None, randoms=(ts.to(dev), noise.to(dev), ctx.to(dev)))
        lo.backward()
        opt.step()
        opt.zero_grad()
        ours.append(float(lo))
    dev_rel = [abs(a - b) / abs(b) for a, b in zip(ours, ref)]
    print(f"{variant}: fp32 reference losses {['%.4f' % v for v in ref]}")
    print(f"{variant}: B200 bf16 losses      {['%.4f' % v for v in ours]}  max rel dev {max(dev_rel):.3e}")
    assert max(dev_rel) < 0.05 and sum(dev_rel) / steps < 0.02
    assert sum(ours[-3:]) < sum(ours[:3])              # and it trains


def test_coordattn_non_square_fails_loudly(dev):
    """H != W needs the adaptive_avg_pool1d resampling of the h<->w cross terms (new_scripy.py:118-126), which the gate
    kernels do not implement; the U-Net only ever sees square maps.  No torch fallback: the call raises."""
    from diffusionmodel_b200 import unet as U
    from diffusionmodel_b200._lib import DmB200Error
    mod = U.CoordAttn(32).to(dev)
    with pytest.raises(DmB200Error):
        mod(torch.zeros((2, 12, 20, 32), device=dev, dtype=torch.bfloat16))


@pytest.mark.parametrize("cfg", [(24, 128, 3), (40, 128, 2), (16, 384, 1), (200, 128, 1)], ids=lambda c: "F%d_s%d_b%d" % c)
def test_odd_widths_and_sizes(dev, cfg):
    """Widths that are not multiples of 64 (channel pitches padded to 8, K chunks to 64, mid = C // 16 down to 1) and a
    non-power-of-two image size: train-mode loss against the kernel-matched oracle, finite gradients, two sampling steps."""
    n_feat, size, batch = cfg
    ddpm, sd = build("rdd", n_feat, 5, 700, 3, dev, enhance_with_attn_map=True)
    ddpm.train()
    inp = make_inputs("rdd", batch, 3, size, 5, 700, 3)
    x, c, attn, ts, noise, ctx = (inp[k].to(dev) for k in ("x", "c", "attn_mask", "ts", "noise", "ctx_mask"))
    loss = ddpm(x, c, attn, randoms=(ts, noise, ctx))
    loss.backward()
    torch.cuda.synchronize()
    sched = P.ddpm_schedules(1e-4, 0.02, 700)
    with torch.no_grad():
        lo = P.ddpm_loss({k: v.clone() for k, v in sd.items()}, sched, inp["x"], inp["c"], inp["attn_mask"], inp["ts"],
                         inp["noise"], inp["ctx_mask"], variant="rdd", n_T=700, training=True, attn_map=inp["attn_mask"],
                         operand_dtype=torch.bfloat16, store_dtype=torch.bfloat16)
    print(f"F={n_feat} size={size} B={batch}: loss {float(loss):.5f}  kernel-matched oracle {float(lo):.5f}")
    assert abs(float(loss) - float(lo)) < 3e-3 * abs(float(lo))
    g = grads_of(ddpm)
    assert len(g) > 250 and all(torch.isfinite(v).all() for v in g.values())
    ddpm.eval()
    out = ddpm.sample(5, (3, size, size), dev, guide_w=2.0, steps=2)
    assert out.shape == (5, 3, size, size) and torch.isfinite(out).all()
