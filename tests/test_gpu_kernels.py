"""Kernel-level parity (GPU): every C-ABI entry point against a CPU restatement of the aten op it
replaces, on seeded inputs.  Convolutions are checked against fp32 CPU convolutions of the
bf16-rounded operands (precision-matched: the kernels round operands to bf16 and accumulate in fp32).
Tolerances (rel-L2): bf16-stored outputs 4e-3 (one bf16 rounding), fp32 outputs 2e-4.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import ref_port as P

pytestmark = pytest.mark.gpu

BF16_TOL = 4e-3
F32_TOL = 2e-4


def bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


def nhwc(x, dev, ld=None):
    """fp32 NCHW (CPU) -> bf16 NHWC on the device with zero pad lanes."""
    n, c, h, w = x.shape
    ld = ld or (c + 7) // 8 * 8
    out = torch.zeros((n, h, w, ld), dtype=torch.bfloat16)
    out[..., :c] = x.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out.to(dev)


def nchw(y, c):
    return y[..., :c].float().cpu().permute(0, 3, 1, 2).contiguous()


def rel(a, b):
    return P.rel_l2(a, b)


CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride, pad
    (2, 16, 16, 64, 64, 3, 1, 1),
    (1, 32, 32, 3, 16, 3, 1, 1),
    (2, 16, 16, 16, 3, 3, 1, 1),
    (2, 16, 16, 48, 80, 1, 1, 0),
    (2, 16, 16, 32, 32, 4, 2, 1),
    (3, 28, 28, 16, 16, 3, 1, 1),
    (2, 14, 14, 32, 16, 3, 1, 1),
    (5, 7, 7, 32, 32, 3, 1, 1),
    (1, 32, 32, 320, 272, 3, 1, 1),
    (1, 128, 128, 192, 192, 3, 1, 1),
    (1, 64, 64, 128, 128, 4, 2, 1),
    (1, 256, 256, 64, 64, 3, 1, 1),
    # odd M-tile counts (one 128-pixel tile per image): the CTA-pair kernels run a phantom tile in the last pair
    (9, 8, 16, 64, 64, 1, 1, 0),
    (9, 16, 32, 64, 96, 4, 2, 1),
    (11, 16, 8, 32, 64, 3, 1, 1),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv2d_fwd_bwd(dev, case):
    from diffusionmodel_b200 import ops
    n, h, w, cin, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(hash(case) % 10000)
    x = bf(torch.randn(n, cin, h, w, generator=g))
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    b = torch.randn(cout, generator=g) * 0.1
    xr = x.clone().requires_grad_(True)
    wr = bf(wt).requires_grad_(True)
    br = b.clone().requires_grad_(True)
    y_ref = F.conv2d(xr, wr, br, stride, pad)
    ho, wo = y_ref.shape[2:]
    dy = bf(torch.randn(n, cout, ho, wo, generator=g))
    y_ref.backward(dy)

    xd = nhwc(x, dev).requires_grad_(True)
    wd = torch.nn.Parameter(wt.to(dev))
    bd = torch.nn.Parameter(b.to(dev))
    y, stats = ops.conv2d(xd, wd, bd, ops.WeightPack(), stride=stride, pad=pad, want_stats=True)
    assert rel(nchw(y, cout), y_ref) < BF16_TOL
    if y.shape[3] > cout:
        assert float(y[..., cout:].detach().float().abs().max()) == 0.0          # pad lanes are zero
    # fused BatchNorm statistics: per-channel sum / sum of squares of the outputs AS STORED (bf16-rounded), per-CTA rows
    s = stats.sum(0).double().cpu()
    ys = y[..., :cout].detach().double().cpu()
    assert float((s[0] - ys.sum((0, 1, 2))).abs().max()) < 1e-5 * float(ys.abs().sum((0, 1, 2)).max())
    assert rel(s[1], (ys * ys).sum((0, 1, 2))) < 1e-5
    assert rel(s[1], (y_ref.detach() ** 2).sum((0, 2, 3))) < 1e-3
    y.backward(nhwc(dy, dev))
    torch.cuda.synchronize()
    assert rel(nchw(xd.grad, cin), xr.grad) < BF16_TOL
    assert rel(wd.grad.cpu(), wr.grad) < F32_TOL
    assert rel(bd.grad.cpu(), br.grad) < F32_TOL


def test_conv2d_f32_head_and_dual_source(dev):
    """cat(x0, x1) -> conv without materialising the concat; fp32 output for the prediction head."""
    from diffusionmodel_b200 import ops
    g = torch.Generator().manual_seed(5)
    n, h, w, c0, c1, cout = 2, 16, 16, 16, 16, 3
    x0 = bf(torch.randn(n, c0, h, w, generator=g))
    x1 = bf(torch.randn(n, c1, h, w, generator=g))
    wt = torch.randn(cout, c0 + c1, 3, 3, generator=g) / math.sqrt((c0 + c1) * 9)
    b = torch.randn(cout, generator=g) * 0.1
    x0r, x1r = x0.clone().requires_grad_(True), x1.clone().requires_grad_(True)
    wr = bf(wt).requires_grad_(True)
    y_ref = F.conv2d(torch.cat((x0r, x1r), 1), wr, b, 1, 1)
    dy = torch.randn(n, cout, h, w, generator=g)
    y_ref.backward(bf(dy))
    x0d, x1d = nhwc(x0, dev).requires_grad_(True), nhwc(x1, dev).requires_grad_(True)
    wd = torch.nn.Parameter(wt.to(dev))
    bd = torch.nn.Parameter(b.to(dev))
    y, _ = ops.conv2d(x0d, wd, bd, ops.WeightPack(), x1=x1d, c1=c1, stride=1, pad=1, out_f32=True)
    assert y.dtype == torch.float32 and y.shape[3] == 4
    assert rel(nchw(y, cout), y_ref) < F32_TOL
    gy = torch.zeros_like(y)
    gy[..., :cout] = bf(dy).permute(0, 2, 3, 1).to(dev)
    y.backward(gy)
    assert rel(nchw(x0d.grad, c0), x0r.grad) < BF16_TOL
    assert rel(nchw(x1d.grad, c1), x1r.grad) < BF16_TOL
    assert rel(wd.grad.cpu(), wr.grad) < F32_TOL


@pytest.mark.parametrize("case", [(2, 16, 16, 64, 64), (1, 32, 32, 192, 192), (1, 48, 24, 40, 24), (3, 16, 8, 16, 16),
                                  (1, 128, 128, 192, 192), (1, 64, 64, 320, 272)], ids=lambda c: "x".join(map(str, c)))
def test_conv3x3_halo_kernel(dev, case):
    """The input-patch-reuse 3x3 kernel (nine shifted descriptor views of one 10x18 halo patch) against the
    fp32 CPU convolution and against the generic per-tap implicit-GEMM kernel (same products, accumulated
    chunk-major instead of tap-major, so equal up to fp32 summation order / one bf16 rounding)."""
    from diffusionmodel_b200 import _lib, ops
    n, h, w, cin, cout = case
    g = torch.Generator().manual_seed(43)
    x = bf(torch.randn(n, cin, h, w, generator=g))
    wt = torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    b = torch.randn(cout, generator=g) * 0.1
    y_ref = F.conv2d(x, bf(wt), b, 1, 1)
    wd, bd = torch.nn.Parameter(wt.to(dev)), torch.nn.Parameter(b.to(dev))
    outs = {}
    for name, key5 in (("halo", 0), ("halo_single_cta", 2), ("generic", 1)):   # 0: CTA pairs where the tile width allows
        _lib.debug_set(5, key5)
        try:
            with torch.no_grad():
                y, stats = ops.conv2d(nhwc(x, dev), wd, bd, ops.WeightPack(), stride=1, pad=1, want_stats=True)
            torch.cuda.synchronize()
        finally:
            _lib.debug_set(5, 0)
        outs[name] = (y.clone(), stats.sum(0).cpu())
        assert rel(nchw(y, cout), y_ref) < BF16_TOL, name
    for name in ("halo", "halo_single_cta"):
        assert rel(outs[name][0].float(), outs["generic"][0].float()) < 1e-3, name
        assert rel(outs[name][1], outs["generic"][1]) < 1e-5, name


@pytest.mark.parametrize("version", [1, 2, 3])
@pytest.mark.parametrize("case", [(2, 16, 16, 64, 512, 3, 1, 1), (1, 32, 32, 192, 192, 3, 1, 1), (2, 16, 16, 96, 128, 4, 2, 1),
                                  (4, 8, 8, 320, 640, 1, 1, 0), (3, 10, 12, 40, 24, 3, 1, 1)],
                         ids=lambda c: "x".join(map(str, c)))
def test_wgrad_both_kernels(dev, case, version):
    """The weight-gradient kernels (1: M = Cout tiles, 2: M = im2col boxes with two accumulators per CTA, 3: CTA
    pairs with one accumulator per CTA) on shapes any of them may be chosen for."""
    from diffusionmodel_b200 import _lib, ops
    n, h, w, cin, cout, k, stride, pad = case
    g = torch.Generator().manual_seed(41)
    x = bf(torch.randn(n, cin, h, w, generator=g))
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    xr = x.clone()
    wr = bf(wt).requires_grad_(True)
    y_ref = F.conv2d(xr, wr, None, stride, pad)
    dy = bf(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)
    wd = torch.nn.Parameter(wt.to(dev))
    _lib.debug_set(4, version)
    try:
        y, _ = ops.conv2d(nhwc(x, dev), wd, None, ops.WeightPack(), stride=stride, pad=pad)
        y.backward(nhwc(dy, dev))
        torch.cuda.synchronize()
    finally:
        _lib.debug_set(4, 0)
    assert rel(wd.grad.cpu(), wr.grad) < F32_TOL


@pytest.mark.parametrize("case", [(4, 2, 2, 128, 128, 8), (3, 1, 1, 32, 32, 7), (2, 7, 7, 64, 16, 2), (2, 2, 2, 72, 96, 8)],
                         ids=lambda c: "x".join(map(str, c)))
def test_conv_transpose(dev, case):
    from diffusionmodel_b200 import ops
    n, h, w, cin, cout, k = case
    g = torch.Generator().manual_seed(7)
    x = bf(torch.randn(n, cin, h, w, generator=g))
    wt = torch.randn(cin, cout, k, k, generator=g) / math.sqrt(cin)
    b = torch.randn(cout, generator=g) * 0.1
    xr = x.clone().requires_grad_(True)
    wr = bf(wt).requires_grad_(True)
    br = b.clone().requires_grad_(True)
    y_ref = F.conv_transpose2d(xr, wr, br, k)
    dy = bf(torch.randn(y_ref.shape, generator=g))
    y_ref.backward(dy)
    xd = nhwc(x, dev).requires_grad_(True)
    wd, bd = torch.nn.Parameter(wt.to(dev)), torch.nn.Parameter(b.to(dev))
    y = ops.conv_transpose(xd, wd, bd, ops.WeightPack(), k)
    assert rel(nchw(y, cout), y_ref) < BF16_TOL
    y.backward(nhwc(dy, dev))
    assert rel(nchw(xd.grad, cin), xr.grad) < BF16_TOL
    assert rel(wd.grad.cpu(), wr.grad) < F32_TOL
    assert rel(bd.grad.cpu(), br.grad) < F32_TOL


@pytest.mark.parametrize("passes", [1, 4, 11])
def test_conv_transpose_queued_weight_gradient(dev, passes):
    """ConvTranspose weights owned by FusedAdamW that see <= 64 pixels per backward queue their wgrad operands
    (ops._RankGrad) and contract them with one GEMM at flush: the gradient after `passes` backward passes (11 > the
    queue depth of 8 forces an intermediate flush) must equal the sum of the per-pass gradients."""
    from diffusionmodel_b200 import ops, FusedAdamW
    n, h, w, cin, cout, k = 2, 2, 2, 64, 32, 8
    g = torch.Generator().manual_seed(37)
    wt = torch.randn(cin, cout, k, k, generator=g) / math.sqrt(cin)
    wr = bf(wt).requires_grad_(True)
    wd = torch.nn.Parameter(wt.to(dev))
    opt = FusedAdamW([wd], lr=1e-3)
    pack = ops.WeightPack()
    dx_ref = []
    for i in range(passes):
        x = bf(torch.randn(n, cin, h, w, generator=g))
        dy = bf(torch.randn(n, cout, h * k, w * k, generator=g))
        xr = x.clone().requires_grad_(True)
        F.conv_transpose2d(xr, wr, None, k).backward(dy)
        xd = nhwc(x, dev).requires_grad_(True)
        ops.conv_transpose(xd, wd, None, pack, k).backward(nhwc(dy, dev))
        assert rel(nchw(xd.grad, cin), xr.grad) < BF16_TOL
    entry = ops._rank_grad_entry(wd, n, h, w, cin, cin, k * k * cout)
    assert isinstance(entry, ops._RankGrad) and entry.count == (passes if passes <= 8 else passes - 8)
    opt.flush()
    assert entry.count == 0
    assert rel(wd.grad.cpu(), wr.grad) < F32_TOL
    opt.zero_grad()
    assert float(wd.grad.abs().max()) == 0.0


@pytest.mark.parametrize("case", [(16, 1536, 8192), (5, 100, 2144), (1, 8, 32), (16, 72, 6144)], ids=lambda c: "x".join(map(str, c)))
def test_skinny_gemm(dev, case):
    """dm_skinny_gemm (split-K mma.sync GEMM for <= 16 rows, the up0 data gradient) against an fp32 matmul of the
    bf16 operands; ragged M / N, a partial last K slice, and pad lanes of the output left untouched."""
    from diffusionmodel_b200 import ops, _lib
    m, n, k = case
    g = torch.Generator().manual_seed(31)
    a = bf(torch.randn(m, k, generator=g))
    w = bf(torch.randn(n, k, generator=g)) / math.sqrt(k)
    w = bf(w)
    ref = a @ w.t()
    ad, wd = a.to(torch.bfloat16).to(dev), w.to(torch.bfloat16).to(dev)
    ldo = (n + 7) // 8 * 8 + 8
    out = torch.full((m, ldo), 7.0, dtype=torch.bfloat16, device=dev)
    scratch = torch.empty(_lib.fn("dm_skinny_gemm_scratch")(n, k), device=dev, dtype=torch.float32)
    ops.call("dm_skinny_gemm", ops._p(ad), k, ops._p(wd), k, ops._p(out), ldo, ops._p(scratch), m, n, k, ops._stream())
    assert rel(out[:, :n].float().cpu(), ref) < BF16_TOL
    assert float((out[:, n:].float() - 7.0).abs().max()) == 0.0


@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("shape", [(4, 24, 16, 16), (2, 192, 32, 32), (3, 20, 7, 7)])
def test_bn_gelu(dev, shape, training):
    """Sequential(Conv2d, BatchNorm2d, GELU) against the reference ops incl. running-stat update."""
    from diffusionmodel_b200 import ops, unet
    n, c, h, w = shape
    g = torch.Generator().manual_seed(11)
    seq = torch.nn.Sequential(torch.nn.Conv2d(c, c, 3, 1, 1), torch.nn.BatchNorm2d(c), torch.nn.GELU())
    with torch.no_grad():
        seq[1].weight.copy_(0.5 + torch.rand(c, generator=g)); seq[1].bias.copy_(torch.randn(c, generator=g) * 0.2)
        seq[1].running_mean.copy_(torch.randn(c, generator=g) * 0.1); seq[1].running_var.copy_(0.5 + torch.rand(c, generator=g))
        seq[0].weight.copy_(bf(seq[0].weight))
    import copy
    ref = copy.deepcopy(seq).train(training)
    x = bf(torch.randn(n, c, h, w, generator=g))
    xr = x.clone().requires_grad_(True)
    z_ref = ref(xr)
    dz = bf(torch.randn(z_ref.shape, generator=g))
    z_ref.backward(dz)
    mod = seq.to(dev).train(training)
    xd = nhwc(x, dev).requires_grad_(True)
    z = unet.conv_bn_act(xd, mod)
    assert rel(nchw(z, c), z_ref) < 6e-3
    z.backward(nhwc(dz, dev))
    assert rel(nchw(xd.grad, c), xr.grad) < 1.2e-2
    assert rel(mod[1].weight.grad.cpu(), ref[1].weight.grad) < 8e-3
    assert rel(mod[1].bias.grad.cpu(), ref[1].bias.grad) < 8e-3
    assert rel(mod[0].weight.grad.cpu(), ref[0].weight.grad) < 1.2e-2
    if training:
        assert rel(mod[1].running_mean.cpu(), ref[1].running_mean) < 2e-3
        assert rel(mod[1].running_var.cpu(), ref[1].running_var) < 2e-3
        assert int(mod[1].num_batches_tracked) == int(ref[1].num_batches_tracked)


@pytest.mark.parametrize("act", ["relu", "gelu"])
@pytest.mark.parametrize("shape", [(2, 64, 16, 16), (3, 16, 7, 7), (2, 192, 32, 32)])
def test_group_norm_act(dev, shape, act):
    from diffusionmodel_b200 import ops
    n, c, h, w = shape
    g = torch.Generator().manual_seed(13)
    gn = torch.nn.GroupNorm(8, c)
    with torch.no_grad():
        gn.weight.copy_(0.5 + torch.rand(c, generator=g)); gn.bias.copy_(torch.randn(c, generator=g) * 0.2)
    import copy
    ref = copy.deepcopy(gn)
    x = bf(torch.randn(n, c, h, w, generator=g) * 1.5 + 0.3)
    xr = x.clone().requires_grad_(True)
    z_ref = (F.relu if act == "relu" else F.gelu)(ref(xr))
    dz = bf(torch.randn(z_ref.shape, generator=g))
    z_ref.backward(dz)
    gn = gn.to(dev)
    xd = nhwc(x, dev).requires_grad_(True)
    z = ops.gn_act(xd, gn, ops.ACT_RELU if act == "relu" else ops.ACT_GELU)
    assert rel(nchw(z, c), z_ref) < BF16_TOL
    z.backward(nhwc(dz, dev))
    assert rel(nchw(xd.grad, c), xr.grad) < BF16_TOL
    assert rel(gn.weight.grad.cpu(), ref.weight.grad) < 1e-3
    assert rel(gn.bias.grad.cpu(), ref.bias.grad) < 1e-3


def test_upcat_film_pool(dev):
    from diffusionmodel_b200 import ops
    g = torch.Generator().manual_seed(17)
    n, ca, cb, h, w = 2, 16, 24, 8, 8
    a = bf(torch.randn(n, ca, h, w, generator=g)); b = bf(torch.randn(n, cb, h, w, generator=g))
    ce = torch.randn(n, ca, generator=g); te = torch.randn(n, ca, generator=g)
    ar, br_, cer, ter = (t.clone().requires_grad_(True) for t in (a, b, ce, te))
    fa = cer[:, :, None, None] * ar + ter[:, :, None, None]
    up_ref = F.interpolate(torch.cat((fa, br_), 1), scale_factor=2, mode="bilinear", align_corners=True)
    dy = bf(torch.randn(up_ref.shape, generator=g))
    up_ref.backward(dy)
    ad, bd = nhwc(a, dev).requires_grad_(True), nhwc(b, dev).requires_grad_(True)
    ced, ted = ce.to(dev).requires_grad_(True), te.to(dev).requires_grad_(True)
    up = ops.upcat(ops.film(ad, ced, ted, ca), bd, ca, cb)
    assert rel(nchw(up, ca + cb), up_ref) < 6e-3
    up.backward(nhwc(dy, dev))
    assert rel(nchw(ad.grad, ca), ar.grad) < 8e-3
    assert rel(nchw(bd.grad, cb), br_.grad) < BF16_TOL
    assert rel(ced.grad.cpu(), cer.grad) < 8e-3
    assert rel(ted.grad.cpu(), ter.grad) < 8e-3
    # AvgPool2d(k)+GELU and MaxPool2d(2)
    x = bf(torch.randn(2, 16, 16, 16, generator=g))
    for k in (8, 2):
        xr = x.clone().requires_grad_(True)
        ref = F.gelu(F.avg_pool2d(xr, k))
        d = bf(torch.randn(ref.shape, generator=g))
        ref.backward(d)
        xd = nhwc(x, dev).requires_grad_(True)
        out = ops.avgpool_act(xd, 16, k, ops.ACT_GELU)
        assert rel(nchw(out, 16), ref) < BF16_TOL
        out.backward(nhwc(d, dev))
        assert rel(nchw(xd.grad, 16), xr.grad) < BF16_TOL
    xr = x.clone().requires_grad_(True)
    ref = F.max_pool2d(xr, 2)
    d = bf(torch.randn(ref.shape, generator=g))
    ref.backward(d)
    xd = nhwc(x, dev).requires_grad_(True)
    out = ops.maxpool2(xd, 16)
    assert torch.equal(nchw(out, 16), ref.detach())
    out.backward(nhwc(d, dev))
    assert torch.equal(nchw(xd.grad, 16), xr.grad)


@pytest.mark.parametrize("shape", [(2, 8, 8, 16, 24), (1, 5, 7, 8, 3), (3, 1, 1, 8, 8), (2, 16, 16, 192, 192), (1, 6, 4, 40, 20)],
                         ids=lambda s: "x".join(map(str, s)))
def test_upcat_quad_kernels_bit_identical_to_per_pixel_form(dev, shape):
    """The 2x2-quad upsample+cat kernels (one source block per output quad / 6x6 window per input block) must
    reproduce the one-thread-per-pixel kernels bit for bit, forward and backward, incl. odd sizes and ragged channels."""
    from diffusionmodel_b200 import ops, _lib
    n, h, w, ca, cb = shape
    g = torch.Generator().manual_seed(23)
    a = nhwc(torch.randn(n, ca, h, w, generator=g), dev)
    b = nhwc(torch.randn(n, cb, h, w, generator=g), dev)
    dy = None
    res = {}
    for mode in (1, 0):
        _lib.debug_set(9, mode)
        try:
            ad, bd = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
            up = ops.upcat(ad, bd, ca, cb)
            if dy is None:
                dy = torch.randn(up.shape, generator=g).to(torch.bfloat16).to(dev)
            up.backward(dy)
            torch.cuda.synchronize()
            res[mode] = (up.detach().clone(), ad.grad.clone(), bd.grad.clone())
        finally:
            _lib.debug_set(9, 0)
    for x, y in zip(res[0], res[1]):
        assert torch.equal(x, y)
    ref = F.interpolate(torch.cat((nchw(a, ca), nchw(b, cb)), 1), scale_factor=2, mode="bilinear", align_corners=True)
    assert rel(nchw(res[0][0], ca + cb), ref) < BF16_TOL


@pytest.mark.parametrize("case", [(4, 1, 1536, 1536, 1, 0, True), (4, 5, 768, 768, 1, 0, True), (30, 192, 12, 192, 1, 3, False),
                                  (4, 1536, 96, 1536, 1, 3, False), (130, 10, 256, 256, 1, 0, True), (3, 40, 7, 33, 2, 2, True),
                                  (4, 3080, 40, 64, 1, 0, True), (5, 1541, 9, 12, 1, 3, True), (12, 3072, 192, 3072, 1, 0, True)],
                         ids=lambda c: "x".join(map(str, c)))
def test_linear_mlp_kernels(dev, case):
    """dm_linear_act_fwd/bwd (SEBlock.fc, EmbedFC.model) against the same two-layer MLP in fp32 torch on the CPU:
    outputs, input gradient, and parameter gradients accumulated (+=) into pre-filled buffers."""
    from diffusionmodel_b200 import ops
    n, cin, hid, cout, act1, act2, bias = case
    g = torch.Generator().manual_seed(29)
    acts = {0: lambda v: v, 1: F.gelu, 2: F.relu, 3: torch.sigmoid}
    x = torch.randn(n, cin, generator=g)
    w1 = torch.randn(hid, cin, generator=g) / math.sqrt(cin)
    w2 = torch.randn(cout, hid, generator=g) / math.sqrt(hid)
    b1 = torch.randn(hid, generator=g) if bias else None
    b2 = torch.randn(cout, generator=g) if bias else None
    dy = torch.randn(n, cout, generator=g)
    ps = [t.clone().requires_grad_(True) for t in (x, w1, w2)] + [t.clone().requires_grad_(True) if t is not None else None for t in (b1, b2)]
    xr, w1r, w2r, b1r, b2r = ps
    yr = acts[act2](F.linear(acts[act1](F.linear(xr, w1r, b1r)), w2r, b2r))
    yr.backward(dy)
    d = [t.to(dev) if t is not None else None for t in (x, w1, b1, w2, b2)]
    xd, w1d, b1d, w2d, b2d = d
    pre = {}
    for name, t in (("w1", w1d), ("b1", b1d), ("w2", w2d), ("b2", b2d)):
        if t is not None:
            t.requires_grad_(True)
            t.grad = torch.randn(t.shape, generator=g).to(dev)
            pre[name] = t.grad.clone()
    y, saved = ops.mlp2_fwd(xd, w1d.detach(), None if b1d is None else b1d.detach(), w2d.detach(),
                            None if b2d is None else b2d.detach(), act1, act2)
    assert rel(y.cpu(), yr.detach()) < F32_TOL
    dx = ops.mlp2_bwd(saved, dy.to(dev), w1d, b1d, w2d, b2d, act1, act2, True)
    assert rel(dx.cpu(), xr.grad) < F32_TOL
    for name, t, r in (("w1", w1d, w1r), ("b1", b1d, b1r), ("w2", w2d, w2r), ("b2", b2d, b2r)):
        if t is not None:
            assert rel((t.grad - pre[name]).cpu(), r.grad) < F32_TOL, name
    # autograd wrapper without an input gradient (EmbedFC: t and the one-hot class need none)
    for t in (w1d, b1d, w2d, b2d):
        if t is not None:
            t.grad = None
    out = ops._Mlp2.apply(xd, act1, act2, w1d, b1d, w2d, b2d)
    out.backward(dy.to(dev))
    assert rel(w1d.grad.cpu(), w1r.grad) < F32_TOL and rel(w2d.grad.cpu(), w2r.grad) < F32_TOL


def test_local_enhancer_mask_bit_exact(dev):
    """(mask > 1.2) index set must be bit-exact incl. the threshold itself, its neighbours and NaN."""
    from diffusionmodel_b200 import ops
    g = torch.Generator().manual_seed(19)
    n, c, h, w = 2, 16, 8, 8
    x = bf(torch.randn(n, c, h, w, generator=g)); y = bf(torch.randn(n, c, h, w, generator=g))
    mask = torch.rand(n, h, w, generator=g) * 3
    mask[0, 0, 0] = 1.2
    mask[0, 0, 1] = torch.nextafter(torch.tensor(1.2), torch.tensor(2.0))
    mask[0, 0, 2] = torch.nextafter(torch.tensor(1.2), torch.tensor(0.0))
    mask[0, 0, 3] = float("nan")
    mask[0, 0, 4] = float("inf")
    high = (mask > 1.2).float().unsqueeze(1)
    ref = x + y * high
    xd, yd = nhwc(x, dev).requires_grad_(True), nhwc(y, dev).requires_grad_(True)
    out = ops.mask_fma(xd, yd, mask.to(dev), 1.2, c)
    got = nchw(out, c)
    assert torch.equal(got, bf(ref))
    # the index set itself: where y != 0 the output differs from x exactly on the high pixels
    sel = (got != x).any(1)
    assert torch.equal(sel, (high[:, 0] > 0) & (y != 0).any(1))
    d = bf(torch.randn(n, c, h, w, generator=g))
    out.backward(nhwc(d, dev))
    assert torch.equal(nchw(yd.grad, c), bf(d * high))
    assert torch.equal(nchw(xd.grad, c), d)


def test_q_sample_loss_reverse_step(dev):
    from diffusionmodel_b200 import ops
    g = torch.Generator().manual_seed(23)
    n, c, h, w, n_T = 4, 3, 16, 16, 700
    sched = P.ddpm_schedules(1e-4, 0.02, n_T)
    x = torch.rand(n, c, h, w, generator=g) * 2 - 1
    noise = torch.randn(n, c, h, w, generator=g)
    ts = torch.randint(1, n_T + 1, (n,), generator=g)
    xt_ref = P.q_sample(sched, x, ts, noise)
    xt = ops.q_sample(x.to(dev), noise.to(dev), sched["sqrtab"].to(dev), sched["sqrtmab"].to(dev), ts.to(dev))
    assert torch.equal(nchw(xt, c), bf(xt_ref))                     # bit-exact up to the bf16 store
    # weighted loss fwd/bwd, incl. exact-threshold mask values
    mask = P.synth_attn_mask(n, h, g)
    mask[0, 0, 0], mask[0, 0, 1] = 1.2, 0.8
    pred = torch.randn(n, c, h, w, generator=g)
    pr = pred.clone().requires_grad_(True)
    loss_ref = P.weighted_loss(noise, pr, mask)
    loss_ref.backward()
    pd = torch.zeros(n, h, w, 4)
    pd[..., :c] = pred.permute(0, 2, 3, 1)
    pd = pd.to(dev).requires_grad_(True)
    loss = ops.ddpm_loss(pd, noise.to(dev), mask.to(dev))
    assert abs(float(loss) - float(loss_ref)) < 2e-6 * abs(float(loss_ref)) + 1e-7
    loss.backward()
    assert rel(pd.grad[..., :c].cpu().permute(0, 3, 1, 2), pr.grad) < 1e-6
    # plain MSE (MNIST)
    pr2 = pred.clone().requires_grad_(True)
    l2 = F.mse_loss(noise, pr2); l2.backward()
    pd2 = pd.detach().clone().requires_grad_(True)
    l2d = ops.ddpm_loss(pd2, noise.to(dev), None)
    assert abs(float(l2d) - float(l2)) < 2e-6 * float(l2)
    l2d.backward()
    assert rel(pd2.grad[..., :c].cpu().permute(0, 3, 1, 2), pr2.grad) < 1e-6
    # CFG combine + reverse step: bit-exact in fp32 given the same eps
    ns = 2
    eps = torch.randn(2 * ns, c, h, w, generator=g)
    xi = torch.randn(ns, c, h, w, generator=g)
    z = torch.randn(ns, c, h, w, generator=g)
    for i, zz in ((350, z), (1, 0)):
        ref = P.reverse_step(sched, xi, eps[:ns], eps[ns:], zz, i, 2.0)
        ed = torch.zeros(2 * ns, h, w, 4)
        ed[..., :c] = eps.permute(0, 2, 3, 1)
        x_out, xt_next = ops.cfg_reverse_step(ed.to(dev), xi.to(dev), zz.to(dev) if i > 1 else None, 2.0,
                                              float(sched["oneover_sqrta"][i]), float(sched["mab_over_sqrtmab"][i]),
                                              float(sched["sqrt_beta_t"][i]))
        assert torch.equal(x_out.cpu(), ref)
        assert torch.equal(nchw(xt_next, c), bf(torch.cat([ref, ref], 0)))


def test_fused_adamw_matches_torch(dev):
    from diffusionmodel_b200 import FusedAdamW
    g = torch.Generator().manual_seed(29)
    shapes = [(33, 7), (5,), (16, 3, 3, 3), (1,)]
    ps = [torch.randn(s, generator=g) for s in shapes]
    ref = [torch.nn.Parameter(p.clone()) for p in ps]
    mine = [torch.nn.Parameter(p.clone().to(dev)) for p in ps]
    o_ref = torch.optim.AdamW(ref, lr=1e-2, weight_decay=1e-2)
    o = FusedAdamW(mine, lr=1e-2, weight_decay=1e-2, max_grad_norm=1.0)
    for _ in range(3):
        grads = [torch.randn(s, generator=g) for s in shapes]
        for p, gr in zip(ref, grads):
            p.grad = gr.clone()
        for p, gr in zip(mine, grads):
            p.grad.copy_(gr.to(dev))
        torch.nn.utils.clip_grad_norm_(ref, 1.0)
        o_ref.step(); o.step(); o.zero_grad()
    for a, b in zip(mine, ref):
        assert rel(a.detach().cpu(), b.detach()) < 1e-5


def test_cached_batch_prep_bit_exact(dev):
    """diffusionmodel_b200.data.CachedCrackBatches (dm_prep_batch) against the reference CrackDataset outputs in
    tests/golden/crack_items.npz: normalised images and attention masks bit-exact, flip decisions honoured, labels."""
    import os
    import numpy as np
    from diffusionmodel_b200 import data
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "crack_items.npz"))
    size = g["u8"].shape[1]
    boxes = torch.tensor([data.scale_box(*[int(v) for v in bw[:4]], int(bw[4]), int(bw[5]), size) for bw in g["box_wh"]],
                         dtype=torch.int32)
    cache = data.CachedCrackBatches(torch.from_numpy(g["u8"]), torch.from_numpy(g["label"]), boxes, dev)
    idx = torch.arange(len(cache))
    x, c, m = cache.batch(idx, flips=torch.from_numpy(g["flip"]))
    assert torch.equal(x.cpu(), torch.from_numpy(g["x"])) and torch.equal(m.cpu(), torch.from_numpy(g["mask"]))
    assert torch.equal(c.cpu(), torch.from_numpy(g["label"]).long())
    # default flips: torch.rand(B) < 0.5 from the host generator (what torchvision's RandomHorizontalFlip draws per sample)
    gen = torch.Generator().manual_seed(3)
    expect = torch.rand(len(cache), generator=torch.Generator().manual_seed(3)) < 0.5
    x2, _, m2 = cache.batch(idx, generator=gen)
    for i in range(len(cache)):
        ref = torch.from_numpy(g["x"][i]) if bool(expect[i]) == bool(g["flip"][i]) else torch.from_numpy(g["x"][i]).flip(2)
        assert torch.equal(x2[i].cpu(), ref)
    assert torch.equal(m2.cpu(), torch.from_numpy(g["mask"]))          # the reference never flips the mask
    # ragged request: repeated / out-of-order indices, empty batch
    x3, c3, _ = cache.batch([4, 0, 4], flips=[0, 1, 1])
    assert x3.shape[0] == 3 and int(c3[1]) == int(g["label"][0])
    x4, c4, m4 = cache.batch([], flips=[])
    assert x4.shape[0] == 0 and m4.shape[0] == 0


def test_image_metrics_ssim_psnr(dev):
    """dm_image_metrics against the reference's ImageMetrics values (golden) and the oracle on larger random pairs:
    per-image [-1,1] -> [0,1] mapping, identical pair -> +inf PSNR, evaluate_batch means."""
    import math
    import os
    import numpy as np
    from diffusionmodel_b200 import metrics
    from oracle import ref_port as P
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "image_metrics.npz"))
    a, b = torch.from_numpy(g["a"]), torch.from_numpy(g["b"])
    out = metrics.ssim_psnr(a.to(dev), b.to(dev)).cpu()
    for i in range(a.shape[0]):
        assert abs(float(out[i, 0]) - float(g["ssim"][i])) < 2e-5 * max(1.0, abs(float(g["ssim"][i]))), i
        if math.isinf(float(g["psnr"][i])):
            assert math.isinf(float(out[i, 1])) and float(out[i, 1]) > 0
        else:
            assert abs(float(out[i, 1]) - float(g["psnr"][i])) < 1e-4 * abs(float(g["psnr"][i])), i
    gen = torch.Generator().manual_seed(2)
    x = torch.rand(3, 3, 256, 256, generator=gen) * 2 - 1
    y = (x + 0.05 * torch.randn(x.shape, generator=gen)).clamp(-1, 1)
    out = metrics.ssim_psnr(x.to(dev), y.to(dev)).cpu()
    for i in range(3):
        assert abs(float(out[i, 0]) - float(P.calc_ssim(x[i], y[i]))) < 2e-5
        assert abs(float(out[i, 1]) - float(P.calc_psnr(x[i], y[i]))) < 1e-3
    m = metrics.evaluate_batch(x.to(dev), y.to(dev))
    assert abs(m["ssim"] - float(out[:, 0].double().mean())) < 1e-6 and abs(m["psnr"] - float(out[:, 1].double().mean())) < 1e-4
    assert metrics.evaluate_batch(x[:2].to(dev), y.to(dev)) == {}


def test_abi_empty_inputs_and_argument_errors(dev):
    """C-ABI edge behaviour: empty batches are a no-op that returns DM_OK and touches nothing; malformed arguments come
    back as a negative status with a message (raised as DmB200Error by the ctypes mirror), never as a launch."""
    from diffusionmodel_b200 import _lib, ops
    from diffusionmodel_b200._lib import DmB200Error
    P_, st = ops._p, ops._stream()
    c = 24
    y = torch.randn(2, 4, 4, c, device=dev).to(torch.bfloat16)
    z = torch.full_like(y, 7.0)
    ones, zeros = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    n0 = _lib.launch_count()
    assert ops.call("dm_bn_act_fwd", P_(y), c, P_(zeros), P_(ones), P_(ones), P_(zeros), P_(z), c, 0, c, 1, st) == 0
    assert ops.call("dm_upcat_fwd", P_(y), c, c, P_(y), c, c, P_(z), 2 * c, 0, 4, 4, st) == 0
    x = torch.randn(4, 8, device=dev)
    w = torch.randn(16, 8, device=dev)
    out = torch.full((4, 16), 7.0, device=dev)
    assert ops.call("dm_linear_act_fwd", P_(x), P_(w), None, None, P_(out), 0, 8, 16, 0, st) == 0
    torch.cuda.synchronize()
    assert _lib.launch_count() == n0                                     # nothing was launched
    assert float(z.float().min()) == 7.0 and float(out.min()) == 7.0      # and nothing written
    with pytest.raises(DmB200Error, match="dm_bn_act_fwd"):
        ops.call("dm_bn_act_fwd", P_(y), 12, P_(zeros), P_(ones), P_(ones), P_(zeros), P_(z), c, 32, c, 1, st)   # pitch % 8
    with pytest.raises(DmB200Error, match="dm_bn_stats"):
        ops.call("dm_bn_stats", P_(y), c, P_(zeros), c, 0, c, st)
    with pytest.raises(DmB200Error, match="act must be"):
        ops.call("dm_linear_act_fwd", P_(x), P_(w), None, None, P_(out), 4, 8, 16, 7, st)
    a = torch.zeros(17, 64, device=dev, dtype=torch.bfloat16)
    with pytest.raises(DmB200Error, match="M must be"):
        ops.call("dm_skinny_gemm", P_(a), 64, P_(a), 64, P_(a), 64, P_(x), 17, 17, 64, st)
    with pytest.raises(DmB200Error, match="divisible"):
        ops.call("dm_gn_act_fwd", P_(y), c, P_(ones), P_(zeros), P_(z), c, P_(zeros), P_(zeros), P_(zeros), 2, 16, c, 7, 1e-5, 1, st)
    assert _lib.launch_count() == n0


def test_linear_fwd_scalar_and_vector_paths_agree(dev):
    """dm_linear_act_fwd takes 16-byte weight loads when Cin % 4 == 0 and both operands are 16-byte aligned, scalar loads
    otherwise: the same matrix at an aligned and at a 4-byte-offset address must give the same rows (fp32 tolerance:
    the two paths add in a different order), and both must match torch."""
    from diffusionmodel_b200 import ops
    P_, st = ops._p, ops._stream()
    g = torch.Generator().manual_seed(31)
    for n, cin, cout in ((4, 1536, 96), (30, 200, 40), (9, 772, 17)):
        x = torch.randn(n, cin, generator=g)
        w = torch.randn(cout, cin, generator=g) / math.sqrt(cin)
        b = torch.randn(cout, generator=g)
        ref = F.gelu(F.linear(x, w, b))
        xd, bd = x.to(dev), b.to(dev)
        buf = torch.zeros(cout * cin + 1, device=dev)
        outs = []
        for off in (0, 1):
            wd = buf[off:off + cout * cin].view(cout, cin)
            wd.copy_(w)
            assert (wd.data_ptr() % 16 == 0) == (off == 0)
            pre, yv = torch.empty(n, cout, device=dev), torch.empty(n, cout, device=dev)
            ops.call("dm_linear_act_fwd", P_(xd), P_(wd), P_(bd), P_(pre), P_(yv), n, cin, cout, 1, st)
            outs.append(yv.cpu())
            assert rel(yv.cpu(), ref) < F32_TOL
        assert rel(outs[0], outs[1]) < 1e-5


def test_fid_matches_reference_value(dev):
    """metrics.calc_fid / evaluate_batch (ImageMetrics.calc_fid, new_scripy.py:1146-1187) with the golden's stub feature
    network: batching, batch-level range decision, resize and the Frechet distance against the reference's value."""
    import os
    import numpy as np
    from diffusionmodel_b200 import metrics
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "fid.npz"))
    proj = P.fid_stub_projection(int(g["proj_seed"])).to(dev)
    real, gen = torch.from_numpy(g["real"]).to(dev), torch.from_numpy(g["gen"]).to(dev)
    feature_fn = lambda x: x.flatten(1) @ proj
    fid = metrics.calc_fid(real, gen, feature_fn, batch_size=8)
    assert fid == pytest.approx(float(g["fid"]), rel=1e-4, abs=1e-5)
    fd = metrics.frechet_distance(torch.from_numpy(g["feats_real"]).to(dev), torch.from_numpy(g["feats_gen"]).to(dev))
    assert fd == pytest.approx(float(g["fid"]), rel=1e-6, abs=1e-8)          # eigenvalue route vs scipy sqrtm: 4e-8
    m = metrics.evaluate_batch(real, gen, feature_fn=feature_fn)
    assert set(m) == {"fid", "ssim", "psnr"} and m["fid"] == pytest.approx(fid)
    m = metrics.evaluate_batch(real, gen)                     # no feature network: the reference's failure value
    assert math.isnan(m["fid"]) and "ssim" in m
    assert "fid" not in metrics.evaluate_batch(real[:4], gen[:4])         # fewer than 10 samples: no FID (:1266)


def test_shared_skip_upcat_and_shared_second_source_conv(dev):
    """The sampling loop's CFG halves share the encoder's skip tensors: dm_upcat_fwd_shared and the dual-source conv with a
    second source of fewer samples must equal the same ops on an explicitly repeated tensor, bit for bit."""
    from diffusionmodel_b200 import ops
    g = torch.Generator().manual_seed(12)
    n, nb, h, w, ca, cb = 6, 3, 8, 8, 16, 24
    a = nhwc(bf(torch.randn(n, ca, h, w, generator=g)), dev)
    b = nhwc(bf(torch.randn(nb, cb, h, w, generator=g)), dev)
    with torch.no_grad():
        shared = ops.upcat(a, b, ca, cb)
        full = ops.upcat(a, torch.cat([b, b], 0), ca, cb)
    assert shared.shape == full.shape and torch.equal(shared, full)
    cout, c0, c1 = 16, 16, 24
    x0 = nhwc(bf(torch.randn(n, c0, h, w, generator=g)), dev)
    wt = torch.nn.Parameter((torch.randn(cout, c0 + c1, 3, 3, generator=g) / 19).to(dev))
    bias = torch.nn.Parameter(torch.randn(cout, generator=g).to(dev) * 0.1)
    with torch.no_grad():
        ys, _ = ops.conv2d(x0, wt, bias, ops.WeightPack(), x1=b, c1=c1, stride=1, pad=1)
        yf, _ = ops.conv2d(x0, wt, bias, ops.WeightPack(), x1=torch.cat([b, b], 0), c1=c1, stride=1, pad=1)
    assert torch.equal(ys, yf)
    from diffusionmodel_b200._lib import DmB200Error
    with pytest.raises(DmB200Error):                       # forward-only
        ops.upcat(a.clone().requires_grad_(True), b, ca, cb)
    with pytest.raises(DmB200Error):                       # the skip batch must tile the batch
        with torch.no_grad():
            ops.upcat(a, b[:2].contiguous()[:, :, :, :].repeat(2, 1, 1, 1)[:4], ca, cb)
