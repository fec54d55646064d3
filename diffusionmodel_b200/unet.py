"""ContextUnet denoisers on the sm_100a kernels: drop-in mirrors of the reference modules.

``ContextUnet`` (enhanced: CoordAttn + SE + LocalEnhancer) mirrors new_scripy.py:270-356 and
``MnistContextUnet`` mirrors MNIST_script.py:119-187.  Constructor signatures, attribute names and the
full ``state_dict`` layout (parameter / buffer names, NCHW fp32 shapes, construction order and
therefore default initialisation under a given seed) are the reference's; only ``forward`` differs:
it runs NHWC bf16 activations through the kernels of libdm_b200.so via ``ops``.

The ``nn.Conv2d`` / ``nn.BatchNorm2d`` / ... children are used as parameter containers only -- their
own ``forward`` (aten/cuDNN) is never called on the hot path, and there is no CPU fallback.
"""
from __future__ import annotations

import weakref

import torch
import torch.nn as nn

from . import ops
from .ops import ACT_GELU, ACT_NONE, ACT_RELU

RES_SCALE = 1.0 / 1.414          # the reference divides by 1.414, not sqrt(2) (new_scripy.py:205)
HIGH_THRESH = 1.2                # Cfg.HIGH_THRESH (new_scripy.py:31)


def _pack_of(conv):
    pk = conv.__dict__.get("_dm_pack")
    if pk is None:
        pk = ops.WeightPack()
        conv.__dict__["_dm_pack"] = pk
    return pk


def conv(x, m, *, x1=None, c1=0, want_stats=False, out_f32=False, bias_grad_by_norm=False, add_bias=True):
    """Run an nn.Conv2d container through the implicit-GEMM kernel."""
    return ops.conv2d(x, m.weight, m.bias, _pack_of(m), x1=x1, c1=c1, stride=m.stride[0], pad=m.padding[0],
                      want_stats=want_stats, out_f32=out_f32, bias_grad_by_norm=bias_grad_by_norm, add_bias=add_bias)


def _folded_bn(cv, bn):
    """Eval-mode BatchNorm folded onto the conv output: y*scale + shift with the running statistics
    (constants while sampling).  Cached per module until a parameter or buffer changes."""
    ts = [bn.weight, bn.bias, bn.running_mean, bn.running_var] + ([cv.bias] if cv.bias is not None else [])
    stamp = tuple(t._version for t in ts) + tuple(t.data_ptr() for t in ts) + (ops._weights_epoch, ops._bn_stats_epoch)
    hit = bn.__dict__.get("_dm_fold")
    if hit is not None and hit[0] == stamp:
        return hit[1], hit[2]
    with torch.no_grad():
        scale = (bn.weight.float() * torch.rsqrt(bn.running_var.float() + bn.eps)).contiguous()
        shift = bn.bias.float() - bn.running_mean.float() * scale
        if cv.bias is not None:
            shift = shift + cv.bias.float() * scale
        if hit is not None and hit[1].shape == scale.shape and hit[1].device == scale.device:
            hit[1].copy_(scale); hit[2].copy_(shift)      # in place: captured sampling graphs keep their pointers
            scale, shift = hit[1], hit[2]
        else:
            shift = shift.contiguous()
    bn.__dict__["_dm_fold"] = (stamp, scale, shift)
    _FOLDED[bn] = cv
    return scale, shift


_FOLDED = weakref.WeakKeyDictionary()        # BatchNorm -> its conv, for every folded pair seen so far


def refresh_folded_norms():
    """Bring every cached folded scale/shift up to date IN PLACE (captured sampling graphs read them)."""
    for bn, cv in list(_FOLDED.items()):
        _folded_bn(cv, bn)


def conv_bn_act(x, seq, act=ACT_GELU, **kw):
    """Sequential(Conv2d, BatchNorm2d, GELU): conv with fused statistics, then one normalise+activate pass
    (training / autograd); in no-grad eval mode the norm and activation ride in the conv epilogue."""
    cv, bn = seq[0], seq[1]
    if not bn.training and not torch.is_grad_enabled():
        scale, shift = _folded_bn(cv, bn)
        return ops.conv2d_fused_eval(x, cv.weight, _pack_of(cv), shift, scale, act, stride=cv.stride[0],
                                     pad=cv.padding[0], **kw)
    # train mode: batch normalisation cancels the conv bias exactly, so the GEMM epilogue skips it
    y, stats = conv(x, cv, want_stats=bn.training and ops.conv_takes_stats(cv.out_channels), bias_grad_by_norm=cv.bias is not None,
                    add_bias=not bn.training, **kw)
    return ops.bn_act(y, stats, bn, act, conv_bias=cv.bias, bias_outside=bn.training and cv.bias is not None)


# ------------------------------------------------------------------------------------------ blocks
class SEBlock(nn.Module):
    """Parameter container for the squeeze-excite MLP (new_scripy.py:143-158)."""

    def __init__(self, channels, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(
            nn.Linear(channels, channels // reduction, bias=False), nn.GELU(),
            nn.Linear(channels // reduction, channels, bias=False), nn.Sigmoid())


class ResConvBlock(nn.Module):
    """Two conv-BN-GELU units; residual variants add the input (or the first unit's output when the
    channel count changes), optionally through an SE gate, and scale by 1/1.414
    (new_scripy.py:176-209; MNIST_script.py:31-65 with ``use_se=False``)."""

    def __init__(self, in_ch, out_ch, is_res=False, use_se=True):
        super().__init__()
        self.same_ch = in_ch == out_ch
        self.is_res = is_res
        self.out_ch = out_ch
        self.conv1 = nn.Sequential(nn.Conv2d(in_ch, out_ch, 3, 1, 1), nn.BatchNorm2d(out_ch), nn.GELU())
        self.conv2 = nn.Sequential(nn.Conv2d(out_ch, out_ch, 3, 1, 1), nn.BatchNorm2d(out_ch), nn.GELU())
        if use_se:
            self.se = SEBlock(out_ch) if is_res else None

    def forward(self, x):
        c = self.out_ch
        if not self.is_res:
            return conv_bn_act(conv_bn_act(x, self.conv1), self.conv2)
        if self.same_ch:
            x, res = ops.fork(x, c)
            x1 = conv_bn_act(x, self.conv1)
        else:
            x1, res = ops.fork(conv_bn_act(x, self.conv1), c)
        x2 = conv_bn_act(x1, self.conv2)
        se = getattr(self, "se", None)
        return ops.se_residual(x2, res, c, RES_SCALE, se.fc if se is not None else None)


class CoordAttn(nn.Module):
    """Coordinate attention (new_scripy.py:70-140): row/column mean pooling and the final gating pass are bandwidth
    kernels; the C/16-wide gate network (two 1x1 convs + BatchNorm + GELU, the h<->w cross projections, two 1x1 convs +
    sigmoid) runs in fp32 on the dm_ca_gates kernels.  Square feature maps only."""

    def __init__(self, channel, reduction=16):
        super().__init__()
        mid = channel // reduction
        self.channel = channel
        self.pool_h = nn.AdaptiveAvgPool2d((None, 1))
        self.pool_w = nn.AdaptiveAvgPool2d((1, None))
        self.conv1_h = nn.Conv2d(channel, mid, kernel_size=1)
        self.conv1_w = nn.Conv2d(channel, mid, kernel_size=1)
        self.bn1_h = nn.BatchNorm2d(mid)
        self.bn1_w = nn.BatchNorm2d(mid)
        self.act = nn.GELU()
        self.h2w_proj = nn.Conv2d(mid, mid, kernel_size=1)
        self.w2h_proj = nn.Conv2d(mid, mid, kernel_size=1)
        self.gamma_h = nn.Parameter(torch.zeros(1))
        self.gamma_w = nn.Parameter(torch.zeros(1))
        self.conv_h = nn.Conv2d(mid, channel, kernel_size=1)
        self.conv_w = nn.Conv2d(mid, channel, kernel_size=1)
        self.sigmoid = nn.Sigmoid()
        self.alpha = nn.Parameter(torch.zeros(1))
        self.beta = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        return ops.coord_attn(x, self.channel, self)


class LocalEnhancer(nn.Module):
    """x + conv(GELU(GN(conv(x)))) * (mask > thresh) (new_scripy.py:161-174)."""

    def __init__(self, in_ch, high_thresh=HIGH_THRESH):
        super().__init__()
        self.high_thresh = high_thresh
        self.in_ch = in_ch
        self.conv = nn.Sequential(nn.Conv2d(in_ch, in_ch, kernel_size=3, padding=1), nn.GroupNorm(8, in_ch),
                                  nn.GELU(), nn.Conv2d(in_ch, in_ch, kernel_size=3, padding=1))

    def forward(self, x, mask):
        x, res = ops.fork(x, self.in_ch)
        y, _ = conv(x, self.conv[0])
        y = ops.gn_act(y, self.conv[1], ACT_GELU)
        y, _ = conv(y, self.conv[3])
        return ops.mask_fma(res, y, mask, float(self.high_thresh), self.in_ch)


class UnetDown(nn.Module):
    """1x1 compress + BN + GELU, 1x1 expand, 3x3 + BN + GELU, residual SE block, 4x4 stride-2 conv
    (new_scripy.py:211-235)."""

    def __init__(self, in_ch, out_ch, compress_ratio=4):
        super().__init__()
        mid = in_ch // compress_ratio
        self.channel_compress = nn.Sequential(nn.Conv2d(in_ch, mid, 1), nn.BatchNorm2d(mid), nn.GELU())
        self.ch_adjust = nn.Conv2d(mid, out_ch, 1)
        self.down = nn.Sequential(nn.Conv2d(out_ch, out_ch, 3, padding=1), nn.BatchNorm2d(out_ch), nn.GELU(),
                                  ResConvBlock(out_ch, out_ch, is_res=True),
                                  nn.Conv2d(out_ch, out_ch, 4, stride=2, padding=1))

    def forward(self, x):
        x = conv_bn_act(x, self.channel_compress)
        x, _ = conv(x, self.ch_adjust)
        x = conv_bn_act(x, self.down)          # children 0,1 are the conv and its BatchNorm
        x = self.down[3](x)
        y, _ = conv(x, self.down[4])
        return y


class UnetUp(nn.Module):
    """cat -> bilinear x2 -> 3x3 conv -> two plain conv blocks (new_scripy.py:237-253)."""

    def __init__(self, in_ch, out_ch):
        super().__init__()
        self.model = nn.Sequential(
            nn.Sequential(nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True),
                          nn.Conv2d(in_ch, out_ch, 3, padding=1)),
            ResConvBlock(out_ch, out_ch), ResConvBlock(out_ch, out_ch))

    def forward(self, x, skip, cx, cs):
        u = ops.upcat(x, skip, cx, cs)
        y, _ = conv(u, self.model[0][1])
        return self.model[2](self.model[1](y))


class EmbedFC(nn.Module):
    """Linear-GELU-Linear on a scalar / one-hot input (new_scripy.py:255-268): [N, <=n_classes] fp32 rows through
    the weight-stationary dm_linear_act_fwd/bwd kernels (2 launches forward, 2 backward)."""

    def __init__(self, input_dim, emb_dim):
        super().__init__()
        self.input_dim = input_dim
        self.model = nn.Sequential(nn.Linear(input_dim, emb_dim), nn.GELU(), nn.Linear(emb_dim, emb_dim))

    def forward(self, x):
        return ops.embed_fc(x.view(-1, self.input_dim), self.model[0], self.model[2])


def _head(x, x0, seq, c_each):
    """conv3x3(cat(x, x0)) + GroupNorm + ReLU + conv3x3 -> fp32 NHWC (new_scripy.py:310-315,355); the
    concat is never materialised: the conv kernel reads both sources."""
    y, _ = conv(x, seq[0], x1=x0, c1=c_each)
    y = ops.gn_act(y, seq[1], ACT_RELU)
    y, _ = conv(y, seq[3], out_f32=True)
    return y


# ------------------------------------------------------------------------------------------ denoisers
class ContextUnet(nn.Module):
    """Enhanced context U-Net (new_scripy.py:270-356).  ``forward(x, c, t, ctx_mask)`` keeps the
    reference signature; the optional ``attn_map`` ([B,H,W]) feeds LocalEnhancer the attention map the
    reference call site meant to pass (new_scripy.py:353 passes ctx_mask, which contributes exactly 0):
    without it the shipped behaviour (+0) is reproduced."""

    variant = "rdd"

    def __init__(self, in_ch=3, n_feat=192, n_classes=10):
        super().__init__()
        self.in_ch, self.n_feat, self.n_classes = in_ch, n_feat, n_classes
        f = n_feat
        self.init_conv = ResConvBlock(in_ch, f, is_res=True)
        self.down1 = UnetDown(f, f)
        self.down2 = UnetDown(f, 2 * f)
        self.down3 = UnetDown(2 * f, 4 * f)
        self.down4 = UnetDown(4 * f, 8 * f)
        self.ca1 = CoordAttn(f)
        self.ca2 = CoordAttn(2 * f)
        self.ca3 = CoordAttn(4 * f)
        self.ca4 = CoordAttn(8 * f)
        self.to_vec = nn.Sequential(nn.AvgPool2d(8), nn.GELU())
        self.time_emb1 = EmbedFC(1, 8 * f)
        self.time_emb2 = EmbedFC(1, 4 * f)
        self.ctx_emb1 = EmbedFC(n_classes, 8 * f)
        self.ctx_emb2 = EmbedFC(n_classes, 4 * f)
        self.up0 = nn.Sequential(nn.ConvTranspose2d(8 * f, 8 * f, 8, 8), nn.GroupNorm(8, 8 * f), nn.ReLU())
        self.up1 = UnetUp(16 * f, 4 * f)
        self.up2 = UnetUp(8 * f, 2 * f)
        self.up3 = UnetUp(4 * f, f)
        self.up4 = UnetUp(2 * f, f)
        self.local_enhance = LocalEnhancer(f)
        self.out = nn.Sequential(nn.Conv2d(2 * f, f, 3, 1, 1), nn.GroupNorm(8, f), nn.ReLU(),
                                 nn.Conv2d(f, self.in_ch, 3, 1, 1))
        ops.mark_conv2d_weights(self)

    def trunk_front(self, x):
        """init_conv ... down3 with CoordAttn (new_scripy.py:318-326): the first three skip tensors and ``d3_in``, the
        tensor that enters down4.  x: bf16 NHWC."""
        f = self.n_feat
        x0, x0_skip = ops.fork(self.init_conv(x), f)
        d1, d1s = ops.fork(self.ca1(self.down1(x0)), f)
        d2, d2s = ops.fork(self.ca2(self.down2(d1)), 2 * f)
        d3, d3s = ops.fork(self.ca3(self.down3(d2)), 4 * f)
        return dict(x0=x0_skip, d1=d1s, d2=d2s, d3=d3s, d3_in=d3)

    def trunk_back(self, d3_in):
        """down4 + CoordAttn + to_vec (new_scripy.py:327-332): 101 M of the trunk's 135 M parameters."""
        f = self.n_feat
        d4, d4s = ops.fork(self.ca4(self.down4(d3_in)), 8 * f)
        return dict(d4=d4s, hidden=ops.avgpool_act(d4, 8 * f, 8, ACT_GELU))

    def trunk(self, x):
        """Everything up to the pooled bottleneck: the skip tensors and ``hidden``.  Every parameter used here precedes
        ``time_emb1`` in ``parameters()`` order."""
        t = self.trunk_front(x)
        t.update(self.trunk_back(t.pop("d3_in")))
        return t

    def grad_ready_regions(self):
        """For a backward pass run in three stages (decoder side; trunk_back; trunk_front): after stage k the gradients of
        the parameters in ``[first, end)`` -- given as (first parameter, first parameter after the region or None) -- are
        final.  Stage 0: time_emb1 ... out (62 % of the bytes); stage 1: down4 (29 %)."""
        return [(next(self.time_emb1.parameters()), None), (next(self.down4.parameters()), next(self.ca1.parameters()))]

    def first_decoder_param(self):
        """First parameter (in ``parameters()`` order) that ``trunk`` does not use: everything from here on belongs to the
        decoder side (embeddings, up0, up blocks, LocalEnhancer, head)."""
        return next(self.time_emb1.parameters())

    def up0_of(self, hidden):
        """up0 = ConvTranspose2d(8f, 8f, 8, 8) + GroupNorm + ReLU (new_scripy.py:297-301,347)."""
        u1 = ops.conv_transpose(hidden, self.up0[0].weight, self.up0[0].bias, _pack_of(self.up0[0]), 8)
        return ops.gn_act(u1, self.up0[1], ACT_RELU)

    def encode(self, x):
        """Everything that does not depend on (c, t, ctx_mask): trunk + up0 (new_scripy.py:318-332,347).  x: bf16 NHWC."""
        enc = self.trunk(x)
        enc["u1"] = self.up0_of(enc.pop("hidden"))
        return enc

    def decode(self, enc, c, t, ctx_mask, attn_map=None):
        """Embeddings, FiLM, up1..up4, LocalEnhancer, head (new_scripy.py:334-355) -> eps as fp32 NHWC."""
        f = self.n_feat
        c1h = ops.ctx_onehot(c, ctx_mask, self.n_classes, flip=False)             # no flip (new_scripy.py:337-340)
        t = t.to(torch.float32)
        cemb1, temb1 = self.ctx_emb1(c1h), self.time_emb1(t)
        cemb2, temb2 = self.ctx_emb2(c1h), self.time_emb2(t)
        u2 = self.up1(ops.film(enc["u1"], cemb1, temb1, 8 * f), enc["d4"], 8 * f, 8 * f)
        u3 = self.up2(ops.film(u2, cemb2, temb2, 4 * f), enc["d3"], 4 * f, 4 * f)
        u4 = self.up3(u3, enc["d2"], 2 * f, 2 * f)
        u5 = self.up4(u4, enc["d1"], f, f)
        if attn_map is not None:
            u5 = self.local_enhance(u5, attn_map)
        # attn_map None: the shipped call site adds conv(...) * 0 -- identical output, no side effects
        return _head(u5, enc["x0"], self.out, f)

    def forward_nhwc(self, x, c, t, ctx_mask, attn_map=None):
        """x: bf16 NHWC; returns eps as fp32 NHWC (pitch 4 for 3 channels)."""
        with ops.batched_counters(self):
            return self.decode(self.encode(x), c, t, ctx_mask, attn_map)

    def forward(self, x, c, t, ctx_mask, attn_map=None):
        if x.shape[2] % 128 or x.shape[3] % 128:
            raise RuntimeError("ContextUnet: image size must be a multiple of 128 (AvgPool2d(8) after four halvings)")
        y = self.forward_nhwc(ops.to_nhwc(x), c, t, ctx_mask, attn_map)
        return ops.to_nchw_f32(y, self.in_ch)


class MnistUnetDown(nn.Module):
    """ResidualConvBlock + MaxPool2d(2) (MNIST_script.py:68-78)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.out_channels = out_channels
        self.model = nn.Sequential(ResConvBlock(in_channels, out_channels, use_se=False), nn.MaxPool2d(2))

    def forward(self, x):
        return ops.maxpool2(self.model[0](x), self.out_channels)


class MnistUnetUp(nn.Module):
    """cat -> ConvTranspose2d(2,2) -> two conv blocks (MNIST_script.py:81-97)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.model = nn.Sequential(nn.ConvTranspose2d(in_channels, out_channels, 2, 2),
                                   ResConvBlock(out_channels, out_channels, use_se=False),
                                   ResConvBlock(out_channels, out_channels, use_se=False))

    def forward(self, x, skip, cx, cs):
        u = _Cat.apply(x, skip, cx, cs)      # materialised concat (small tensors at 7x7 / 14x14)
        ct = self.model[0]
        y = ops.conv_transpose(u, ct.weight, ct.bias, _pack_of(ct), 2)
        return self.model[2](self.model[1](y))


class _Cat(torch.autograd.Function):
    """Channel concat of two NHWC activations (cx multiple of 8) through the axpby kernel."""

    @staticmethod
    def forward(ctx, a, b, ca, cb):
        n, h, w, _ = a.shape
        out = torch.empty((n, h, w, ca + ops.r8(cb)), device=a.device, dtype=torch.bfloat16)
        st = ops._stream()
        ops.call("dm_axpby", ops._p(a), a.stride(2), None, 0, ops._p(out), out.stride(2), n * h * w, ca, 1.0, 0.0, st)
        ops.call("dm_axpby", ops._p(b), b.stride(2), None, 0, ops._p(out[..., ca:]), out.stride(2), n * h * w, cb, 1.0, 0.0, st)
        ctx.cfg = (ca, cb, a.shape[3], b.shape[3])
        return out

    @staticmethod
    def backward(ctx, g):
        ca, cb, wa, wb = ctx.cfg
        ga = g[..., :ca]
        gb = g[..., ca:]
        if wa != ca or wb != gb.shape[3]:
            raise RuntimeError("cat backward: unexpected padded operand widths")
        return ga, gb, None, None


class MnistContextUnet(nn.Module):
    """Original minDiffusion-style context U-Net (MNIST_script.py:119-187)."""

    variant = "mnist"

    def __init__(self, in_channels, n_feat=256, n_classes=10):
        super().__init__()
        self.in_channels, self.n_feat, self.n_classes = in_channels, n_feat, n_classes
        f = n_feat
        self.init_conv = ResConvBlock(in_channels, f, is_res=True, use_se=False)
        self.down1 = MnistUnetDown(f, f)
        self.down2 = MnistUnetDown(f, 2 * f)
        self.to_vec = nn.Sequential(nn.AvgPool2d(7), nn.GELU())
        self.timeembed1 = EmbedFC(1, 2 * f)
        self.timeembed2 = EmbedFC(1, 1 * f)
        self.contextembed1 = EmbedFC(n_classes, 2 * f)
        self.contextembed2 = EmbedFC(n_classes, 1 * f)
        self.up0 = nn.Sequential(nn.ConvTranspose2d(2 * f, 2 * f, 7, 7), nn.GroupNorm(8, 2 * f), nn.ReLU())
        self.up1 = MnistUnetUp(4 * f, f)
        self.up2 = MnistUnetUp(2 * f, f)
        self.out = nn.Sequential(nn.Conv2d(2 * f, f, 3, 1, 1), nn.GroupNorm(8, f), nn.ReLU(),
                                 nn.Conv2d(f, self.in_channels, 3, 1, 1))
        ops.mark_conv2d_weights(self)

    @property
    def in_ch(self):
        return self.in_channels

    def trunk(self, x):
        """init_conv, down1, down2, to_vec (MNIST_script.py:157-161): independent of (c, t, mask)."""
        f = self.n_feat
        x0, x0s = ops.fork(self.init_conv(x), f)
        d1, d1s = ops.fork(self.down1(x0), f)
        d2, d2s = ops.fork(self.down2(d1), 2 * f)
        hidden = ops.avgpool_act(d2, 2 * f, 7, ACT_GELU)
        return dict(x0=x0s, d1=d1s, d2=d2s, hidden=hidden)

    def first_decoder_param(self):
        return next(self.timeembed1.parameters())

    def grad_ready_regions(self):
        return [(next(self.timeembed1.parameters()), None)]

    def up0_of(self, hidden):
        """up0 = ConvTranspose2d(2f, 2f, 7, 7) + GroupNorm + ReLU (MNIST_script.py:139-144,176)."""
        u1 = ops.conv_transpose(hidden, self.up0[0].weight, self.up0[0].bias, _pack_of(self.up0[0]), 7)
        return ops.gn_act(u1, self.up0[1], ACT_RELU)

    def encode(self, x):
        enc = self.trunk(x)
        enc["u1"] = self.up0_of(enc.pop("hidden"))
        return enc

    def decode(self, enc, c, t, context_mask, attn_map=None):
        f = self.n_feat
        c1h = ops.ctx_onehot(c, context_mask, self.n_classes, flip=True)          # flip and negate (MNIST_script.py:170)
        t = t.to(torch.float32)
        cemb1, temb1 = self.contextembed1(c1h), self.timeembed1(t)
        cemb2, temb2 = self.contextembed2(c1h), self.timeembed2(t)
        u2 = self.up1(ops.film(enc["u1"], cemb1, temb1, 2 * f), enc["d2"], 2 * f, 2 * f)
        u3 = self.up2(ops.film(u2, cemb2, temb2, f), enc["d1"], f, f)
        return _head(u3, enc["x0"], self.out, f)

    def forward_nhwc(self, x, c, t, context_mask, attn_map=None):
        with ops.batched_counters(self):
            return self.decode(self.encode(x), c, t, context_mask)

    def forward(self, x, c, t, context_mask):
        y = self.forward_nhwc(ops.to_nhwc(x), c, t, context_mask)
        return ops.to_nchw_f32(y, self.in_channels)
