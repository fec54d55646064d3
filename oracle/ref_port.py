"""CPU oracle: a functional restatement of the reference diffusion hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``diffusionmodel_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and only as the checker or the
timed CPU baseline -- never as the product path.

What it restates (file:line under /root/reference):
  * ``ddpm_schedules``                         new_scripy.py:358-384 == MNIST_script.py:190-216
  * enhanced ``ContextUnet.forward``           new_scripy.py:317-356  (variant "rdd")
  * original ``ContextUnet.forward``           MNIST_script.py:155-187 (variant "mnist")
  * ``DDPM.forward`` (weighted loss / MSE)     new_scripy.py:401-439 / MNIST_script.py:234-252
  * ``DDPM.sample`` CFG reverse loop           new_scripy.py:441-477 / MNIST_script.py:254-300

All arithmetic lives in PyTorch (un-pinned dependency of the reference; this
container has torch 2.11.0).  The port is written as plain functions over a
``state_dict`` (reference key names) using the same ``torch.nn.functional`` calls
the reference's ``nn.Module``s dispatch to, so on CPU it is bit-identical to the
imported reference; ``oracle/make_golden.py`` asserts exactly that and writes the
fixtures under ``tests/golden/`` (the reference ships no tests or golden vectors
of its own -- parity is pinned by those generated fixtures).

The one deliberate deviation: ``new_scripy.py:353`` passes ``ctx_mask`` where an
attention map is intended (it only broadcasts for degenerate shapes and then
contributes exactly +0).  ``attn_map=None`` reproduces the shipped behaviour
(LocalEnhancer adds 0); passing a ``[B,H,W]`` map exercises the intended path.

``operand_dtype=torch.bfloat16`` rounds the inputs and weights of every
conv / conv-transpose / linear to bf16 (fp32 accumulate, everything else fp32):
the precision-matched oracle of SURVEY.md Appendix D.

``store_dtype=torch.bfloat16`` (with ``operand_dtype=torch.bfloat16``) is the KERNEL-MATCHED oracle: the same
algorithm with a rounding wherever the B200 path keeps a tensor in bf16 between two kernels -- the output of every
tensor-core convolution, of every norm+activation pass, of the SE / CoordAttn / FiLM / upsample / mask passes, x_t --
in the forward pass and, at the same tensors, on the gradient in the backward pass; the [N, C]-sized layers the
kernels run in fp32 (SE and EmbedFC linears, the CoordAttn gate network) keep fp32 operands; a convolution feeding a
batch-statistics BatchNorm is stored WITHOUT its bias (the norm cancels it).  Train-mode
BatchNorm amplifies every rounding difference with depth (Appendix D), so only an oracle that rounds where the kernels
round can be compared tightly at full depth; with it the remaining differences are fp32 summation order.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

HIGH_THRESH = 1.2   # new_scripy.py:31
MID_THRESH = 0.8    # new_scripy.py:32
HIGH_WEIGHT = 3.0   # new_scripy.py:33
MID_WEIGHT = 1.0    # new_scripy.py:34
LOW_WEIGHT = 0.5    # new_scripy.py:35
FEAT_CONSIST_WEIGHT = 2.0  # new_scripy.py:36


# --------------------------------------------------------------------------- schedules
def ddpm_schedules(beta1: float, beta2: float, T: int) -> dict:
    """new_scripy.py:358-384 -- seven fp32 tables with T+1 entries."""
    assert beta1 < beta2 < 1.0
    beta_t = (beta2 - beta1) * torch.arange(0, T + 1, dtype=torch.float32) / T + beta1
    alpha_t = 1 - beta_t
    alphabar_t = torch.cumsum(torch.log(alpha_t), dim=0).exp()
    sqrtmab = torch.sqrt(1 - alphabar_t)
    return {
        "alpha_t": alpha_t,
        "oneover_sqrta": 1 / torch.sqrt(alpha_t),
        "sqrt_beta_t": torch.sqrt(beta_t),
        "alphabar_t": alphabar_t,
        "sqrtab": torch.sqrt(alphabar_t),
        "sqrtmab": sqrtmab,
        "mab_over_sqrtmab": (1 - alpha_t) / sqrtmab,
    }


# --------------------------------------------------------------------------- primitives
class _StoreRound(torch.autograd.Function):
    """A tensor kept in a narrow dtype between two kernels: rounded going forward, its gradient rounded going back."""

    @staticmethod
    def forward(ctx, t, dt):
        ctx.dt = dt
        return t.to(dt).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dt).to(torch.float32), None


class _GradRound(torch.autograd.Function):
    """Identity forward; the gradient is rounded (the fp32 head output whose gradient the GEMMs take in bf16)."""

    @staticmethod
    def forward(ctx, t, dt):
        ctx.dt = dt
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dt).to(torch.float32), None


class _Ctx:
    """Carries the state dict, train/eval flag and operand / storage rounding through the port."""

    def __init__(self, sd, training, operand_dtype=None, tap=None, store_dtype=None):
        self.sd = sd
        self.training = training
        self.od = operand_dtype
        self.tap = tap  # optional dict: name -> intermediate tensor
        self.sdt = store_dtype

    def rnd(self, t):
        return t if self.od is None else t.to(self.od).to(torch.float32)

    def rnd_small(self, t):
        """Operands of the [N, C]-sized fp32 layers (linears, CoordAttn gate network): fp32 in the kernel-matched mode."""
        return t if (self.od is None or self.sdt is not None) else t.to(self.od).to(torch.float32)

    def st(self, t):
        """A tensor the B200 path stores in ``store_dtype`` between kernels (identity in the plain modes)."""
        return t if self.sdt is None else _StoreRound.apply(t, self.sdt)

    def p(self, name):
        return self.sd[name]

    def record(self, name, t):
        if self.tap is not None:
            self.tap[name] = t
        return t


def _conv(cx, pre, x, stride=1, padding=0, before_batch_stats=False):
    """``before_batch_stats``: this conv feeds a BatchNorm.  Kernel-matched train mode stores it without the bias."""
    if cx.sdt is not None and before_batch_stats and cx.training:
        y = cx.st(F.conv2d(cx.rnd(x), cx.rnd(cx.p(pre + ".weight")), None, stride, padding))
        return y + cx.p(pre + ".bias").view(1, -1, 1, 1)
    return cx.st(F.conv2d(cx.rnd(x), cx.rnd(cx.p(pre + ".weight")), cx.p(pre + ".bias"), stride, padding))


def _conv_small(cx, pre, x):
    """1x1 convs of the CoordAttn gate network on the pooled [N, C, L, 1] rows (fp32 kernels on the B200 path)."""
    return F.conv2d(cx.rnd_small(x), cx.rnd_small(cx.p(pre + ".weight")), cx.p(pre + ".bias"))


def _convT(cx, pre, x, stride):
    return cx.st(F.conv_transpose2d(cx.rnd(x), cx.rnd(cx.p(pre + ".weight")), cx.p(pre + ".bias"), stride))


def _linear(cx, pre, x, bias=True):
    return F.linear(cx.rnd_small(x), cx.rnd_small(cx.p(pre + ".weight")), cx.p(pre + ".bias") if bias else None)


def _bn(cx, pre, x):
    """nn.BatchNorm2d: eps 1e-5, momentum 0.1; train mode updates the running buffers in ``sd``."""
    sd = cx.sd
    if cx.training and (pre + ".num_batches_tracked") in sd:
        sd[pre + ".num_batches_tracked"] += 1
    return F.batch_norm(x, sd[pre + ".running_mean"], sd[pre + ".running_var"],
                        sd[pre + ".weight"], sd[pre + ".bias"], cx.training, 0.1, 1e-5)


def _gn(cx, pre, x, groups=8):
    return F.group_norm(x, groups, cx.p(pre + ".weight"), cx.p(pre + ".bias"), 1e-5)


def _conv_bn_gelu(cx, pre, x):
    """Sequential(Conv2d 3x3 p1, BatchNorm2d, GELU): new_scripy.py:183-187."""
    return cx.st(F.gelu(_bn(cx, pre + ".1", _conv(cx, pre + ".0", x, 1, 1, before_batch_stats=True))))


# --------------------------------------------------------------------------- blocks
def se_block(cx, pre, x):
    """new_scripy.py:143-158."""
    b, c = x.shape[:2]
    y = F.adaptive_avg_pool2d(x, 1).squeeze(-1).squeeze(-1)
    y = torch.sigmoid(_linear(cx, pre + ".fc.2", F.gelu(_linear(cx, pre + ".fc.0", y, False)), False))
    return x * y.view(b, c, 1, 1)          # stored only after the residual add (res_conv_block)


def res_conv_block(cx, pre, x, is_res, has_se):
    """new_scripy.py:176-209 (has_se=True) / MNIST_script.py:31-65 (has_se=False)."""
    x1 = _conv_bn_gelu(cx, pre + ".conv1", x)
    x2 = _conv_bn_gelu(cx, pre + ".conv2", x1)
    if not is_res:
        return x2
    if has_se:
        x2 = se_block(cx, pre + ".se", x2)
    same = cx.p(pre + ".conv1.0.weight").shape[0] == cx.p(pre + ".conv1.0.weight").shape[1]
    out = (x + x2) if same else (x1 + x2)
    return cx.st(out / 1.414)


def coord_attn(cx, pre, x):
    """new_scripy.py:97-140."""
    n, c, h, w = x.shape
    x_h = F.adaptive_avg_pool2d(x, (None, 1))
    x_w = F.adaptive_avg_pool2d(x, (1, None))
    x_h = F.gelu(_bn(cx, pre + ".bn1_h", _conv_small(cx, pre + ".conv1_h", x_h)))
    x_w = F.gelu(_bn(cx, pre + ".bn1_w", _conv_small(cx, pre + ".conv1_w", x_w)))
    h2w = _conv_small(cx, pre + ".h2w_proj", x_h).permute(0, 1, 3, 2)
    w2h = _conv_small(cx, pre + ".w2h_proj", x_w).permute(0, 1, 3, 2)
    h2w = F.adaptive_avg_pool2d(h2w, (1, w))
    w2h = F.adaptive_avg_pool2d(w2h, (h, 1))
    x_h = x_h + torch.sigmoid(cx.p(pre + ".gamma_h")) * w2h
    x_w = x_w + torch.sigmoid(cx.p(pre + ".gamma_w")) * h2w
    a_h = torch.sigmoid(_conv_small(cx, pre + ".conv_h", x_h))
    a_w = torch.sigmoid(_conv_small(cx, pre + ".conv_w", x_w))
    alpha = torch.sigmoid(cx.p(pre + ".alpha"))
    beta = torch.sigmoid(cx.p(pre + ".beta"))
    wsum = alpha + beta + 1e-8
    return cx.st(x * ((alpha / wsum) * a_h + (beta / wsum) * a_w))


def local_enhancer(cx, pre, x, mask, high_thresh=HIGH_THRESH):
    """new_scripy.py:161-174 with a [B,H,W] attention map."""
    high = (mask > high_thresh).float().unsqueeze(1)
    y = _conv(cx, pre + ".conv.0", x, 1, 1)
    y = cx.st(F.gelu(_gn(cx, pre + ".conv.1", y)))
    y = _conv(cx, pre + ".conv.3", y, 1, 1)
    return cx.st(x + y * high)


def unet_down_rdd(cx, pre, x):
    """new_scripy.py:211-235."""
    x = cx.st(F.gelu(_bn(cx, pre + ".channel_compress.1", _conv(cx, pre + ".channel_compress.0", x, before_batch_stats=True))))
    x = _conv(cx, pre + ".ch_adjust", x)
    x = cx.st(F.gelu(_bn(cx, pre + ".down.1", _conv(cx, pre + ".down.0", x, 1, 1, before_batch_stats=True))))
    x = res_conv_block(cx, pre + ".down.3", x, True, True)
    return _conv(cx, pre + ".down.4", x, 2, 1)


def unet_up_rdd(cx, pre, x, skip):
    """new_scripy.py:237-253."""
    x = torch.cat((x, skip), 1)
    x = cx.st(F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True))
    x = _conv(cx, pre + ".model.0.1", x, 1, 1)
    x = res_conv_block(cx, pre + ".model.1", x, False, False)
    return res_conv_block(cx, pre + ".model.2", x, False, False)


def unet_down_mnist(cx, pre, x):
    """MNIST_script.py:68-78."""
    return cx.st(F.max_pool2d(res_conv_block(cx, pre + ".model.0", x, False, False), 2))


def unet_up_mnist(cx, pre, x, skip):
    """MNIST_script.py:81-97."""
    x = _convT(cx, pre + ".model.0", torch.cat((x, skip), 1), 2)
    x = res_conv_block(cx, pre + ".model.1", x, False, False)
    return res_conv_block(cx, pre + ".model.2", x, False, False)


def embed_fc(cx, pre, x, input_dim):
    """new_scripy.py:255-268."""
    x = x.view(-1, input_dim)
    return _linear(cx, pre + ".model.2", F.gelu(_linear(cx, pre + ".model.0", x)))


def up0(cx, pre, hidden, k):
    """ConvTranspose2d(k, k) + GroupNorm(8) + ReLU: new_scripy.py:297-301 / MNIST_script.py:139-144."""
    return cx.st(F.relu(_gn(cx, pre + ".1", _convT(cx, pre + ".0", hidden, k))))


def out_head(cx, pre, x):
    """conv3x3 + GroupNorm(8) + ReLU + conv3x3: new_scripy.py:310-315."""
    y = cx.st(F.relu(_gn(cx, pre + ".1", _conv(cx, pre + ".0", x, 1, 1))))
    w, b = cx.rnd(cx.p(pre + ".3.weight")), cx.p(pre + ".3.bias")
    y = F.conv2d(cx.rnd(y), w, b, 1, 1)                          # the head's output stays fp32 ...
    return y if cx.sdt is None else _GradRound.apply(y, cx.sdt)  # ... its gradient reaches the GEMMs in bf16


# --------------------------------------------------------------------------- denoisers
def unet_forward(sd, x, c, t, ctx_mask, *, variant, training, attn_map=None,
                 operand_dtype=None, prefix="", tap=None, store_dtype=None):
    """ContextUnet.forward.  ``sd`` uses the reference's parameter names (optionally under
    ``prefix``, e.g. "nn_model.").  In train mode the BatchNorm running buffers in ``sd`` are
    updated in place exactly as the reference's modules would."""
    if prefix:
        sd = _PrefixView(sd, prefix)
    cx = _Ctx(sd, training, operand_dtype, tap, store_dtype)
    x = cx.st(x)                            # x_t reaches the first conv in the storage dtype
    if variant == "rdd":
        return _unet_rdd(cx, x, c, t, ctx_mask, attn_map)
    if variant == "mnist":
        return _unet_mnist(cx, x, c, t, ctx_mask)
    raise ValueError(variant)


class _PrefixView(dict):
    def __init__(self, sd, prefix):
        super().__init__()
        self._sd, self._p = sd, prefix

    def __getitem__(self, k):
        return self._sd[self._p + k]

    def __setitem__(self, k, v):
        self._sd[self._p + k] = v

    def __contains__(self, k):
        return (self._p + k) in self._sd


def _unet_rdd(cx, x, c, t, ctx_mask, attn_map):
    """new_scripy.py:317-356."""
    n_feat = cx.p("init_conv.conv1.0.weight").shape[0]
    n_classes = cx.p("ctx_emb1.model.0.weight").shape[1]
    x0 = cx.record("init_conv", res_conv_block(cx, "init_conv", x, True, True))
    d1 = cx.record("down1", coord_attn(cx, "ca1", unet_down_rdd(cx, "down1", x0)))
    d2 = cx.record("down2", coord_attn(cx, "ca2", unet_down_rdd(cx, "down2", d1)))
    d3 = cx.record("down3", coord_attn(cx, "ca3", unet_down_rdd(cx, "down3", d2)))
    d4 = cx.record("down4", coord_attn(cx, "ca4", unet_down_rdd(cx, "down4", d3)))
    hidden = cx.st(F.gelu(F.avg_pool2d(d4, 8)))
    c1h = F.one_hot(c.long(), num_classes=n_classes).type(torch.float)
    c1h = c1h * ctx_mask[:, None].repeat(1, n_classes)          # :337-340, no flip
    cemb1 = embed_fc(cx, "ctx_emb1", c1h, n_classes).view(-1, n_feat * 8, 1, 1)
    temb1 = embed_fc(cx, "time_emb1", t, 1).view(-1, n_feat * 8, 1, 1)
    cemb2 = embed_fc(cx, "ctx_emb2", c1h, n_classes).view(-1, n_feat * 4, 1, 1)
    temb2 = embed_fc(cx, "time_emb2", t, 1).view(-1, n_feat * 4, 1, 1)
    u1 = cx.record("up0", up0(cx, "up0", hidden, 8))
    u2 = cx.record("up1", unet_up_rdd(cx, "up1", cx.st(cemb1 * u1 + temb1), d4))
    u3 = cx.record("up2", unet_up_rdd(cx, "up2", cx.st(cemb2 * u2 + temb2), d3))
    u4 = cx.record("up3", unet_up_rdd(cx, "up3", u3, d2))
    u5 = cx.record("up4", unet_up_rdd(cx, "up4", u4, d1))
    if attn_map is None:
        # shipped call site (:353) contributes exactly +0 wherever it runs at all
        attn_map = torch.zeros(x.shape[0], x.shape[2], x.shape[3], dtype=x.dtype, device=x.device)
    u5 = cx.record("local_enhance", local_enhancer(cx, "local_enhance", u5, attn_map))
    return out_head(cx, "out", torch.cat((u5, x0), 1))


def _unet_mnist(cx, x, c, t, context_mask):
    """MNIST_script.py:155-187."""
    n_feat = cx.p("init_conv.conv1.0.weight").shape[0]
    n_classes = cx.p("contextembed1.model.0.weight").shape[1]
    x0 = cx.record("init_conv", res_conv_block(cx, "init_conv", x, True, False))
    d1 = cx.record("down1", unet_down_mnist(cx, "down1", x0))
    d2 = cx.record("down2", unet_down_mnist(cx, "down2", d1))
    hidden = cx.st(F.gelu(F.avg_pool2d(d2, 7)))
    c1h = F.one_hot(c, num_classes=n_classes).type(torch.float)
    m = context_mask[:, None].repeat(1, n_classes)
    m = (-1 * (1 - m))                                          # :170 flip and negate
    c1h = c1h * m
    cemb1 = embed_fc(cx, "contextembed1", c1h, n_classes).view(-1, n_feat * 2, 1, 1)
    temb1 = embed_fc(cx, "timeembed1", t, 1).view(-1, n_feat * 2, 1, 1)
    cemb2 = embed_fc(cx, "contextembed2", c1h, n_classes).view(-1, n_feat, 1, 1)
    temb2 = embed_fc(cx, "timeembed2", t, 1).view(-1, n_feat, 1, 1)
    u1 = cx.record("up0", up0(cx, "up0", hidden, 7))
    u2 = cx.record("up1", unet_up_mnist(cx, "up1", cx.st(cemb1 * u1 + temb1), d2))
    u3 = cx.record("up2", unet_up_mnist(cx, "up2", cx.st(cemb2 * u2 + temb2), d1))
    return out_head(cx, "out", torch.cat((u3, x0), 1))


# --------------------------------------------------------------------------- DDPM
def draw_train_randoms(x, c, n_T, drop_prob, variant):
    """The reference's RNG draw order inside DDPM.forward (new_scripy.py:405-413,
    MNIST_script.py:239-249): randint on the CPU generator, randn_like(x), bernoulli."""
    ts = torch.randint(1, n_T + 1, (x.shape[0],))
    noise = torch.randn_like(x)
    if variant == "rdd":
        ctx_mask = torch.bernoulli(torch.ones_like(c, dtype=torch.float) * (1 - drop_prob))
    else:
        ctx_mask = torch.bernoulli(torch.zeros_like(c) + drop_prob)
    return ts, noise, ctx_mask


def q_sample(sched, x, ts, noise):
    """x_t = sqrtab[t] x + sqrtmab[t] eps: new_scripy.py:408-411."""
    return sched["sqrtab"][ts, None, None, None] * x + sched["sqrtmab"][ts, None, None, None] * noise


def weighted_loss(noise, pred, attn_mask):
    """new_scripy.py:417-437 (mask repeated to exactly 3 channels, :418)."""
    m = attn_mask.unsqueeze(1).repeat(1, 3, 1, 1)
    w = torch.where(m > HIGH_THRESH, torch.tensor(HIGH_WEIGHT),
                    torch.where(m > MID_THRESH, torch.tensor(MID_WEIGHT), torch.tensor(LOW_WEIGHT)))
    loss = (noise - pred) ** 2
    high = (m > HIGH_THRESH).float()
    feat = torch.mean(torch.abs(pred * high - noise * high)) * FEAT_CONSIST_WEIGHT
    return (loss * w).mean() + feat


def ddpm_loss(sd, sched, x, c, attn_mask, ts, noise, ctx_mask, *, variant, n_T, training=True,
              attn_map=None, operand_dtype=None, prefix="nn_model.", store_dtype=None):
    """DDPM.forward with the random draws passed in (see draw_train_randoms)."""
    x_t = q_sample(sched, x, ts, noise)
    pred = unet_forward(sd, x_t, c, ts / n_T, ctx_mask, variant=variant, training=training,
                        attn_map=attn_map, operand_dtype=operand_dtype, prefix=prefix, store_dtype=store_dtype)
    if variant == "rdd":
        return weighted_loss(noise, pred, attn_mask)
    return F.mse_loss(noise, pred)                               # MNIST_script.py:252 (arg order kept)


def reverse_step(sched, x_i, eps1, eps2, z, i, guide_w):
    """new_scripy.py:468-475."""
    eps = (1 + guide_w) * eps1 - guide_w * eps2
    return sched["oneover_sqrta"][i] * (x_i - eps * sched["mab_over_sqrtmab"][i]) + sched["sqrt_beta_t"][i] * z


def ddpm_sample(sd, sched, x_T, zs, guide_w, *, variant, n_T, n_classes, steps=None,
                operand_dtype=None, prefix="nn_model.", store=None, trace=None):
    """CFG reverse loop (new_scripy.py:441-477 / MNIST_script.py:254-300) with x_T and the
    per-step noises ``zs[i]`` (i = n_T .. 2) supplied by the caller.  ``steps`` truncates the
    loop to the first ``steps`` iterations (for tests); ``trace`` (a list) receives x_i after every step."""
    n = x_T.shape[0]
    ncls = 10 if variant == "mnist" else n_classes              # MNIST_script.py:262 hard-codes 10
    c_i = torch.arange(0, ncls).repeat(int(n / ncls)).repeat(2)
    ctx_mask = torch.zeros_like(c_i)
    ctx_mask[n:] = 1.0
    x_i = x_T
    done = 0
    for i in range(n_T, 0, -1):
        t_is = torch.tensor([i / n_T]).repeat(n, 1, 1, 1).repeat(2, 1, 1, 1)
        z = zs[i] if i > 1 else 0
        eps = unet_forward(sd, x_i.repeat(2, 1, 1, 1), c_i, t_is, ctx_mask, variant=variant,
                           training=False, operand_dtype=operand_dtype, prefix=prefix)
        x_i = reverse_step(sched, x_i, eps[:n], eps[n:], z, i, guide_w)
        if store is not None and (i % 20 == 0 or i == n_T or i < 8):
            store.append(x_i.detach().clone())
        if trace is not None:
            trace.append(x_i.detach().clone())
        done += 1
        if steps is not None and done >= steps:
            break
    return x_i


# --------------------------------------------------------------------------- synthetic inputs
def synth_attn_mask(batch, size, gen):
    """Attention map as CrackDataset.__getitem__ builds it (new_scripy.py:535-546): 0.5, lower
    half 1.0, one random bbox 3.0."""
    m = torch.full((batch, size, size), 0.5)
    m[:, size // 2:, :] = MID_WEIGHT
    for b in range(batch):
        xs = torch.randint(0, size, (2,), generator=gen).sort().values
        ys = torch.randint(0, size, (2,), generator=gen).sort().values
        m[b, int(ys[0]):int(ys[1]), int(xs[0]):int(xs[1])] = HIGH_WEIGHT
    return m


def rel_l2(a, b):
    a = a.detach().double().flatten()
    b = b.detach().double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


# --------------------------------------------------------------------------------------- input pipeline (8f rank 3)
def crack_item(image_u8, box_xyxy, orig_wh, flip, img_size, low=0.5, mid=1.0, high=3.0, mean=0.5, std=0.5):
    """CrackDataset.__getitem__ (new_scripy.py:516-551) after decode + Resize, with the transform tail of :683-688:
    ``image_u8`` [H,W,3] uint8 (already resized) -> (x fp32 [3,H,W], attn_mask fp32 [H,W]).  ``flip``: the
    RandomHorizontalFlip decision.  The mask is built before / independently of the transform (:535-549), i.e. it is not
    flipped; box coordinates use Python round() and the clamp of :542-545."""
    xmin, ymin, xmax, ymax = box_xyxy
    ow, oh = orig_wh
    attn = torch.ones((img_size, img_size)) * low                                   # :535
    attn[img_size // 2:, :] = mid                                                   # :538-539
    cl = lambda v: max(0, min(img_size - 1, v))
    xs0, xs1 = cl(round(xmin * img_size / ow)), cl(round(xmax * img_size / ow))     # :542-543
    ys0, ys1 = cl(round(ymin * img_size / oh)), cl(round(ymax * img_size / oh))     # :544-545
    attn[ys0:ys1, xs0:xs1] = high                                                   # :546
    img = image_u8
    if flip:
        img = img.flip(1)                                                           # RandomHorizontalFlip, :685
    x = img.permute(2, 0, 1).to(torch.float32).div(255)                             # ToTensor, :686
    x = (x - mean) / std                                                            # Normalize, :687
    return x, attn


# --------------------------------------------------------------------------------------- sample quality (8f rank 4)
def calc_ssim(img1, img2):
    """ImageMetrics.calc_ssim (new_scripy.py:1189-1222): global-statistics SSIM in numpy float32."""
    import numpy as np
    if img1.min() < 0:
        img1 = (img1 + 1) / 2
    if img2.min() < 0:
        img2 = (img2 + 1) / 2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    a, b = img1.cpu().numpy(), img2.cpu().numpy()
    mu1, mu2 = np.mean(a), np.mean(b)
    s1, s2 = np.std(a), np.std(b)
    s12 = np.mean((a - mu1) * (b - mu2))
    return ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 ** 2 + mu2 ** 2 + C1) * (s1 ** 2 + s2 ** 2 + C2))


def calc_psnr(img1, img2):
    """ImageMetrics.calc_psnr (new_scripy.py:1224-1251)."""
    import numpy as np
    if img1.min() < 0:
        img1 = (img1 + 1) / 2
    if img2.min() < 0:
        img2 = (img2 + 1) / 2
    mse = torch.mean((img1 - img2) ** 2).item()
    if mse == 0:
        return float("inf")
    return 20 * np.log10(1.0 / np.sqrt(mse))


def fid_preprocess(images):
    """ImageMetrics._extract_features up to the Inception call (new_scripy.py:1133-1142): map the BATCH to [0,1] when its
    minimum is negative, resize to 299 x 299 (bilinear, align_corners=False)."""
    if images.min() < 0:
        images = (images + 1) / 2
    if images.shape[2] != 299 or images.shape[3] != 299:
        images = F.interpolate(images, size=(299, 299), mode="bilinear", align_corners=False)
    return images


def fid_from_features(real_feats, gen_feats):
    """The Frechet half of ImageMetrics.calc_fid (new_scripy.py:1168-1187) on two [N, D] numpy feature matrices."""
    import numpy as np
    from scipy import linalg
    mu_real = np.mean(real_feats, axis=0)
    sigma_real = np.cov(real_feats, rowvar=False)
    mu_gen = np.mean(gen_feats, axis=0)
    sigma_gen = np.cov(gen_feats, rowvar=False)
    diff = mu_real - mu_gen
    covmean = linalg.sqrtm(sigma_real.dot(sigma_gen))      # the reference passes disp=False (SciPy < 1.16 signature: value, error)
    if np.iscomplexobj(covmean):
        covmean = covmean.real
    return diff.dot(diff) + np.trace(sigma_real + sigma_gen - 2 * covmean)


def fid_stub_projection(seed, dim=16):
    """The fixed [3*299*299, dim] projection that stands in for Inception-v3 in the FID fixture (tests/golden/fid.npz)."""
    return torch.randn(3 * 299 * 299, dim, generator=torch.Generator().manual_seed(int(seed))) / 300.0
