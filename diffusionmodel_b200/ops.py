"""Autograd operators over the sm_100a kernels.

Activations travel as bf16 NHWC tensors ``[N, H, W, ld]`` (``ld`` = channel pitch, a multiple of 8,
pad lanes zero; channel slices of such tensors are valid operands).  Parameters stay fp32 in the
reference's layouts; their gradients are accumulated by the kernels directly into ``param.grad``
(the Functions return ``None`` for parameter inputs), which avoids ~400 tiny accumulation kernels
per backward pass.
"""
from __future__ import annotations

import ctypes
import weakref

import torch

from . import _lib
from ._lib import call

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
_weights_epoch = 0            # bumped by the fused optimizer (it writes parameters behind torch's back)
_bn_stats_epoch = 0           # bumped whenever kernels may have updated BatchNorm running statistics in place


def bump_weights_epoch():
    global _weights_epoch
    _weights_epoch += 1


def bump_bn_stats_epoch():
    global _bn_stats_epoch
    _bn_stats_epoch += 1


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# --------------------------------------------------------------------------------------- GEMM-native weight storage
# The fused optimizer may store a Conv2d weight [Cout, Cin, kh, kw] in the implicit GEMM's own order
# [Cout][kh][kw][Cin] (the parameter becomes a permuted view: same shape, same values, same state_dict).  Then
# the wgrad GEMM accumulates straight into ``param.grad``'s memory, and the bf16 copy the optimizer writes next
# to the update IS the forward weight pack -- no scatter and no pack kernel per step for those weights.
_conv2d_weights = {}                    # id(param) -> weakref: parameters known to be nn.Conv2d weights (ConvTranspose2d
                                        # weights are [Cin, Cout, k, k]); keyed by id because tensors compare elementwise
_native = {}                            # weight.data_ptr() -> _NativeWeight


class _NativeWeight:
    __slots__ = ("ref", "shadow", "stamp")

    def __init__(self, param, shadow):
        self.ref, self.shadow, self.stamp = weakref.ref(param), shadow, None


def mark_conv2d_weights(module):
    """Tell the optimizer which 4-D parameters are Conv2d weights (called by the U-Net constructors)."""
    for m in module.modules():
        if isinstance(m, torch.nn.Conv2d):
            _conv2d_weights[id(m.weight)] = weakref.ref(m.weight)


def native_strides(shape):
    cout, cin, kh, kw = shape
    return (kh * kw * cin, 1, kw * cin, cin)


def wants_native(p):
    """Conv2d weights with a multiple of 64 input channels and a real filter footprint: the packed GEMM layout
    then has no padding, i.e. it is a permutation of the parameter."""
    r = _conv2d_weights.get(id(p))
    return r is not None and r() is p and p.dim() == 4 and p.shape[1] % 64 == 0 and p.shape[2] * p.shape[3] > 1


def register_native(param, shadow2d):
    _native[param.data_ptr()] = _NativeWeight(param, shadow2d)


def release_registries():
    """Forget every parameter registered by an optimizer (GEMM-native shadows, deferred packed gradients): lets a model
    and its optimizer be garbage-collected before another one is built in the same process."""
    _native.clear()
    _deferred.clear()
    _conv2d_weights.clear()


def _native_of(w):
    e = _native.get(w.data_ptr())
    if e is None:
        return None
    p = e.ref()
    if p is None or p.shape != w.shape or tuple(w.stride()) != native_strides(w.shape):
        del _native[w.data_ptr()]
        return None
    return e


def natives_fresh():
    """The optimizer has just written every registered shadow together with its update."""
    for key in list(_native):
        e = _native[key]
        p = e.ref()
        if p is None or p.data_ptr() != key:
            del _native[key]
        else:
            e.stamp = (p._version, _weights_epoch)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def r8(c):
    return (c + 7) // 8 * 8


def r64(c):
    return (c + 63) // 64 * 64


def _chk(t, name="tensor"):
    if t.device.type != "cuda":
        raise _lib.DmB200Error(f"{name}: the hot path runs on CUDA only (got {t.device}); there is no CPU fallback")
    if t.device.index != torch.cuda.current_device():
        # raw pointers, the reduction workspace and the launching stream all belong to the CURRENT device (one process per GPU)
        raise _lib.DmB200Error(f"{name}: tensor lives on {t.device} but the current device is cuda:{torch.cuda.current_device()}; "
                               "call torch.cuda.set_device() first (one process per GPU)")
    if t.dtype != torch.bfloat16 or t.dim() != 4 or t.stride(3) != 1:
        raise _lib.DmB200Error(f"{name}: expected a bf16 NHWC activation, got {t.dtype} {tuple(t.shape)} {t.stride()}")
    n, h, w, _ = t.shape
    ld = t.stride(2)
    if t.stride(1) != w * ld or t.stride(0) != h * w * ld or ld % 8 or t.data_ptr() % 16:
        raise _lib.DmB200Error(f"{name}: unsupported strides {t.stride()} for shape {tuple(t.shape)}")
    return ld


def new_act(n, h, w, c, device, dtype=torch.bfloat16):
    return torch.empty((n, h, w, r8(c)), device=device, dtype=dtype)


def grad_buf(p):
    """fp32 accumulation target for a parameter's gradient."""
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


# --------------------------------------------------------------------------------------- weight packs
class WeightPack:
    """bf16 K-major packs of one conv weight for the implicit-GEMM kernels, rebuilt IN PLACE when the fp32
    master parameter changes (torch-side in-place update or the fused optimizer epoch), so their
    addresses stay valid inside captured CUDA graphs."""

    def __init__(self):
        self.cache = {}
        self.weight = None
        _all_packs.add(self)

    def get(self, w, kind, **kw):
        key = (kind, tuple(sorted(kw.items())))
        stamp = (w._version, _weights_epoch, w.data_ptr())
        hit = self.cache.get(key)
        if hit is not None and hit[0] == stamp:
            return hit[1]
        self.weight = weakref.ref(w)
        out = _PACKERS[kind](w.detach(), hit[1] if hit is not None else None, **kw)
        self.cache[key] = (stamp, out)
        return out

    def refresh(self):
        w = self.weight() if self.weight is not None else None
        if w is not None:
            for kind, kw in list(self.cache):
                self.get(w, kind, **dict(kw))


_all_packs = weakref.WeakSet()


def refresh_weight_packs():
    """Re-pack every cached weight pack whose master changed (called by the optimizer right after its
    update, so no pack kernel runs inside a later forward/backward -- or inside a replayed graph)."""
    for pk in list(_all_packs):
        pk.refresh()


def _taps(offs):
    arr = (ctypes.c_longlong * 64)(*offs, *([0] * (64 - len(offs))))
    return arr


def _pack(w, out, rows, cols, tap_offs, s_row, s_col, c_split, cols_k, row_len, tap_major):
    total_rows = rows * (len(tap_offs) if tap_major else 1)
    if out is None or tuple(out.shape) != (total_rows, row_len):
        out = torch.empty((total_rows, row_len), device=w.device, dtype=torch.bfloat16)
    call("dm_pack_weight", _p(w), _p(out), rows, cols, len(tap_offs), _taps(tap_offs), s_row, s_col, c_split,
         cols_k, row_len, 1 if tap_major else 0, _stream())
    return out


def _cols_k(cin, c_split):
    return r64(cin) if not c_split else r64(c_split) + r64(cin - c_split)


def conv_geom(w, c_split=0):
    cout, cin, kh, kw = w.shape
    ck = _cols_k(cin, c_split)
    return cout, cin, kh, kw, ck


def _fresh_shadow(w, e):
    """bf16 [Cout, kh*kw*Cin] copy of a GEMM-native weight, re-cast only when the parameter changed behind the
    optimizer's back (load_state_dict, broadcast, manual edits)."""
    stamp = (w._version, _weights_epoch)
    if e.stamp != stamp:
        cout, cin, kh, kw = w.shape
        e.shadow.copy_(w.permute(0, 2, 3, 1).reshape(cout, kh * kw * cin))
        e.stamp = stamp
    return e.shadow


def _pack_transposed(w, e, out, rows_out, row_len, dst_offs):
    cout, cin, kh, kw = w.shape
    if out is None or tuple(out.shape) != (rows_out, row_len):
        out = torch.empty((rows_out, row_len), device=w.device, dtype=torch.bfloat16)
    call("dm_pack_transpose", _p(_fresh_shadow(w, e)), _p(out), cout, cin, kh * kw, _taps(dst_offs), row_len, _stream())
    return out


def _pack_fwd(w, out=None, c_split=0):
    cout, cin, kh, kw, ck = conv_geom(w, c_split)
    e = _native_of(w)
    if e is not None and ck == cin:
        return _fresh_shadow(w, e)      # GEMM-native storage: the forward pack is the optimizer's bf16 shadow
    s0, s1, s2, s3 = w.stride()
    offs = [r * s2 + s * s3 for r in range(kh) for s in range(kw)]
    return _pack(w, out, cout, cin, offs, s0, s1, c_split, ck, kh * kw * ck, False)


def _pack_dgrad(w, out=None):
    """stride-1 data gradient = conv with the 180-degree rotated kernel and Cin/Cout swapped."""
    cout, cin, kh, kw = w.shape
    ck = r64(cout)
    e = _native_of(w)
    if e is not None:
        # source tap (r, s) lands at rotated position (kh-1-r, kw-1-s) of the data-gradient filter
        dst = [((kh - 1 - r) * kw + (kw - 1 - s)) * ck for r in range(kh) for s in range(kw)]
        return _pack_transposed(w, e, out, cin, kh * kw * ck, dst)
    s0, s1, s2, s3 = w.stride()
    offs = [(kh - 1 - r) * s2 + (kw - 1 - s) * s3 for r in range(kh) for s in range(kw)]
    return _pack(w, out, cin, cout, offs, s1, s0, 0, ck, kh * kw * ck, False)


def _pack_s2dgrad(w, out=None):
    """k=4, s=2, p=1 data gradient: four output-parity phases of 2x2 taps (see dm_conv2d_s2_dgrad)."""
    cout, cin, kh, kw = w.shape
    assert kh == 4 and kw == 4
    ck = r64(cout)
    s0, s1, s2, s3 = w.stride()
    rsel = {0: (1, 3), 1: (0, 2)}
    e = _native_of(w)
    if e is not None:
        dst = [0] * 16
        for ph in range(2):
            for pw in range(2):
                for j, (r, s_) in enumerate((r, s_) for r in rsel[ph] for s_ in rsel[pw]):
                    dst[r * kw + s_] = (ph * 2 + pw) * cin * 4 * ck + j * ck
        return _pack_transposed(w, e, out, 4 * cin, 4 * ck, dst)
    if out is None or tuple(out.shape) != (4 * cin, 4 * ck):
        out = torch.empty((4 * cin, 4 * ck), device=w.device, dtype=torch.bfloat16)
    i = 0
    for ph in range(2):
        for pw in range(2):
            offs = [r * s2 + s * s3 for r in rsel[ph] for s in rsel[pw]]
            _pack(w, out[i * cin:(i + 1) * cin], cin, cout, offs, s1, s0, 0, ck, 4 * ck, False)
            i += 1
    return out


def _pack_convt_fwd(w, out=None):
    """ConvTranspose2d weight [Cin, Cout, k, k] -> rows (tap, co), K = ci."""
    cin, cout, k, _ = w.shape
    offs = [t for t in range(k * k)]
    return _pack(w, out, cout, cin, offs, k * k, cout * k * k, 0, r64(cin), r64(cin), True)


def _pack_convt_dgrad(w, out=None):
    """rows ci, K = (tap, co): a 1x1 conv over the space-to-depth'd output gradient."""
    cin, cout, k, _ = w.shape
    offs = [t for t in range(k * k)]
    return _pack(w, out, cin, cout, offs, cout * k * k, k * k, 0, cout, r64(k * k * cout), False)


def _pack_im2col(w, out=None):
    """[Cout, Cin, 3, 3] with Cin*9 <= 32 as a 1x1 weight over the im2col'd input: K column = ci*9 + tap,
    i.e. the parameter's own memory order, padded to one 64-wide K block."""
    cout, cin, kh, kw = w.shape
    return _pack(w, out, cout, cin * kh * kw, [0], cin * kh * kw, 1, 0, 64, 64, False)


_PACKERS = {"im2col": _pack_im2col, "fwd": _pack_fwd, "dgrad": _pack_dgrad, "s2dgrad": _pack_s2dgrad,
            "convt_fwd": _pack_convt_fwd, "convt_dgrad": _pack_convt_dgrad}


# --------------------------------------------------------------------------------------- weight gradients
class _PackedGrad:
    """Persistent fp32 accumulator of one conv weight's gradient in the GEMM's packed layout.

    The wgrad kernel red.adds into it on every micro-batch; ``flush`` scatters it into the parameter's
    NCHW ``.grad`` (+=) and re-zeroes it in the same pass.  Owned by the optimizer (FusedAdamW), which
    flushes once per optimizer step instead of once per backward (ACCUM_STEPS x fewer scatter passes and
    no per-backward zero-fill of a weight-sized buffer)."""

    __slots__ = ("weight", "dwp", "args", "dirty")

    def __init__(self, weight, shape):
        self.weight, self.args, self.dirty = weight, None, False
        self.dwp = torch.zeros(shape, device=weight.device, dtype=torch.float32)

    def flush(self):
        if self.dirty:
            call("dm_unpack_wgrad", _p(self.dwp), _p(grad_buf(self.weight)), *self.args, 1, _stream())
            self.dirty = False


class _RankGrad:
    """Deferred weight gradient of a ConvTranspose2d(k, stride=k) that sees only a handful of input pixels per
    backward pass (up0 on the 2x2 bottleneck: 16 pixels against a 151 M-element weight, new_scripy.py:297-301).
    Its wgrad GEMM has K = 16 and is bound by the read-modify-write of the 604 MB fp32 gradient, so instead of one such
    pass per micro-batch the operands (the space-to-depth'd output gradient and the input) of up to CAP backward passes
    are queued -- 3 MB each -- and contracted by ONE GEMM with K = 16 * passes when the optimizer flushes.
    Inside a CUDA-graph capture the operands go to fixed staging buffers; ``after_replay`` moves them into the queue."""

    CAP = 8
    __slots__ = ("weight", "geom", "ss", "xs", "stage_s", "stage_x", "count", "dirty")

    def __init__(self, weight, n, hin, win, cin, ldx, kc):
        dev = weight.device
        self.weight, self.geom = weight, (n, hin, win, cin, ldx, kc)
        self.ss = torch.empty((self.CAP * n, hin, win, kc), device=dev, dtype=torch.bfloat16)
        self.xs = torch.empty((self.CAP * n, hin, win, ldx), device=dev, dtype=torch.bfloat16)
        self.stage_s = self.stage_x = None
        self.count, self.dirty = 0, False

    def _slot(self):
        if self.count == self.CAP:
            self.flush()
        n, k = self.geom[0], self.count
        self.count += 1
        self.dirty = True
        return self.ss[k * n:(k + 1) * n], self.xs[k * n:(k + 1) * n]

    def targets(self):
        """(space-to-depth destination, input destination) for the backward pass being recorded."""
        if _capture_log is not None:
            if self.stage_s is None:
                n = self.geom[0]
                self.stage_s, self.stage_x = torch.empty_like(self.ss[:n]), torch.empty_like(self.xs[:n])
            _capture_log.append(self)
            return self.stage_s, self.stage_x
        return self._slot()

    def after_replay(self):
        s, x = self._slot()
        s.copy_(self.stage_s)
        x.copy_(self.stage_x)

    def flush(self):
        if self.count:
            n, hin, win, cin, ldx, kc = self.geom
            call("dm_conv2d_wgrad", _p(self.ss), kc, kc, None, 0, 0, _p(self.xs), ldx, _p(grad_buf(self.weight)),
                 n * self.count, hin, win, cin, 1, 1, 1, 0, _stream())
        self.count, self.dirty = 0, False

    def discard(self):
        self.count, self.dirty = 0, False


def _rank_grad_entry(weight, n, hin, win, cin, ldx, kc):
    """The queue for ``weight`` if the optimizer registered it for deferred gradients (defer_weight_grads), else None."""
    slot = _deferred.get(weight.data_ptr())
    if slot is None:
        return None
    p = slot[0]()
    if p is None or p.shape != weight.shape:
        return None
    e = slot[1]
    if not isinstance(e, _RankGrad) or e.geom != (n, hin, win, cin, ldx, kc):
        if e is not None:
            e.flush()
        e = slot[1] = _RankGrad(weight, n, hin, win, cin, ldx, kc)
    return e


_deferred = {}          # weight.data_ptr() -> [weakref(param), _PackedGrad | None]
_capture_log = None     # while a train step is being captured: the _PackedGrad entries it accumulates into


def defer_weight_grads(params):
    """Register parameters whose conv weight gradients may stay packed until ``flush_weight_grads``."""
    import weakref
    for p in params:
        if p.dim() == 4:
            _deferred[p.data_ptr()] = [weakref.ref(p), None]


def _live_entries():
    for key in list(_deferred):
        ref, e = _deferred[key]
        p = ref()
        if p is None or p.data_ptr() != key:
            del _deferred[key]              # parameter gone or re-homed: forget it
        elif e is not None:
            yield e


class capture_log:
    """Context manager used while CUDA-graph capturing a backward pass: collects the packed-gradient
    accumulators the captured kernels write, so each replay can mark them dirty again."""

    def __enter__(self):
        global _capture_log
        _capture_log = []
        return _capture_log

    def __exit__(self, *exc):
        global _capture_log
        _capture_log = None


def flush_weight_grads():
    for e in _live_entries():
        e.flush()


def discard_weight_grads():
    """zero_grad(): drop packed gradients that were never flushed."""
    for e in _live_entries():
        if isinstance(e, _RankGrad):
            e.discard()
        elif e.dirty:
            e.dwp.zero_()
            e.dirty = False


def _wgrad_into(weight, shape, unpack_args, run):
    """Run ``run(dwp)`` (the wgrad GEMM accumulating into ``dwp``) and deliver the result to weight.grad:
    immediately, or -- for parameters registered by the optimizer -- at the next flush."""
    key = weight.data_ptr()
    slot = _deferred.get(key)
    if slot is not None:
        p = slot[0]()
        if p is None or p.shape != weight.shape:
            del _deferred[key]
            slot = None
    if slot is not None:
        e = slot[1]
        if e is None or e.dwp.shape != torch.Size(shape):
            e = slot[1] = _PackedGrad(weight, shape)
        run(e.dwp)
        e.args, e.dirty = unpack_args, True
        if _capture_log is not None:
            _capture_log.append(e)
        return
    dwp = torch.zeros(shape, device=weight.device, dtype=torch.float32)
    run(dwp)
    call("dm_unpack_wgrad", _p(dwp), _p(grad_buf(weight)), *unpack_args, 0, _stream())


# --------------------------------------------------------------------------------------- convolution
def conv_stat_rows(n, ho, wo, cout):
    return _lib.fn("dm_conv2d_fwd_stat_rows")(n, ho, wo, cout)


_stats_max_cout = None


def conv_takes_stats(cout):
    """True when the conv epilogue can also produce the BatchNorm statistics of its output (FUSED_CONV_STATS and the
    per-CTA slab fits: up to dm_conv2d_fwd_stats_max_cout() = 1536 channels; wider layers keep the dm_bn_stats pass)."""
    global _stats_max_cout
    if _stats_max_cout is None:
        _stats_max_cout = _lib.fn("dm_conv2d_fwd_stats_max_cout")()
    return FUSED_CONV_STATS and cout <= _stats_max_cout


class _Conv2d(torch.autograd.Function):
    """nn.Conv2d (new_scripy.py:184 etc.) on one or two channel-concatenated NHWC sources."""

    @staticmethod
    def forward(ctx, x0, x1, weight, bias, pack, c0, c1, stride, pad, want_stats, out_f32, bias_grad_by_norm, add_bias):
        ld0 = _chk(x0, "conv input")
        ld1 = _chk(x1, "conv input 2") if x1 is not None else 0
        n, hin, win, _ = x0.shape
        cout, cin, kh, kw = weight.shape
        assert cin == c0 + c1, (cin, c0, c1)
        ho = (hin + 2 * pad - kh) // stride + 1
        wo = (win + 2 * pad - kw) // stride + 1
        wpk = pack.get(weight, "fwd", c_split=c0 if x1 is not None else 0)
        scale, act = None, ACT_NONE
        if out_f32:
            y = torch.empty((n, ho, wo, (cout + 3) // 4 * 4), device=x0.device, dtype=torch.float32)
        else:
            y = new_act(n, ho, wo, cout, x0.device)
        stats = None
        if want_stats:
            stats = torch.empty((conv_stat_rows(n, ho, wo, cout), 2, cout), device=x0.device, dtype=torch.float32)
        call("dm_conv2d_fwd", _p(x0), c0, ld0, _p(x1), c1, ld1, _p(wpk), _p(bias) if add_bias else None, _p(scale), act,
             _p(y), y.stride(2), int(out_f32), _p(stats), cout, n, hin, win, cout, kh, kw, stride, pad, _stream())
        ctx.save_for_backward(x0, x1, weight, bias)
        ctx.pack, ctx.geom = pack, (c0, c1, stride, pad, out_f32, bias_grad_by_norm)
        ctx.mark_non_differentiable(*([stats] if stats is not None else []))
        return (y, stats) if want_stats else (y, None)

    @staticmethod
    def backward(ctx, dy, _dstats):
        x0, x1, weight, bias = ctx.saved_tensors
        c0, c1, stride, pad, out_f32, bias_grad_by_norm = ctx.geom
        cout, cin, kh, kw = weight.shape
        n, hin, win, _ = x0.shape
        ho, wo = dy.shape[1], dy.shape[2]
        st = _stream()
        if out_f32:      # fp32 head output: its gradient arrives as fp32 NHWC; the GEMMs take bf16
            dyb = new_act(n, ho, wo, cout, dy.device)
            dy = dy.contiguous()
            call("dm_cast_nhwc", _p(dy), dy.stride(2), _p(dyb), dyb.stride(2), n * ho * wo, cout, st)
            dy = dyb
        lddy = _chk(dy, "conv grad")
        # bias gradient: column sums of dy (a following BatchNorm's backward produces it instead)
        if bias is not None and not bias_grad_by_norm:
            call("dm_colsum", _p(dy), lddy, _p(grad_buf(bias)), n * ho * wo, cout, st)
        # weight gradient: packed fp32 [Cout][taps][Cin_k], scattered (+=) into the NCHW parameter grad
        ck = _cols_k(cin, c0 if x1 is not None else 0)

        def run_wgrad(dwp, stream=st):
            call("dm_conv2d_wgrad", _p(x0), c0, x0.stride(2), _p(x1), c1, x1.stride(2) if x1 is not None else 0, _p(dy),
                 lddy, _p(dwp), n, hin, win, cout, kh, kw, stride, pad, stream)
        gw = grad_buf(weight)
        if ck == cin and _native_of(weight) is not None and tuple(gw.stride()) == native_strides(weight.shape):
            # GEMM-native storage: param.grad's memory is the packed accumulator
            run_wgrad(gw)
        else:
            s0, s1, s2, s3 = gw.stride()
            offs = [r * s2 + s * s3 for r in range(kh) for s in range(kw)]
            unpack = (cout, cin, kh * kw, _taps(offs), s0, s1, c0 if x1 is not None else 0, ck, kh * kw * ck, 0)
            _wgrad_into(weight, (cout, kh * kw * ck), unpack, run_wgrad)
        # data gradient
        dx0 = dx1 = None
        need0 = ctx.needs_input_grad[0]
        need1 = x1 is not None and ctx.needs_input_grad[1]
        if need0 or need1:
            # one gradient tensor for both sources; the halves are channel slices of it
            width = x0.shape[3] + (x1.shape[3] if x1 is not None else 0)
            dx = torch.empty((n, hin, win, width), device=dy.device, dtype=torch.bfloat16)
            if stride == 1:
                wd = ctx.pack.get(weight, "dgrad")
                call("dm_conv2d_fwd", _p(dy), cout, lddy, None, 0, 0, _p(wd), None, None, 0, _p(dx), dx.stride(2), 0, None, 0,
                     n, ho, wo, cin, kh, kw, 1, kh - 1 - pad, st)
            else:
                wd = ctx.pack.get(weight, "s2dgrad")
                call("dm_conv2d_s2_dgrad", _p(dy), cout, lddy, _p(wd), _p(dx), cin, dx.stride(2), n, ho, wo, st)
            if x1 is None:
                dx0 = dx
            else:
                dx0, dx1 = dx[..., :c0], dx[..., c0:]
        return dx0, dx1, None, None, None, None, None, None, None, None, None, None, None


class _Im2colConv3x3(torch.autograd.Function):
    """3x3 / pad 1 convolution of a <=3-channel image (the U-Net's first conv) as im2col + 1x1 GEMM."""

    @staticmethod
    def forward(ctx, x, weight, bias, pack, want_stats, bias_grad_by_norm, add_bias):
        ldx = _chk(x, "conv input")
        n, h, w, _ = x.shape
        cout, cin = weight.shape[0], weight.shape[1]
        st = _stream()
        xi = torch.empty((n, h, w, 32), device=x.device, dtype=torch.bfloat16)
        call("dm_im2col3x3", _p(x), ldx, _p(xi), n, h, w, cin, st)
        wpk = pack.get(weight, "im2col")
        y = new_act(n, h, w, cout, x.device)
        stats = None
        if want_stats:
            stats = torch.empty((conv_stat_rows(n, h, w, cout), 2, cout), device=x.device, dtype=torch.float32)
        call("dm_conv2d_fwd", _p(xi), cin * 9, 32, None, 0, 0, _p(wpk), _p(bias) if add_bias else None, None, 0, _p(y),
             y.stride(2), 0, _p(stats), cout, n, h, w, cout, 1, 1, 1, 0, st)
        ctx.save_for_backward(xi, weight, bias)
        ctx.pack, ctx.cfg = pack, bias_grad_by_norm
        ctx.mark_non_differentiable(*([stats] if stats is not None else []))
        return (y, stats) if want_stats else (y, None)

    @staticmethod
    def backward(ctx, dy, _dstats):
        xi, weight, bias = ctx.saved_tensors
        cout, cin = weight.shape[0], weight.shape[1]
        n, h, w, _ = xi.shape
        lddy = _chk(dy, "conv grad")
        st = _stream()
        if bias is not None and not ctx.cfg:
            call("dm_colsum", _p(dy), lddy, _p(grad_buf(bias)), n * h * w, cout, st)
        unpack = (cout, cin * 9, 1, _taps([0]), cin * 9, 1, 0, 64, 64, 0)
        _wgrad_into(weight, (cout, 64), unpack, lambda dwp: call(
            "dm_conv2d_wgrad", _p(xi), cin * 9, 32, None, 0, 0, _p(dy), lddy, _p(dwp), n, h, w, cout, 1, 1, 1, 0, st))
        dx = None
        if ctx.needs_input_grad[0]:
            wd = ctx.pack.get(weight, "dgrad")
            dx = new_act(n, h, w, cin, dy.device)
            call("dm_conv2d_fwd", _p(dy), cout, lddy, None, 0, 0, _p(wd), None, None, 0, _p(dx), dx.stride(2), 0, None, 0,
                 n, h, w, cin, 3, 3, 1, 1, st)
        return dx, None, None, None, None, None, None


def conv2d_fused_eval(x0, weight, pack, shift, scale, act, *, x1=None, c1=0, stride=1, pad=0):
    """Inference-only conv with an eval-mode BatchNorm + activation folded into the epilogue:
    act(conv(x) * scale + shift) in ONE kernel (sampling loop; no autograd tape)."""
    c0 = weight.shape[1] - c1
    ld0 = _chk(x0, "conv input")
    ld1 = _chk(x1, "conv input 2") if x1 is not None else 0
    n, hin, win, _ = x0.shape
    cout, cin, kh, kw = weight.shape
    if x1 is None and cin <= 3 and (kh, kw, stride, pad) == (3, 3, 1, 1):       # first conv: im2col + 1x1
        xi = torch.empty((n, hin, win, 32), device=x0.device, dtype=torch.bfloat16)
        call("dm_im2col3x3", _p(x0), ld0, _p(xi), n, hin, win, cin, _stream())
        y = new_act(n, hin, win, cout, x0.device)
        call("dm_conv2d_fwd", _p(xi), cin * 9, 32, None, 0, 0, _p(pack.get(weight, "im2col")), _p(shift), _p(scale), act,
             _p(y), y.stride(2), 0, None, 0, n, hin, win, cout, 1, 1, 1, 0, _stream())
        return y
    ho = (hin + 2 * pad - kh) // stride + 1
    wo = (win + 2 * pad - kw) // stride + 1
    wpk = pack.get(weight, "fwd", c_split=c0 if x1 is not None else 0)
    y = new_act(n, ho, wo, cout, x0.device)
    call("dm_conv2d_fwd", _p(x0), c0, ld0, _p(x1), c1, ld1, _p(wpk), _p(shift), _p(scale), act, _p(y), y.stride(2), 0,
         None, 0, n, hin, win, cout, kh, kw, stride, pad, _stream())
    return y


def _conv2d_shared_x1(x0, x1, weight, bias, pack, c0, c1, stride, pad, add_bias):
    """Forward-only dual-source conv whose second source holds fewer samples than the first and is shared cyclically by
    the batch (the CFG halves of a sampling batch share the encoder's x0, new_scripy.py:355): one launch per group of
    len(x1) samples, no duplicated copy of x1."""
    if torch.is_grad_enabled() and (x0.requires_grad or x1.requires_grad or weight.requires_grad):
        raise _lib.DmB200Error("conv2d: a shared second source (fewer samples than the batch) is forward-only")
    ld0, ld1 = _chk(x0, "conv input"), _chk(x1, "conv input 2")
    n, hin, win, _ = x0.shape
    nb = x1.shape[0]
    if n % nb or x1.shape[1:3] != x0.shape[1:3]:
        raise _lib.DmB200Error(f"conv2d: second source {tuple(x1.shape)} does not tile the batch {tuple(x0.shape)}")
    cout, cin, kh, kw = weight.shape
    ho = (hin + 2 * pad - kh) // stride + 1
    wo = (win + 2 * pad - kw) // stride + 1
    wpk = pack.get(weight, "fwd", c_split=c0)
    y = new_act(n, ho, wo, cout, x0.device)
    for g in range(n // nb):
        xg, yg = x0[g * nb:(g + 1) * nb], y[g * nb:(g + 1) * nb]
        call("dm_conv2d_fwd", _p(xg), c0, ld0, _p(x1), c1, ld1, _p(wpk), _p(bias) if add_bias else None, None, ACT_NONE, _p(yg),
             y.stride(2), 0, None, cout, nb, hin, win, cout, kh, kw, stride, pad, _stream())
    return y


def conv2d(x0, weight, bias, pack, *, x1=None, c0=None, c1=0, stride=1, pad=0, want_stats=False, out_f32=False,
           bias_grad_by_norm=False, add_bias=True):
    c0 = weight.shape[1] - c1 if c0 is None else c0
    if (x1 is None and weight.shape[1] <= 3 and tuple(weight.shape[2:]) == (3, 3) and stride == 1 and pad == 1
            and not out_f32):
        return _Im2colConv3x3.apply(x0, weight, bias, pack, want_stats, bias_grad_by_norm, add_bias)
    if x1 is not None and (c0 % 8 or x0.shape[3] != c0):
        raise _lib.DmB200Error("dual-source conv needs a tight first source with a multiple-of-8 channel count")
    if x1 is not None and x1.shape[0] != x0.shape[0]:
        return _conv2d_shared_x1(x0, x1, weight, bias, pack, c0, c1, stride, pad, add_bias), None
    return _Conv2d.apply(x0, x1, weight, bias, pack, c0, c1, stride, pad, want_stats, out_f32, bias_grad_by_norm, add_bias)


class _ConvT(torch.autograd.Function):
    """nn.ConvTranspose2d(k, stride=k) (new_scripy.py:298; MNIST_script.py:88,141)."""

    @staticmethod
    def forward(ctx, x, weight, bias, pack, k):
        ldx = _chk(x, "convT input")
        n, hin, win, _ = x.shape
        cin, cout = weight.shape[0], weight.shape[1]
        wpk = pack.get(weight, "convt_fwd")
        y = new_act(n, hin * k, win * k, cout, x.device)
        call("dm_convt_fwd", _p(x), cin, ldx, _p(wpk), _p(bias), _p(y), y.stride(2), n, hin, win, cout, k, _stream())
        ctx.save_for_backward(x, weight, bias)
        ctx.pack, ctx.k = pack, k
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, bias = ctx.saved_tensors
        k = ctx.k
        cin, cout = weight.shape[0], weight.shape[1]
        n, hin, win, _ = x.shape
        lddy = _chk(dy, "convT grad")
        st = _stream()
        if bias is not None:
            call("dm_colsum", _p(dy), lddy, _p(grad_buf(bias)), n * hin * k * win * k, cout, st)
        kc = k * k * cout
        # weight gradient.  With the k x k output blocks gathered onto channels in the PARAMETER's order
        # (co*k*k + tap) and the GEMM roles swapped (rows = ci, columns = (co, tap)), the product
        # dW[ci][(co, tap)] = sum_pixels x[ci] * dy[(co, tap)] lands in the [Cin, Cout, k, k] gradient itself:
        # the wgrad kernel red.adds straight into p.grad -- no packed detour for the 151 M-parameter up0.
        rank = None
        if kc % 64 == 0 and n * hin * win <= 64:
            rank = _rank_grad_entry(weight, n, hin, win, cin, x.stride(2), kc)
        if rank is not None:
            # few pixels against a huge weight: queue the operands, one GEMM per optimizer step (see _RankGrad)
            s2c, xq = rank.targets()
            call("dm_space_to_depth", _p(dy), lddy, _p(s2c), s2c.stride(2), n, hin, win, cout, k, 1, st)
            xq.copy_(x)
        else:
            s2c = new_act(n, hin, win, kc, dy.device)
            call("dm_space_to_depth", _p(dy), lddy, _p(s2c), s2c.stride(2), n, hin, win, cout, k, 1, st)
            if kc % 64 == 0:
                call("dm_conv2d_wgrad", _p(s2c), kc, s2c.stride(2), None, 0, 0, _p(x), x.stride(2), _p(grad_buf(weight)),
                     n, hin, win, cin, 1, 1, 1, 0, st)
            else:                       # K columns not a multiple of 64 (e.g. 7x7x16): padded staging buffer
                ckc = r64(kc)
                unpack = (cin, kc, 1, _taps([0]), kc, 1, 0, ckc, ckc, 0)
                _wgrad_into(weight, (cin, ckc), unpack, lambda dwp: call(
                    "dm_conv2d_wgrad", _p(s2c), kc, s2c.stride(2), None, 0, 0, _p(x), x.stride(2), _p(dwp), n, hin, win,
                    cin, 1, 1, 1, 0, st))
        del s2c
        dx = None
        if ctx.needs_input_grad[0]:
            # data gradient: a 1x1 conv over the tap-major gather [N, hin, win, (tap, co)]
            wd = ctx.pack.get(weight, "convt_dgrad")
            s2d = new_act(n, hin, win, kc, dy.device)
            call("dm_space_to_depth", _p(dy), lddy, _p(s2d), s2d.stride(2), n, hin, win, cout, k, 0, st)
            dx = new_act(n, hin, win, cin, dy.device)
            m = n * hin * win
            if m <= 16 and kc % 32 == 0 and kc >= 4096:
                # a handful of pixels against a huge K (up0 on the 2x2 bottleneck: 16 x 98304 x 1536): as a conv this is
                # 9 output tiles; the skinny kernel splits K over the grid and streams the weight pack once at HBM rate
                if dx.shape[3] != cin:
                    dx.zero_()
                scratch = torch.empty(_lib.fn("dm_skinny_gemm_scratch")(cin, kc), device=dy.device, dtype=torch.float32)
                call("dm_skinny_gemm", _p(s2d), s2d.stride(2), _p(wd), wd.stride(0), _p(dx), dx.stride(2), _p(scratch), m, cin,
                     kc, st)
            else:
                call("dm_conv2d_fwd", _p(s2d), kc, s2d.stride(2), None, 0, 0, _p(wd), None, None, 0, _p(dx), dx.stride(2), 0,
                     None, 0, n, hin, win, cin, 1, 1, 1, 0, st)
        return dx, None, None, None, None


def conv_transpose(x, weight, bias, pack, k):
    return _ConvT.apply(x, weight, bias, pack, k)


# --------------------------------------------------------------------------------------- normalisation
class _BnAct(torch.autograd.Function):
    """BatchNorm2d (eps 1e-5, momentum 0.1) + activation (new_scripy.py:185-186)."""

    @staticmethod
    def forward(ctx, y, stats, gamma, beta, rmean, rvar, conv_bias, c, training, act, momentum, eps, bias_outside):
        ldy = _chk(y, "bn input")
        n, h, w, _ = y.shape
        mean = torch.empty(c, device=y.device, dtype=torch.float32)
        invstd = torch.empty(c, device=y.device, dtype=torch.float32)
        st = _stream()
        if training:
            call("dm_bn_finalize", _p(stats), stats.shape[0], stats.shape[2], c, float(n * h * w), _p(mean), _p(invstd),
                 _p(rmean), _p(rvar), momentum, eps, _p(conv_bias) if bias_outside else None, st)
        else:
            call("dm_bn_finalize", None, 0, 0, c, 1.0, _p(mean), _p(invstd), _p(rmean), _p(rvar), momentum, eps, None, st)
        z = torch.empty_like(y)
        call("dm_bn_act_fwd", _p(y), ldy, _p(mean), _p(invstd), _p(gamma), _p(beta), _p(z), z.stride(2), n * h * w, c,
             act, st)
        ctx.save_for_backward(y, mean, invstd, gamma, beta, conv_bias)
        ctx.cfg = (c, training, act)
        return z

    @staticmethod
    def backward(ctx, dz):
        y, mean, invstd, gamma, beta, conv_bias = ctx.saved_tensors
        c, training, act = ctx.cfg
        lddz = _chk(dz, "bn grad")
        n, h, w, _ = y.shape
        dy = torch.empty_like(y)
        npix = n * h * w
        scratch = torch.empty(_lib.fn("dm_bn_act_bwd_scratch")(npix, c), device=y.device, dtype=torch.float32)
        call("dm_bn_act_bwd", _p(dz), lddz, _p(y), y.stride(2), _p(mean), _p(invstd), _p(gamma), _p(beta), _p(dy),
             dy.stride(2), _p(grad_buf(gamma)), _p(grad_buf(beta)),
             _p(grad_buf(conv_bias)) if conv_bias is not None else None, _p(scratch), npix, c, act, int(training),
             _stream())
        return dy, None, None, None, None, None, None, None, None, None, None, None, None


_counters_batched = False     # True while a model forward bumps all num_batches_tracked buffers in one launch
FUSED_CONV_STATS = False      # True: BatchNorm statistics (of the bf16-rounded outputs) from the conv epilogue.  Measured on one box
                              # against the separate dm_bn_stats pass (tools/ab_micro_step.py): -0.5..-1 % per micro-step with
                              # shared-memory atomics (run-to-run different last bits), +-0.1 % with the deterministic
                              # per-lane-quarter slabs that are in the kernel now -- the transposing shuffle reduction
                              # (62 SHFL + 124 SEL per 32 x 32 chunk) costs what the saved pass over y gains.  Off.


class batched_counters:
    """Inside this context ``bn_act`` leaves ``num_batches_tracked`` alone: the caller has already advanced
    every train-mode BatchNorm's counter with one multi-tensor add (41 single-element kernels otherwise)."""

    def __init__(self, module):
        self.bufs = [m.num_batches_tracked for m in module.modules()
                     if isinstance(m, torch.nn.BatchNorm2d) and m.training and m.num_batches_tracked is not None
                     and not getattr(m, "_dm_counts_itself", False)]

    def __enter__(self):
        global _counters_batched
        if self.bufs:
            torch._foreach_add_(self.bufs, 1)
        self.prev, _counters_batched = _counters_batched, True

    def __exit__(self, *exc):
        global _counters_batched
        _counters_batched = self.prev


def bn_stats(y, c):
    """Per-block partial sums of y and y^2 over all pixels (train-mode BatchNorm statistics)."""
    ldy = _chk(y, "bn input")
    n, h, w, _ = y.shape
    rows = _lib.fn("dm_bn_stats_rows")(n * h * w, c)
    part = torch.empty((rows, 2, c), device=y.device, dtype=torch.float32)
    call("dm_bn_stats", _p(y), ldy, _p(part), c, n * h * w, c, _stream())
    return part


def bn_act(y, stats, bn, act, conv_bias=None, bias_outside=False):
    """``conv_bias``: bias parameter of the convolution that produced ``y`` (called with
    ``bias_grad_by_norm=True``); its gradient is produced by this norm's backward.
    ``bias_outside``: train mode only -- that conv did NOT add its bias to ``y`` (batch normalisation
    cancels a per-channel constant exactly); the bias then only enters the tracked running mean."""
    training = bn.training
    if training and bn.num_batches_tracked is not None and not _counters_batched:
        bn.num_batches_tracked.add_(1)
    if training:
        bump_bn_stats_epoch()
    if training and stats is None:
        stats = bn_stats(y, bn.num_features)
    return _BnAct.apply(y, stats, bn.weight, bn.bias, bn.running_mean, bn.running_var, conv_bias, bn.num_features,
                        training, act, float(bn.momentum), float(bn.eps), bool(bias_outside and training))


class _GnAct(torch.autograd.Function):
    """GroupNorm(G, C) + activation (new_scripy.py:167-168,299-300,312-313)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, c, groups, eps, act):
        ldx = _chk(x, "gn input")
        n, h, w, _ = x.shape
        mean = torch.empty(n * groups, device=x.device, dtype=torch.float32)
        rstd = torch.empty(n * groups, device=x.device, dtype=torch.float32)
        scratch = torch.empty(_lib.fn("dm_gn_scratch")(n, h * w, c), device=x.device, dtype=torch.float32)
        z = torch.empty_like(x)
        call("dm_gn_act_fwd", _p(x), ldx, _p(gamma), _p(beta), _p(z), z.stride(2), _p(mean), _p(rstd), _p(scratch), n,
             h * w, c, groups, eps, act, _stream())
        ctx.save_for_backward(x, mean, rstd, gamma, beta, scratch)
        ctx.cfg = (c, groups, act)
        return z

    @staticmethod
    def backward(ctx, dz):
        x, mean, rstd, gamma, beta, scratch = ctx.saved_tensors
        c, groups, act = ctx.cfg
        lddz = _chk(dz, "gn grad")
        n, h, w, _ = x.shape
        dx = torch.empty_like(x)
        call("dm_gn_act_bwd", _p(dz), lddz, _p(x), x.stride(2), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(dx),
             dx.stride(2), _p(grad_buf(gamma)), _p(grad_buf(beta)), _p(scratch), n, h * w, c, groups, act, _stream())
        return dx, None, None, None, None, None, None


def gn_act(x, gn, act):
    return _GnAct.apply(x, gn.weight, gn.bias, gn.num_channels, gn.num_groups, float(gn.eps), act)


# --------------------------------------------------------------------------------------- small MLPs
ACT_SIGMOID = 3      # dm_linear_act_* only


def _param_grad_ptr(p):
    return _p(grad_buf(p)) if p is not None and p.requires_grad else None


def mlp2_fwd(x, w1, b1, w2, b2, act1, act2):
    """act2(act1(x W1^T + b1) W2^T + b2) on fp32 rows [N, Cin] through dm_linear_act_fwd (two launches).
    Returns (y, saved) with saved = what mlp2_bwd needs."""
    n, cin = x.shape
    hid, cout = w1.shape[0], w2.shape[0]
    st = _stream()
    pre1 = torch.empty((n, hid), device=x.device, dtype=torch.float32)
    h = torch.empty_like(pre1)
    y = torch.empty((n, cout), device=x.device, dtype=torch.float32)
    pre2 = torch.empty_like(y) if act2 in (ACT_GELU, ACT_RELU) else None
    call("dm_linear_act_fwd", _p(x), _p(w1), _p(b1), _p(pre1), _p(h), n, cin, hid, act1, st)
    call("dm_linear_act_fwd", _p(h), _p(w2), _p(b2), _p(pre2), _p(y), n, hid, cout, act2, st)
    return y, (x, pre1, h, y if act2 == ACT_SIGMOID else pre2)


def mlp2_bwd(saved, dy, w1, b1, w2, b2, act1, act2, need_dx):
    """Backward of mlp2_fwd: parameter gradients are added straight into the parameters' .grad memory; returns the
    input gradient (or None).  Three launches (two without an input gradient), no atomics."""
    x, pre1, h, aux2 = saved
    n, cin = x.shape
    hid, cout = w1.shape[0], w2.shape[0]
    st = _stream()
    np2 = _lib.fn("dm_linear_bwd_parts")(cout)
    parts2 = torch.empty((np2, n, hid), device=x.device, dtype=torch.float32)
    call("dm_linear_act_bwd", _p(dy), 1, _p(aux2), act2, _p(h), _p(w2), _param_grad_ptr(w2), _param_grad_ptr(b2),
         _p(parts2), n, hid, cout, st)
    np1 = _lib.fn("dm_linear_bwd_parts")(hid)
    parts1 = torch.empty((np1, n, cin), device=x.device, dtype=torch.float32) if need_dx else None
    call("dm_linear_act_bwd", _p(parts2), np2, _p(pre1), act1, _p(x), _p(w1), _param_grad_ptr(w1), _param_grad_ptr(b1),
         _p(parts1), n, cin, hid, st)
    if not need_dx:
        return None
    if np1 == 1:
        return parts1[0]
    dx = torch.empty((n, cin), device=x.device, dtype=torch.float32)
    call("dm_sum_parts", _p(parts1), np1, _p(dx), n * cin, st)
    return dx


def _lin_param(p):
    if p is None:
        return None
    if p.dtype != torch.float32 or not p.is_contiguous():
        raise _lib.DmB200Error("linear parameters must be contiguous fp32")
    return p


class _Mlp2(torch.autograd.Function):
    """EmbedFC.model = Linear - GELU - Linear (new_scripy.py:259-263; MNIST_script.py:107-111) on [N, input_dim] rows."""

    @staticmethod
    def forward(ctx, x, act1, act2, w1, b1, w2, b2):
        x = x.detach().to(torch.float32).contiguous()
        if not x.is_cuda:
            raise _lib.DmB200Error("mlp2: CUDA tensors only (no CPU fallback)")
        y, saved = mlp2_fwd(x, w1.detach(), b1.detach() if b1 is not None else None, w2.detach(),
                            b2.detach() if b2 is not None else None, act1, act2)
        ctx.saved = saved
        ctx.cfg = (act1, act2, w1, b1, w2, b2)
        return y

    @staticmethod
    def backward(ctx, dy):
        act1, act2, w1, b1, w2, b2 = ctx.cfg
        dx = mlp2_bwd(ctx.saved, dy.to(torch.float32).contiguous(), w1, b1, w2, b2, act1, act2, ctx.needs_input_grad[0])
        return dx, None, None, None, None, None, None


def ctx_onehot(c, ctx_mask, n_classes, flip):
    """one_hot(c) * ctx_mask (new_scripy.py:337-340) or one_hot(c) * -(1 - ctx_mask) (``flip``, MNIST_script.py:165-171)
    as fp32 rows [N, n_classes]; no gradient (both inputs are data)."""
    if not c.is_cuda:
        raise _lib.DmB200Error("ctx_onehot: CUDA tensors only (no CPU fallback)")
    c = c.long().contiguous()
    if c.dim() != 1 or ctx_mask.shape != c.shape:
        raise RuntimeError(f"ctx_onehot: labels {tuple(c.shape)} and context mask {tuple(ctx_mask.shape)} must be equal-length vectors")
    if ctx_mask.dtype not in (torch.float32, torch.int64):
        ctx_mask = ctx_mask.to(torch.float32)
    ctx_mask = ctx_mask.contiguous()
    out = torch.empty((c.shape[0], n_classes), device=c.device, dtype=torch.float32)
    call("dm_ctx_onehot", _p(c), _p(ctx_mask), int(ctx_mask.dtype == torch.int64), _p(out), c.shape[0], n_classes,
         int(flip), _stream())
    return out


def embed_fc(x, lin1, lin2):
    """EmbedFC forward on [N, input_dim] rows (new_scripy.py:265-268)."""
    return _Mlp2.apply(x, ACT_GELU, ACT_NONE, _lin_param(lin1.weight), _lin_param(lin1.bias), _lin_param(lin2.weight),
                       _lin_param(lin2.bias))


class _SeResidual(torch.autograd.Function):
    """out = (res + x2 * gate(x2)) * scale, gate = SEBlock MLP on the global average pool
    (new_scripy.py:154-158,196-205).  w1 None: plain (res + x2) * scale (MNIST_script.py:57-61)."""

    @staticmethod
    def forward(ctx, x2, res, c, scale, w1, w2):
        ld2, ldr = _chk(x2, "se x2"), _chk(res, "se residual")
        n, h, w, _ = x2.shape
        st = _stream()
        gate = None
        ctx.small = None
        if w1 is not None:
            pooled = torch.empty((n, c), device=x2.device, dtype=torch.float32)
            call("dm_pool_nhw", _p(x2), ld2, _p(pooled), n, h * w, c, 1.0 / (h * w), st)
            gate, ctx.small = mlp2_fwd(pooled, w1.detach(), None, w2.detach(), None, ACT_GELU, ACT_SIGMOID)
        out = torch.empty_like(x2)
        call("dm_se_apply_fwd", _p(x2), ld2, _p(gate), _p(res), ldr, _p(out), out.stride(2), n, h * w, c, scale, st)
        ctx.save_for_backward(x2, gate)
        ctx.cfg = (c, scale, w1, w2)
        return out

    @staticmethod
    def backward(ctx, dout):
        x2, gate = ctx.saved_tensors
        c, scale, w1, w2 = ctx.cfg
        lddo = _chk(dout, "se grad")
        n, h, w, _ = x2.shape
        st = _stream()
        dpool = None
        if ctx.small is not None:
            dgate = torch.empty((n, c), device=x2.device, dtype=torch.float32)
            call("dm_pool_prod_nhw", _p(dout), lddo, _p(x2), x2.stride(2), _p(dgate), n, h * w, c, scale, st)
            dpool = mlp2_bwd(ctx.small, dgate, w1, None, w2, None, ACT_GELU, ACT_SIGMOID, True)
        dx2 = torch.empty_like(x2)
        dres = torch.empty_like(x2)
        call("dm_se_apply_bwd", _p(dout), lddo, _p(gate), _p(dpool), _p(dx2), dx2.stride(2), _p(dres), dres.stride(2), n,
             h * w, c, scale, st)
        return dx2, dres, None, None, None, None


def se_residual(x2, res, c, scale, se_fc=None):
    if se_fc is None:
        return _SeResidual.apply(x2, res, c, scale, None, None)
    return _SeResidual.apply(x2, res, c, scale, _lin_param(se_fc[0].weight), _lin_param(se_fc[2].weight))


class _CaGatesC(ctypes.Structure):
    """DmCaGates (include/dm_b200.h)."""
    _PTRS = ("xh xw w1_h w1_w b1_h b1_w bn_g_h bn_g_w bn_b_h bn_b_w bn_rm_h bn_rm_w bn_rv_h bn_rv_w wp_h2w wp_w2h "
             "bp_h2w bp_w2h wc_h wc_w bc_h bc_w gamma_h gamma_w alpha beta u part stat t t0 ah aw").split()
    _fields_ = [(k, ctypes.c_void_p) for k in _PTRS] + [(k, ctypes.c_int) for k in ("R", "C", "m", "nblk", "training")] + [
        ("eps", ctypes.c_float), ("momentum", ctypes.c_float)]


class _CaGatesGradC(ctypes.Structure):
    """DmCaGatesGrad (include/dm_b200.h)."""
    _PTRS = ("d_ah d_aw d_xh d_xw dh dt du dz part scal g_w1_h g_w1_w g_b1_h g_b1_w g_bn_g_h g_bn_g_w g_bn_b_h g_bn_b_w g_wp_h2w "
             "g_wp_w2h g_bp_h2w g_bp_w2h g_wc_h g_wc_w g_bc_h g_bc_w g_gamma_h g_gamma_w g_alpha g_beta").split()
    _fields_ = [(k, ctypes.c_void_p) for k in _PTRS]


def _ca_param_map(mod):
    """struct field stem -> parameter of a unet.CoordAttn module."""
    return {"w1_h": mod.conv1_h.weight, "w1_w": mod.conv1_w.weight, "b1_h": mod.conv1_h.bias, "b1_w": mod.conv1_w.bias,
            "bn_g_h": mod.bn1_h.weight, "bn_g_w": mod.bn1_w.weight, "bn_b_h": mod.bn1_h.bias, "bn_b_w": mod.bn1_w.bias,
            "wp_h2w": mod.h2w_proj.weight, "wp_w2h": mod.w2h_proj.weight, "bp_h2w": mod.h2w_proj.bias,
            "bp_w2h": mod.w2h_proj.bias, "wc_h": mod.conv_h.weight, "wc_w": mod.conv_w.weight, "bc_h": mod.conv_h.bias,
            "bc_w": mod.conv_w.bias, "gamma_h": mod.gamma_h, "gamma_w": mod.gamma_w, "alpha": mod.alpha, "beta": mod.beta}


class _CoordAttnFused(torch.autograd.Function):
    """x * (alpha' * a_h + beta' * a_w) (new_scripy.py:97-140): directional pooling (dm_ca_pool), the C/16-wide gate
    network (dm_ca_gates_fwd/bwd, fp32) and the gating pass (dm_ca_gate_fwd/bwd); H == W."""

    @staticmethod
    def forward(ctx, x, c, mod, *params):
        ldx = _chk(x, "coordattn input")
        n, h, w, _ = x.shape
        st = _stream()
        dev = x.device
        m, r = mod.conv1_h.weight.shape[0], n * h
        f32 = dict(device=dev, dtype=torch.float32)
        xh, xw = torch.empty((n, h, c), **f32), torch.empty((n, w, c), **f32)
        call("dm_ca_pool", _p(x), ldx, None, 0, _p(xh), _p(xw), n, h, w, c, 1.0 / w, 1.0 / h, st)
        nblk = -(-r // _lib.fn("dm_ca_gates_rows_per_block")())
        training = mod.bn1_h.training
        work = {"u": torch.empty((2, r, m), **f32), "part": torch.empty((nblk, 2, 2, m), **f32),
                "stat": torch.empty((2, 2, m), **f32), "t": torch.empty((2, r, m), **f32), "t0": torch.empty((2, r, m), **f32),
                "ah": torch.empty((n, h, c), **f32), "aw": torch.empty((n, w, c), **f32), "xh": xh, "xw": xw}
        pm = _ca_param_map(mod)
        for k, v in pm.items():
            if v.dtype != torch.float32 or not v.is_contiguous():
                raise _lib.DmB200Error(f"CoordAttn parameter {k}: expected contiguous fp32")
        S = _CaGatesC()
        for k, v in {**pm, **work, "bn_rm_h": mod.bn1_h.running_mean, "bn_rm_w": mod.bn1_w.running_mean,
                     "bn_rv_h": mod.bn1_h.running_var, "bn_rv_w": mod.bn1_w.running_var}.items():
            setattr(S, k, v.data_ptr())
        S.R, S.C, S.m, S.nblk, S.training = r, c, m, nblk, int(training)
        S.eps, S.momentum = float(mod.bn1_h.eps), float(mod.bn1_h.momentum)
        if training:
            bump_bn_stats_epoch()
            if not _counters_batched:
                torch._foreach_add_([mod.bn1_h.num_batches_tracked, mod.bn1_w.num_batches_tracked], 1)
        call("dm_ca_gates_fwd", ctypes.addressof(S), st)
        out = torch.empty_like(x)
        call("dm_ca_gate_fwd", _p(x), ldx, _p(work["ah"]), _p(work["aw"]), _p(out), out.stride(2), n, h, w, c, st)
        ctx.save_for_backward(x)
        ctx.ca = (c, mod, S, work, pm)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        c, mod, S, work, pm = ctx.ca
        lddo = _chk(dout, "coordattn grad")
        n, h, w, _ = x.shape
        st = _stream()
        f32 = dict(device=x.device, dtype=torch.float32)
        dah, daw = torch.empty((n, h, c), **f32), torch.empty((n, w, c), **f32)
        call("dm_ca_pool", _p(dout), lddo, _p(x), x.stride(2), _p(dah), _p(daw), n, h, w, c, 1.0, 1.0, st)
        scal = mod.__dict__.get("_dm_ca_scal")
        if scal is None or scal.device != x.device:
            scal = mod.__dict__["_dm_ca_scal"] = torch.zeros(4, **f32)      # the backward kernels re-arm it
        G = _CaGatesGradC()
        tmp = {"d_ah": dah, "d_aw": daw, "d_xh": torch.empty((n, h, c), **f32), "d_xw": torch.empty((n, w, c), **f32),
               "dh": torch.empty_like(work["u"]), "dt": torch.empty_like(work["u"]), "du": torch.empty_like(work["u"]),
               "dz": torch.empty((2, n * h, c), **f32),
               "part": torch.empty_like(work["part"]), "scal": scal}
        for k, v in tmp.items():
            setattr(G, k, v.data_ptr())
        for k, v in pm.items():
            setattr(G, "g_" + k, grad_buf(v).data_ptr())
        call("dm_ca_gates_bwd", ctypes.addressof(S), ctypes.addressof(G), st)
        dx = torch.empty_like(x)
        call("dm_ca_gate_bwd", _p(dout), lddo, _p(work["ah"]), _p(work["aw"]), _p(tmp["d_xh"]), _p(tmp["d_xw"]), _p(dx),
             dx.stride(2), n, h, w, c, st)
        return (dx, None, None) + (None,) * len(pm)


def coord_attn(x, c, mod):
    """CoordAttn.forward (new_scripy.py:97-140) on a square feature map.  H != W would need the adaptive_avg_pool1d
    resampling of the h<->w cross terms (:118-126), which the gate kernels do not implement: fail loudly."""
    if x.shape[1] != x.shape[2]:
        raise _lib.DmB200Error(f"CoordAttn: square feature maps only (got {x.shape[1]}x{x.shape[2]}); the U-Net's inputs are "
                               "square (Cfg.IMG_SIZE, new_scripy.py:25)")
    pm = _ca_param_map(mod)
    return _CoordAttnFused.apply(x, c, mod, *pm.values())


# --------------------------------------------------------------------------------------- resampling / glue
class _Upcat(torch.autograd.Function):
    """cat((a, b), 1) -> Upsample(x2, bilinear, align_corners=True) (new_scripy.py:242,251)."""

    @staticmethod
    def forward(ctx, a, b, ca, cb):
        lda, ldb = _chk(a, "upcat a"), _chk(b, "upcat b")
        n, h, w, _ = a.shape
        out = new_act(n, 2 * h, 2 * w, ca + cb, a.device)
        call("dm_upcat_fwd", _p(a), lda, ca, _p(b), ldb, cb, _p(out), out.stride(2), n, h, w, _stream())
        ctx.cfg = (ca, cb, n, h, w)
        return out

    @staticmethod
    def backward(ctx, dout):
        ca, cb, n, h, w = ctx.cfg
        lddo = _chk(dout, "upcat grad")
        da = new_act(n, h, w, ca, dout.device)
        db = new_act(n, h, w, cb, dout.device)
        call("dm_upcat_bwd", _p(dout), lddo, _p(da), da.stride(2), ca, _p(db), db.stride(2), cb, n, h, w, _stream())
        return da, db, None, None


def _upcat_shared(a, b, ca, cb):
    """Forward-only upcat with a skip tensor ``b`` of fewer samples than ``a``, shared cyclically (sample n reads
    b[n % len(b)]): the CFG halves of a sampling batch share the encoder's skip tensors without duplicating them."""
    if torch.is_grad_enabled() and (a.requires_grad or b.requires_grad):
        raise _lib.DmB200Error("upcat: a shared skip tensor (fewer samples than the batch) is forward-only")
    lda, ldb = _chk(a, "upcat a"), _chk(b, "upcat b")
    n, h, w, _ = a.shape
    if n % b.shape[0] or b.shape[1:3] != a.shape[1:3]:
        raise _lib.DmB200Error(f"upcat: skip batch {tuple(b.shape)} does not tile the batch {tuple(a.shape)}")
    out = new_act(n, 2 * h, 2 * w, ca + cb, a.device)
    call("dm_upcat_fwd_shared", _p(a), lda, ca, _p(b), ldb, cb, b.shape[0], _p(out), out.stride(2), n, h, w, _stream())
    return out


def upcat(a, b, ca, cb):
    if b.shape[0] != a.shape[0]:
        return _upcat_shared(a, b, ca, cb)
    return _Upcat.apply(a, b, ca, cb)


class _Film(torch.autograd.Function):
    """cemb * x + temb broadcast over pixels (new_scripy.py:348-349)."""

    @staticmethod
    def forward(ctx, x, ce, te, c):
        ldx = _chk(x, "film input")
        n, h, w, _ = x.shape
        ce, te = ce.contiguous().float(), te.contiguous().float()
        if tuple(ce.shape) != (n, c) or tuple(te.shape) != (n, c):          # the reference's broadcast would raise here too
            raise RuntimeError(f"film: embeddings {tuple(ce.shape)} / {tuple(te.shape)} do not match the batch ({n}, {c})")
        out = torch.empty_like(x)
        call("dm_film_fwd", _p(x), ldx, _p(ce), _p(te), _p(out), out.stride(2), n, h * w, c, _stream())
        ctx.save_for_backward(x, ce)
        ctx.c = c
        return out

    @staticmethod
    def backward(ctx, dout):
        x, ce = ctx.saved_tensors
        c = ctx.c
        lddo = _chk(dout, "film grad")
        n, h, w, _ = x.shape
        dx = torch.empty_like(x)
        dce = torch.empty((n, c), device=x.device, dtype=torch.float32)
        dte = torch.empty((n, c), device=x.device, dtype=torch.float32)
        call("dm_film_bwd", _p(dout), lddo, _p(x), x.stride(2), _p(ce), _p(dx), dx.stride(2), _p(dce), _p(dte), n, h * w,
             c, _stream())
        return dx, dce, dte, None


def film(x, ce, te, c):
    return _Film.apply(x, ce, te, c)


class _AvgPoolAct(torch.autograd.Function):
    """AvgPool2d(k) + GELU (new_scripy.py:290; MNIST_script.py:132)."""

    @staticmethod
    def forward(ctx, x, c, k, act):
        ldx = _chk(x, "avgpool input")
        n, h, w, _ = x.shape
        out = new_act(n, h // k, w // k, c, x.device)
        call("dm_avgpool_act_fwd", _p(x), ldx, _p(out), out.stride(2), n, h, w, c, k, act, _stream())
        ctx.save_for_backward(x)
        ctx.cfg = (c, k, act)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        c, k, act = ctx.cfg
        lddo = _chk(dout, "avgpool grad")
        n, h, w, _ = x.shape
        dx = torch.zeros_like(x) if (h % k or w % k) else torch.empty_like(x)
        call("dm_avgpool_act_bwd", _p(dout), lddo, _p(x), x.stride(2), _p(dx), dx.stride(2), n, h, w, c, k, act, _stream())
        return dx, None, None, None


def avgpool_act(x, c, k, act):
    return _AvgPoolAct.apply(x, c, k, act)


class _MaxPool2(torch.autograd.Function):
    """MaxPool2d(2) (MNIST_script.py:74)."""

    @staticmethod
    def forward(ctx, x, c):
        ldx = _chk(x, "maxpool input")
        n, h, w, _ = x.shape
        out = new_act(n, h // 2, w // 2, c, x.device)
        call("dm_maxpool2_fwd", _p(x), ldx, _p(out), out.stride(2), n, h, w, c, _stream())
        ctx.save_for_backward(x)
        ctx.c = c
        return out

    @staticmethod
    def backward(ctx, dout):
        (x,) = ctx.saved_tensors
        lddo = _chk(dout, "maxpool grad")
        n, h, w, _ = x.shape
        dx = torch.zeros_like(x) if (h % 2 or w % 2) else torch.empty_like(x)
        call("dm_maxpool2_bwd", _p(dout), lddo, _p(x), x.stride(2), _p(dx), dx.stride(2), n, h, w, ctx.c, _stream())
        return dx, None


def maxpool2(x, c):
    return _MaxPool2.apply(x, c)


class _MaskFma(torch.autograd.Function):
    """x + y * (mask > thresh): the LocalEnhancer attention-mask weighting (new_scripy.py:173-174)."""

    @staticmethod
    def forward(ctx, x, y, mask, thresh, c):
        ldx, ldy = _chk(x, "mask_fma x"), _chk(y, "mask_fma y")
        n, h, w, _ = x.shape
        mask = mask.contiguous().float()
        out = torch.empty_like(x)
        call("dm_mask_fma", _p(x), ldx, _p(y), ldy, _p(mask), thresh, _p(out), out.stride(2), n * h * w, c, _stream())
        ctx.save_for_backward(mask)
        ctx.cfg = (thresh, c)
        return out

    @staticmethod
    def backward(ctx, dout):
        (mask,) = ctx.saved_tensors
        thresh, c = ctx.cfg
        lddo = _chk(dout, "mask_fma grad")
        n, h, w, _ = dout.shape
        dy = torch.empty_like(dout)
        call("dm_mask_fma", None, 0, _p(dout), lddo, _p(mask), thresh, _p(dy), dy.stride(2), n * h * w, c, _stream())
        return dout, dy, None, None, None


def mask_fma(x, y, mask, thresh, c):
    return _MaskFma.apply(x, y, mask, thresh, c)


class _Fork(torch.autograd.Function):
    """Use one activation twice; the two incoming gradients are summed by one fused kernel."""

    @staticmethod
    def forward(ctx, x, c):
        ctx.c = c
        return x.view_as(x), x.view_as(x)

    @staticmethod
    def backward(ctx, g1, g2):
        if g1 is None or g2 is None:
            return (g1 if g2 is None else g2), None
        ld1, ld2 = _chk(g1, "fork grad"), _chk(g2, "fork grad")
        n, h, w, _ = g1.shape
        out = new_act(n, h, w, ctx.c, g1.device)
        call("dm_axpby", _p(g1), ld1, _p(g2), ld2, _p(out), out.stride(2), n * h * w, ctx.c, 1.0, 1.0, _stream())
        return out, None


def fork(x, c):
    return _Fork.apply(x, c)


# --------------------------------------------------------------------------------------- layout boundary
class _ToNhwc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if x.device.type != "cuda":
            raise _lib.DmB200Error("the hot path runs on CUDA only; there is no CPU fallback")
        x = x.contiguous().float()
        n, c, h, w = x.shape
        y = new_act(n, h, w, c, x.device)
        call("dm_nchw_to_nhwc", _p(x), _p(y), y.stride(2), 0, n, c, h, w, _stream())
        ctx.c = c
        return y

    @staticmethod
    def backward(ctx, dy):
        ld = _chk(dy, "input grad")
        n, h, w, _ = dy.shape
        dx = torch.empty((n, ctx.c, h, w), device=dy.device, dtype=torch.float32)
        call("dm_nhwc_to_nchw", _p(dy), 0, ld, _p(dx), n, ctx.c, h, w, _stream())
        return dx


def to_nhwc(x):
    return _ToNhwc.apply(x)


class _ToNchwF32(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, c):
        n, h, w, ld = y.shape
        out = torch.empty((n, c, h, w), device=y.device, dtype=torch.float32)
        call("dm_nhwc_to_nchw", _p(y), int(y.dtype == torch.float32), y.stride(2), _p(out), n, c, h, w, _stream())
        ctx.cfg = (c, ld, y.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        c, ld, dtype = ctx.cfg
        dout = dout.contiguous().float()
        n, _, h, w = dout.shape
        dy = torch.empty((n, h, w, ld), device=dout.device, dtype=dtype)
        call("dm_nchw_to_nhwc", _p(dout), _p(dy), ld, int(dtype == torch.float32), n, c, h, w, _stream())
        return dy, None


def to_nchw_f32(y, c):
    return _ToNchwF32.apply(y, c)


# --------------------------------------------------------------------------------------- DDPM kernels
LOSS_CFG = dict(hi_t=1.2, mid_t=0.8, hi_w=3.0, mid_w=1.0, lo_w=0.5, fcw=2.0)   # new_scripy.py:31-36


def q_sample(x, noise, sqrtab, sqrtmab, ts):
    """x_t as the bf16 NHWC tensor the first conv consumes (new_scripy.py:408-411)."""
    n, c, h, w = x.shape
    if noise.shape != x.shape or ts.numel() != n or not x.is_cuda:
        raise RuntimeError(f"q_sample: x {tuple(x.shape)}, noise {tuple(noise.shape)}, timesteps {tuple(ts.shape)} do not match "
                           "(or are not CUDA tensors: there is no CPU fallback)")
    xt = new_act(n, h, w, c, x.device)
    call("dm_q_sample", _p(x.contiguous()), _p(noise.contiguous()), _p(sqrtab), _p(sqrtmab), _p(ts.contiguous()), _p(xt),
         xt.stride(2), n, c, h, w, _stream())
    return xt


class _DdpmLoss(torch.autograd.Function):
    """Attention-weighted noise-prediction loss (new_scripy.py:417-437); mask=None: MSE (MNIST_script.py:252)."""

    @staticmethod
    def forward(ctx, pred, noise, mask, cfg):
        n, h, w, ldp = pred.shape
        c = noise.shape[1]
        loss = torch.empty((), device=pred.device, dtype=torch.float32)
        scratch = torch.empty(8, device=pred.device, dtype=torch.float32)
        vals = [cfg[k] for k in ("hi_t", "mid_t", "hi_w", "mid_w", "lo_w", "fcw")]
        call("dm_ddpm_loss_fwd", _p(pred), pred.stride(2), _p(noise), _p(mask), _p(loss), _p(scratch), n, c, h, w, *vals,
             _stream())
        ctx.save_for_backward(pred, noise, mask)
        ctx.vals = vals
        return loss

    @staticmethod
    def backward(ctx, gout):
        pred, noise, mask = ctx.saved_tensors
        n, h, w, _ = pred.shape
        c = noise.shape[1]
        dpred = torch.empty_like(pred)
        gout = gout.contiguous().float()
        call("dm_ddpm_loss_bwd", _p(pred), pred.stride(2), _p(noise), _p(mask), _p(gout), _p(dpred), dpred.stride(2), n, c,
             h, w, *ctx.vals, _stream())
        return dpred, None, None, None


def ddpm_loss(pred_nhwc_f32, noise, mask):
    noise = noise.contiguous().float()
    mask = mask.contiguous().float() if mask is not None else None
    n, h, w, _ = pred_nhwc_f32.shape
    if noise.dim() != 4 or (noise.shape[0], noise.shape[2], noise.shape[3]) != (n, h, w) or (
            mask is not None and tuple(mask.shape) != (n, h, w)):
        raise RuntimeError(f"ddpm_loss: prediction {tuple(pred_nhwc_f32.shape)} (NHWC), noise {tuple(noise.shape)} and attention "
                           f"mask {None if mask is None else tuple(mask.shape)} do not match")
    return _DdpmLoss.apply(pred_nhwc_f32, noise, mask, LOSS_CFG)


def cfg_reverse_step(eps_nhwc_f32, x, z, guide_w, a, b, s, want_next=True):
    """One classifier-free-guided reverse step (new_scripy.py:468-475).  Returns (x_next fp32 NCHW,
    doubled bf16 NHWC batch for the next denoiser call)."""
    n, c, h, w = x.shape
    x_out = torch.empty_like(x)
    xt = new_act(2 * n, h, w, c, x.device) if want_next else None
    call("dm_cfg_reverse_step", _p(eps_nhwc_f32), eps_nhwc_f32.stride(2), _p(x), _p(z), _p(x_out), _p(xt),
         xt.stride(2) if xt is not None else 8, float(guide_w), float(a), float(b), float(s), n, c, h, w, _stream())
    return x_out, xt


def cfg_reverse_step_dev(eps_nhwc_f32, x, z, coef4, x_out, xt_out):
    """Graph-capturable reverse step: (guide_w, oneover_sqrta, mab_over_sqrtmab, sqrt_beta) come from the 4-float
    device tensor ``coef4``; writes ``x_out`` (fp32 NCHW) and ``xt_out`` (doubled bf16 NHWC batch)."""
    n, c, h, w = x.shape
    call("dm_cfg_reverse_step_dev", _p(eps_nhwc_f32), eps_nhwc_f32.stride(2), _p(x), _p(z), _p(x_out), _p(xt_out),
         xt_out.stride(2), _p(coef4), n, c, h, w, _stream())


def cfg_reverse_step_w(eps_nhwc_f32, x, z, coef4, wvec, x_out, xt_out):
    """Graph-capturable reverse step with one guidance scale per trajectory: ``wvec`` [n] fp32 on the device,
    ``coef4[1:]`` = (oneover_sqrta, mab_over_sqrtmab, sqrt_beta)."""
    n, c, h, w = x.shape
    if eps_nhwc_f32.shape[0] != 2 * n or z.shape != x.shape or wvec.numel() != n or coef4.numel() != 4:
        raise RuntimeError("cfg_reverse_step_w: eps must hold 2n samples, z match x, one guidance scale per trajectory")
    call("dm_cfg_reverse_step_w", _p(eps_nhwc_f32), eps_nhwc_f32.stride(2), _p(x), _p(z), _p(x_out), _p(xt_out),
         xt_out.stride(2), _p(coef4), _p(wvec), n, c, h, w, _stream())


# --------------------------------------------------------------------------------------- profiling hook
class _Profile:
    """CUDA-event timing of every C-ABI call on the launching stream (bench.py's roofline numbers)."""

    def __init__(self):
        self.records = []

    @staticmethod
    def _flops(name, a):
        if name in ("dm_conv2d_fwd", "dm_conv2d_wgrad"):
            if name == "dm_conv2d_fwd":
                c0, c1, n, hin, win, cout, kh, kw, stride, pad = a[1], a[4], a[15], a[16], a[17], a[18], a[19], a[20], a[21], a[22]
            else:
                c0, c1, n, hin, win, cout, kh, kw, stride, pad = a[1], a[4], a[9], a[10], a[11], a[12], a[13], a[14], a[15], a[16]
            ho, wo = (hin + 2 * pad - kh) // stride + 1, (win + 2 * pad - kw) // stride + 1
            return 2.0 * n * ho * wo * cout * (c0 + c1) * kh * kw
        if name == "dm_conv2d_s2_dgrad":
            return 2.0 * a[7] * a[8] * a[9] * a[1] * a[5] * 16
        if name == "dm_convt_fwd":
            return 2.0 * a[7] * a[8] * a[9] * a[1] * a[10] * a[11] * a[11]
        return 0.0

    @staticmethod
    def _bytes(name, a):
        """Algorithmic HBM bytes of one call of a bandwidth-bound entry point (DESIGN.md section 3 per-element figures:
        every operand read once, every result written once, bf16 activations), 0 for the GEMM entries."""
        E = lambda *idx: float(_prod(a[i] for i in idx))
        if name == "dm_bn_stats":
            return 2 * E(4, 5)
        if name == "dm_bn_act_fwd":
            return 4 * E(8, 9)
        if name == "dm_bn_act_bwd":
            return 10 * E(14, 15)
        if name == "dm_gn_act_fwd":
            return 6 * E(9, 10, 11)
        if name == "dm_gn_act_bwd":
            return 10 * E(13, 14, 15)
        if name == "dm_pool_nhw":
            return 2 * E(3, 4, 5)
        if name == "dm_pool_prod_nhw":
            return 4 * E(5, 6, 7)
        if name == "dm_se_apply_fwd":
            return 6 * E(7, 8, 9)
        if name == "dm_se_apply_bwd":
            return 6 * E(8, 9, 10)
        if name == "dm_ca_pool":
            return (2 if a[2] is None else 4) * E(6, 7, 8, 9)
        if name == "dm_ca_gate_fwd":
            return 4 * E(6, 7, 8, 9)
        if name == "dm_ca_gate_bwd":
            return 4 * E(8, 9, 10, 11)
        if name == "dm_upcat_fwd":
            return 10 * E(8, 9, 10) * (a[2] + a[5])
        if name == "dm_upcat_bwd":
            return 10 * E(8, 9, 10) * (a[4] + a[7])
        if name == "dm_film_fwd":
            return 4 * E(6, 7, 8)
        if name == "dm_film_bwd":
            return 6 * E(9, 10, 11)
        if name == "dm_avgpool_act_fwd":
            return 2 * E(4, 5, 6, 7) * (1 + 1.0 / (a[8] * a[8]))
        if name == "dm_avgpool_act_bwd":
            return 2 * E(6, 7, 8, 9) * (2 + 1.0 / (a[10] * a[10]))
        if name == "dm_maxpool2_fwd":
            return 2 * E(4, 5, 6, 7) * 1.25
        if name == "dm_maxpool2_bwd":
            return 2 * E(6, 7, 8, 9) * 2.25
        if name == "dm_mask_fma":
            return (4 if a[0] is None else 6) * E(8, 9)
        if name == "dm_axpby":
            return (4 if a[2] is None else 6) * E(6, 7)
        if name == "dm_colsum":
            return 2 * E(3, 4)
        if name == "dm_q_sample":
            return 10 * E(7, 8, 9, 10)
        if name == "dm_ddpm_loss_fwd":
            return (8 + 4.0 / a[7]) * E(6, 7, 8, 9)
        if name == "dm_ddpm_loss_bwd":
            return (12 + 4.0 / a[8]) * E(7, 8, 9, 10)
        if name == "dm_cfg_reverse_step":
            return 20 * E(11, 12, 13, 14) + 4 * a[6] * E(11, 13, 14)
        if name == "dm_cfg_reverse_step_dev":
            return 20 * E(8, 9, 10, 11) + 4 * a[6] * E(8, 10, 11)
        if name == "dm_cfg_reverse_step_w":
            return 20 * E(9, 10, 11, 12) + 4 * a[6] * E(9, 11, 12)
        if name == "dm_im2col3x3":
            return (2 * a[1] + 64) * E(3, 4, 5)
        if name == "dm_space_to_depth":
            return 4 * E(4, 5, 6, 7, 8, 8)
        if name in ("dm_nchw_to_nhwc", "dm_nhwc_to_nchw"):
            return 6 * E(4, 5, 6, 7) if name == "dm_nchw_to_nhwc" else 6 * E(4, 5, 6, 7)
        if name == "dm_cast_nhwc":
            return 6 * E(4, 5)
        if name == "dm_linear_act_fwd":
            return 4 * E(6, 7)
        if name == "dm_linear_act_bwd":
            return 12 * E(10, 11)
        if name == "dm_skinny_gemm":
            return 2 * E(8, 9)
        if name == "dm_sumsq":
            return 4 * E(1)
        if name == "dm_adamw":
            return 28 * E(4)
        if name == "dm_adamw_bf16":
            return 30 * E(5)
        if name == "dm_pack_transpose":
            return 4 * E(2, 3, 4)
        if name == "dm_pack_weight":
            return 6 * E(2, 3, 4)
        if name == "dm_unpack_wgrad":
            return 12 * E(2, 3, 4)
        if name in ("dm_ca_gates_fwd", "dm_ca_gates_bwd"):
            g = _CaGatesC.from_address(a[0])
            once = 16.0 * g.C * g.m + 16.0 * g.R * g.C          # both directions: two C x m weights, rows in, rows out (fp32)
            return once if name == "dm_ca_gates_fwd" else 3 * once
        return 0.0

    def call(self, name, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = _lib.call(name, *args)
        e1.record()
        fl = self._flops(name, args)
        self.records.append((name, fl, 0.0 if fl else self._bytes(name, args), e0, e1))
        return rc

    def summary(self):
        groups = {"dm_conv2d_fwd": "conv_gemm", "dm_conv2d_s2_dgrad": "conv_gemm", "dm_convt_fwd": "conv_gemm",
                  "dm_conv2d_wgrad": "wgrad_gemm"}
        out = {}
        for name, fl, nb, e0, e1 in self.records:
            g = groups.get(name, name)
            d = out.setdefault(g, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
            d["ms"] += e0.elapsed_time(e1)
            d["flops"] += fl
            d["bytes"] += nb
            d["n"] += 1
        return out


def _prod(it):
    r = 1
    for v in it:
        r *= v
    return r


def enable_profile():
    global call
    prof = _Profile()
    call = prof.call
    return prof


def disable_profile():
    global call
    call = _lib.call
