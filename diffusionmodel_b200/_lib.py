"""ctypes binding of libdm_b200.so (the C-ABI declared in include/dm_b200.h).

There is no CPU fallback: importing the package works without a GPU (so the CPU test-suite can
check the exported symbols), but any compute call without the library or without a CUDA device
raises immediately.
"""
from __future__ import annotations

import ctypes
import os
import re

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DM_B200_LIB") or os.path.join(_PKG, "libdm_b200.so")    # env override: A/B runs of two builds
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "dm_b200.h")

_lib = None


class DmB200Error(RuntimeError):
    pass


def declared_symbols() -> list:
    """Every function name declared in include/dm_b200.h."""
    with open(HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dm_[a-z0-9_]+)\s*\(", src)))


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DmB200Error(
                f"{LIB_PATH} is missing: build it with `python -m diffusionmodel_b200.build` "
                "(there is no CPU or PyTorch fallback for the hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.dm_last_error.restype = ctypes.c_char_p
        _lib.dm_debug_set.argtypes = [ctypes.c_int, ctypes.c_longlong]
        _lib.dm_debug_set.restype = None
        _lib.dm_launch_count.restype = ctypes.c_longlong
    return _lib


_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_longlong
_F = ctypes.c_float
_D = ctypes.c_double

# argument signatures, in header order (p = pointer, i = int, l = long long, f = float, d = double)
_SIGS = {
    "dm_conv2d_fwd": "pii pii p p pi pii pi iiiiiiii p",
    "dm_conv2d_fwd_stat_rows": "iiii",
    "dm_conv2d_fwd_stats_max_cout": "",
    "dm_conv2d_s2_dgrad": "pii p pii iii p",
    "dm_convt_fwd": "pii p p pi iiiii p",
    "dm_conv2d_wgrad": "pii pii pi p iiiiiiii p",
    "dm_pack_weight": "p p iii p ll iil i p",
    "dm_unpack_wgrad": "p p iii p ll iil i i p",
    "dm_nchw_to_nhwc": "p pi i iiii p",
    "dm_cast_nhwc": "pi pi l i p",
    "dm_nhwc_to_nchw": "pii p iiii p",
    "dm_im2col3x3": "pi p iiii p",
    "dm_space_to_depth": "pi pi iiiii i p",
    "dm_bn_stats_rows": "li",
    "dm_bn_stats": "pi pi l i p",
    "dm_bn_finalize": "p iii d pp pp ff p p",
    "dm_bn_act_fwd": "pi pppp pi l ii p",
    "dm_bn_act_bwd": "pi pi pppp pi pp p p l iii p",
    "dm_bn_act_bwd_scratch": "li",
    "dm_gn_scratch": "iii",
    "dm_gn_act_fwd": "pi pp pi pp p iiii f i p",
    "dm_gn_act_bwd": "pi pi pppp pi pp p iiii i p",
    "dm_pool_nhw": "pi p iii f p",
    "dm_pool_prod_nhw": "pi pi p iii f p",
    "dm_se_apply_fwd": "pi p pi pi iii f p",
    "dm_se_apply_bwd": "pi p p pi pi iii f p",
    "dm_linear_act_fwd": "ppp pp iii i p",
    "dm_ctx_onehot": "pp i p iii p",
    "dm_linear_bwd_parts": "i",
    "dm_linear_act_bwd": "pi pi pp pp p iii p",
    "dm_sum_parts": "pi p l p",
    "dm_skinny_gemm_scratch": "ii",
    "dm_skinny_gemm": "pl pl pi p iii p",
    "dm_ca_pool": "pi pi pp iiii ff p",
    "dm_ca_gate_fwd": "pi pp pi iiii p",
    "dm_ca_gate_bwd": "pi pp pp pi iiii p",
    "dm_upcat_fwd": "pii pii pi iii p",
    "dm_upcat_fwd_shared": "pii pii i pi iii p",
    "dm_upcat_bwd": "pi pii pii iii p",
    "dm_film_fwd": "pi pp pi iii p",
    "dm_film_bwd": "pi pi p pi pp iii p",
    "dm_avgpool_act_fwd": "pi pi iiiii i p",
    "dm_avgpool_act_bwd": "pi pi pi iiiii i p",
    "dm_maxpool2_fwd": "pi pi iiii p",
    "dm_maxpool2_bwd": "pi pi pi iiii p",
    "dm_mask_fma": "pi pi p f pi l i p",
    "dm_axpby": "pi pi pi l i ff p",
    "dm_colsum": "pi p l i p",
    "dm_q_sample": "pp pp p pi iiii p",
    "dm_ddpm_loss_fwd": "pi p p p p iiii ffffff p",
    "dm_ddpm_loss_bwd": "pi p p p pi iiii ffffff p",
    "dm_cfg_reverse_step": "pi p p p pi ffff iiii p",
    "dm_cfg_reverse_step_dev": "pi p p p pi p iiii p",
    "dm_cfg_reverse_step_w": "pi p p p pi p p iiii p",
    "dm_ca_gates_rows_per_block": "",
    "dm_ca_gates_fwd": "p p",
    "dm_ca_gates_bwd": "p p p",
    "dm_prep_batch": "ppp pp iii fffff p",
    "dm_image_metrics": "ppp i l p",
    "dm_sumsq": "p l p p",
    "dm_pack_transpose": "pp iii p l p",
    "dm_adamw": "pppp l fffffff p f p",
    "dm_adamw_bf16": "ppppp l fffffff p f p",
}
_RET_LL = {"dm_bn_act_bwd_scratch", "dm_gn_scratch", "dm_skinny_gemm_scratch"}          # size queries return long long, not a status
_CT = {"p": _P, "i": _I, "l": _L, "f": _F, "d": _D}
_bound = {}


def fn(name):
    f = _bound.get(name)
    if f is None:
        f = getattr(lib(), name)
        f.argtypes = [_CT[ch] for ch in _SIGS[name].replace(" ", "")]
        f.restype = _L if name in _RET_LL else _I
        _bound[name] = f
    return f


_workspace = None


def _ensure_workspace():
    """One 32 MiB device scratch buffer for the reduction kernels' partial sums (dm_set_workspace).  It lives on the
    device that is current at the first call and is used in stream order: the library serves ONE device and ONE compute
    stream per process (the data-parallel layout is one process per GPU); ops._chk rejects tensors of another device."""
    global _workspace
    if _workspace is None:
        import torch
        _workspace = torch.empty(32 << 20, dtype=torch.uint8, device="cuda")
        f = lib().dm_set_workspace
        f.argtypes, f.restype = [_P, _L], _I
        f(ctypes.c_void_p(_workspace.data_ptr()), _workspace.numel())


def call(name, *args):
    if _workspace is None:
        _ensure_workspace()
    rc = fn(name)(*args)
    if rc != 0:
        msg = lib().dm_last_error().decode(errors="replace")
        raise DmB200Error(f"{name} failed (rc={rc}): {msg}")
    return rc


def launch_count() -> int:
    return int(lib().dm_launch_count())


def kernel_count(name: str) -> int:
    """Launches so far of the kernel variant ``name`` (dm_kernel_count)."""
    f = lib().dm_kernel_count
    f.argtypes, f.restype = [ctypes.c_char_p], ctypes.c_longlong
    return int(f(name.encode()))


def last_kernel():
    """(name, parameter) of the most recent conv / weight-gradient / skinny launch (dm_last_kernel)."""
    f = lib().dm_last_kernel
    f.argtypes, f.restype = [ctypes.POINTER(ctypes.c_int)], ctypes.c_char_p
    v = ctypes.c_int(0)
    return f(ctypes.byref(v)).decode(), int(v.value)


class kernel_counts:
    """Context manager: ``with kernel_counts() as k: ...; k["conv3x3_halo2"]`` = launches of that variant inside."""
    NAMES = ("conv_gemm", "conv_gemm2", "conv3x3_halo", "conv3x3_halo2", "wgrad_gemm", "wgrad2_gemm", "wgrad3_pair", "skinny_gemm")

    def __enter__(self):
        self._before = {n: kernel_count(n) for n in self.NAMES}
        self.delta = {}
        return self

    def __exit__(self, *exc):
        self.delta = {n: kernel_count(n) - self._before[n] for n in self.NAMES}

    def __getitem__(self, name):
        return self.delta[name]


class sm_limit:
    """``with sm_limit(132): ...``: the persistent GEMM grids launched (or graph-captured) inside leave SMs free for a
    kernel running concurrently on another stream (dm_set_sm_limit)."""

    def __init__(self, sms):
        self.sms = int(sms)

    def __enter__(self):
        if lib().dm_set_sm_limit(self.sms) != 0:
            raise DmB200Error(lib().dm_last_error().decode())

    def __exit__(self, *exc):
        lib().dm_set_sm_limit(148)


def debug_set(key: int, value: int) -> None:
    lib().dm_debug_set(key, value)
