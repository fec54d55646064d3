"""Optimizer side of the train step: global grad-norm clip + AdamW as two kernels over flat buffers.

Replaces ``scaler.unscale_ / clip_grad_norm_(params, 1.0) / AdamW.step`` (new_scripy.py:797-803):
all parameters are re-homed as views of one flat fp32 buffer (names, shapes and state_dict unchanged),
their gradients as views of a second one, so the clip and the update are one ``dm_sumsq`` and one
``dm_adamw`` launch and a data-parallel gradient all-reduce is one NCCL call on one tensor.
"""
from __future__ import annotations

import ctypes

import torch

from . import ops


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class FusedAdamW(torch.optim.Optimizer):
    """AdamW (decoupled weight decay, bias correction: torch.optim.AdamW semantics) with optional
    global-norm clipping folded into the update.  ``param_groups[0]['lr']`` stays schedulable."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=0.0,
                 defer_weight_grads=True, gemm_native_weights=True):
        params = [p for p in params if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.max_grad_norm = float(max_grad_norm)
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedAdamW runs on CUDA only; there is no CPU fallback")
        self._params = params
        self._offsets = []
        off = 0
        for p in params:
            self._offsets.append(off)
            off += (p.numel() + 7) // 8 * 8          # 16-byte aligned slots in the fp32 AND the bf16 shadow buffer (TMA)
        self._n = off
        # Conv2d weights the GEMMs can use as stored (ops.wants_native): kept in [Cout][kh][kw][Cin] order inside the
        # flat buffers; the parameter, its gradient and its state_dict entry are permuted views of that memory
        self._native = [bool(gemm_native_weights and ops.wants_native(p)) for p in params]
        self.flat_param = torch.zeros(off, device=dev, dtype=torch.float32)
        self.flat_bf16 = torch.zeros(off, device=dev, dtype=torch.bfloat16) if any(self._native) else None
        self.flat_grad = torch.zeros(off, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(off, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(off, device=dev, dtype=torch.float32)
        self._gnorm_sq = torch.zeros(1, device=dev, dtype=torch.float32)
        self._step = 0
        with torch.no_grad():
            for p, o, nat in zip(params, self._offsets, self._native):
                view = self._view(self.flat_param, o, p, nat)
                view.copy_(p.data)
                p.data = view
                if nat:
                    cout, cin, kh, kw = p.shape
                    ops.register_native(p, self.flat_bf16[o:o + p.numel()].view(cout, kh * kw * cin))
        self._attach_grads()
        # conv weight gradients stay in the wgrad GEMM's packed layout across the accumulation window and
        # are scattered into the flat gradient once per step (see ops._PackedGrad)
        if defer_weight_grads:
            ops.defer_weight_grads(params)

    @staticmethod
    def _view(flat, o, p, native):
        if native:
            cout, cin, kh, kw = p.shape
            return flat[o:o + p.numel()].view(cout, kh, kw, cin).permute(0, 3, 1, 2)
        return flat[o:o + p.numel()].view_as(p)

    # ------------------------------------------------------------------ checkpointing (new_scripy.py:736-741)
    def state_dict(self):
        """``torch.optim.AdamW`` layout: ``{'state': {i: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups': [...]}`` with
        per-parameter moments in the parameter's own (reference, NCHW-contiguous) shape -- GEMM-native moments are
        permuted back -- so a checkpoint written here resumes under ``torch.optim.AdamW`` and vice versa."""
        state = {}
        if self._step > 0:
            for i, (p, o, nat) in enumerate(zip(self._params, self._offsets, self._native)):
                state[i] = {"step": torch.tensor(float(self._step)),
                            "exp_avg": self._view(self.exp_avg, o, p, nat).contiguous().clone(),
                            "exp_avg_sq": self._view(self.exp_avg_sq, o, p, nat).contiguous().clone()}
        groups = [{**{k: v for k, v in g.items() if k != "params"}, "params": list(range(len(self._params)))}
                  for g in self.param_groups]
        return {"state": state, "param_groups": groups, "max_grad_norm": self.max_grad_norm}

    @torch.no_grad()
    def load_state_dict(self, sd):
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self._params):
            raise ValueError("FusedAdamW.load_state_dict: expected one parameter group with "
                             f"{len(self._params)} parameters, got {[len(g['params']) for g in groups]}")
        state = sd["state"]
        steps = set()
        for i, (p, o, nat) in enumerate(zip(self._params, self._offsets, self._native)):
            st = state.get(groups[0]["params"][i], state.get(i))
            if st is None:
                self._view(self.exp_avg, o, p, nat).zero_()
                self._view(self.exp_avg_sq, o, p, nat).zero_()
                continue
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"FusedAdamW.load_state_dict: moment {i} has shape {tuple(st['exp_avg'].shape)}, "
                                 f"parameter has {tuple(p.shape)}")
            self._view(self.exp_avg, o, p, nat).copy_(st["exp_avg"])
            self._view(self.exp_avg_sq, o, p, nat).copy_(st["exp_avg_sq"])
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"FusedAdamW.load_state_dict: parameters disagree on the step count: {sorted(steps)}")
        self._step = steps.pop() if steps else 0
        for k, v in groups[0].items():
            if k != "params":
                self.param_groups[0][k] = v
        if "max_grad_norm" in sd:
            self.max_grad_norm = float(sd["max_grad_norm"])

    def flush(self):
        """Make ``p.grad`` (the flat gradient buffer) complete: scatter any packed weight gradients."""
        self._attach_grads()
        ops.flush_weight_grads()

    def _attach_grads(self):
        for p, o, nat in zip(self._params, self._offsets, self._native):
            gv = self._view(self.flat_grad, o, p, nat)
            if p.grad is None or p.grad.data_ptr() != gv.data_ptr():
                if p.grad is not None:
                    gv.copy_(p.grad)
                p.grad = gv

    def zero_grad(self, set_to_none: bool = False):
        ops.discard_weight_grads()
        self.flat_grad.zero_()
        self._attach_grads()

    @torch.no_grad()
    def grad_norm(self):
        """Global L2 norm of the (flat) gradient as a device scalar."""
        self.flush()
        self._gnorm_sq.zero_()
        ops.call("dm_sumsq", _p(self.flat_grad), self._n, _p(self._gnorm_sq), ops._stream())
        return self._gnorm_sq.sqrt()

    @torch.no_grad()
    def step(self, closure=None):
        self.flush()
        g = self.param_groups[0]
        self._step += 1
        b1, b2 = g["betas"]
        st = ops._stream()
        gn = None
        if self.max_grad_norm > 0:
            self._gnorm_sq.zero_()
            ops.call("dm_sumsq", _p(self.flat_grad), self._n, _p(self._gnorm_sq), st)
            gn = _p(self._gnorm_sq)
        hyper = (float(g["lr"]), float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]),
                 1.0 - b1 ** self._step, 1.0 - b2 ** self._step, gn, self.max_grad_norm, st)
        if self.flat_bf16 is not None:
            ops.call("dm_adamw_bf16", _p(self.flat_param), _p(self.flat_grad), _p(self.exp_avg), _p(self.exp_avg_sq),
                     _p(self.flat_bf16), self._n, *hyper)
        else:
            ops.call("dm_adamw", _p(self.flat_param), _p(self.flat_grad), _p(self.exp_avg), _p(self.exp_avg_sq), self._n,
                     *hyper)
        ops.bump_weights_epoch()
        ops.natives_fresh()
        ops.refresh_weight_packs()
