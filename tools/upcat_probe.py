"""Dev probe: cat + bilinear x2 (dm_upcat_fwd / dm_upcat_bwd) bandwidth at the training and the sampling batch, quad
kernels (default) against the one-thread-per-pixel form (dm_debug_set(9, 1)).  python tools/upcat_probe.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusionmodel_b200 import ops, _lib

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[2], out


for n in (4, 30):
    for (h, c) in ((128, 192), (64, 384), (32, 768), (16, 1536)):
        a = torch.randn(n, h, h, c, device=dev).to(torch.bfloat16).requires_grad_(True)
        b = torch.randn(n, h, h, c, device=dev).to(torch.bfloat16).requires_grad_(True)
        line = f"n={n:2d} {h}x{h}x{c}+{c}:"
        for mode in (1, 0):
            _lib.debug_set(9, mode)
            ms, out = timed(lambda: ops.upcat(a, b, c, c))
            gb = (a.numel() + b.numel() + out.numel()) * 2 / 1e9
            dy = torch.randn_like(out)
            msb, _ = timed(lambda: torch.autograd.grad(out, (a, b), dy, retain_graph=True))
            line += f"  {'pixel' if mode else 'quad '} fwd {ms*1e3:7.1f} us {gb/ms*1e3:6.0f} GB/s  bwd {msb*1e3:7.1f} us {gb/msb*1e3:6.0f} GB/s |"
            del out, dy
        _lib.debug_set(9, 0)
        print(line, flush=True)
        del a, b
