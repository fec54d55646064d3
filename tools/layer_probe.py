"""Dev probe: forward, data-gradient and weight-gradient GEMMs of single 3x3 layers at the step's shapes, timed alone with
CUDA events (median of REPS, L2 flushed before every launch), with the kernel variant the dispatcher chose.
python tools/layer_probe.py [cin:cout:hw[:k:stride] ...]      (batch 4; PROBE_NCU=1: two passes only, for an ncu capture)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusionmodel_b200 import ops, _lib

dev = torch.device("cuda:0")
shapes = [tuple(int(v) for v in s.split(":")) for s in sys.argv[1:]] or [
    (192, 192, 256), (384, 384, 128), (768, 768, 64), (1536, 1536, 32), (768, 768, 32), (384, 384, 64), (192, 192, 128),
    (3072, 768, 32), (1536, 384, 64), (768, 192, 128)]
NCU = os.environ.get("PROBE_NCU") == "1"
REPS = 2 if NCU else 7
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
real = _lib.call
recs = []


def timed_call(name, *args):
    if not name.startswith("dm_conv2d"):
        return real(name, *args)
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rc = real(name, *args); e1.record()
    recs.append((name, _lib.last_kernel(), e0, e1, ops._Profile._flops(name, args)))
    return rc


ops.call = timed_call
if os.environ.get("PROBE_BN"):
    _lib.debug_set(3, int(os.environ["PROBE_BN"]))          # dev: forced tile width
if os.environ.get("PROBE_SINGLE") == "1":
    _lib.debug_set(5, 2)                                    # dev: no CTA pairs
g = torch.Generator(device=dev).manual_seed(1)
for spec in shapes:
    cin, cout, hw = spec[:3]
    k, stride = (spec[3], spec[4]) if len(spec) == 5 else (3, 1)
    n = 4
    x = torch.randn(n, hw, hw, cin, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
    w = torch.nn.Parameter(torch.randn(cout, cin, k, k, device=dev, generator=g) / (k * cin ** 0.5))
    b = torch.nn.Parameter(torch.zeros(cout, device=dev))
    dy = torch.randn(n, hw // stride, hw // stride, cout, device=dev, generator=g).to(torch.bfloat16)
    pack = ops.WeightPack()
    recs.clear()
    for rep in range(REPS):
        y, _ = ops.conv2d(x, w, b, pack, stride=stride, pad=(k - 1) // 2 if stride == 1 else 1, bias_grad_by_norm=True, add_bias=False)
        y.backward(dy)
        w.grad = None; x.grad = None
    torch.cuda.synchronize()
    per, k = {}, len(recs) // REPS
    for i, (name, kern, e0, e1, fl) in enumerate(recs[k:]):            # first pass = warm-up (weight packs)
        per.setdefault((i % k, name, kern), []).append((e0.elapsed_time(e1), fl))
    for (pos, name, kern), v in per.items():
        ts = sorted(t for t, _ in v); t = ts[len(ts) // 2]
        print(f"{cin:5d}->{cout:<5d} @{hw:<4d} k{k}s{stride} {pos} {name:20s} {kern[0]:14s} {kern[1]:4d}  {t * 1e3:8.1f} us  {v[0][1] / t / 1e9:7.0f} TF/s", flush=True)
    del x, w, b, dy, y, pack
print("ok")
