"""dev: weight-gradient kernel variant x split-K sweep on the cfg2 layers that run below 1.1 PFLOP/s (graph replay of 20
launches per configuration).  python tools/wgrad_sweep.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusionmodel_b200 import _lib, ops
dev = torch.device("cuda:0")
p = ops._p
SHAPES = [  # cin, cout, hw, k, stride, pad
    (768, 768, 32, 3, 1, 1), (384, 384, 64, 3, 1, 1), (192, 192, 128, 3, 1, 1), (1536, 1536, 32, 3, 1, 1),
    (768, 768, 64, 3, 1, 1), (1536, 1536, 32, 4, 2, 1), (192, 192, 256, 3, 1, 1), (384, 384, 128, 3, 1, 1)]
N = 4
for cin, cout, hw, k, stride, pad in SHAPES:
    ho = (hw + 2 * pad - k) // stride + 1
    x = torch.randn(N, hw, hw, cin, device=dev).to(torch.bfloat16)
    dy = torch.randn(N, ho, ho, cout, device=dev).to(torch.bfloat16)
    dwp = torch.zeros(cout, k * k * cin, device=dev)
    flops = 2.0 * N * ho * ho * cout * cin * k * k
    def run():
        ops.call("dm_conv2d_wgrad", p(x), cin, cin, None, 0, 0, p(dy), cout, p(dwp), N, hw, hw, cout, k, k, stride, pad, ops._stream())
    res = []
    for ver in (0, 1, 2, 3):
        for splits in (0, 1, 2, 3, 4, 6, 8, 12, 16, 24, 32):
            _lib.debug_set(4, ver); _lib.debug_set(2, splits)
            try:
                run(); torch.cuda.synchronize()
                name, used = _lib.last_kernel()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(20):
                        run()
                g.replay(); torch.cuda.synchronize()
                ts = []
                for _ in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                us = sorted(ts)[1] * 1e3 / 20
                res.append((us, ver, splits, name, used))
            except Exception as e:          # noqa: BLE001
                pass
    _lib.debug_set(4, 0); _lib.debug_set(2, 0)
    auto = [r for r in res if r[1] == 0 and r[2] == 0][0]
    best = min(res)
    print(f"{cin}->{cout} @{hw}^2 k{k}s{stride}: auto {auto[3]} splits={auto[4]} {auto[0]:.1f} us {flops / auto[0] / 1e6:.0f} TF/s | "
          f"best {best[3]} (ver {best[1]}) splits={best[4]} {best[0]:.1f} us {flops / best[0] / 1e6:.0f} TF/s")
    for r in sorted(res)[:4]:
        print(f"      {r[3]:12s} ver={r[1]} forced_splits={r[2]:2d} used={r[4]:3d} {r[0]:7.1f} us")
