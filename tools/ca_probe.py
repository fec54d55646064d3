"""Dev probe: one CoordAttn forward+backward at a cfg2 shape (for ncu).  python tools/ca_probe.py [C L N]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusionmodel_b200 import unet as U

c, l, n = (int(v) for v in (sys.argv[1:4] + ["1536", "16", "4"][len(sys.argv) - 1:]))
dev = torch.device("cuda:0")
torch.manual_seed(0)
mod = U.CoordAttn(c).to(dev).train()
x = torch.randn(n, l, l, c, device=dev).to(torch.bfloat16).requires_grad_(True)
for _ in range(3):
    y = mod(x)
    y.backward(torch.randn_like(y))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    y = mod(x)
    y.backward(torch.randn_like(y))
e1.record(); torch.cuda.synchronize()
print(f"C={c} L={l} N={n}: {e0.elapsed_time(e1) / 20 * 1000:.1f} us per fwd+bwd (eager, incl. host)")
