"""Dev probe for ncu: one launch each (after one warm-up) of the quad upsample+cat kernels, the skinny split-K GEMM of the
up0 data gradient and the MLP kernels at their Cfg-default sizes.  python tools/new_kernels_probe.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusionmodel_b200 import ops, _lib

dev = torch.device("cuda:0")
P_, st = ops._p, ops._stream()
n, h, c = 4, 128, 192
a = torch.randn(n, h, h, c, device=dev).to(torch.bfloat16)
b = torch.randn(n, h, h, c, device=dev).to(torch.bfloat16)
up = torch.empty((n, 2 * h, 2 * h, 2 * c), device=dev, dtype=torch.bfloat16)
da, db = torch.empty_like(a), torch.empty_like(b)
m, nn, k = 16, 1536, 98304
A = torch.randn(m, k, device=dev).to(torch.bfloat16)
W = torch.randn(nn, k, device=dev).to(torch.bfloat16)
out = torch.empty((m, nn), device=dev, dtype=torch.bfloat16)
scr = torch.empty(_lib.fn("dm_skinny_gemm_scratch")(nn, k), device=dev)
x = torch.randn(4, 1536, device=dev)
w1 = torch.randn(1536, 1536, device=dev) / 40
w2 = torch.randn(1536, 1536, device=dev) / 40
b1 = torch.zeros(1536, device=dev); b2 = torch.zeros(1536, device=dev)
for t in (w1, w2, b1, b2):
    t.requires_grad_(True); t.grad = torch.zeros_like(t)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for rep in range(2):
    flush.zero_()
    ops.call("dm_upcat_fwd", P_(a), c, c, P_(b), c, c, P_(up), 2 * c, n, h, h, st)
    flush.zero_()
    ops.call("dm_upcat_bwd", P_(up), 2 * c, P_(da), c, c, P_(db), c, c, n, h, h, st)
    flush.zero_()
    ops.call("dm_skinny_gemm", P_(A), k, P_(W), k, P_(out), nn, P_(scr), m, nn, k, st)
    flush.zero_()
    y, saved = ops.mlp2_fwd(x, w1.detach(), b1.detach(), w2.detach(), b2.detach(), 1, 0)
    ops.mlp2_bwd(saved, torch.ones_like(y), w1, b1, w2, b2, 1, 0, False)
torch.cuda.synchronize()
print("ok")
