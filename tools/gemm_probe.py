"""Dev probe for ncu: the dominant layer of the step (3x3, 192 -> 192 on 4 x 256 x 256) forward, data gradient and weight
gradient, and its BatchNorm + GELU passes, one launch each after one warm-up, L2 flushed in between.
python tools/gemm_probe.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusionmodel_b200 import ops

dev = torch.device("cuda:0")
n, h, c = 4, 256, 192
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(n, h, h, c, device=dev, generator=g).to(torch.bfloat16).requires_grad_(True)
w = torch.nn.Parameter(torch.randn(c, c, 3, 3, device=dev, generator=g) / 41.6)
b = torch.nn.Parameter(torch.zeros(c, device=dev))
bn = torch.nn.BatchNorm2d(c).to(dev).train()
dy = torch.randn(n, h, h, c, device=dev, generator=g).to(torch.bfloat16)
pack = ops.WeightPack()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for rep in range(2):
    flush.zero_()
    y, _ = ops.conv2d(x, w, b, pack, stride=1, pad=1, bias_grad_by_norm=True, add_bias=False)
    flush.zero_()
    z = ops.bn_act(y, None, bn, ops.ACT_GELU, conv_bias=b, bias_outside=True)
    flush.zero_()
    z.backward(dy)
torch.cuda.synchronize()
print("ok")
