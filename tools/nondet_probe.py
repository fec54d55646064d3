"""Dev probe: where does run-to-run nondeterminism enter in train mode? (block-level forward hooks)"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusionmodel_b200 as D
from diffusionmodel_b200 import ops, unet
from oracle import ref_port as P
from oracle.synth import make_inputs
from tests.test_gpu_model import build
dev = torch.device('cuda:0')
variant, f, size, b = "rdd", 32, 128, 4
inp = make_inputs(variant, b, 3, size, 5, 700, 3)
x, c, attn, ts, noise, ctx = (inp[k].to(dev) for k in ("x", "c", "attn_mask", "ts", "noise", "ctx_mask"))
runs = []
for rep in range(2):
    ddpm, _ = build(variant, f, 5, 700, 3, dev, enhance_with_attn_map=True)
    ddpm.train()
    rec = {}
    def mk(name):
        def hook(m, i, o):
            if torch.is_tensor(o): rec[name] = o.detach().float().cpu()
        return hook
    for name, m in ddpm.nn_model.named_modules():
        if isinstance(m, (unet.ResConvBlock, unet.UnetDown, unet.UnetUp, unet.CoordAttn, unet.LocalEnhancer)):
            m.register_forward_hook(mk(name))
    lo = ddpm(x, c, attn, randoms=(ts, noise, ctx)); lo.backward(); torch.cuda.synchronize()
    rec["loss"] = lo.detach().cpu()
    grads = {k: p.grad.detach().cpu().clone() for k, p in ddpm.named_parameters() if p.grad is not None}
    runs.append((rec, grads))
for k in runs[0][0]:
    a, bb = runs[0][0][k], runs[1][0][k]
    print(f"fwd {k:40s} rel-L2 {P.rel_l2(a, bb):.3e}  n_diff {(a != bb).sum().item()}/{a.numel()}")
worst = sorted(((P.rel_l2(runs[0][1][k], runs[1][1][k]), k) for k in runs[0][1] if float(runs[0][1][k].norm()) > 0), reverse=True)
for e, k in worst[:12]: print(f"grad {k:50s} {e:.3e}")
for e, k in worst[-8:]: print(f"grad {k:50s} {e:.3e}")
# single conv + BN + GELU unit, train mode, twice
torch.manual_seed(0)
seq = torch.nn.Sequential(torch.nn.Conv2d(64, 64, 3, 1, 1), torch.nn.BatchNorm2d(64), torch.nn.GELU()).to(dev).train()
xx = torch.randn(4, 64, 64, 64, device=dev).to(torch.bfloat16)
outs = []
for rep in range(2):
    xi = xx.clone().requires_grad_(True)
    for p in seq.parameters(): p.grad = None
    z = unet.conv_bn_act(xi, seq); z.backward(torch.ones_like(z) * 0.01 * z.detach()); torch.cuda.synchronize()
    outs.append((z.detach().float().cpu(), xi.grad.float().cpu(), seq[0].weight.grad.cpu().clone(), seq[1].weight.grad.cpu().clone()))
for i, n in enumerate(("z", "dx", "dW", "dgamma")):
    print("unit", n, "rel-L2", P.rel_l2(outs[0][i], outs[1][i]), "n_diff", (outs[0][i] != outs[1][i]).sum().item())
