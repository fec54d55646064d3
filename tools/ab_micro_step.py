"""Dev A/B on ONE box: the graphed cfg2 training micro-step timed with a module-level switch off and on, alternating
(boxes of the pool differ by more than most single changes, so both arms run in the same process).
python tools/ab_micro_step.py ops.FUSED_CONV_STATS"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusionmodel_b200 as D
from diffusionmodel_b200 import ops
import bench

switch = sys.argv[1] if len(sys.argv) > 1 else "ops.FUSED_CONV_STATS"
mod, name = switch.split(".")
target = {"ops": ops}[mod]
dev = torch.device("cuda:0")
C = bench.CFG
torch.manual_seed(0)
net = D.ContextUnet(C["in_ch"], C["n_feat"], C["n_classes"])
ddpm = D.DDPM(net, C["betas"], C["n_T"], dev, C["drop_prob"], enhance_with_attn_map=True).to(dev).train()
opt = D.FusedAdamW(ddpm.parameters(), lr=C["lr"], weight_decay=C["wd"], max_grad_norm=1.0)
gen = torch.Generator().manual_seed(100)
x, c, m = (t.to(dev) for t in bench.synth_batch(gen, C["batch"], C["img"], C["n_classes"]))
graphs = {}
for v in (False, True):
    setattr(target, name, v)
    graphs[v] = ddpm.capture_train_step(x, c, m, loss_scale=0.25)
    opt.zero_grad()


def ms(step, n=20):
    for _ in range(3):
        step(x, c, m)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step(x, c, m)
    e1.record()
    torch.cuda.synchronize()
    opt.zero_grad()
    return e0.elapsed_time(e1) / n


for rnd in range(3):
    a, b = ms(graphs[False]), ms(graphs[True])
    print(f"round {rnd}: {switch}=False {a:.3f} ms   True {b:.3f} ms   ({(b / a - 1) * 100:+.2f} %)", flush=True)
print("losses", float(graphs[False](x, c, m)), float(graphs[True](x, c, m)))
