"""One-off evidence run (not part of the suite): the benchmarked training configuration (new_scripy Cfg defaults: n_feat 192,
3 x 256 x 256, micro-batch 4 x accum 4, clip 1.0, AdamW lr 1e-4 wd 1e-5, train-mode BatchNorm) for STEPS optimizer steps from
identical weights, data and per-micro-batch random draws, twice: the bf16 B200 path (CUDA-graphed micro-steps + FusedAdamW)
and the fp32 reference loop (oracle/ref_port.py + torch.optim.AdamW + clip_grad_norm_, new_scripy.py:784-803) executed on the
GPU with TF32 off.  Prints both loss curves (mean over the accumulation window) and their relative deviation.

    python tools/loss_curve_cfg2.py [out.json] [steps]"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusionmodel_b200 as D
from oracle import ref_port as P
from oracle.synth import make_inputs
from tests.test_gpu_model import build
import bench

dev = torch.device("cuda:0")
out_path = sys.argv[1] if len(sys.argv) > 1 else None
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
C = bench.CFG
n_T, batch, accum, size, ncls = C["n_T"], C["batch"], C["accum"], C["img"], C["n_classes"]
ddpm, sd = build("rdd", C["n_feat"], ncls, n_T, 31, dev, enhance_with_attn_map=True)
ddpm.train()
# `accum` different micro-batches, reused every step (a fixed tiny dataset), fresh random draws every micro-step
data = [make_inputs("rdd", batch, 3, size, ncls, n_T, 100 + k) for k in range(accum)]
g = torch.Generator().manual_seed(77)
draws = [[(torch.randint(1, n_T + 1, (batch,), generator=g), torch.randn(batch, 3, size, size, generator=g),
           torch.bernoulli(torch.full((batch,), 0.9), generator=g)) for _ in range(accum)] for _ in range(steps)]

# ---- fp32 reference loop on the GPU (TF32 off)
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
sd_o = {k: v.to(dev) for k, v in sd.items()}
params = [v.requires_grad_(True) for k, v in sd_o.items() if v.is_floating_point() and k.startswith("nn_model.") and "running" not in k]
opt_o = torch.optim.AdamW(params, lr=C["lr"], weight_decay=C["wd"])
sched = {k: v.to(dev) for k, v in P.ddpm_schedules(*C["betas"], n_T).items()}
data_d = [{k: v.to(dev) for k, v in d.items()} for d in data]
ref, t0 = [], time.time()
for s in range(steps):
    opt_o.zero_grad()
    tot = 0.0
    for k in range(accum):
        ts, noise, ctx = (t.to(dev) for t in draws[s][k])
        d = data_d[k]
        lo = P.ddpm_loss(sd_o, sched, d["x"], d["c"], d["attn_mask"], ts, noise, ctx, variant="rdd", n_T=n_T, training=True,
                         attn_map=d["attn_mask"]) / accum
        lo.backward()
        tot += float(lo.detach())
    torch.nn.utils.clip_grad_norm_(params, 1.0)
    opt_o.step()
    ref.append(tot)
torch.cuda.synchronize()
t_ref = time.time() - t0
del sd_o, params, opt_o
torch.cuda.empty_cache()

# ---- ours: graphed micro-steps + FusedAdamW
opt = D.FusedAdamW(ddpm.parameters(), lr=C["lr"], weight_decay=C["wd"], max_grad_norm=1.0)
d0 = data_d[0]
micro = ddpm.capture_train_step(d0["x"], d0["c"], d0["attn_mask"], loss_scale=1.0 / accum)
opt.zero_grad()
ours, t0 = [], time.time()
for s in range(steps):
    tot = 0.0
    for k in range(accum):
        ts, noise, ctx = (t.to(dev) for t in draws[s][k])
        d = data_d[k]
        tot += float(micro(d["x"], d["c"], d["attn_mask"], randoms=(ts, noise, ctx)).detach())
    opt.step()
    opt.zero_grad()
    ours.append(tot)
torch.cuda.synchronize()
t_ours = time.time() - t0
rel = [abs(a - b) / abs(b) for a, b in zip(ours, ref)]
print(f"fp32 reference loop on the GPU ({t_ref:.0f} s): {['%.4f' % v for v in ref]}")
print(f"bf16 B200 path ({t_ours:.1f} s incl. host-side loss reads): {['%.4f' % v for v in ours]}")
print(f"relative deviation per step: {['%.2e' % v for v in rel]}   max {max(rel):.2e}  mean {sum(rel) / len(rel):.2e}")
if out_path:
    json.dump({"config": {**{k: v for k, v in C.items()}, "steps": steps, "reference": "ref_port + torch.optim.AdamW on cuda, fp32, TF32 off"},
               "reference_losses": ref, "b200_losses": ours, "rel_dev": rel}, open(out_path, "w"), indent=1)
