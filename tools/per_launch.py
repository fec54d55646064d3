"""Per-launch table of one eager training micro-step at Cfg defaults: CUDA events around every C-ABI call, grouped by
(entry point, integer arguments = the layer's geometry).  python tools/per_launch.py [out.json]"""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusionmodel_b200 as D
from diffusionmodel_b200 import ops, _lib
import bench

dev = torch.device("cuda:0")
torch.manual_seed(0)
C = bench.CFG
net = D.ContextUnet(C["in_ch"], C["n_feat"], C["n_classes"])
ddpm = D.DDPM(net, C["betas"], C["n_T"], dev, C["drop_prob"], enhance_with_attn_map=True).to(dev).train()
opt = D.FusedAdamW(ddpm.parameters(), lr=C["lr"], weight_decay=C["wd"], max_grad_norm=1.0)
gen = torch.Generator().manual_seed(100)
x, c, m = (t.to(dev) for t in bench.synth_batch(gen, C["batch"], C["img"], C["n_classes"]))


def micro():
    (ddpm(x, c, m) / C["accum"]).backward()


for _ in range(3):
    micro()
opt.flush(); opt.step(); opt.zero_grad()
micro()
recs = []
real = _lib.call


def timed_call(name, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rc = real(name, *args); e1.record()
    ints = tuple(a for a in args if isinstance(a, int) and not isinstance(a, bool))
    recs.append((name, ints, e0, e1, ops._Profile._flops(name, args), ops._Profile._bytes(name, args)))
    return rc


ops.call = timed_call
REPS = 3
for _ in range(REPS):
    micro()
torch.cuda.synchronize()
ops.call = real
groups = {}
for name, ints, e0, e1, fl, nb in recs:
    g = groups.setdefault((name, ints), [0.0, 0, 0.0, 0.0])
    g[0] += e0.elapsed_time(e1) / REPS; g[1] += 1; g[2] += fl / REPS; g[3] += (0.0 if fl else nb) / REPS
rows = sorted(([k[0], list(k[1]), round(v[0], 4), v[1] // REPS, v[2], v[3]] for k, v in groups.items()), key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"total {tot:.2f} ms per micro-step over {sum(r[3] for r in rows)} launches")
for r in rows[:90]:
    rate = f"{r[4] / r[2] / 1e9:7.0f} TF/s" if r[4] else (f"{r[5] / r[2] / 1e6:7.0f} GB/s" if r[5] else " " * 12)
    print(f"{r[2]:8.3f} ms  x{r[3]:<3d} {rate} {r[0]:22s} {r[1]}")
if len(sys.argv) > 1:
    json.dump({"total_ms": tot, "rows": rows}, open(sys.argv[1], "w"))
