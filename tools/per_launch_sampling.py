"""Per-launch table of one eager CFG reverse step at the benchmarked sampling configuration (n_feat 192, 256 x 256, 15
trajectories, shared encoder): CUDA events around every C-ABI call, grouped by (entry point, geometry).
python tools/per_launch_sampling.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusionmodel_b200 as D
from diffusionmodel_b200 import ops, _lib

dev = torch.device("cuda:0")
torch.manual_seed(0)
ddpm = D.DDPM(D.ContextUnet(3, 192, 5), (1e-4, 0.02), 700, dev, 0.1).to(dev).eval()
ddpm.sample_noise = "device"
ddpm.graph_sampling = False
ddpm.sample(15, (3, 256, 256), dev, guide_w=2.0, steps=2)
recs = []
real = _lib.call


def timed_call(name, *args):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rc = real(name, *args); e1.record()
    ints = tuple(a for a in args if isinstance(a, int) and not isinstance(a, bool))
    recs.append((name, ints, _lib.last_kernel() if name.startswith("dm_conv") else ("", 0), e0, e1,
                 ops._Profile._flops(name, args), ops._Profile._bytes(name, args)))
    return rc


ops.call = timed_call
REPS = 3
ddpm.sample(15, (3, 256, 256), dev, guide_w=2.0, steps=REPS)
torch.cuda.synchronize()
ops.call = real
groups = {}
for name, ints, kern, e0, e1, fl, nb in recs:
    g = groups.setdefault((name, ints, kern), [0.0, 0, 0.0, 0.0])
    g[0] += e0.elapsed_time(e1) / REPS; g[1] += 1; g[2] += fl / REPS; g[3] += (0.0 if fl else nb) / REPS
rows = sorted(([k[0], list(k[1]), k[2], round(v[0], 4), v[1] // REPS, v[2], v[3]] for k, v in groups.items()), key=lambda r: -r[3])
tot = sum(r[3] for r in rows)
print(f"total {tot:.2f} ms per reverse step over {sum(r[4] for r in rows)} launches")
for r in rows[:60]:
    rate = f"{r[5] / r[3] / 1e9:7.0f} TF/s" if r[5] else (f"{r[6] / r[3] / 1e6:7.0f} GB/s" if r[6] else " " * 12)
    print(f"{r[3]:8.3f} ms  x{r[4]:<3d} {rate} {r[0]:22s} {r[2][0]:14s}{r[2][1]:4d} {r[1]}")
