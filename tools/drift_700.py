"""One-off evidence run (not part of the suite: ~2 min of host time): bf16 B200 sampling against the fp32 oracle over the
FULL n_T = 700 CFG reverse trajectory of the reference schedule (new_scripy.py:441-477), n_feat 16, 128 x 128, 5 classes,
guide_w 2, identical injected noise.  Prints the drift curve; python tools/drift_700.py [out.json]"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_port as P
from tests.test_gpu_model import build

dev = torch.device("cuda:0")
n_feat, size, ncls, n_T, seed, w = 16, 128, 5, 700, 23, 2.0
ddpm, sd = build("rdd", n_feat, ncls, n_T, seed, dev)
ddpm.eval()
g = torch.Generator().manual_seed(seed)
x_T = torch.randn(ncls, 3, size, size, generator=g)
zs = {i: torch.randn(ncls, 3, size, size, generator=g) for i in range(n_T, 1, -1)}
sched = P.ddpm_schedules(1e-4, 0.02, n_T)
trace = []
t0 = time.time()
with torch.no_grad():
    P.ddpm_sample(sd, sched, x_T, zs, w, variant="rdd", n_T=n_T, n_classes=ncls, trace=trace)
t_cpu = time.time() - t0
marks = [1, 10, 50, 100, 200, 300, 400, 500, 600, 700]
drift = {}
for k in marks:
    out = ddpm.sample(ncls, (3, size, size), dev, guide_w=w, steps=k, noise=(x_T, zs))
    drift[k] = P.rel_l2(out.cpu(), trace[k - 1])
print(f"fp32 oracle: {t_cpu:.0f} s on the host; final |x| rms {float(trace[-1].pow(2).mean().sqrt()):.3f}")
print("bf16 B200 sampling vs fp32 oracle, rel-L2 of x_i after k reverse steps:", {k: f"{v:.2e}" for k, v in drift.items()})
if len(sys.argv) > 1:
    json.dump({"config": dict(n_feat=n_feat, size=size, n_classes=ncls, n_T=n_T, guide_w=w, seed=seed), "rel_l2_by_step": drift},
              open(sys.argv[1], "w"), indent=1)
