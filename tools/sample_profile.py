"""Dev tool: the CFG reverse step at the benchmarked sampling configuration (n_feat 192, 256 x 256, 15 trajectories): graphed
ms per step, host enqueue time, and the eager per-entry-point breakdown (ops.enable_profile).  python tools/sample_profile.py"""
import sys, os, json, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusionmodel_b200 as D
from diffusionmodel_b200 import ops
dev = torch.device('cuda:0')
torch.manual_seed(0)
ddpm = D.DDPM(D.ContextUnet(3, 192, 5), (1e-4, 0.02), 700, dev, 0.1).to(dev).eval()
ddpm.sample_noise = "device"
ddpm.sample(15, (3, 256, 256), dev, guide_w=2.0, steps=3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ddpm.sample(15, (3, 256, 256), dev, guide_w=2.0, steps=10); e1.record(); torch.cuda.synchronize()
print("ms per reverse step", e0.elapsed_time(e1) / 10)
# host-only enqueue time
real = ops.call; ops.call = lambda *a, **k: 0
t0 = time.perf_counter(); ddpm.sample(15, (3, 256, 256), dev, guide_w=2.0, steps=10); torch.cuda.synchronize(); t1 = time.perf_counter()
ops.call = real
print("host enqueue ms per reverse step", (t1 - t0) * 100)
ddpm.graph_sampling = False      # CUDA events cannot be recorded inside a graph replay
ddpm.sample(15, (3, 256, 256), dev, guide_w=2.0, steps=2)
prof = ops.enable_profile()
ddpm.sample(15, (3, 256, 256), dev, guide_w=2.0, steps=2)
torch.cuda.synchronize(); ops.disable_profile()
agg = prof.summary()
tot = 0
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    tot += v["ms"]; print(f"{k:24s} {v['ms']/2:8.3f} ms/step n={v['n']//2}  {v['flops']/max(v['ms'],1e-9)/1e9:8.1f} TF/s")
print("sum per step", tot / 2)
