"""dev: board power and SM clock of three phases run alone for ~3 s each (graph replays): the dominant conv forward, the
BatchNorm forward + backward passes on the same tensor, and the whole training micro-step.  python tools/power_probe.py"""
import os, subprocess, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusionmodel_b200 as D
from diffusionmodel_b200 import _lib, ops
import bench

dev = torch.device("cuda:0")
Q = "clocks.sm,power.draw.instant,power.draw.average"


class Sampler:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "--id=0", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                  stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.p.stdout:
            try:
                f = [float(v) for v in line.split(",")]
                self.rows.append((time.time(), f))
            except Exception:
                pass

    def window(self, t0, t1):
        r = [f for t, f in self.rows if t0 + 0.5 <= t <= t1]
        if not r:
            return None
        med = lambda i: sorted(x[i] for x in r)[len(r) // 2]
        return {"sm_mhz": med(0), "power_instant_w": med(1), "power_avg_w": med(2), "samples": len(r)}


def loop(graph, seconds):
    torch.cuda.synchronize()
    t0 = time.time()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < seconds:
        for _ in range(10):
            graph.replay()
        n += 10
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return t0, time.time(), e0.elapsed_time(e1) / n


s = Sampler()
P_ = ops._p
n, h, c = 4, 256, 192
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(n, h, h, c, device=dev, generator=g).to(torch.bfloat16)
w = torch.nn.Parameter(torch.randn(c, c, 3, 3, device=dev, generator=g) / 41.6)
pack = ops.WeightPack()
with torch.no_grad():
    ops.conv2d(x, w, None, pack, stride=1, pad=1)
gconv = torch.cuda.CUDAGraph()
with torch.no_grad(), torch.cuda.graph(gconv):
    for _ in range(20):
        ops.conv2d(x, w, None, pack, stride=1, pad=1)
t0, t1, ms = loop(gconv, 3.0)
flops = 2.0 * n * h * h * c * c * 9 * 20
print("conv forward 192->192 @256^2 alone:", s.window(t0, t1), f"{flops / ms / 1e9:.0f} TFLOP/s")

P = n * h * h
y = torch.randn(n, h, h, c, device=dev, generator=g).to(torch.bfloat16); dz = torch.randn(n, h, h, c, device=dev, generator=g).to(torch.bfloat16)
z = torch.empty_like(y)
mean, inv, ga, be = torch.zeros(c, device=dev), torch.ones(c, device=dev), torch.ones(c, device=dev), torch.zeros(c, device=dev)
dga, dbe = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
scr = torch.empty(_lib.fn("dm_bn_act_bwd_scratch")(P, c), device=dev)
part = torch.empty((_lib.fn("dm_bn_stats_rows")(P, c), 2, c), device=dev)
def bn():
    ops.call("dm_bn_stats", P_(y), c, P_(part), c, P, c, ops._stream())
    ops.call("dm_bn_act_fwd", P_(y), c, P_(mean), P_(inv), P_(ga), P_(be), P_(z), c, P, c, 1, ops._stream())
    ops.call("dm_bn_act_bwd", P_(dz), c, P_(y), c, P_(mean), P_(inv), P_(ga), P_(be), P_(z), c, P_(dga), P_(dbe), None, P_(scr), P, c, 1, 1, ops._stream())
bn(); torch.cuda.synchronize()
gbn = torch.cuda.CUDAGraph()
with torch.cuda.graph(gbn):
    for _ in range(10):
        bn()
t0, t1, ms = loop(gbn, 3.0)
print("BatchNorm statistics + apply + backward chain alone:", s.window(t0, t1), f"{16.0 * P * c * 10 / ms / 1e6:.0f} GB/s algorithmic")

torch.manual_seed(0)
C = bench.CFG
net = D.ContextUnet(C["in_ch"], C["n_feat"], C["n_classes"])
ddpm = D.DDPM(net, C["betas"], C["n_T"], dev, C["drop_prob"], enhance_with_attn_map=True).to(dev).train()
opt = D.FusedAdamW(ddpm.parameters(), lr=C["lr"], weight_decay=C["wd"], max_grad_norm=1.0)
gen = torch.Generator().manual_seed(100)
xb, cb, mb = (t.to(dev) for t in bench.synth_batch(gen, C["batch"], C["img"], C["n_classes"]))
step = ddpm.capture_train_step(xb, cb, mb, loss_scale=0.25)
opt.zero_grad()
t0, t1, ms = loop(step.graph, 4.0)
print("training micro-step graph alone:", s.window(t0, t1), f"{ms:.2f} ms per micro-step, {4038.5 * 4 / ms:.0f} TFLOP/s")
s.p.terminate()
