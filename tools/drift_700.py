"""One-off evidence run (not part of the suite): bf16 B200 sampling against the fp32 oracle over the FULL n_T = 700 CFG
reverse trajectory of the reference schedule (new_scripy.py:441-477), 5 classes, guide_w 2, identical injected noise.
Prints the drift curve.

    python tools/drift_700.py [out.json]                       n_feat 16, 128 x 128, oracle on the host cores (~2 min)
    python tools/drift_700.py out.json 128 28 cuda mnist       cfg1: the MNIST variant, n_feat 128, 40 trajectories, n_T = 400
    python tools/drift_700.py out.json 192 256 cuda            the benchmarked configuration; the oracle (same ref_port code)
                                                               runs on the GPU in fp32 with TF32 off (~4 min), and once more
                                                               with TF32 convolutions on -- how the reference itself samples on
                                                               a GPU (new_scripy.py:867-872: no autocast around sample())"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_port as P
from tests.test_gpu_model import build

dev = torch.device("cuda:0")
out_path = sys.argv[1] if len(sys.argv) > 1 else None
n_feat = int(sys.argv[2]) if len(sys.argv) > 2 else 16
size = int(sys.argv[3]) if len(sys.argv) > 3 else 128
where = sys.argv[4] if len(sys.argv) > 4 else "cpu"
variant = sys.argv[5] if len(sys.argv) > 5 else "rdd"
ncls, n_T, seed, w = (5, 700, 23, 2.0) if variant == "rdd" else (10, 400, 23, 2.0)
in_ch, n_traj = (3, ncls) if variant == "rdd" else (1, 4 * ncls)
ddpm, sd = build(variant, n_feat, ncls, n_T, seed, dev)
ddpm.eval()
g = torch.Generator().manual_seed(seed)
x_T = torch.randn(n_traj, in_ch, size, size, generator=g)
zs = {i: torch.randn(n_traj, in_ch, size, size, generator=g) for i in range(n_T, 1, -1)}
sched = P.ddpm_schedules(1e-4, 0.02, n_T)
marks = [k for k in (1, 10, 50, 100, 200, 300, 400, 500, 600, 700) if k <= n_T]


def oracle(tf32):
    trace = []
    t0 = time.time()
    with torch.no_grad():
        if where == "cuda":
            old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
            torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = tf32
            try:
                with torch.device(dev):                    # the oracle's index / time tensors are created on the device too
                    P.ddpm_sample({k: v.to(dev) for k, v in sd.items()}, {k: v.to(dev) for k, v in sched.items()}, x_T.to(dev),
                                  {i: z.to(dev) for i, z in zs.items()}, w, variant=variant, n_T=n_T, n_classes=ncls, trace=trace)
                torch.cuda.synchronize()
            finally:
                torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        else:
            P.ddpm_sample(sd, sched, x_T, zs, w, variant=variant, n_T=n_T, n_classes=ncls, trace=trace)
    return {k: trace[k - 1].cpu() for k in marks}, time.time() - t0


ref, t_ref = oracle(False)
print(f"fp32 oracle on {where}: {t_ref:.0f} s; final |x| rms {float(ref[n_T].pow(2).mean().sqrt()):.3f}", flush=True)
drift = {}
for k in marks:
    out = ddpm.sample(n_traj, (in_ch, size, size), dev, guide_w=w, steps=k, noise=(x_T, zs))
    out = out[0] if isinstance(out, tuple) else out            # MNIST_script.py:300 returns (x_i, x_i_store)
    drift[k] = P.rel_l2(out.cpu(), ref[k])
print("bf16 B200 sampling vs fp32 oracle, rel-L2 of x_i after k reverse steps:", {k: f"{v:.2e}" for k, v in drift.items()}, flush=True)
res = {"config": dict(variant=variant, n_feat=n_feat, size=size, n_classes=ncls, trajectories=n_traj, n_T=n_T, guide_w=w, seed=seed, oracle_on=where),
       "rel_l2_by_step": drift}
if where == "cuda":
    ref_tf32, t_tf32 = oracle(True)
    res["reference_precision_tf32_convs_rel_l2_by_step"] = {k: P.rel_l2(ref_tf32[k], ref[k]) for k in marks}
    print(f"the oracle with TF32 convolutions ({t_tf32:.0f} s) vs the fp32 oracle:",
          {k: f"{v:.2e}" for k, v in res["reference_precision_tf32_convs_rel_l2_by_step"].items()})
if out_path:
    json.dump(res, open(out_path, "w"), indent=1)
