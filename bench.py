"""Headline benchmark: DDPM train step of the enhanced ContextUnet (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode train|sample]

One "step" = one optimizer step of the reference training loop at Cfg defaults
(new_scripy.py:22-67,777-803): ACCUM_STEPS=4 micro-batches of BATCH_SIZE=4 synthetic 3x256x256 images
(fwd + bwd each), then global-norm clip (1.0) + AdamW(lr 1e-4, wd 1e-5) -- 16 images per step per GPU,
n_feat=192, n_classes=5, n_T=700, bf16 operands / fp32 accumulate, LocalEnhancer executed with the
attention map (so the reference's full 1346 GFLOP/img forward is computed, none of it skipped).

value : img/s with the step's inputs already resident in HBM (CUDA events, max over ranks).
e2e   : the same through the public module API with HOST (pinned) buffers: every micro-batch is
        copied host->device and its loss read back device->host inside the timed region.
N > 1 : data parallel, one process per GPU (torchrun), fixed per-GPU batch (weak scaling), one NCCL
        all-reduce of the flat gradient per optimizer step.
--impl reference : the CPU oracle port of the reference (oracle/ref_port.py; the Python reference
        itself cannot travel to the GPU box) on ALL host cores of the box, bounded sample per step.
Reported beside the headline (rank 0, N = 1): cpu_baseline (the oracle on the host cores), gpu_eager_baseline (the
same oracle code -- the reference's own aten calls -- on this GPU through cuDNN/cuBLAS eager: fp32/TF32, autocast fp16
as the reference trains, autocast bf16), roofline.hbm_kernels (every bandwidth kernel class alone), roofline.non_gemm
(all non-GEMM launches of one step: algorithmic bytes / time / HBM peak), sampling sweep (samples_per_class 1/3/8/16,
batched guidance scales 2/4/6 = BASELINE.json configs[2]) and bounded cfg1 / cfg5 figures under "extra".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = dict(n_feat=192, in_ch=3, n_T=700, betas=(1e-4, 0.02), drop_prob=0.1, batch=4, accum=4, lr=1e-4, wd=1e-5,
           img=256, n_classes=5)
GFLOP_PER_IMG_FWD = 1346.17          # SURVEY.md 8(d), hooks on the imported reference
GFLOP_PER_IMG_TRAIN = 4038.5


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            try:
                pw.append(float(f[2]))
            except Exception:
                pass
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort(); pw.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w": pw[len(pw) // 2] if pw else None,
                "power_w_max": pw[-1] if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def synth_batch(gen, batch, img, n_classes):
    """Synthetic road-damage-shaped batch: images in [-1,1], labels, attention map built like
    CrackDataset.__getitem__ (new_scripy.py:535-546): 0.5, lower half 1.0, one random bbox 3.0."""
    x = torch.rand(batch, 3, img, img, generator=gen) * 2 - 1
    c = torch.randint(0, n_classes, (batch,), generator=gen)
    m = torch.full((batch, img, img), 0.5)
    m[:, img // 2:, :] = 1.0
    for b in range(batch):
        xs = torch.randint(0, img, (2,), generator=gen).sort().values
        ys = torch.randint(0, img, (2,), generator=gen).sort().values
        m[b, int(ys[0]):int(ys[1]) + 1, int(xs[0]):int(xs[1]) + 1] = 3.0
    return x, c, m


# ------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """CPU oracle port of the reference train micro-step, bounded sample: ONE image fwd+bwd per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_port as P
    import diffusionmodel_b200 as D
    torch.manual_seed(0)
    cores = host_threads()                 # torchrun exports OMP_NUM_THREADS=1: take every core this process may use
    net = D.ContextUnet(CFG["in_ch"], CFG["n_feat"], CFG["n_classes"])          # parameter container only (CPU)
    sd = {"nn_model." + k: v.detach().clone() for k, v in net.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    sched = P.ddpm_schedules(*CFG["betas"], CFG["n_T"])
    gen = torch.Generator().manual_seed(0)
    x, c, m = synth_batch(gen, 1, CFG["img"], CFG["n_classes"])

    def step():
        ts, noise, ctx = P.draw_train_randoms(x, c, CFG["n_T"], CFG["drop_prob"], "rdd")
        for v in sd.values():
            v.grad = None
        loss = P.ddpm_loss(sd, sched, x, c, m, ts, noise, ctx, variant="rdd", n_T=CFG["n_T"], training=True, attn_map=m)
        loss.backward()
        return float(loss.detach())
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = 1.0 / dt
    sample = "1 image fwd+bwd per step (F=192, 3x256x256, fp32, CPU oracle port of new_scripy.DDPM.forward), no optimizer"
    line = {"impl": "reference", "metric": "ddpm_train_imgs_per_s", "value": val, "unit": "img/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": val, "unit": "img/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def host_threads():
    """Use all host cores the process is allowed on (only rank 0 runs CPU work; the other ranks idle)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


def workload_config(n):
    return {"workload": "new_scripy.ContextUnet DDPM train step, Cfg defaults (n_feat=192, 3x256x256, n_T=700, "
                        "5 classes), batch 4 x accum 4 per GPU, clip 1.0 + AdamW; LocalEnhancer fed the attention map",
            "global_batch": CFG["batch"] * CFG["accum"] * n, "micro_batch": CFG["batch"], "accum_steps": CFG["accum"],
            "parallelism": f"dp{n}",
            "reference_arm": "--impl reference = CPU oracle port, 1 image fwd+bwd per step, no optimizer, normalised per image",
            "cuda_graph": os.environ.get("DM_BENCH_GRAPH", "1") != "0",
            "allreduce": ("last micro-step's backward in 3 graphs: decoder-side gradients (62 % of the bytes) all-reduced under the "
                          "trunk's backward, down4's (29 %) under init_conv..down3's, the rest after it" if n > 1 and os.environ.get("DM_BENCH_OVERLAP", "1") != "0" else
                          ("one NCCL all-reduce of the flat fp32 gradient after the last micro-step" if n > 1 else "none (1 GPU)")), "l2": "per-step working set (>5 GB of activations) is far larger than the 126 MB L2"}


# ------------------------------------------------------------------------------------------ our arm
def cpu_baseline_sample():
    """Oracle port timed on this box's host cores: one image fwd+bwd per run, about 10 s of CPU work."""
    from oracle import ref_port as P
    import diffusionmodel_b200 as D
    cores = host_threads()
    net = D.ContextUnet(CFG["in_ch"], CFG["n_feat"], CFG["n_classes"])
    sd = {"nn_model." + k: v.detach().clone() for k, v in net.state_dict().items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    sched = P.ddpm_schedules(*CFG["betas"], CFG["n_T"])
    gen = torch.Generator().manual_seed(0)
    x, c, m = synth_batch(gen, 1, CFG["img"], CFG["n_classes"])
    ts, noise, ctx = P.draw_train_randoms(x, c, CFG["n_T"], CFG["drop_prob"], "rdd")

    def one():
        for v in sd.values():
            v.grad = None
        loss = P.ddpm_loss(sd, sched, x, c, m, ts, noise, ctx, variant="rdd", n_T=CFG["n_T"], training=True, attn_map=m)
        loss.backward()
    one()                                   # untimed: first-touch page faults, thread-pool start
    runs, t0 = 0, time.perf_counter()
    while runs < 8 and (runs < 2 or time.perf_counter() - t0 < 10.0):      # about 10 s of CPU work
        one()
        runs += 1
    dt = (time.perf_counter() - t0) / runs
    return {"value": 1.0 / dt, "unit": "img/s", "cores": cores, "kind": "port",
            "sample": f"1 image fwd+bwd (F=192, 3x256x256, fp32) through oracle/ref_port.py, mean of {runs} runs after one "
                      "warm-up, no optimizer"}



def _time_cuda(fn, warm, reps):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def gpu_eager_baseline(dev):
    """the reference's own code path on THIS GPU.  The reference ships no kernels of its own: its modules
    dispatch to aten -> cuDNN / cuBLAS; oracle/ref_port.py makes the same aten calls in the same order (pinned bit-exact to
    the imported reference on the CPU), so running it on ``cuda`` is the reference's eager GPU path.  Timed: one cfg2 train
    micro-step (B=4 forward + backward, no optimizer) and one n=15 CFG reverse step (doubled batch of 30, eval, no grad),
    in the reference's precisions: fp32 with cuDNN TF32 convolutions (torch defaults, how it samples, new_scripy.py:1041),
    autocast fp16 (how it trains, :784) and autocast bf16."""
    from oracle import ref_port as P
    import diffusionmodel_b200 as D
    torch.manual_seed(0)
    net = D.ContextUnet(CFG["in_ch"], CFG["n_feat"], CFG["n_classes"])
    sd = {"nn_model." + k: v.detach().to(dev) for k, v in net.state_dict().items()}
    del net
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    sched = {k: v.to(dev) for k, v in P.ddpm_schedules(*CFG["betas"], CFG["n_T"]).items()}
    gen = torch.Generator().manual_seed(0)
    x, c, m = (t.to(dev) for t in synth_batch(gen, CFG["batch"], CFG["img"], CFG["n_classes"]))
    ts = torch.randint(1, CFG["n_T"] + 1, (CFG["batch"],), generator=gen).to(dev)
    noise = torch.randn(x.shape, generator=gen).to(dev)
    ctx = torch.ones(CFG["batch"], device=dev)
    n_samp = 3 * CFG["n_classes"]
    xs = torch.randn(n_samp, 3, CFG["img"], CFG["img"], device=dev)
    c_i = torch.arange(0, CFG["n_classes"], device=dev).repeat(3).repeat(2)
    cm = torch.zeros_like(c_i)
    cm[n_samp:] = 1.0
    t_is = torch.full((2 * n_samp, 1, 1, 1), 0.5, device=dev)

    def train_micro():
        for v in sd.values():
            v.grad = None
        loss = P.ddpm_loss(sd, sched, x, c, m, ts, noise, ctx, variant="rdd", n_T=CFG["n_T"], training=True, attn_map=m)
        loss.backward()

    def sample_step():
        with torch.no_grad():
            eps = P.unet_forward(sd, xs.repeat(2, 1, 1, 1), c_i, t_is, cm, variant="rdd", training=False, prefix="nn_model.")
            P.reverse_step(sched, xs, eps[:n_samp], eps[n_samp:], torch.randn_like(xs), 350, 2.0)
    out = {"what": "oracle/ref_port.py (the reference's aten calls) on cuda, PyTorch eager + cuDNN/cuBLAS", "torch": torch.__version__,
           "cudnn_allow_tf32": bool(torch.backends.cudnn.allow_tf32), "matmul_allow_tf32": bool(torch.backends.cuda.matmul.allow_tf32)}
    import contextlib
    for name, ctxmgr in (("fp32_tf32", contextlib.nullcontext), ("autocast_fp16", lambda: torch.autocast("cuda", dtype=torch.float16)),
                         ("autocast_bf16", lambda: torch.autocast("cuda", dtype=torch.bfloat16))):
        with ctxmgr():
            ms_t = _time_cuda(train_micro, 2, 3)
            ms_s = _time_cuda(sample_step, 2, 3)
        out[name] = {"train_micro_step_ms": ms_t, "train_imgs_per_s": CFG["batch"] / (ms_t * 1e-3),
                     "reverse_step_ms_n15": ms_s, "sampled_imgs_per_s": n_samp / (CFG["n_T"] * ms_s * 1e-3)}
    for v in sd.values():
        v.grad = None
    del sd
    torch.cuda.empty_cache()
    return out


def extra_configs(dev, steps):
    """Bounded figures for the other BASELINE.json configs, outside the headline: cfg1 (MNIST ContextUnet F=128, batch 128:
    optimizer step + a short CFG loop at n_sample=40) and cfg5 (stress: F=384, 3x256x256, batch 16: micro-step + AdamW, and a
    reverse step at n=15), through the same public API."""
    import diffusionmodel_b200 as D
    out = {}

    def train_fig(ddpm, batch_fn, n_img, gflop_img):
        opt = D.FusedAdamW(ddpm.parameters(), lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0)
        step = ddpm.capture_train_step(*batch_fn())
        opt.zero_grad()
        b = batch_fn()

        def one():
            step(*b)
            opt.step()
            opt.zero_grad()
        ms = _time_cuda(one, 3, steps)
        return {"ms_per_step": ms, "imgs_per_s": n_img / (ms * 1e-3), "tflops": gflop_img * n_img / ms}, opt
    # ---- cfg1
    torch.manual_seed(0)
    ddpm = D.DDPM(D.MnistContextUnet(1, 128, 10), (1e-4, 0.02), 400, dev, 0.1).to(dev).train()
    g = torch.Generator().manual_seed(1)
    fig, opt = train_fig(ddpm, lambda: (torch.rand(128, 1, 28, 28, generator=g).to(dev), torch.randint(0, 10, (128,), generator=g).to(dev)),
                         128, 8.24)
    ddpm.eval()
    ddpm.sample_noise = "device"
    ddpm.sample(40, (1, 28, 28), dev, guide_w=2.0, steps=3)
    ms = _time_cuda(lambda: ddpm.sample(40, (1, 28, 28), dev, guide_w=2.0, steps=10), 1, 1) / 10
    fig.update({"workload": "MNIST_script ContextUnet F=128, 1x28x28, batch 128: fwd + bwd + clip + AdamW (graphed micro-step)",
                "sample_ms_per_reverse_step_n40": ms, "sampled_imgs_per_s_n40": 40 / (400 * ms * 1e-3)})
    out["cfg1_mnist_f128_b128"] = fig
    del ddpm, opt
    torch.cuda.empty_cache()
    # ---- cfg5
    free, _ = torch.cuda.mem_get_info()
    if free < 120e9:
        out["cfg5_f384_b16"] = {"skipped": f"only {free / 1e9:.0f} GB free"}
        return out
    torch.manual_seed(0)
    ddpm = D.DDPM(D.ContextUnet(3, 384, 5), (1e-4, 0.02), 700, dev, 0.1, enhance_with_attn_map=True).to(dev).train()
    g = torch.Generator().manual_seed(2)
    fig, opt = train_fig(ddpm, lambda: tuple(t.to(dev) for t in synth_batch(g, 16, 256, 5)), 16, 16146.0)
    ddpm.eval()
    ddpm.sample_noise = "device"
    ddpm.sample(15, (3, 256, 256), dev, guide_w=2.0, steps=3)
    ms = _time_cuda(lambda: ddpm.sample(15, (3, 256, 256), dev, guide_w=2.0, steps=5), 1, 1) / 5
    fig.update({"workload": "new_scripy ContextUnet n_feat=384 (1.41 B parameters), 3x256x256, batch 16: fwd + bwd + clip + AdamW",
                "sample_ms_per_reverse_step_n15": ms, "sampled_imgs_per_s_n15": 15 / (700 * ms * 1e-3),
                "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9})
    out["cfg5_f384_b16"] = fig
    del ddpm, opt
    torch.cuda.empty_cache()
    return out


def sampling_sweep(ddpm, dev, world):
    """BASELINE.json configs[2]: the CFG reverse step over samples_per_class 1/3/8/16 (x 5 classes) at one guidance scale,
    and the three scales 2/4/6 as ONE trajectory batch (DDPM.sample(guide_w=[2, 4, 6])) against three sequential runs."""
    rows = []
    size = (3, CFG["img"], CFG["img"])

    def ms_per_step(n, gw, steps=4):
        ddpm.sample(n, size, dev, guide_w=gw, steps=2)                    # capture + allocator
        return _time_cuda(lambda: ddpm.sample(n, size, dev, guide_w=gw, steps=steps), 1, 1) / steps
    for spc in (1, 3, 8, 16):
        n = spc * CFG["n_classes"]
        ms = ms_per_step(n, 2.0)
        rows.append({"samples_per_class": spc, "trajectories": n, "guide_scales": [2.0], "ms_per_reverse_step": ms,
                     "imgs_per_s_per_gpu": n / (CFG["n_T"] * ms * 1e-3), "executed_tflops": n * (649.7 + 2 * (696.5 - 86.97)) / ms})
    for spc in (1, 3):
        n = spc * CFG["n_classes"]
        ms = ms_per_step(n, [2.0, 4.0, 6.0])
        seq = next(r for r in rows if r["samples_per_class"] == spc)["ms_per_reverse_step"] * 3
        rows.append({"samples_per_class": spc, "trajectories": 3 * n, "guide_scales": [2.0, 4.0, 6.0], "batched": True,
                     "ms_per_reverse_step": ms, "imgs_per_s_per_gpu": 3 * n / (CFG["n_T"] * ms * 1e-3),
                     "executed_tflops": 3 * n * (649.7 + 2 * (696.5 - 86.97)) / ms, "three_sequential_runs_ms": seq,
                     "speedup_vs_sequential": seq / ms})
    ddpm._sample_graphs.clear()
    torch.cuda.empty_cache()
    return rows


def non_gemm_aggregate(agg, hbm_peak):
    """Every launch of one eager step that is not a conv / weight-gradient GEMM: summed CUDA-event time against summed
    algorithmic bytes (ops._Profile._bytes: DESIGN.md section 3 per-element figures), and the same per entry point."""
    rows, ms, nb = [], 0.0, 0.0
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        if k in ("conv_gemm", "wgrad_gemm"):
            continue
        ms += v["ms"]
        nb += v["bytes"]
        gbs = v["bytes"] / v["ms"] / 1e6 if v["ms"] > 0 else 0.0
        rows.append({"entry": k, "launches": v["n"], "ms_per_step": round(v["ms"], 3), "algorithmic_GB": round(v["bytes"] / 1e9, 3),
                     "GBps": round(gbs, 1), "frac": round(gbs / hbm_peak, 3)})
    gbs = nb / ms / 1e6 if ms > 0 else 0.0
    return {"ms_per_step": ms, "algorithmic_GB_per_step": nb / 1e9, "GBps": gbs, "frac": gbs / hbm_peak, "peak": hbm_peak,
            "note": "eager launches with CUDA events around each C-ABI call (launch gaps of the small kernels included)",
            "entries": rows}


def hbm_kernel_table(dev, hbm_peak):
    """Achieved algorithmic GB/s of the bandwidth-bound kernel classes on the dominant activation shape of the step
    (4 x 256 x 256 x 192 bf16 = 100.7 MB per tensor).  Each kernel is launched 9 times back to back, rotating over 3
    operand sets (>= 600 MB in total, far above the 126 MB L2, so every launch reads cold data), between one pair of
    CUDA events; the nine launches are replayed as one CUDA graph so that the figure does not depend on how fast this
    box's host can issue 20 us kernels through ctypes (median of five replays); bytes = the DESIGN.md per-element figures x elements."""
    from diffusionmodel_b200 import _lib, ops
    P_ = ops._p
    n, h, c = CFG["batch"], CFG["img"], CFG["n_feat"]
    P = n * h * h
    E = P * c
    g = torch.Generator(device=dev).manual_seed(5)
    sets = []
    for _ in range(3):
        d = dict(y=torch.randn(n, h, h, c, device=dev, generator=g).to(torch.bfloat16),
                 dz=torch.randn(n, h, h, c, device=dev, generator=g).to(torch.bfloat16))
        d["z"] = torch.empty_like(d["y"])
        d["a"] = torch.randn(n, h // 2, h // 2, c, device=dev, generator=g).to(torch.bfloat16)
        d["b"] = torch.randn(n, h // 2, h // 2, c, device=dev, generator=g).to(torch.bfloat16)
        d["up"] = torch.empty((n, h, h, 2 * c), device=dev, dtype=torch.bfloat16)
        d["da"], d["db"] = torch.empty_like(d["a"]), torch.empty_like(d["b"])
        sets.append(d)
    mean, inv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
    ga, be = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    dga, dbe = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    scr = torch.empty(_lib.fn("dm_bn_act_bwd_scratch")(P, c), device=dev)
    part = torch.empty((_lib.fn("dm_bn_stats_rows")(P, c), 2, c), device=dev)
    gate = torch.rand(n, c, device=dev)
    ns = 15
    eps = torch.randn(2 * ns, h, h, 4, device=dev)
    xs, zs = torch.randn(ns, 3, h, h, device=dev), torch.randn(ns, 3, h, h, device=dev)
    xo, xt = torch.empty_like(xs), torch.empty((2 * ns, h, h, 8), device=dev, dtype=torch.bfloat16)
    gnm, gnr = torch.empty(n * 8, device=dev), torch.empty(n * 8, device=dev)
    gscr = torch.empty(_lib.fn("dm_gn_scratch")(n, h * h, c), device=dev)
    xh, xw = torch.empty(n, h, c, device=dev), torch.empty(n, h, c, device=dev)
    ah, aw = torch.rand(n, h, c, device=dev), torch.rand(n, h, c, device=dev)
    pooled = torch.empty(n, c, device=dev)
    mask = torch.rand(n, h, h, device=dev) * 2
    # small-tensor kernels at the shapes the step uses them on
    x4, nz4 = torch.randn(n, 3, h, h, device=dev), torch.randn(n, 3, h, h, device=dev)
    sab, smab = torch.rand(CFG["n_T"] + 1, device=dev), torch.rand(CFG["n_T"] + 1, device=dev)
    ts4 = torch.randint(1, CFG["n_T"] + 1, (n,), device=dev)
    xt4 = torch.empty(n, h, h, 8, device=dev, dtype=torch.bfloat16)
    pred4, dpred4 = torch.randn(n, h, h, 4, device=dev), torch.empty(n, h, h, 4, device=dev)
    loss_s, lscr, gout = torch.empty((), device=dev), torch.empty(8, device=dev), torch.ones((), device=dev)
    lv = (1.2, 0.8, 3.0, 1.0, 0.5, 2.0)
    c8 = 8 * c
    fx = torch.randn(n, 16, 16, c8, device=dev).to(torch.bfloat16)
    fo = torch.empty_like(fx)
    ce, te = torch.randn(n, c8, device=dev), torch.randn(n, c8, device=dev)
    kernels = [
        ("bn_stats_kernel", "dm_bn_stats", 2 * E,
         lambda d: ops.call("dm_bn_stats", P_(d["y"]), c, P_(part), c, P, c, ops._stream())),
        ("bn_fwd_kernel (BatchNorm + GELU)", "dm_bn_act_fwd", 4 * E,
         lambda d: ops.call("dm_bn_act_fwd", P_(d["y"]), c, P_(mean), P_(inv), P_(ga), P_(be), P_(d["z"]), c, P, c, 1, ops._stream())),
        ("bn_bwd_reduce + finalize + apply", "dm_bn_act_bwd", 10 * E,
         lambda d: ops.call("dm_bn_act_bwd", P_(d["dz"]), c, P_(d["y"]), c, P_(mean), P_(inv), P_(ga), P_(be), P_(d["z"]), c,
                            P_(dga), P_(dbe), None, P_(scr), P, c, 1, 1, ops._stream())),
        ("gn_fwd (GroupNorm(8) + GELU: stats + apply)", "dm_gn_act_fwd", 6 * E,
         lambda d: ops.call("dm_gn_act_fwd", P_(d["y"]), c, P_(ga), P_(be), P_(d["z"]), c, P_(gnm), P_(gnr), P_(gscr), n, h * h, c,
                            8, 1e-5, 1, ops._stream())),
        ("gn_bwd (reduce + apply)", "dm_gn_act_bwd", 10 * E,
         lambda d: ops.call("dm_gn_act_bwd", P_(d["dz"]), c, P_(d["y"]), c, P_(gnm), P_(gnr), P_(ga), P_(be), P_(d["z"]), c,
                            P_(dga), P_(dbe), P_(gscr), n, h * h, c, 8, 1, ops._stream())),
        ("ew_kernel<SeFwd> (SE gate + residual)", "dm_se_apply_fwd", 6 * E,
         lambda d: ops.call("dm_se_apply_fwd", P_(d["y"]), c, P_(gate), P_(d["dz"]), c, P_(d["z"]), c, n, h * h, c, 0.7072, ops._stream())),
        ("nc_reduce (SE global average pool)", "dm_pool_nhw", 2 * E,
         lambda d: ops.call("dm_pool_nhw", P_(d["y"]), c, P_(pooled), n, h * h, c, 1.0 / (h * h), ops._stream())),
        ("ca_pool (CoordAttn row + column means)", "dm_ca_pool", 2 * E,
         lambda d: ops.call("dm_ca_pool", P_(d["y"]), c, None, 0, P_(xh), P_(xw), n, h, h, c, 1.0 / h, 1.0 / h, ops._stream())),
        ("ca_gate_fwd (x * (a_h + a_w))", "dm_ca_gate_fwd", 4 * E,
         lambda d: ops.call("dm_ca_gate_fwd", P_(d["y"]), c, P_(ah), P_(aw), P_(d["z"]), c, n, h, h, c, ops._stream())),
        ("ca_gate_bwd", "dm_ca_gate_bwd", 4 * E,
         lambda d: ops.call("dm_ca_gate_bwd", P_(d["dz"]), c, P_(ah), P_(aw), P_(xh), P_(xw), P_(d["z"]), c, n, h, h, c, ops._stream())),
        ("mask_fma (LocalEnhancer x + y * (mask > 1.2))", "dm_mask_fma", 6 * E,
         lambda d: ops.call("dm_mask_fma", P_(d["y"]), c, P_(d["dz"]), c, P_(mask), 1.2, P_(d["z"]), c, P, c, ops._stream())),
        ("upcat_fwd_quad_kernel (cat + bilinear x2)", "dm_upcat_fwd", (E // 4 * 2 + 2 * E) * 2,
         lambda d: ops.call("dm_upcat_fwd", P_(d["a"]), c, c, P_(d["b"]), c, c, P_(d["up"]), 2 * c, n, h // 2, h // 2, ops._stream())),
        ("upcat_bwd_quad_kernel", "dm_upcat_bwd", (E // 4 * 2 + 2 * E) * 2,
         lambda d: ops.call("dm_upcat_bwd", P_(d["up"]), 2 * c, P_(d["da"]), c, c, P_(d["db"]), c, c, n, h // 2, h // 2, ops._stream())),
        ("film_fwd (cemb * x + temb, 4x16x16x1536: 3 MB, launch-latency class)", "dm_film_fwd", 4 * n * 256 * c8,
         lambda d: ops.call("dm_film_fwd", P_(fx), c8, P_(ce), P_(te), P_(fo), c8, n, 256, c8, ops._stream())),
        ("q_sample (4x3x256x256: 8 MB, launch-latency class)", "dm_q_sample", 10 * n * 3 * h * h,
         lambda d: ops.call("dm_q_sample", P_(x4), P_(nz4), P_(sab), P_(smab), P_(ts4), P_(xt4), 8, n, 3, h, h, ops._stream())),
        ("ddpm_loss_fwd (4x3x256x256, launch-latency class)", "dm_ddpm_loss_fwd", (8 + 4.0 / 3) * n * 3 * h * h,
         lambda d: ops.call("dm_ddpm_loss_fwd", P_(pred4), 4, P_(nz4), P_(mask), P_(loss_s), P_(lscr), n, 3, h, h, *lv, ops._stream())),
        ("ddpm_loss_bwd (4x3x256x256, launch-latency class)", "dm_ddpm_loss_bwd", (12 + 4.0 / 3) * n * 3 * h * h,
         lambda d: ops.call("dm_ddpm_loss_bwd", P_(pred4), 4, P_(nz4), P_(mask), P_(gout), P_(dpred4), 4, n, 3, h, h, *lv, ops._stream())),
        ("cfg_reverse_step_kernel (n=15)", "dm_cfg_reverse_step", ns * 3 * h * h * 20 + 2 * ns * h * h * 16,
         lambda d: ops.call("dm_cfg_reverse_step", P_(eps), 4, P_(xs), P_(zs), P_(xo), P_(xt), 8, 2.0, 1.01, 0.02, 0.1, ns, 3,
                            h, h, ops._stream())),
    ]
    rows = []
    for kname, entry, nbytes, fn in kernels:
        for d in sets:
            fn(d)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(9):
                fn(sets[i % 3])
        graph.replay()
        torch.cuda.synchronize()
        reps = []
        for _ in range(5):                      # median of five replays (clocks move under the power cap)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            reps.append(e0.elapsed_time(e1))
        us = sorted(reps)[2] / 9 * 1e3
        gbs = nbytes / us / 1e3
        rows.append({"kernel": kname, "entry": entry, "us": round(us, 1), "algorithmic_MB": round(nbytes / 1e6, 1),
                     "GBps": round(gbs, 1), "frac": round(gbs / hbm_peak, 3)})
    return rows


def run_ours(args):
    import diffusionmodel_b200 as D
    from diffusionmodel_b200 import _lib, ops, parallel
    rank, local, world = parallel.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)                                  # identical initial weights on every rank
    net = D.ContextUnet(CFG["in_ch"], CFG["n_feat"], CFG["n_classes"])
    ddpm = D.DDPM(net, CFG["betas"], CFG["n_T"], dev, CFG["drop_prob"], enhance_with_attn_map=True)
    ddpm.to(dev).train()
    opt = D.FusedAdamW(ddpm.parameters(), lr=CFG["lr"], weight_decay=CFG["wd"], max_grad_norm=1.0)
    parallel.broadcast_parameters(opt.flat_param, [b for b in ddpm.buffers()])
    torch.manual_seed(1234 + rank)                        # per-rank data / noise streams
    gen = torch.Generator().manual_seed(100 + rank)
    accum, batch = CFG["accum"], CFG["batch"]
    host = [tuple(t.pin_memory() for t in synth_batch(gen, batch, CFG["img"], CFG["n_classes"])) for _ in range(accum)]
    resident = [tuple(t.to(dev) for t in hb) for hb in host]
    h2d = sum(t.numel() * t.element_size() for hb in host for t in hb)

    def micro_eager(x, c, m):
        loss = ddpm(x, c, m) / accum
        loss.backward()
        return loss

    # the public fast path: one CUDA graph per micro-step (forward + backward), see DDPM.capture_train_step
    use_graph = os.environ.get("DM_BENCH_GRAPH", "1") != "0"
    micro = micro_eager
    last, red = None, None
    if use_graph:
        micro = ddpm.capture_train_step(*resident[0], loss_scale=1.0 / accum)
        if world > 1 and os.environ.get("DM_BENCH_OVERLAP", "1") != "0":
            # N > 1: the last micro-step of the window as two graphs, the all-reduce of the decoder-side gradients (62 % of
            # the bytes) issued between them and hidden behind the trunk's backward (parallel.OverlappedGradReduce)
            last = ddpm.capture_train_step(*resident[0], loss_scale=1.0 / accum, split_backward=True,
                                           trunk_sm_limit=148 - parallel.NCCL_CTAS)
            red = parallel.OverlappedGradReduce(opt, net.grad_ready_regions())
        opt.zero_grad()

    def reduce_and_update():
        if red is not None:
            red.finish()
        else:
            opt.flush()
            parallel.allreduce_mean_(opt.flat_grad)
        opt.step()
        opt.zero_grad()

    def step_resident():
        for i, (x, c, m) in enumerate(resident):
            if last is not None and i == accum - 1:
                last(x, c, m, between=red.reduce_ready)
            else:
                micro(x, c, m)
        reduce_and_update()

    def step_e2e():
        # pinned host batches -> device every micro-step; the running loss (new_scripy.py:789 reads it per micro-batch
        # for the progress bar) is accumulated on the device and read back once per optimizer step, so the host never
        # stalls the launch queue inside the accumulation window
        tot = None
        for i, (hx, hc, hm) in enumerate(host):
            x, c, m = hx.to(dev, non_blocking=True), hc.to(dev, non_blocking=True), hm.to(dev, non_blocking=True)
            if last is not None and i == accum - 1:
                l = last(x, c, m, between=red.reduce_ready).detach()
            else:
                l = micro(x, c, m).detach()
            tot = l.clone() if tot is None else tot + l
        reduce_and_update()
        return float(tot)                                 # device->host read of the step's loss

    def timed(fn, steps):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        return parallel.max_over_ranks(e0.elapsed_time(e1), dev)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    ms = timed(step_resident, args.steps)
    launches = _lib.launch_count() - l0
    if use_graph:                                         # kernels inside the replayed graphs run without a C-ABI call
        launches += micro.kernels_per_replay * accum * args.steps
    clocks = sampler.stop() if rank == 0 else None
    imgs = accum * batch * world * args.steps
    value = imgs / (ms * 1e-3)
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    e2e = imgs / (ms_e2e * 1e-3)

    # roofline of the dominant kernel class (conv_gemm_kernel: fwd + dgrad implicit GEMMs), CUDA events
    # around every launch of one extra instrumented step on the launching stream
    micro, last, red = micro_eager, None, None            # CUDA events cannot be recorded inside a graph replay
    step_resident()
    prof = ops.enable_profile()
    step_resident()
    torch.cuda.synchronize()
    ops.disable_profile()
    agg = prof.summary()
    if rank == 0 and os.environ.get("DM_BENCH_BREAKDOWN"):
        with open(os.environ["DM_BENCH_BREAKDOWN"], "w") as f:
            json.dump({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}, f, indent=1)
    # ---- sampled img/s: the CFG reverse loop (new_scripy.py:441-477) on this rank's shard of trajectories:
    # samples_per_class=3 x 5 classes (the CLI default --samples 3), guide_w=2.0, timed over a bounded number
    # of the n_T=700 identical reverse steps; device-side noise (the reference's per-step CPU randn + H2D is
    # kept as DDPM(sample_noise="reference") for parity runs)
    fast = os.environ.get("DM_BENCH_FAST") == "1"          # profiling runs: skip the sampling loop and the CPU baseline
    ddpm.eval()
    ddpm.sample_noise = "device"
    n_samp, s_steps = 3 * CFG["n_classes"], (1 if fast else 2 * max(args.steps, 5))
    # warm-up: weight packs / BatchNorm folds, graph capture, and one call of the timed length so the caching allocator
    # has settled (a one-off cudaMalloc/cudaFree inside a 10-step timed call once read as +9 ms per step)
    ddpm.sample(n_samp, (3, CFG["img"], CFG["img"]), dev, guide_w=2.0, steps=3)
    ddpm.sample(n_samp, (3, CFG["img"], CFG["img"]), dev, guide_w=2.0, steps=s_steps)
    torch.cuda.synchronize()
    ms_samp = timed(lambda: ddpm.sample(n_samp, (3, CFG["img"], CFG["img"]), dev, guide_w=2.0, steps=s_steps), 1) / s_steps
    gf_exec = n_samp * (649.7 + 2 * (696.5 - 86.97))      # shared encoder once, decoder twice, LocalEnhancer(+0) skipped
    sampling = {"n_sample_per_gpu": n_samp, "guide_w": 2.0, "n_T": CFG["n_T"], "steps_timed": s_steps,
                "ms_per_reverse_step": ms_samp, "imgs_per_s": world * n_samp / (CFG["n_T"] * ms_samp * 1e-3),
                "noise": "device", "shared_encoder_cfg": True, "executed_tflops": gf_exec / ms_samp,
                "reference_schedule_tflops": n_samp * 2 * 1346.17 / ms_samp}
    ddpm.train()

    pk, src = peaks()
    line = None
    if rank == 0:
        conv = agg.get("conv_gemm", {"ms": 0.0, "flops": 0.0, "n": 0})
        wg = agg.get("wgrad_gemm", {"ms": 0.0, "flops": 0.0, "n": 0})
        ach = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else 0.0
        peak = pk["bf16_tflops_sustained"]
        roof = {"bound": "tensor",
                "kernel": "conv3x3_halo2_kernel / conv_gemm2_kernel (CTA pairs) + conv3x3_halo_kernel + conv_gemm_kernel: fwd + data-gradient implicit GEMMs, all layers of one step",
                "achieved": ach, "peak": peak, "peak_source": f"{src} bf16_tflops_sustained", "unit": "TFLOP/s",
                "frac": ach / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch on the dominant layer (3x3, 192->192,
                # 4x256x256: 100.7 MB in, 100.7 MB out algorithmic), profiles/r02_dominant_layer_ncu_full.txt
                "traffic": 152.1e6, "traffic_unit": "B/launch (ncu --set full, dominant layer, profiles/r02_dominant_layer_ncu_full.txt: dram read 101.5 MB + write 50.7 MB, the rest of the 100.7 MB output still in L2 at kernel end; algorithmic 201.3 MB)",
                "launches": conv["n"], "ms_per_step": conv["ms"],
                "wgrad": {"achieved": wg["flops"] / (wg["ms"] * 1e-3) / 1e12 if wg["ms"] > 0 else 0.0,
                          "ms_per_step": wg["ms"], "launches": wg["n"]},
                "other_kernels_ms_per_step": sum(v["ms"] for k, v in agg.items() if k not in ("conv_gemm", "wgrad_gemm")),
                "non_gemm": non_gemm_aggregate(agg, pk["hbm_gbs"]),
                "step_tflops": GFLOP_PER_IMG_TRAIN * accum * batch / 1e3 / (ms / args.steps * 1e-3) / 1e0 / 1e0}
        roof["step_frac_of_peak"] = roof["step_tflops"] / peak
        if world == 1 and use_graph and roof["non_gemm"].get("algorithmic_GB_per_step"):
            # what the non-GEMM launches cost where the step actually runs (CUDA graphs, 1.2 us between dependent kernels):
            # the graphed step minus the event-timed totals of the two GEMM classes.  The GEMMs run no faster inside the
            # graph (power cap), so this is an upper bound on the non-GEMM time, a lower bound on its bandwidth.
            in_graph = ms / args.steps - conv["ms"] - wg["ms"]
            gb = roof["non_gemm"]["algorithmic_GB_per_step"]
            roof["non_gemm"]["in_graph_estimate"] = {
                "ms_per_step": in_graph, "GBps": gb / (in_graph * 1e-3), "frac": gb / (in_graph * 1e-3) / pk["hbm_gbs"],
                "how": "graphed ms_per_step - fwd/dgrad GEMM ms - wgrad GEMM ms (both event-timed in the eager instrumented step)"}
        # the bandwidth-bound kernel classes against the measured HBM copy peak (cfg_reverse_step works on 23.6 MB: L2-sized)
        roof["hbm_kernels"] = {"peak": pk["hbm_gbs"], "peak_source": f"{src} hbm_gbs", "unit": "GB/s",
                               "shape": "4x256x256x192 bf16 NHWC (100.7 MB per tensor), cold operands",
                               "rows": hbm_kernel_table(dev, pk["hbm_gbs"])}
        ad = agg.get("dm_adamw_bf16") or agg.get("dm_adamw")
        if ad and ad["ms"] > 0:
            nb = 30.0 * opt._n
            roof["hbm_kernels"]["rows"].append({"kernel": "adamw_bf16_kernel (clip + AdamW + bf16 weight copy, whole model)",
                                                "entry": "dm_adamw_bf16", "us": round(ad["ms"] * 1e3 / ad["n"], 1),
                                                "algorithmic_MB": round(nb / 1e6, 1), "GBps": round(nb / (ad["ms"] / ad["n"]) / 1e6, 1),
                                                "frac": round(nb / (ad["ms"] / ad["n"]) / 1e6 / pk["hbm_gbs"], 3)})
        cpu = cpu_baseline_sample() if (world == 1 and not fast) else None
        eager = extra = None
        if world == 1 and not fast:
            # the legs below are reported beside the headline; a failure in one of them must not lose the line
            def guarded(fn, *a):
                try:
                    return fn(*a)
                except Exception as e:          # noqa: BLE001
                    torch.cuda.empty_cache()
                    return {"error": f"{type(e).__name__}: {e}"[:300]}
            ddpm.eval()
            sampling["sweep"] = guarded(sampling_sweep, ddpm, dev, world)
            ddpm.train()
            # drop the cfg2 model before the baselines allocate theirs (closures above share these cells)
            del micro, opt, ddpm, net, resident
            ops.release_registries()
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            eager = guarded(gpu_eager_baseline, dev)
            extra = guarded(extra_configs, dev, 5)
        line = {"metric": "ddpm_train_imgs_per_s", "value": value, "unit": "img/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(world),
                "e2e": {"value": e2e, "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
                "sampling": sampling}
        if cpu is not None:
            line["cpu_baseline"] = cpu
            line["gpu_eager_baseline"] = eager
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
