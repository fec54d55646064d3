"""`--mode train/generate` command line of the reference (new_scripy.py:1292-1321) over the B200 hot path.

    python -m diffusionmodel_b200.cli --mode train [--epochs E --steps_per_epoch S]
    python -m diffusionmodel_b200.cli --mode generate --ckpt CKPT --guide_scales 2 4 --samples 3 [--no_eval]
    torchrun --nproc-per-node N -m diffusionmodel_b200.cli --mode train        # data parallel, one rank per GPU

Flag names, defaults and the missing-checkpoint behaviour are the reference's.  The drivers are the thin
caller contract of SURVEY.md 2.1 (train_model new_scripy.py:659-943, gen_samples :945-1108): accumulation
over ACCUM_STEPS micro-batches, clip 1.0, AdamW(lr 1e-4, wd 1e-5), CosineAnnealingWarmRestarts(10, 2, 3e-5),
checkpoints as ``{'epoch', 'model_state_dict', 'optimizer_state_dict', 'scheduler_state_dict', 'loss', 'metrics'}``
(:736-743; ``--resume CKPT`` continues from one, optimizer moments and scheduler position included).  The reference's dataset
(./cropped_images, VOC XML) is not shipped with it: ``--data DIR`` reads that layout through the device-side batch
pipeline (data.py), otherwise batches are synthetic road-damage-shaped tensors
(images in [-1,1], labels, attention map 0.5 / 1.0 lower half / 3.0 box, new_scripy.py:535-546); FID/SSIM
evaluation (new_scripy.py:1111-1290): ``generate --data DIR`` compares the samples with real images (SSIM / PSNR on the
device; FID needs a feature network, see metrics.py) and writes ``quality_metrics.json`` like the reference; ``--no_eval`` skips it.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import torch

from . import DDPM, ContextUnet, FusedAdamW, parallel


class Cfg:                      # new_scripy.py:22-67 (only what the drivers read)
    N_FEAT, IN_CH, N_T, BETAS, DROP_PROB = 192, 3, 700, (1e-4, 0.02), 0.1
    BATCH_SIZE, ACCUM_STEPS, LR, WD, N_EPOCH = 4, 4, 1e-4, 1e-5, 400
    SAVE_DIR, SAMPLE_DIR = "./output/diffusion/", "./output/samples/"
    GUIDE_SCALES, SAMPLES_PER_CLASS, IMG_SIZE = [2.0, 4.0], 3, 256


def synth_batch(gen, batch, img, n_classes):
    x = torch.rand(batch, 3, img, img, generator=gen) * 2 - 1
    c = torch.randint(0, n_classes, (batch,), generator=gen)
    m = torch.full((batch, img, img), 0.5)
    m[:, img // 2:, :] = 1.0
    for b in range(batch):
        xs = torch.randint(0, img, (2,), generator=gen).sort().values
        ys = torch.randint(0, img, (2,), generator=gen).sort().values
        m[b, int(ys[0]):int(ys[1]) + 1, int(xs[0]):int(xs[1]) + 1] = 3.0
    return x, c, m


def build(n_classes, device, n_feat=Cfg.N_FEAT, enhance_with_attn_map=False):
    """Default = the shipped semantics (LocalEnhancer contributes +0 in training AND in sampling, new_scripy.py:353).
    ``enhance_with_attn_map=True`` feeds the attention map to LocalEnhancer during training; ``DDPM.sample`` has no
    attention map to give it, so a model trained that way is sampled without the enhancer term."""
    net = ContextUnet(in_ch=Cfg.IN_CH, n_feat=n_feat, n_classes=n_classes)
    return DDPM(nn_model=net, betas=Cfg.BETAS, n_T=Cfg.N_T, device=device, drop_prob=Cfg.DROP_PROB,
                enhance_with_attn_map=enhance_with_attn_map).to(device)


def save_ckpt(path, ddpm, optim, sched, epoch, loss, metrics_log=None):
    """The reference's checkpoint dict (new_scripy.py:730-744)."""
    torch.save({"epoch": epoch, "model_state_dict": ddpm.state_dict(), "optimizer_state_dict": optim.state_dict(),
                "scheduler_state_dict": sched.state_dict(), "loss": loss, "metrics": metrics_log or {}}, path)


def train_model(args):
    rank, local, world = parallel.init_from_env()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    torch.manual_seed(0)
    cache = None
    if args.data:                      # the reference's ./cropped_images layout: decoded + resized once, cached on the device
        from .data import CachedCrackBatches
        cache = CachedCrackBatches.from_voc_dir(args.data, args.img, device)
        args.n_classes = len(cache.classes)                      # new_scripy.py:692
    ddpm = build(args.n_classes, device, args.n_feat, args.enhance_attn_map).train()
    optim = FusedAdamW(ddpm.parameters(), lr=Cfg.LR, weight_decay=Cfg.WD, max_grad_norm=1.0)
    sched = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(optim, T_0=10, T_mult=2, eta_min=3e-5)
    first_epoch = 0
    if args.resume:
        ck = torch.load(args.resume, map_location="cpu")
        ddpm.load_state_dict(ck["model_state_dict"])
        if "optimizer_state_dict" in ck:
            optim.load_state_dict(ck["optimizer_state_dict"])
        if "scheduler_state_dict" in ck:
            sched.load_state_dict(ck["scheduler_state_dict"])
        first_epoch = int(ck.get("epoch", -1)) + 1
    parallel.broadcast_parameters(optim.flat_param, list(ddpm.buffers()))
    gen = torch.Generator().manual_seed(100 + rank)
    torch.manual_seed(1234 + rank)
    step_fn = last_fn = red = None
    os.makedirs(Cfg.SAVE_DIR, exist_ok=True)
    ema = None
    for ep in range(first_epoch, args.epochs):
        t0, ema, seen = time.time(), None, 0
        for step in range(args.steps_per_epoch):
            if cache is not None:      # random batch of the cached set; flip / normalise / mask rasterisation in one kernel
                x, c, m = cache.batch(torch.randint(0, len(cache), (Cfg.BATCH_SIZE,), generator=gen), generator=gen)
            else:
                x, c, m = (t.to(device, non_blocking=True) for t in synth_batch(gen, Cfg.BATCH_SIZE, args.img, args.n_classes))
            closes = (step + 1) % Cfg.ACCUM_STEPS == 0 or step + 1 == args.steps_per_epoch      # :795
            if step_fn is None and not args.no_graph:
                step_fn = ddpm.capture_train_step(x, c, m, loss_scale=1.0 / Cfg.ACCUM_STEPS)
                if world > 1:      # last micro-step of a window: staged backward, gradient all-reduce hidden behind it
                    last_fn = ddpm.capture_train_step(x, c, m, loss_scale=1.0 / Cfg.ACCUM_STEPS, split_backward=True,
                                                      trunk_sm_limit=148 - parallel.NCCL_CTAS)
                    red = parallel.OverlappedGradReduce(optim, ddpm.nn_model.grad_ready_regions())
                optim.zero_grad()
            if step_fn is not None:
                loss = last_fn(x, c, m, between=red.reduce_ready) if (closes and last_fn is not None) else step_fn(x, c, m)
            else:
                loss = ddpm(x, c, m) / Cfg.ACCUM_STEPS              # new_scripy.py:785-786
                ddpm.scaler.scale(loss).backward()                 # :792 (disabled scaler: bf16)
            li = loss.item() * Cfg.ACCUM_STEPS
            ema = li if ema is None else 0.95 * ema + 0.05 * li    # :806-809
            seen += Cfg.BATCH_SIZE * world
            if closes:
                if red is not None:
                    red.finish()
                else:
                    optim.flush()
                    parallel.allreduce_mean_(optim.flat_grad)
                optim.step()                                       # unscale / clip 1.0 / AdamW, :797-801
                optim.zero_grad()
        sched.step()                                               # :848
        if rank == 0:
            dt = time.time() - t0
            print(f"epoch {ep}: loss(ema) {ema:.4f}  lr {optim.param_groups[0]['lr']:.2e}  {seen / dt:.1f} img/s", flush=True)
    if rank == 0:
        path = os.path.join(Cfg.SAVE_DIR, "best_model.pt")
        save_ckpt(path, ddpm, optim, sched, args.epochs - 1, ema)                       # :730-744, final save :926
        print(f"saved {path}")
    return ddpm


def gen_samples(ckpt, n_samples_per_class=3, guide_scales=(2.0, 4.0), eval_quality=True, n_classes=5, n_feat=Cfg.N_FEAT,
                img=Cfg.IMG_SIZE, sequential_scales=False, data=None, feature_fn=None):
    rank, local, world = parallel.init_from_env()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    try:
        checkpoint = torch.load(ckpt, map_location="cpu")          # new_scripy.py:966-969
    except Exception as e:                                         # the reference swallows load errors and returns None
        print(f"Error loading checkpoint: {e}")
        return None
    ddpm = build(n_classes, device, n_feat)
    ddpm.drop_prob = 0.0
    state = checkpoint["model_state_dict"] if isinstance(checkpoint, dict) and "model_state_dict" in checkpoint else checkpoint
    ddpm.load_state_dict(state)                                    # :976-990 (raw-dict fallback)
    ddpm.eval()
    os.makedirs(Cfg.SAMPLE_DIR, exist_ok=True)
    out = {}
    n_sample = parallel.shard_samples(n_samples_per_class * n_classes, n_classes, rank, world)
    # real images for the quality assessment (new_scripy.py:1001-1029): n_samples_per_class * min(n_classes, 4) of them, from
    # the dataset directory if one is given (the reference reads ./cropped_images; nothing is shipped with it)
    real_images = None
    if eval_quality and data:
        from .data import CachedCrackBatches
        cache = CachedCrackBatches.from_voc_dir(data, img, device)
        need = min(len(cache), n_samples_per_class * min(n_classes, 4))
        order = torch.randperm(len(cache), generator=torch.Generator().manual_seed(0))[:need]
        real_images = cache.batch(order, flips=torch.zeros(need, dtype=torch.int32))[0]
    elif eval_quality and rank == 0:
        print("No dataset directory (--data): image-quality assessment skipped")
    quality_metrics = {}
    # the reference loops over the guidance scales (:1036-1041); here all scales run as ONE trajectory batch unless
    # --sequential_scales asks for the reference's loop (same trajectories, S times the kernel launches)
    groups = [[w] for w in guide_scales] if sequential_scales else [list(guide_scales)]
    for ws in groups:
        t0 = time.time()
        xs = ddpm.sample(n_sample, (3, img, img), device, guide_w=ws) if n_sample else []
        torch.cuda.synchronize()
        for w, x_gen in zip(ws, xs):
            torch.save(x_gen.cpu(), os.path.join(Cfg.SAMPLE_DIR, f"samples_w{w}_rank{rank}.pt"))
            out[w] = x_gen
            if real_images is not None and len(real_images) > 0:                 # :1067-1078
                from . import metrics
                k = min(len(real_images), len(x_gen))
                m = metrics.evaluate_batch(real_images[:k], x_gen[:k], feature_fn=feature_fn)
                quality_metrics[w] = m
                if rank == 0:
                    print(f"Image quality metrics (w={w}):")
                    for name, value in m.items():
                        print(f"  {name.upper()}: {value:.4f}")
        if rank == 0:
            print(f"guide_w={ws}: {n_sample} samples/rank per scale in {time.time() - t0:.1f}s", flush=True)
    if quality_metrics and rank == 0:                                            # :1085-1101
        import json
        path = os.path.join(Cfg.SAMPLE_DIR, "quality_metrics.json")
        with open(path, "w") as f:
            json.dump({str(k): {kk: float(vv) for kk, vv in v.items()} for k, v in quality_metrics.items()}, f, indent=2)
        print(f"Quality metrics saved to: {path}")
    out["quality_metrics"] = quality_metrics
    return out


def main(argv=None):
    parser = argparse.ArgumentParser(description="Enhanced Diffusion Model Training/Generation")
    parser.add_argument("--mode", type=str, default="train", choices=["train", "generate"],
                        help="Mode: 'train' for training, 'generate' for sample generation")
    parser.add_argument("--ckpt", type=str, default=None, help="Checkpoint path for generation mode")
    parser.add_argument("--guide_scales", type=float, nargs="+", default=Cfg.GUIDE_SCALES, help="Guidance scales for generation")
    parser.add_argument("--samples", type=int, default=Cfg.SAMPLES_PER_CLASS, help="Number of samples per class")
    parser.add_argument("--no_eval", action="store_true", help="Skip image quality evaluation")
    # additions (the reference hard-codes these in Cfg / the dataset)
    parser.add_argument("--epochs", type=int, default=Cfg.N_EPOCH)
    parser.add_argument("--steps_per_epoch", type=int, default=64)
    parser.add_argument("--n_classes", type=int, default=5)
    parser.add_argument("--n_feat", type=int, default=Cfg.N_FEAT)
    parser.add_argument("--img", type=int, default=Cfg.IMG_SIZE)
    parser.add_argument("--sequential_scales", action="store_true",
                        help="generate: one sampling run per guidance scale like the reference, instead of one batch of all scales")
    parser.add_argument("--resume", type=str, default=None, help="checkpoint to continue training from")
    parser.add_argument("--enhance_attn_map", action="store_true",
                        help="feed the attention map to LocalEnhancer while training (the shipped call site passes ctx_mask: +0)")
    parser.add_argument("--no_graph", action="store_true", help="eager launches instead of the CUDA-graphed micro-step")
    parser.add_argument("--data", type=str, default=None,
                        help="dataset root in the reference's layout (images/<class>/*.jpg + annotations/*.xml); synthetic batches if omitted")
    args = parser.parse_args(argv)
    if args.mode == "train":
        train_model(args)
    elif args.mode == "generate":
        if args.ckpt is None:
            print("Error: Checkpoint path required for generation mode")
            parser.print_help()
            sys.exit(1)
        gen_samples(args.ckpt, n_samples_per_class=args.samples, guide_scales=args.guide_scales,
                    eval_quality=not args.no_eval, n_classes=args.n_classes, n_feat=args.n_feat, img=args.img,
                    sequential_scales=args.sequential_scales, data=args.data)


if __name__ == "__main__":
    main()
