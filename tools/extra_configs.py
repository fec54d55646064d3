"""Dev tool: the other BASELINE.json configurations on one B200 (results -> gpurun_out/extra_configs.json).

  cfg1  MNIST_script ContextUnet DDPM, 1x28x28, 10 classes, n_feat=128, batch 128: train step + 20 sampling steps
  cfg3  new_scripy CFG sampling, samples_per_class in {1, 3, 8}, guide_w in {2, 4, 6}: ms per reverse step
  cfg5  stress shape: 3x256x256, n_feat=384 (doubled), batch 8: train micro-step
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import diffusionmodel_b200 as D  # noqa: E402


def timed(fn, n=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda:0")
    out = {}
    which = sys.argv[1:] or ["cfg1", "cfg3", "cfg5"]
    if "cfg1" in which:
        torch.manual_seed(0)
        ddpm = D.DDPM(D.MnistContextUnet(1, 128, 10), (1e-4, 0.02), 400, dev, 0.1).to(dev).train()
        opt = D.FusedAdamW(ddpm.parameters(), lr=1e-4, weight_decay=0.0)
        x, c = torch.rand(128, 1, 28, 28, device=dev), torch.randint(0, 10, (128,), device=dev)
        step = ddpm.capture_train_step(x, c)
        opt.zero_grad()

        def train():
            step(x, c); opt.step(); opt.zero_grad()
        ms = timed(train, 20, 5)
        ddpm.eval(); ddpm.sample_noise = "device"
        ms_s = timed(lambda: ddpm.sample(40, (1, 28, 28), dev, guide_w=2.0, steps=20), 2, 1) / 20
        out["cfg1_mnist_f128_b128"] = {"train_ms_per_step": ms, "train_img_per_s": 128 / ms * 1e3,
                                      "sample_n40_ms_per_reverse_step": ms_s, "sampled_img_per_s_400_steps": 40 / (400 * ms_s) * 1e3}
        print(out["cfg1_mnist_f128_b128"], flush=True)
        del ddpm, opt, step
    if "cfg3" in which:
        torch.manual_seed(0)
        ddpm = D.DDPM(D.ContextUnet(3, 192, 5), (1e-4, 0.02), 700, dev, 0.1).to(dev).eval()
        ddpm.sample_noise = "device"
        res = {}
        for spc in (1, 3, 8):
            for w in (2.0, 4.0, 6.0) if spc == 3 else (2.0,):
                n = spc * 5
                ms = timed(lambda: ddpm.sample(n, (3, 256, 256), dev, guide_w=w, steps=5), 1, 1) / 5
                res[f"samples_per_class={spc},guide_w={w}"] = {"n_sample": n, "ms_per_reverse_step": ms,
                                                                "imgs_per_s_700_steps": n / (700 * ms) * 1e3}
                print(spc, w, res[f"samples_per_class={spc},guide_w={w}"], flush=True)
        out["cfg3_sampling"] = res
        del ddpm
    if "cfg5" in which:
        torch.manual_seed(0)
        b = 8
        ddpm = D.DDPM(D.ContextUnet(3, 384, 5), (1e-4, 0.02), 700, dev, 0.1, enhance_with_attn_map=True).to(dev).train()
        opt = D.FusedAdamW(ddpm.parameters(), lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0)
        x = torch.rand(b, 3, 256, 256, device=dev) * 2 - 1
        c = torch.randint(0, 5, (b,), device=dev)
        m = torch.full((b, 256, 256), 0.5, device=dev); m[:, 128:] = 1.0; m[:, 40:90, 60:200] = 3.0

        def micro():
            ddpm(x, c, m).backward()
        ms = timed(micro, 3, 2)
        opt.step(); opt.zero_grad(); torch.cuda.synchronize()
        out["cfg5_f384_b8"] = {"micro_step_ms": ms, "img_per_s_fwd_bwd": b / ms * 1e3,
                               "tflops": 3 * 5382.0 * b / ms, "max_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
        print(out["cfg5_f384_b8"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "extra_configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
