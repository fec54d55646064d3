"""Per-layer kernel timings on the B200 (dev tool; results go to gpurun_out/microbench.json).

Times every distinct convolution shape of the cfg2 ContextUnet (F=192, 256x256, batch 4) through the
C-ABI -- forward, data gradient, weight gradient -- plus the bandwidth kernels at the dominant shapes,
with CUDA events on the launching stream, L2 flushed between iterations.

    python tools/microbench.py [--only conv|ew] [--iters 5] [--batch 4]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from diffusionmodel_b200 import _lib, ops  # noqa: E402

# (H, Cin, Cout, k, stride, pad, count per forward)
F = 192
CONVS = [
    (256, F, F, 3, 1, 1, 11), (256, 2 * F, F, 3, 1, 1, 2), (128, 2 * F, 2 * F, 3, 1, 1, 3), (128, F, F, 3, 1, 1, 4),
    (128, 4 * F, F, 3, 1, 1, 1), (64, 4 * F, 4 * F, 3, 1, 1, 3), (64, 2 * F, 2 * F, 3, 1, 1, 4),
    (64, 8 * F, 2 * F, 3, 1, 1, 1), (32, 8 * F, 8 * F, 3, 1, 1, 3), (32, 4 * F, 4 * F, 3, 1, 1, 4),
    (32, 16 * F, 4 * F, 3, 1, 1, 1), (256, 3, F, 3, 1, 1, 1), (256, F, 3, 3, 1, 1, 1),
    (256, F, F, 4, 2, 1, 1), (128, 2 * F, 2 * F, 4, 2, 1, 1), (64, 4 * F, 4 * F, 4, 2, 1, 1), (32, 8 * F, 8 * F, 4, 2, 1, 1),
    (256, F, F // 4, 1, 1, 0, 1), (256, F // 4, F, 1, 1, 0, 1), (32, 4 * F, F, 1, 1, 0, 1), (32, F, 8 * F, 1, 1, 0, 1),
]


def timer(fn, iters, flush):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--out", default="microbench.json")
    ap.add_argument("--ablate", action="store_true", help="halo-kernel timing decomposition (no MMA / no TMA / no epilogue)")
    ap.add_argument("--bn_sweep", action="store_true", help="forced tile widths, full kernel and MMA-only")
    ap.add_argument("--ab", action="store_true", help="also time the alternative kernels (generic conv, wgrad v1/v2)")
    ap.add_argument("--shapes", type=int, default=0, help="only the first N conv shapes")
    ap.add_argument("--pick", default="", help="comma-separated indices into the conv shape table")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    n = a.batch
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    g = torch.Generator(device=dev).manual_seed(0)
    if a.only in ("", "conv"):
        picked = [CONVS[int(i)] for i in a.pick.split(",")] if a.pick else (CONVS[:a.shapes] if a.shapes else CONVS)
        for (h, cin, cout, k, s, p, cnt) in picked:
            x = torch.randn(n, h, h, ops.r8(cin), device=dev, generator=g).to(torch.bfloat16)
            w = torch.nn.Parameter(torch.randn(cout, cin, k, k, device=dev, generator=g) / (cin * k * k) ** 0.5)
            b = torch.nn.Parameter(torch.zeros(cout, device=dev))
            pack = ops.WeightPack()
            ho = (h + 2 * p - k) // s + 1
            xr = x.requires_grad_(True)
            y, _ = ops.conv2d(xr, w, b, pack, stride=s, pad=p, want_stats=True)
            dy = torch.randn_like(y)
            fl = 2.0 * n * ho * ho * cout * cin * k * k
            wpk = pack.get(w, "fwd", c_split=0)
            stats = torch.empty((ops.conv_stat_rows(n, ho, ho, cout), 2, cout), device=dev)
            st = ops._stream()
            P_ = ops._p

            def fwd():
                ops.call("dm_conv2d_fwd", P_(x), cin, x.stride(2), None, 0, 0, P_(wpk), P_(b), None, 0, P_(y), y.stride(2), 0,
                         P_(stats), cout, n, h, h, cout, k, k, s, p, st)
            def fwd_nostat():
                ops.call("dm_conv2d_fwd", P_(x), cin, x.stride(2), None, 0, 0, P_(wpk), P_(b), None, 0, P_(y), y.stride(2), 0,
                         None, 0, n, h, h, cout, k, k, s, p, st)
            dx = torch.empty_like(x)
            if s == 1:
                wd = pack.get(w, "dgrad")

                def dgrad():
                    ops.call("dm_conv2d_fwd", P_(dy), cout, dy.stride(2), None, 0, 0, P_(wd), None, None, 0, P_(dx), dx.stride(2), 0,
                             None, 0, n, ho, ho, cin, k, k, 1, k - 1 - p, st)
            else:
                wd = pack.get(w, "s2dgrad")

                def dgrad():
                    ops.call("dm_conv2d_s2_dgrad", P_(dy), cout, dy.stride(2), P_(wd), P_(dx), cin, dx.stride(2), n, ho, ho, st)
            ck = ops._cols_k(cin, 0)
            dwp = torch.zeros((cout, k * k * ck), device=dev)

            def wgrad():
                ops.call("dm_conv2d_wgrad", P_(x), cin, x.stride(2), None, 0, 0, P_(dy), dy.stride(2), P_(dwp), n, h, h,
                         cout, k, k, s, p, st)
            r = {"shape": f"{h}x{h} {cin}->{cout} k{k}s{s}", "count": cnt, "gflop": fl / 1e9}
            variants = [("fwd", fwd), ("fwd_nostat", fwd_nostat), ("dgrad", dgrad), ("wgrad", wgrad)]
            if a.ab:
                variants += [("fwd_nostat_single", fwd_nostat), ("dgrad_single", dgrad),
                             ("fwd_generic", fwd), ("fwd_nostat_generic", fwd_nostat), ("dgrad_generic", dgrad),
                             ("wgrad_v1", wgrad), ("wgrad_v2", wgrad), ("wgrad_v3", wgrad), ("wgrad_v4", wgrad)]
            if a.ablate and k == 3 and s == 1:
                for mask in (0, 2, 4, 6):
                    variants.append((f"dgrad_abl{mask}", dgrad))
            if a.ablate:
                variants += [("fwd_nostat_abl4", fwd_nostat), ("dgrad_abl4", dgrad)]
            if a.bn_sweep and k == 3 and s == 1:
                for bn_ in (64, 96, 128, 192, 256):
                    if bn_ <= ops.r8(cin) * 2:
                        variants += [(f"dgrad_bn{bn_}_abl6", dgrad), (f"dgrad_bn{bn_}_abl0", dgrad)]
            for name, fn in variants:
                if name.startswith("wgrad_v"):
                    _lib.debug_set(4, int(name[-1]))
                if name.endswith("_generic"):
                    _lib.debug_set(5, 1)
                if name.endswith("_single"):
                    _lib.debug_set(5, 2)
                if "_abl" in name:
                    _lib.debug_set(7, int(name.split("_abl")[1]))
                if "_bn" in name:
                    _lib.debug_set(3, int(name.split("_bn")[1].split("_")[0]))
                ms = timer(fn, a.iters, flush)
                _lib.debug_set(4, 0)
                _lib.debug_set(5, 0)
                _lib.debug_set(7, 0)
                _lib.debug_set(3, 0)
                r[name + "_ms"] = round(ms, 4)
                r[name + "_tflops"] = round(fl / ms / 1e9, 1)
            print(json.dumps(r), flush=True)
            rows.append(r)
            del x, y, dy, dx, dwp, xr
    if a.only in ("", "ew"):
        for (h, c) in ((256, F), (128, 2 * F), (32, 8 * F)):
            P = n * h * h
            y = torch.randn(n, h, h, c, device=dev, generator=g).to(torch.bfloat16)
            dz = torch.randn(n, h, h, c, device=dev, generator=g).to(torch.bfloat16)
            z = torch.empty_like(y)
            mean = torch.zeros(c, device=dev); inv = torch.ones(c, device=dev)
            ga = torch.ones(c, device=dev); be = torch.zeros(c, device=dev)
            dga = torch.zeros(c, device=dev); dbe = torch.zeros(c, device=dev)
            scr = torch.empty(_lib.fn("dm_bn_act_bwd_scratch")(P, c), device=dev)
            st = ops._stream()
            P_ = ops._p
            bytes_el = P * c * 2

            def bn_fwd():
                ops.call("dm_bn_act_fwd", P_(y), c, P_(mean), P_(inv), P_(ga), P_(be), P_(z), c, P, c, 1, st)

            def bn_bwd():
                ops.call("dm_bn_act_bwd", P_(dz), c, P_(y), c, P_(mean), P_(inv), P_(ga), P_(be), P_(z), c, P_(dga), P_(dbe),
                         None, P_(scr), P, c, 1, 1, st)

            def colsum():
                ops.call("dm_colsum", P_(dz), c, P_(dbe), P, c, st)

            def axpby():
                ops.call("dm_axpby", P_(y), c, P_(dz), c, P_(z), c, P, c, 1.0, 1.0, st)
            rows_ = _lib.fn("dm_bn_stats_rows")(P, c)
            part = torch.empty((rows_, 2, c), device=dev)
            rm = torch.zeros(c, device=dev); rv = torch.ones(c, device=dev)
            pooled = torch.empty((n, c), device=dev)

            def bn_stats():
                ops.call("dm_bn_stats", P_(y), c, P_(part), c, P, c, st)

            def bn_finalize():
                ops.call("dm_bn_finalize", P_(part), rows_, c, c, float(P), P_(mean), P_(inv), P_(rm), P_(rv), 0.1, 1e-5, None, st)

            def pool():
                ops.call("dm_pool_nhw", P_(y), c, P_(pooled), n, h * h, c, 1.0 / (h * h), st)
            def chain_sf():                 # statistics then apply, back to back: does the second read of y hit L2?
                bn_stats(); bn_finalize(); bn_fwd()

            def chain_bwd2():               # two backward calls back to back on the same tensors (dz, y in L2 for the 2nd?)
                bn_bwd(); bn_bwd()
            for name, fn, nb in (("bn_stats", bn_stats, 1), ("bn_finalize", bn_finalize, 0), ("bn_act_fwd", bn_fwd, 2),
                                 ("chain_stats_fwd", chain_sf, 3), ("chain_bwd_x2", chain_bwd2, 10),
                                 ("bn_act_fwd_v1", bn_fwd, 2),
                                 ("bn_act_bwd", bn_bwd, 5), ("colsum", colsum, 1), ("pool_nhw", pool, 1), ("axpby", axpby, 3)):
                if "_v" in name:
                    _lib.debug_set(8, int(name.split("_v")[1]))
                ms = timer(fn, a.iters, flush)
                _lib.debug_set(8, 0)
                r = {"kernel": name, "shape": f"{h}x{h}x{c}", "ms": round(ms, 4), "GBps": round(nb * bytes_el / ms / 1e6, 1)}
                print(json.dumps(r), flush=True)
                rows.append(r)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", a.out), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
