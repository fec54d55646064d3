"""Worker of tests/test_gpu_dp.py: one rank of a 2+-GPU NCCL data-parallel run (launched with torch.distributed.run).

Checks, printing one JSON line from rank 0:
  * after the initial broadcast and after every optimizer step all ranks hold BIT-IDENTICAL flat parameters;
  * the all-reduced flat gradient equals the mean of the per-rank gradients;
  * the overlapped schedule (last micro-step as two graphs, decoder-side gradients all-reduced under the trunk's backward,
    GEMM grids of the trunk's backward capped to 148 - 16 SMs) produces the same averaged gradient as the plain one.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffusionmodel_b200 as D                      # noqa: E402
from diffusionmodel_b200 import parallel             # noqa: E402
from oracle.synth import fill_state_dict_, make_inputs   # noqa: E402


def main():
    rank, local, world = parallel.init_from_env()
    dev = torch.device("cuda", local)
    n_feat, size, n_classes, n_T, accum, steps = 64, 128, 5, 700, 2, 3
    out = {"world": world}

    def build():
        torch.manual_seed(0)
        net = D.ContextUnet(3, n_feat, n_classes)
        ddpm = D.DDPM(net, (1e-4, 0.02), n_T, "cpu", 0.1, enhance_with_attn_map=True)
        sd = {k: v.clone() for k, v in ddpm.state_dict().items()}
        fill_state_dict_(sd, 7 + rank)               # DIFFERENT weights per rank: the broadcast must fix that
        ddpm.load_state_dict(sd)
        ddpm.device = dev
        ddpm.to(dev).eval()                          # running-statistics BatchNorm: deterministic forward
        opt = D.FusedAdamW(ddpm.parameters(), lr=1e-3, weight_decay=1e-2, max_grad_norm=1.0)
        parallel.broadcast_parameters(opt.flat_param, list(ddpm.buffers()))
        return net, ddpm, opt

    def same_on_all_ranks(t):
        ref = t.clone()
        dist.broadcast(ref, 0)
        ok = torch.tensor([int(torch.equal(ref, t))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        return bool(ok.item())

    batches = []
    for i in range(accum):
        inp = make_inputs("rdd", 2, 3, size, n_classes, n_T, 50 + 10 * rank + i)       # per-rank data
        batches.append(tuple(inp[k].to(dev) for k in ("x", "c", "attn_mask", "ts", "noise", "ctx_mask")))

    results = {}
    for mode in ("plain", "overlapped"):
        net, ddpm, opt = build()
        out[f"{mode}_params_identical_after_broadcast"] = same_on_all_ranks(opt.flat_param)
        x, c, m, ts, nz, cx = batches[0]
        micro = ddpm.capture_train_step(x, c, m, loss_scale=1.0 / accum)
        last = red = None
        if mode == "overlapped":
            last = ddpm.capture_train_step(x, c, m, loss_scale=1.0 / accum, split_backward=True,
                                           trunk_sm_limit=148 - parallel.NCCL_CTAS)
            red = parallel.OverlappedGradReduce(opt, net.grad_ready_regions())
        opt.zero_grad()
        identical, grads_identical, mean_ok, first_grad = True, True, True, None
        for s in range(steps):
            for i, (x, c, m, ts, nz, cx) in enumerate(batches):
                if last is not None and i == accum - 1:
                    if s == 0:                          # reference for the mean check: this rank's complete gradient
                        pass
                    last(x, c, m, randoms=(ts, nz, cx), between=red.reduce_ready)
                else:
                    micro(x, c, m, randoms=(ts, nz, cx))
            if mode == "plain":
                opt.flush()
                mine = opt.flat_grad.clone()
                parallel.allreduce_mean_(opt.flat_grad)
                gathered = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(gathered, mine)
                want = torch.stack(gathered).double().mean(0).float()
                mean_ok &= bool(torch.allclose(opt.flat_grad, want, rtol=1e-6, atol=1e-9))
            else:
                red.finish()
            if s == 0:
                first_grad = opt.flat_grad.clone()
            grads_identical &= same_on_all_ranks(opt.flat_grad)
            opt.step()
            opt.zero_grad()
            identical &= same_on_all_ranks(opt.flat_param)
        torch.cuda.synchronize()
        out[f"{mode}_identical_every_step"] = identical
        out[f"{mode}_allreduced_grads_identical"] = grads_identical
        if mode == "plain":
            out["allreduced_gradient_is_the_mean"] = mean_ok
        results[mode] = (first_grad, opt.flat_param.clone(), red.boundary if red is not None else None)
        del micro, last, red, opt, ddpm, net
    g0, p0, _ = results["plain"]
    g1, p1, b = results["overlapped"]
    rel = lambda a, r: float((a.double() - r.double()).norm() / r.double().norm().clamp_min(1e-30))
    out.update(overlap_boundary_fraction=1.0 - b / g0.numel(), first_step_grad_rel_l2=rel(g1, g0),
               first_step_grad_rel_l2_trunk=rel(g1[:b], g0[:b]), first_step_grad_rel_l2_tail=rel(g1[b:], g0[b:]),
               final_param_rel_l2=rel(p1, p0), grad_norm=float(g0.norm()))
    if rank == 0:
        print("DP_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
