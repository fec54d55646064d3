"""DDPM process on the sm_100a kernels: drop-in mirror of the reference ``DDPM`` module.

Mirrors new_scripy.py:358-477 (schedules, weighted-loss training forward, CFG reverse sampling) and
MNIST_script.py:190-300 (MSE loss, ``sample`` also returning the stored trajectory).  The seven
schedule tables are registered as buffers under the reference's names, so reference checkpoints
(``{'model_state_dict': ddpm.state_dict()}``) load unchanged.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops


def ddpm_schedules(beta1, beta2, T):
    """Seven fp32 tables with T+1 entries, indexed 1..T (new_scripy.py:358-384).  Kept as the same torch
    expressions as the reference so the tables are bit-identical."""
    assert beta1 < beta2 < 1.0, "beta1 and beta2 must be in (0, 1)"
    beta_t = (beta2 - beta1) * torch.arange(0, T + 1, dtype=torch.float32) / T + beta1
    alpha_t = 1 - beta_t
    alphabar_t = torch.cumsum(torch.log(alpha_t), dim=0).exp()
    sqrtmab = torch.sqrt(1 - alphabar_t)
    return {
        "alpha_t": alpha_t,
        "oneover_sqrta": 1 / torch.sqrt(alpha_t),
        "sqrt_beta_t": torch.sqrt(beta_t),
        "alphabar_t": alphabar_t,
        "sqrtab": torch.sqrt(alphabar_t),
        "sqrtmab": sqrtmab,
        "mab_over_sqrtmab": (1 - alpha_t) / sqrtmab,
    }


class DDPM(nn.Module):
    """``DDPM(nn_model, betas, n_T, device, drop_prob=0.1)`` (new_scripy.py:386-399).

    ``forward(x, c, attn_mask)`` -> scalar loss; ``forward(x, c)`` for the MNIST denoiser.
    ``sample(n_sample, size, device, guide_w)`` -> images (MNIST variant: ``(images, stored)``).

    Extensions that default to the shipped behaviour:
      * ``enhance_with_attn_map``: feed ``attn_mask`` to LocalEnhancer as the attention map the
        reference call site meant (new_scripy.py:353).  Default False = shipped (+0) semantics.
      * ``sample_noise``: "reference" draws x_T and every step's z from the CPU generator exactly like
        new_scripy.py:445,465 (same stream for a given seed); "device" draws on the GPU.
    """

    def __init__(self, nn_model, betas, n_T, device, drop_prob=0.1, enhance_with_attn_map=False,
                 sample_noise="reference"):
        super().__init__()
        self.nn_model = nn_model.to(device)
        # reference training code drives ddpm.scaler (fp16 AMP); bf16 needs no loss scaling
        self.scaler = torch.amp.GradScaler("cuda", enabled=False)
        for k, v in ddpm_schedules(betas[0], betas[1], n_T).items():
            self.register_buffer(k, v)
        self.n_T = n_T
        self.device = device
        self.drop_prob = drop_prob
        self.loss_mse = nn.MSELoss()
        self.n_classes = self.nn_model.n_classes
        self.variant = getattr(nn_model, "variant", "rdd")
        self.enhance_with_attn_map = enhance_with_attn_map
        self.sample_noise = sample_noise
        self.shared_encoder_cfg = True      # exact in eval mode (see sample()); False = the reference's 2n-batch encoder
        self.graph_sampling = True          # one CUDA graph of the reverse step, replayed n_T times
        self._host_sched = None
        self._sample_graphs = {}

    # ------------------------------------------------------------------ training
    def draw_randoms(self, x, c):
        """RNG draw order of the reference (new_scripy.py:405-413; MNIST_script.py:239-249)."""
        ts = torch.randint(1, self.n_T + 1, (x.shape[0],)).to(self.device)
        noise = torch.randn_like(x)
        if self.variant == "rdd":
            ctx_mask = torch.bernoulli(torch.ones_like(c, dtype=torch.float) * (1 - self.drop_prob)).to(self.device)
        else:
            ctx_mask = torch.bernoulli(torch.zeros_like(c) + self.drop_prob).to(self.device)
        return ts, noise, ctx_mask

    def forward(self, x, c, attn_mask=None, randoms=None):
        ts, noise, ctx_mask = randoms if randoms is not None else self.draw_randoms(x, c)
        x = x.contiguous().float()
        noise = noise.contiguous().float()
        xt = ops.q_sample(x, noise, self.sqrtab, self.sqrtmab, ts.long())
        kw = {}
        if self.variant == "rdd":
            if attn_mask is None:
                raise RuntimeError("DDPM.forward: attn_mask is required (new_scripy.py:401)")
            attn_mask = attn_mask.to(self.device)
            if self.enhance_with_attn_map:
                kw["attn_map"] = attn_mask
        pred = self.nn_model.forward_nhwc(xt, c, ts / self.n_T, ctx_mask, **kw)
        return ops.ddpm_loss(pred, noise, attn_mask if self.variant == "rdd" else None)

    def _forward_cut(self, x, c, attn_mask, randoms):
        """``forward`` with the autograd tape cut between the trunk and the decoder side (up0, embeddings, up1.., head):
        returns (loss, stages) with stages = [(outputs, detached stand-ins), ...] in forward order.  ``loss.backward()`` then
        runs the decoder side's backward only (parameter gradients of everything from ``time_emb1`` on, in ``parameters()``
        order) and leaves the gradients of the last stage's outputs in ``stand_in.grad``;
        ``torch.autograd.backward(outputs, [stand_in.grad ...])``, last stage first, runs the rest (new_scripy: down4 +
        CoordAttn, then init_conv ... down3)."""
        ts, noise, ctx_mask = randoms
        x = x.contiguous().float()
        noise = noise.contiguous().float()
        xt = ops.q_sample(x, noise, self.sqrtab, self.sqrtmab, ts.long())
        kw = {}
        if self.variant == "rdd":
            if attn_mask is None:
                raise RuntimeError("DDPM.forward: attn_mask is required (new_scripy.py:401)")
            if self.enhance_with_attn_map:
                kw["attn_map"] = attn_mask
        net = self.nn_model

        def cut(tensors):
            names = list(tensors)
            outs = [tensors[k] for k in names]
            leaves = [o.detach().requires_grad_(True) for o in outs]
            return dict(zip(names, leaves)), (outs, leaves)
        stages = []                                      # [(outputs, detached stand-ins)], in forward order
        with ops.batched_counters(net):
            if hasattr(net, "trunk_back"):               # trunk in two pieces: a third backward stage
                enc, st0 = cut(net.trunk_front(xt))      # skip tensors: their gradients come from the decoder side,
                stages.append(st0)                       # d3_in's from the trunk_back stage
                back, st1 = cut(net.trunk_back(enc.pop("d3_in")))
                stages.append(st1)
                enc.update(back)
            else:
                enc, st0 = cut(net.trunk(xt))
                stages.append(st0)
            enc["u1"] = net.up0_of(enc.pop("hidden"))
            pred = net.decode(enc, c, ts / self.n_T, ctx_mask, **kw)
        return ops.ddpm_loss(pred, noise, attn_mask if self.variant == "rdd" else None), stages

    def capture_train_step(self, x, c, attn_mask=None, loss_scale=1.0, warmup=2, split_backward=False, trunk_sm_limit=0):
        """CUDA-graph one training micro-step: ``loss = self(x, c, attn_mask) * loss_scale; loss.backward()``.

        Returns ``step(x, c, attn_mask=None, randoms=None) -> loss`` (a static device scalar, valid until the
        next call) that copies the batch into the graph's static inputs, draws the step's randoms exactly
        like ``forward`` (new_scripy.py:405-413) and replays ~500 kernel launches with one graph launch.
        Gradients accumulate into the parameters' ``.grad`` (the optimizer's flat buffer) as in eager mode.
        The warm-up passes accumulate gradients: call ``optimizer.zero_grad()`` afterwards.  BatchNorm running
        statistics, ``num_batches_tracked`` and the CPU / CUDA RNG streams are restored to their state at entry.
        Shapes are fixed to those of the example batch; parameters must not be re-allocated afterwards.

        ``split_backward=True`` captures the step as SEVERAL graphs -- forward + the decoder side's backward, then the
        trunk's backward in one (MNIST) or two (new_scripy: down4 + CoordAttn, then init_conv ... down3) more -- and
        ``step(..., between=fn)`` calls ``fn(k)`` after graph k (except the last): when it runs, the gradients of
        ``nn_model.grad_ready_regions()[k]`` are final (k = 0: every parameter from ``time_emb1`` on, 62 % of the flat
        gradient buffer; k = 1: down4, 29 %), so a data-parallel all-reduce of that region overlaps the rest of the
        backward pass (parallel.OverlappedGradReduce).  ``trunk_sm_limit`` caps the GEMM grids of the later graphs so the
        collective's CTAs find free SMs."""
        from . import _lib
        sx, sc = x.detach().clone().float().contiguous(), c.detach().clone()
        sm = attn_mask.detach().clone().to(self.device) if attn_mask is not None else None
        # the warm-up passes must leave no trace but gradients: BatchNorm running statistics / counters and both RNG
        # streams are put back afterwards, so graphed training starts from the same state as eager training
        cpu_rng, cuda_rng = torch.get_rng_state(), torch.cuda.get_rng_state(sx.device)
        saved_bufs = [(b, b.detach().clone()) for b in self.nn_model.buffers()]
        r = self.draw_randoms(sx, sc)
        st = [t.detach().clone() for t in r]

        def first_stage():
            if not split_backward:
                loss = self.forward(sx, sc, sm, randoms=tuple(st)) * loss_scale
                loss.backward()
                return loss, []
            loss, stages = self._forward_cut(sx, sc, sm, tuple(st))
            loss = loss * loss_scale
            loss.backward()
            return loss, stages[::-1]                 # backward order

        def later_stage(stage):
            outs, leaves = stage
            live = [(o, l.grad) for o, l in zip(outs, leaves) if l.grad is not None]
            with _lib.sm_limit(trunk_sm_limit or 148):
                torch.autograd.backward([o for o, _ in live], [g for _, g in live])
        def warm():
            # in a function so that no local keeps an autograd graph alive afterwards: a live graph pins the parameters'
            # grad accumulators to the warm-up stream, and the captured backward must not wait on an uncaptured stream
            for stg in first_stage()[1]:
                later_stage(stg)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                warm()
        torch.cuda.current_stream().wait_stream(side)
        graphs, touched_by = [torch.cuda.CUDAGraph()], []
        k0 = _lib.launch_count()
        with ops.capture_log() as touched, torch.cuda.graph(graphs[0]):
            loss, stages = first_stage()
        touched_by.append(list(touched))
        for stg in stages:
            g = torch.cuda.CUDAGraph()
            with ops.capture_log() as touched, torch.cuda.graph(g, pool=graphs[0].pool()):
                later_stage(stg)
            graphs.append(g)
            touched_by.append(list(touched))
        kernels = _lib.launch_count() - k0          # library kernels recorded in the graph(s) (run on every replay)
        with torch.no_grad():
            for b, v in saved_bufs:
                b.copy_(v)
        ops.bump_bn_stats_epoch()
        torch.set_rng_state(cpu_rng)
        torch.cuda.set_rng_state(cuda_rng, sx.device)

        def mark(entries):
            for e in entries:
                if hasattr(e, "after_replay"):
                    e.after_replay()              # queued wgrad operands: staging buffers -> this pass's queue slot
                else:
                    e.dirty = True

        def step(x, c, attn_mask=None, randoms=None, between=None):
            rr = randoms if randoms is not None else self.draw_randoms(x, c)
            sx.copy_(x, non_blocking=True)
            sc.copy_(c, non_blocking=True)
            if sm is not None:
                sm.copy_(attn_mask, non_blocking=True)
            for dst, src in zip(st, rr):
                dst.copy_(src, non_blocking=True)
            ops.refresh_weight_packs()        # no-op after FusedAdamW.step; re-packs in place after any other weight change
            for k, g in enumerate(graphs):
                g.replay()
                mark(touched_by[k])
                if between is not None and k + 1 < len(graphs):
                    between(k)                # the gradients of nn_model.grad_ready_regions()[k] are final
            ops.bump_bn_stats_epoch()         # the replay updated BatchNorm running statistics in place
            return loss
        step.graph, step.graphs, step.kernels_per_replay = graphs[0], graphs, kernels
        return step

    # ------------------------------------------------------------------ sampling
    def _reverse_step_graph(self, n_sample, size, c_i, ctx_mask, shared):
        """Capture eps = U-Net(x_t) -> CFG combine + reverse step on static buffers (one graph per batch shape)."""
        dev = c_i.device
        key = (n_sample, tuple(size), bool(shared), str(dev))
        g = self._sample_graphs.get(key)
        if g is not None:
            return g
        st = dict(x=torch.zeros((n_sample, *size), device=dev), z=torch.zeros((n_sample, *size), device=dev),
                  t=torch.zeros(2 * n_sample, device=dev), coef=torch.zeros(4, device=dev), c=c_i.clone(), m=ctx_mask.clone(),
                  w=torch.zeros(n_sample, device=dev))
        st["xt"] = ops.new_act(2 * n_sample, size[1], size[2], size[0], dev).zero_()
        st["x_out"], st["xt_out"] = torch.empty_like(st["x"]), torch.empty_like(st["xt"])

        def body():
            if shared:
                enc = self._shared_encoding(st["xt"][:n_sample])
                eps = self.nn_model.decode(enc, st["c"], st["t"], st["m"])
            else:
                eps = self.nn_model.forward_nhwc(st["xt"], st["c"], st["t"], st["m"])
            ops.cfg_reverse_step_w(eps, st["x"], st["z"], st["coef"], st["w"], st["x_out"], st["xt_out"])
            st["x"].copy_(st["x_out"])
            st["xt"].copy_(st["xt_out"])
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body(); body()                                  # warm-up: weight packs, folded norms, allocator
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            body()
        st["graph"] = graph
        st["sched"] = torch.stack([self.oneover_sqrta, self.mab_over_sqrtmab, self.sqrt_beta_t], 1).to(dev).contiguous()
        self._sample_graphs[key] = st
        return st

    def _sample_graphed(self, x_i, xt, c_i, ctx_mask, n_sample, size, wvec, steps, noise, shared):
        st = self._reverse_step_graph(n_sample, size, c_i, ctx_mask, shared)
        from . import unet
        unet.refresh_folded_norms()          # weights / running statistics may have changed since the capture
        ops.refresh_weight_packs()           # ... and the bf16 weight packs the captured kernels read (load_state_dict,
                                             # a torch optimizer, manual edits): stamp compare, re-packed in place
        st["x"].copy_(x_i); st["xt"].copy_(xt); st["c"].copy_(c_i); st["m"].copy_(ctx_mask); st["w"].copy_(wvec)
        store, done, n_T = [], 0, self.n_T
        for i in range(n_T, 0, -1):
            st["t"].fill_(i / n_T)
            st["coef"][1:].copy_(st["sched"][i])
            if i > 1:
                if noise is not None:
                    st["z"].copy_(noise[1][i])
                elif self.sample_noise == "reference":
                    st["z"].copy_(torch.randn(n_sample, *size))          # CPU generator, like new_scripy.py:465
                else:
                    st["z"].normal_()
            else:
                st["z"].zero_()                                           # i == 1: z = 0, no RNG draw
            st["graph"].replay()
            if self.variant == "mnist" and (i % 20 == 0 or i == n_T or i < 8):
                store.append(st["x"].detach().cpu().numpy())
            done += 1
            if steps is not None and done >= steps:
                break
        return st["x"].clone(), store

    def _shared_encoding(self, xt_half):
        """Encoder once for the n trajectories (it sees neither c nor ctx_mask, and t is the same for both CFG halves).  The
        decoder runs on the doubled batch; the skip tensors are NOT duplicated: the upsample+cat kernel and the head's
        dual-source conv read them cyclically (sample n reads skip[n % len(skip)]).  Only u1 (the 16x16 up0 output that
        FiLM scales per sample) is repeated."""
        enc = self.nn_model.encode(xt_half)
        if self.variant == "mnist":          # 28x28 tensors through a materialised concat: repeated like the reference does
            return {k: torch.cat([v, v], 0) for k, v in enc.items()}
        enc["u1"] = torch.cat([enc["u1"], enc["u1"]], 0)
        return enc

    def _sched(self):
        if self._host_sched is None:
            self._host_sched = {k: getattr(self, k).detach().cpu().tolist()
                                for k in ("oneover_sqrta", "mab_over_sqrtmab", "sqrt_beta_t")}
        return self._host_sched

    def _randn(self, shape, device):
        if self.sample_noise == "reference":
            return torch.randn(*shape).to(device)
        return torch.randn(*shape, device=device)

    @torch.no_grad()
    def sample(self, n_sample, size, device, guide_w=0.0, refine_steps=2, steps=None, noise=None):
        """Classifier-free-guided reverse loop (new_scripy.py:441-477; MNIST_script.py:254-300).
        ``steps`` truncates the loop and ``noise=(x_T, {i: z_i})`` injects the noise (tests).

        ``guide_w`` may be a list of S guidance scales: the S x ``n_sample`` trajectories then run as ONE batch (the
        reference loops over the scales, new_scripy.py:1036-1041; at samples_per_class=1 a 5-trajectory batch leaves more
        than half of the machine idle) and a list of S tensors comes back.  Every trajectory is the computation the
        sequential call would do; with ``sample_noise="reference"`` the CPU noise stream is drawn for the whole batch per
        step, i.e. it is not the stream the S sequential calls would consume (inject ``noise`` to compare)."""
        sched = self._sched()
        n_T = self.n_T
        scales = [float(w) for w in guide_w] if isinstance(guide_w, (list, tuple)) else None
        n_each = n_sample
        ws = scales if scales is not None else [float(guide_w)]
        n_sample = n_each * len(ws)
        x_i = (noise[0].to(device) if noise is not None else self._randn((n_sample, *size), device)).float().contiguous()
        if x_i.shape[0] != n_sample:
            raise RuntimeError(f"DDPM.sample: injected x_T holds {x_i.shape[0]} trajectories, expected {n_sample}")
        ncls = 10 if self.variant == "mnist" else self.n_classes        # MNIST_script.py:262 hard-codes 10
        if n_each <= 0 or n_each % ncls:
            # the reference builds c_i = arange(ncls).repeat(n_sample // ncls) (new_scripy.py:447-448) and fails later with a
            # shape mismatch when that is shorter than the batch; fail here, before any kernel sees mismatched rows
            raise RuntimeError(f"DDPM.sample: n_sample ({n_each}) must be a positive multiple of the {ncls} classes")
        c_i = torch.arange(0, ncls, device=device).repeat(int(n_each / ncls)).repeat(len(ws)).repeat(2)
        ctx_mask = torch.zeros_like(c_i)
        ctx_mask[n_sample:] = 1.0
        wvec = torch.tensor(ws, dtype=torch.float32).repeat_interleave(n_each).to(device)
        xt = ops.to_nhwc(torch.cat([x_i, x_i], 0))
        shared = self.shared_encoder_cfg and not self.nn_model.training

        def result(x_fin, store):
            out = list(x_fin.view(len(ws), n_each, *x_fin.shape[1:]).unbind(0)) if scales is not None else x_fin
            return (out, np.array(store)) if self.variant == "mnist" else out
        if self.graph_sampling and not self.nn_model.training and x_i.is_cuda:
            return result(*self._sample_graphed(x_i, xt, c_i, ctx_mask, n_sample, size, wvec, steps, noise, shared))
        store = []
        done = 0
        coef = torch.zeros(4, device=device)
        for i in range(n_T, 0, -1):
            t_is = torch.full((2 * n_sample,), i / n_T, device=device, dtype=torch.float32)
            if i > 1:
                z = (noise[1][i].to(device) if noise is not None else self._randn((n_sample, *size), device))
                z = z.float().contiguous()
            else:
                z = torch.zeros_like(x_i)
            if shared:
                # eval mode: the encoder sees neither c nor ctx_mask and t is the same for both halves, so the
                # two halves of the reference's doubled batch are identical up to up0 -- run it once
                enc = self._shared_encoding(xt[:n_sample])
                eps = self.nn_model.decode(enc, c_i, t_is, ctx_mask)
            else:
                eps = self.nn_model.forward_nhwc(xt, c_i, t_is, ctx_mask)
            coef[1:] = torch.tensor([sched["oneover_sqrta"][i], sched["mab_over_sqrtmab"][i], sched["sqrt_beta_t"][i]])
            x_out, xt = torch.empty_like(x_i), torch.empty_like(xt)
            ops.cfg_reverse_step_w(eps, x_i, z, coef, wvec, x_out, xt)
            x_i = x_out
            if self.variant == "mnist" and (i % 20 == 0 or i == n_T or i < 8):
                store.append(x_i.detach().cpu().numpy())
            done += 1
            if steps is not None and done >= steps:
                break
        return result(x_i, store)
