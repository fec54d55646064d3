"""Data-parallel training on NCCL (GPU, needs >= 2 devices; skipped on a 1-GPU box): SURVEY.md 4 item 8-ii.
Spawns tests/dp_worker.py under torch.distributed.run with 2 ranks."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_nccl_training_keeps_ranks_identical_and_overlap_matches(dev):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    lines = [l for l in r.stdout.splitlines() if l.startswith("DP_RESULT ")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(lines[-1][len("DP_RESULT "):])
    print(res)
    assert res["world"] == 2
    for mode in ("plain", "overlapped"):
        assert res[f"{mode}_params_identical_after_broadcast"] and res[f"{mode}_allreduced_grads_identical"], res
        assert res[f"{mode}_identical_every_step"], res          # needs the deterministic grad-norm reduction (dm_sumsq)
    assert res["allreduced_gradient_is_the_mean"]
    assert 0.5 < res["overlap_boundary_fraction"] < 0.8           # the decoder side is the larger share of the bytes
    # same data, same weights: the two schedules differ only in the split-K / tile schedule of the SM-capped trunk GEMMs
    assert res["first_step_grad_rel_l2"] < 1e-3 and res["final_param_rel_l2"] < 1e-3, res
