"""Dev probe: time per dependent kernel node in a replayed CUDA graph (tiny kernels), i.e. the launch gap the ~2100
kernels of a captured training step pay.  python tools/graph_gap_probe.py"""
import torch
x = torch.zeros(1024, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        x.add_(1)
torch.cuda.current_stream().wait_stream(s)
for n in (200, 2000):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            x.add_(1)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"{n} nodes: {e0.elapsed_time(e1) / 5 / n * 1e3:.2f} us per node")
