"""Stride-2 3x3 dgrad (four parity-class launches) inside a CUDA graph: time per plan run.  Run with and
without GHND_S2_SERIAL=1 to see whether the launches overlap (late griddepcontrol.wait)."""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from hnd_ghnd_object_detectors_b200 import ops, _lib

dt = torch.bfloat16
for (N, H, W, C, K) in [(4, 200, 336, 128, 128), (4, 100, 168, 256, 256), (4, 50, 84, 512, 512)]:
    Ho, Wo = H // 2, W // 2
    dy = torch.randn(N, Ho, Wo, K, device="cuda").to(dt)
    wt = (torch.randn(C, 3, 3, K, device="cuda") * 0.05).to(dt)
    dx = torch.zeros(N, H, W, C, dtype=dt, device="cuda")
    mask = torch.randn(N, H, W, C, device="cuda").to(torch.float16)
    plan = ops.ConvPlan(_lib.CONV_DGRAD, N, H, W, C, K, 3, 3, 2, 1, dy, wt, dx, mask=mask)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        plan.run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(10):
                plan.run()
        g.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            g.replay()
        b.record()
        torch.cuda.synchronize()
    print("dgrad N%d %dx%d C%d K%d 3x3 s2: %.1f us per plan run (%d launches)" % (N, H, W, C, K, a.elapsed_time(b) * 1e3 / 50, plan.n_launches))
