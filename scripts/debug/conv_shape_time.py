"""Time single conv launches of the frozen trunk at the step's shapes (CUDA graph of 8 cold-L2 calls, like
scripts/bench_kernels.py).  GHND_EPI_DEBUG=16 skips the weight loads of the generic producer (wrong results): the
upper bound of what keeping the weight tile resident in shared memory could gain."""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from hnd_ghnd_object_detectors_b200 import ops, _lib

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=8):
    fn()
    torch.cuda.synchronize()
    out = []
    for with_fn in (False, True):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                flush.zero_()
                if with_fn:
                    fn()
        g.replay()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out.append(best)
    return (out[1] - out[0]) / reps * 1e3


def rnd(shape, dt):
    return (torch.randn(shape, device=dev) * 0.1).to(dt)


f16, b16 = torch.float16, torch.bfloat16
cases = [
    ("fwd", 8, 50, 84, 256, 1024, f16, True, False),   # layer3 expansion + residual
    ("fwd", 8, 100, 168, 128, 512, f16, True, False),  # layer2 expansion + residual
    ("fwd", 8, 50, 84, 1024, 256, f16, False, False),  # layer3 reduction
    ("dgrad", 4, 50, 84, 1024, 256, b16, True, True),  # dgrad of the layer3 reduction (+res, mask)
    ("dgrad", 4, 50, 84, 256, 1024, b16, False, True), # dgrad of the layer3 expansion
]
for kind, N, H, W, C, K, dt, res, mask in cases:
    if kind == "fwd":
        x, w, y = rnd((N, H, W, C), dt), rnd((K, 1, 1, C), dt), torch.empty((N, H, W, K), dtype=dt, device=dev)
        r = rnd((N, H, W, K), dt) if res else None
        plan = ops.ConvPlan(_lib.CONV_FWD, N, H, W, C, K, 1, 1, 1, 0, x, w, y, residual=r, relu=True)
    else:
        dy, wt = rnd((N, H, W, K), dt), rnd((C, 1, 1, K), dt)
        dx = torch.empty((N, H, W, C), dtype=dt, device=dev)
        r = rnd((N, H, W, C), dt) if res else None
        m = rnd((N, H, W, C), f16) if mask else None
        plan = ops.ConvPlan(_lib.CONV_DGRAD, N, H, W, C, K, 1, 1, 1, 0, dy, wt, dx, residual=r, mask=m)
    print("%-52s %6.1f us  (%d launches)" % (plan.desc, timed(plan.run), plan.n_launches))
