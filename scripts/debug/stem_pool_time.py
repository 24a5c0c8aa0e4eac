"""Times conv1+pool as one kernel against the conv GEMM + pool kernels (cold L2), teacher+student pair
at the distillation batch and the single stem at the encode batch."""
import sys
import torch
import os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from hnd_ghnd_object_detectors_b200 import ops

dt = torch.float16


def timed(fn, reps=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for N, Hp, Wp, m in [(2, 800, 1344, 2), (4, 800, 1344, 2), (16, 800, 1344, 1), (2, 800, 1344, 1)]:
    packed = torch.randn((N, Hp + 6, Wp + 8, 4), device="cuda").to(dt)
    packed[..., 3] = 0
    w = (torch.randn(64 * m, 7, 32, device="cuda") * 0.1).to(dt)
    b = torch.randn(64 * m, device="cuda") * 0.2
    Hc, Wc = Hp // 2, Wp // 2
    Ho, Wo = (Hc + 1) // 2, (Wc + 1) // 2
    conv = torch.empty((N, Hc, Wc, 64 * m), dtype=dt, device="cuda")
    ys = [torch.empty((N, Ho, Wo, 64), dtype=dt, device="cuda") for _ in range(m)]
    ys2 = [torch.empty_like(y) for y in ys]
    ams = [None] * (m - 1) + [torch.empty((N, Ho, Wo, 64), dtype=torch.uint8, device="cuda")]
    ams2 = [None] * (m - 1) + [torch.empty_like(ams[-1])]
    sp = ops.StemPlan(packed, w, b, conv, N, Hp, Wp)
    fp = ops.StemPoolPlan(packed, w, b, ys2, ams2, N, Hp, Wp)

    def two():
        sp.run()
        for k in range(m):
            ops.maxpool3x3s2(conv, ys[k], ams[k], channels=64, channel_offset=64 * k)

    t2 = timed(two)
    t1 = timed(fp.run)
    same = all(torch.equal(a, c) for a, c in zip(ys, ys2)) and torch.equal(ams[-1], ams2[-1])
    print("N=%d %dx%d models=%d: conv+pool %.1f us  fused %.1f us  (%.2fx)  bitwise %s" % (N, Hp, Wp, m, t2, t1, t2 / t1, same))
    sys.stdout.flush()
