"""Student layer1 dW kernels (tcgen05 wgrad) in a CUDA graph, cold-ish (six different shapes back to back)."""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from hnd_ghnd_object_detectors_b200 import ops

dt = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
shapes = [(4, 201, 337, 256, 256, 0), (4, 202, 338, 128, 256, 0), (4, 203, 339, 64, 128, 0), (4, 202, 338, 256, 64, 1),
          (4, 201, 337, 64, 256, 1), (4, 200, 336, 64, 64, 1)]
for (N, H, W, C, K, pad) in shapes:
    Ho, Wo = H + 2 * pad - 1, W + 2 * pad - 1
    x = torch.randn(N, H, W, C, device="cuda").to(dt)
    dy = torch.randn(N, Ho, Wo, K, device="cuda").to(dt)
    dw = torch.zeros(K, 2, 2, C, device="cuda")
    plan = ops.WgradPlan(N, H, W, C, K, 2, 2, pad, x, dy, dw)
    plan.run()
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv2d_weight(x.float().permute(0, 3, 1, 2), (K, C, 2, 2), dy.float().permute(0, 3, 1, 2), padding=pad)
    err = float((dw.permute(0, 3, 1, 2) - ref).norm() / ref.norm())
    ts = []
    for _ in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        plan.run()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    print("wgrad N%d %dx%d C%d K%d p%d: %.1f us  rel err %.2e" % (N, H, W, C, K, pad, sorted(ts)[3], err))
