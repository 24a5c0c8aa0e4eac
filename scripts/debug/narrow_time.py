"""Bottleneck-side kernels at the distillation batch (4 images) and the encode batch (64): time, achieved GB/s
and agreement with fp32 torch.  GHND_NARROW_TMA=0 selects the direct-load narrow_out kernel."""
import os
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from hnd_ghnd_object_detectors_b200 import ops

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=7):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]


def rel(a, b):
    return float((a - b).norm() / b.norm())


for N in (4, 64):
    H, W, bch = 201, 337, 3
    torch.manual_seed(N)
    x = torch.randn(N, H, W, 64, device="cuda").to(torch.float16)
    w_out = torch.randn(bch, 64, 2, 2, device="cuda") * 0.1       # enc7: 64 -> bch, k2 p1
    w_in = torch.randn(64, bch, 2, 2, device="cuda") * 0.1        # dec2: bch -> 64, k2 p0
    y = ops.conv_narrow_out(x, w_out, 1)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w_out, None, 1, 1)
    e_out = rel(y, ref)
    ws = ops._narrow_ws(64, bch, 2, 2, x.device)
    mm = torch.empty(2 * 1024, device="cuda")
    t_out = timed(lambda: ops.conv_narrow_out(x, w_out, 1, y=y, ws=ws))
    t_mm = timed(lambda: ops.conv_narrow_out(x, w_out, 1, y=y, ws=ws, minmax=mm))
    nb = x.numel() * 2 + y.numel() * 4
    z = torch.randn(N, bch, H + 1, W + 1, device="cuda")
    o = ops.conv_narrow_in(z, w_in, 0)
    e_in = rel(o.float().permute(0, 3, 1, 2), F.conv2d(z, w_in))
    t_in = timed(lambda: ops.conv_narrow_in(z, w_in, 0, y=o, ws=ws))
    nb_in = z.numel() * 4 + o.numel() * 2
    if N == 4:
        g = torch.randn(N, H, W, 64, device="cuda").to(torch.bfloat16)
        dz = ops.conv_narrow_out_dgrad(g, w_in, 0, H + 1, W + 1)
        zz = z.clone().requires_grad_(True)
        (gref,) = torch.autograd.grad(F.conv2d(zz, w_in), zz, g.float().permute(0, 3, 1, 2))
        e_dg = rel(dz, gref)
        t_dg = timed(lambda: ops.conv_narrow_out_dgrad(g, w_in, 0, H + 1, W + 1, dx=dz, ws=ws))
        dw = torch.zeros(64, bch, 2, 2, device="cuda")
        t_wg = timed(lambda: ops.wgrad_narrow(z, g, dw, False, 2, 2, 0))
        print("N=%d  narrow_out dgrad %.1f us (%.0f GB/s, rel %.1e)   wgrad_narrow %.1f us (%.0f GB/s)"
              % (N, t_dg, (g.numel() * 2 + dz.numel() * 4) / t_dg / 1e3, e_dg, t_wg, (g.numel() * 2 + z.numel() * 4) / t_wg / 1e3))
    print("N=%d  narrow_out %.1f us (%.0f GB/s, rel %.1e)  +minmax %.1f us   narrow_in %.1f us (%.0f GB/s, rel %.1e)"
          % (N, t_out, nb / t_out / 1e3, e_out, t_mm, t_in, nb_in / t_in / 1e3, e_in))
    sys.stdout.flush()
