"""Quantizer / dequantizer bandwidth against the tensor size (cold L2): separates the fixed cost of a
launch (prologue, ramp, tail) from the streaming rate.  batch 64 of the split tensor = 13.3 M elements."""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from hnd_ghnd_object_detectors_b200 import ops

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=9):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


empty = timed(lambda: None)
print("empty event pair: %.1f us" % empty)
for n in [64 * 3 * 204 * 340, 4 * 64 * 3 * 204 * 340, 16 * 64 * 3 * 204 * 340]:
    x = torch.randn(n, device=dev)
    q = torch.empty(n, dtype=torch.uint8, device=dev)
    qp = torch.empty(4, dtype=torch.int32, device=dev)
    ws = ops.quantize_ws(n, dev)
    mm = torch.stack([x.min(), x.max()]).contiguous()
    out = torch.empty(n, device=dev)
    t_f = timed(lambda: ops.quantize_u8(x, 8, q=q, qparams=qp, ws=ws))
    q_ref = torch.clamp(torch.round(qp[1].float() + x / qp[0:1].view(torch.float32)), 0, 255).to(torch.uint8)
    ok = bool(torch.equal(q, q_ref))
    t_a = timed(lambda: ops.quantize_u8_minmax(x, mm, 1, 8, q=q, qparams=qp))
    t_d = timed(lambda: ops.dequantize_u8(q, qp, out=out))
    t_c = timed(lambda: out.copy_(x))
    print("n=%.1fM  bytes==torch %s  fused %.1f us (%.0f GB/s)  apply %.1f us (%.0f GB/s)  dequant %.1f us (%.0f GB/s)  copy f32 %.1f us (%.0f GB/s)"
          % (n / 1e6, ok, t_f, 5 * n / t_f / 1e3, t_a, 5 * n / t_a / 1e3, t_d, 5 * n / t_d / 1e3, t_c, 8 * n / t_c / 1e3))
    sys.stdout.flush()
