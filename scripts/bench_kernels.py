"""Micro-benchmark of the HBM-bound kernels at the shapes of the batch-4 GHND step (CUDA events on
the current stream, L2 flushed between iterations by cycling over several copies of the operands).
    python scripts/bench_kernels.py [name-substring ...]
Prints one line per kernel: us, algorithmic MB, GB/s, fraction of the measured HBM peak."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hnd_ghnd_object_detectors_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
PEAK = 6554.6
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
N = int(os.environ.get("GHND_BENCH_N", "4"))
sel = sys.argv[1:]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=8):
    """Device time of one call: a CUDA graph of `reps` x (L2 flush, fn) minus a graph of the flushes
    alone -- no host gaps between launches, cold L2 for every call."""
    if os.environ.get("GHND_BENCH_EAGER"):  # under ncu: plain launches, no graphs
        for _ in range(2):
            flush.zero_()
            fn()
        torch.cuda.synchronize()
        return float("nan")
    fn()
    torch.cuda.synchronize()
    graphs = []
    for with_fn in (False, True):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                flush.zero_()
                if with_fn:
                    fn()
        graphs.append(g)
    out = []
    for g in graphs:
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


def report(name, us, nbytes):
    gbs = nbytes / us / 1e3
    print("%-46s %8.1f us %9.1f MB %8.1f GB/s  %5.1f %%" % (name, us, nbytes / 1e6, gbs, 100 * gbs / PEAK), flush=True)


def want(name):
    return not sel or any(s in name for s in sel)


def rnd(shape, dtype):
    return (torch.randn(shape, device=dev) * 0.5).to(dtype)


f16, bf16 = torch.float16, torch.bfloat16

if want("stem") or want("conv"):
    # the two-model stem (K=128, halo mode) and two output-heavy 1x1 convs of the frozen trunk
    from hnd_ghnd_object_detectors_b200 import _lib
    Hp, Wp = 800, 1344
    packed = rnd((N, Hp + 6, Wp + 8, 4), f16)
    w2 = rnd((128, 7, 32), f16)
    bias2 = torch.randn(128, device=dev)
    y2 = torch.empty((N, Hp // 2, Wp // 2, 128), dtype=f16, device=dev)
    sp = ops.StemPlan(packed, w2, bias2, y2, N, Hp, Wp)
    us = timed(sp.run)
    print("%-46s %8.1f us %9.1f MB out %7.1f TFLOP/s" % ("stem K=128 (4 launches)", us, y2.numel() * 2 / 1e6,
                                                          sp.flops / us / 1e6), flush=True)
    for (n, h, w, c, k, res) in [(N, 200, 336, 64, 256, True), (2 * N, 100, 168, 128, 512, True),
                                 (2 * N, 50, 84, 256, 1024, True), (N, 200, 336, 64, 256, False)]:
        x = rnd((n, h, w, c), f16)
        wt = rnd((k, 1, 1, c), f16)
        b = torch.randn(k, device=dev)
        r_ = rnd((n, h, w, k), f16) if res else None
        y = torch.empty((n, h, w, k), dtype=f16, device=dev)
        cp = ops.ConvPlan(_lib.CONV_FWD, n, h, w, c, k, 1, 1, 1, 0, x, wt, y, bias=b, residual=r_, relu=True)
        us = timed(cp.run)
        nbytes = x.numel() * 2 + y.numel() * 2 * (2 if res else 1)
        print("%-46s %8.1f us %9.1f MB %8.1f GB/s %7.1f TFLOP/s" % (cp.desc[:46], us, nbytes / 1e6,
                                                                     nbytes / us / 1e3, cp.flops / us / 1e6), flush=True)

if want("maxpool"):
    x = rnd((N, 400, 672, 64), f16).relu_()
    y = torch.empty((N, 200, 336, 64), dtype=f16, device=dev)
    am = torch.empty((N, 200, 336, 64), dtype=torch.uint8, device=dev)
    report("maxpool fwd (teacher, no argmax)", timed(lambda: ops.maxpool3x3s2(x, y=y)), x.numel() * 2 + y.numel() * 2)
    report("maxpool fwd (student, argmax)", timed(lambda: ops.maxpool3x3s2(x, y=y, argmax=am)),
           x.numel() * 2 + y.numel() * 3)
    dy = rnd(tuple(y.shape), bf16)
    dx = torch.empty(tuple(x.shape), dtype=bf16, device=dev)
    report("maxpool bwd (+relu mask)", timed(lambda: ops.maxpool3x3s2_bwd(x, am, dy, dx)),
           x.numel() * 2 + dx.numel() * 2 + y.numel() * 3)

if want("bn"):
    for (h, w, c, relu) in [(201, 337, 256, False), (202, 338, 128, True), (203, 339, 64, False),
                            (200, 336, 256, True), (202, 338, 256, True), (201, 337, 64, False)]:
        x = rnd((N, h, w, c), f16)
        dy = rnd((N, h, w, c), bf16)
        dx = torch.empty_like(dy)
        ss = torch.cat([torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.1])
        mi = torch.cat([torch.randn(c, device=dev) * 0.1, torch.rand(c, device=dev) + 0.5])
        sums = torch.zeros(2 * c, dtype=torch.float64, device=dev)
        gamma = torch.ones(c, device=dev)
        dg, db = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
        tag = "%dx%dx%d relu=%d" % (h, w, c, relu)
        report("bn_bwd_reduce " + tag, timed(lambda: ops.bn_bwd_reduce(dy, x, ss, mi, relu, sums)), x.numel() * 4)
        report("bn_bwd_apply  " + tag, timed(lambda: ops.bn_bwd_apply(dy, x, dx, gamma, ss, mi, relu, sums, dg, db)),
               x.numel() * 6)
        y = torch.empty_like(x)
        y2 = torch.empty_like(dy)
        report("bn_apply(+bf16 copy) " + tag, timed(lambda: ops.bn_apply(x, y, ss, relu, y2=y2)), x.numel() * 6)

if want("narrow"):
    bch = 3
    e2 = rnd((N, 203, 339, 64), f16)
    w7 = torch.randn(bch, 64, 2, 2, device=dev) * 0.1
    z = torch.empty((N, bch, 204, 340), device=dev)
    ws = ops._narrow_ws(bch, 64, 2, 2, dev)
    report("narrow_out enc7 (64->3)", timed(lambda: ops.conv_narrow_out(e2, w7, 1, y=z, ws=ws)),
           e2.numel() * 2 + z.numel() * 4)
    w2 = torch.randn(64, bch, 2, 2, device=dev) * 0.1
    raw3 = torch.empty((N, 203, 339, 64), dtype=f16, device=dev)
    pre = torch.cat([torch.rand(bch, device=dev) + 0.5, torch.randn(bch, device=dev) * 0.1])
    report("narrow_in dec2 (3->64, bn+relu pre)", timed(lambda: ops.conv_narrow_in(z, w2, 0, pre=pre, pre_relu=True, y=raw3, ws=ws)),
           raw3.numel() * 2 + z.numel() * 4)
    g3 = rnd((N, 203, 339, 64), bf16)
    gz = torch.empty_like(z)
    report("narrow_out_dgrad dec2", timed(lambda: ops.conv_narrow_out_dgrad(g3, w2, 0, 204, 340, dx=gz, ws=ws)),
           g3.numel() * 2 + gz.numel() * 4)
    ge2 = torch.empty((N, 203, 339, 64), dtype=bf16, device=dev)
    report("narrow_in enc7 dgrad (flip)", timed(lambda: ops.conv_narrow_in(gz, w7, 1, flip=True, y=ge2, ws=ws)),
           ge2.numel() * 2 + gz.numel() * 4)
    dw2 = torch.zeros(64, bch, 2, 2, device=dev)
    dw7 = torch.zeros(bch, 64, 2, 2, device=dev)
    report("wgrad_narrow dec2", timed(lambda: ops.wgrad_narrow(z, g3, dw2, False, 2, 2, 0, pre=pre, pre_relu=True)),
           g3.numel() * 2 + z.numel() * 4)
    e2b = e2.to(bf16)
    report("wgrad_narrow enc7", timed(lambda: ops.wgrad_narrow(gz, e2b, dw7, True, 2, 2, 1)),
           e2.numel() * 2 + gz.numel() * 4)

if want("quant"):
    for n in (4, 16, 64):
        zq = torch.randn(n, 3, 204, 340, device=dev)
        q, qp = ops.quantize_u8(zq)
        report("quantize_u8 batch %d (incl. workspace clear)" % n, timed(lambda: ops.quantize_u8(zq)), zq.numel() * 5)
        report("dequantize_u8 batch %d" % n, timed(lambda: ops.dequantize_u8(q, qp)), zq.numel() * 5)
