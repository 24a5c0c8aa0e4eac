"""Do the streaming BatchNorm-backward kernels run NEXT TO the dW GEMM of the other stream, or after it?
t(GEMM alone), t(BN alone), t(both, two streams): both ~ max = they share the SMs, both ~ sum = they serialise.
Also against a plain torch copy kernel as the streaming partner (control)."""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from hnd_ghnd_object_detectors_b200 import ops

dev = torch.device("cuda", 0)
N, H, W, C, K = 4, 201, 337, 256, 256
Ho, Wo = H - 1, W - 1
x = torch.randn(N, H, W, C, device=dev).to(torch.bfloat16)
dy = torch.randn(N, Ho, Wo, K, device=dev).to(torch.bfloat16)
dw = torch.zeros(K, 2, 2, C, device=dev)
wg = ops.WgradPlan(N, H, W, C, K, 2, 2, 0, x, dy, dw)
# BN backward on a 256-channel tensor of the same size
raw = torch.randn(N, Ho, Wo, K, device=dev).to(torch.float16)
g = torch.randn(N, Ho, Wo, K, device=dev).to(torch.bfloat16)
gx = torch.empty_like(g)
ss = torch.ones(2 * K, device=dev)
mi = torch.ones(2 * K, device=dev)
sums = torch.zeros(2 * K, dtype=torch.float64, device=dev)
gamma = torch.ones(K, device=dev)
dgamma, dbeta = torch.zeros(K, device=dev), torch.zeros(K, device=dev)
big = torch.empty(N * Ho * Wo * K, dtype=torch.bfloat16, device=dev)

def reduce_():
    ops.bn_bwd_reduce(g, raw, ss, mi, True, sums)

def apply_():
    ops.bn_bwd_apply(g, raw, gx, gamma, ss, mi, True, sums, dgamma, dbeta)

def copy_():
    big.copy_(g.view(-1))

s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timed(fa, fb):
    ts = []
    for _ in range(7):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        s1.wait_event(e0)
        s2.wait_event(e0)
        if fa is not None:
            with torch.cuda.stream(s1):
                fa()
                e1.record()
        if fb is not None:
            with torch.cuda.stream(s2):
                fb()
                e2.record()
        torch.cuda.synchronize()
        ta = e0.elapsed_time(e1) * 1e3 if fa is not None else 0.0
        tb = e0.elapsed_time(e2) * 1e3 if fb is not None else 0.0
        ts.append((max(ta, tb), ta, tb))
    ts.sort()
    return ts[3]

for f in (wg.run, reduce_, apply_, copy_):
    f()
torch.cuda.synchronize()
print("alone: dW GEMM %.1f us, bn_bwd_reduce %.1f, bn_bwd_apply %.1f, torch copy %.1f" % (
    timed(wg.run, None)[0], timed(None, reduce_)[0], timed(None, apply_)[0], timed(None, copy_)[0]))
for name, f in (("bn_bwd_reduce", reduce_), ("bn_bwd_apply", apply_), ("torch copy", copy_)):
    t, ta, tb = timed(wg.run, f)
    print("dW GEMM (first) + %-14s: both done after %.1f us (GEMM %.1f, partner %.1f)" % (name, t, ta, tb))
    t, tb, ta = timed(f, wg.run)
    print("%-14s (first) + dW GEMM: both done after %.1f us (GEMM %.1f, partner %.1f)" % (name, t, ta, tb))
