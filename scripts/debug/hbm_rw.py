"""HBM bandwidth by direction on this box: write-only (memset), read-only (sum), copy (read+write)."""
import torch
dev = torch.device("cuda")
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device=dev)
b = torch.empty(n, dtype=torch.uint8, device=dev)
f = torch.empty(n // 4, dtype=torch.float32, device=dev).fill_(1.0)


def t(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best

print("write-only  (memset 1 GiB):      %.0f GB/s" % (n / t(lambda: a.zero_()) / 1e9))
print("read-only   (sum of 1 GiB fp32): %.0f GB/s" % (n / t(lambda: f.sum()) / 1e9))
print("copy        (1 GiB -> 1 GiB):    %.0f GB/s (read + write bytes)" % (2 * n / t(lambda: b.copy_(a)) / 1e9))
h = a.view(torch.float16)
print("write-only  (fill fp16):         %.0f GB/s" % (n / t(lambda: h.fill_(1.0)) / 1e9))
