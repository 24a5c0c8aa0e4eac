"""Does a consumer that sweeps a tensor in the OPPOSITE direction of its producer find the producer's tail in the
126 MB L2?  Producer: b[i] = a[i] chunk by chunk, ascending (reads a, writes b: 2 x size of traffic).  Consumer: c[i] =
b[i] ascending or descending.  Event-timed consumer sweep, chunked launches of torch's copy kernel replayed from CUDA graphs."""
import sys
import torch

dev = torch.device("cuda", 0)
CH = 16
for mb in (34, 69, 137, 275):
    n = mb * (1 << 20) // 2 // CH * CH
    a = torch.randn(n, device=dev).to(torch.bfloat16)
    b, c = torch.empty_like(a), torch.empty_like(a)
    av, bv, cv = a.view(CH, -1), b.view(CH, -1), c.view(CH, -1)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    res = {}
    graphs = {}
    for name, src, dst, order in (("prod", av, bv, range(CH)), ("asc", bv, cv, range(CH)),
                                  ("desc", bv, cv, range(CH - 1, -1, -1))):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in order:
                dst[i].copy_(src[i])
        graphs[name] = g
    for mode in ("asc", "desc", "asc", "desc"):
        ts = []
        for _ in range(7):
            flush.zero_()
            graphs["prod"].replay()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graphs[mode].replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res.setdefault(mode, []).append(sorted(ts)[3])
    print("%4d MB tensor: consumer ascending %s us, descending %s us  (%.2f / %.2f TB/s of read+write)" % (
        mb, ["%.1f" % t for t in res["asc"]], ["%.1f" % t for t in res["desc"]],
        2 * n * 2 / min(res["asc"]) / 1e6, 2 * n * 2 / min(res["desc"]) / 1e6))
    sys.stdout.flush()
