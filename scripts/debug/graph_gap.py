"""Debug: per-node cost of a chain of tiny dependent kernels inside a CUDA graph (launch gap estimate)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hnd_ghnd_object_detectors_b200 import ops
dev = torch.device("cuda")
C = 64
g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev); rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev)
ss = torch.empty(2 * C, device=dev)
x = torch.randn(8 * 1024, device=dev).half(); y = torch.empty_like(x, dtype=torch.bfloat16)
for name, fn in (("bn_eval_params (1 block)", lambda: ops.bn_eval_params(C, g, b, rm, rv, 1e-5, ss)),
                 ("convert16 8K elems", lambda: ops.convert16(x, y))):
    K = 400
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(K):
            fn()
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    print("%-28s graph: %.2f us per node" % (name, e0.elapsed_time(e1) * 1e3 / (10 * K)))
    e0.record()
    for _ in range(K):
        fn()
    e1.record(); torch.cuda.synchronize()
    print("%-28s eager: %.2f us per launch" % (name, e0.elapsed_time(e1) * 1e3 / K))
