"""conv1 dW (tcgen05) at the distillation batch: time per plan run (cold L2) and agreement with fp32 torch."""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from hnd_ghnd_object_detectors_b200 import ops

N, Hp, Wp = 4, 800, 1344
dt = torch.bfloat16
torch.manual_seed(0)
packed = torch.zeros(N, Hp + 6, Wp + 8, 4, device="cuda", dtype=dt)
packed[:, 3:3 + Hp, 3:3 + Wp, :3] = torch.randn(N, Hp, Wp, 3, device="cuda").to(dt)
g = (torch.randn(N, Hp // 2, Wp // 2, 64, device="cuda") * 0.1).to(dt)
scale = torch.rand(64, device="cuda") + 0.5
dw = torch.zeros(64, 3, 7, 7, device="cuda")
plan = ops.StemWgradPlan(packed, g, scale, dw, N, Hp, Wp)
plan.run()
torch.cuda.synchronize()
x = packed[:, 3:3 + Hp, 3:3 + Wp, :3].permute(0, 3, 1, 2).float()
ref = torch.nn.grad.conv2d_weight(x, (64, 3, 7, 7), g.permute(0, 3, 1, 2).float(), stride=2, padding=3) * scale[:, None, None, None]
print("rel err %.2e" % float((dw - ref).norm() / ref.norm()))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(9):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    plan.run()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
print("stem wgrad N%d %dx%d: %.1f us per run (memset + kernel + finish, event-timed)" % (N, Hp, Wp, sorted(ts)[4]))
