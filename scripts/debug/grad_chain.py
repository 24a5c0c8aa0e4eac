"""Debug: stage-by-stage comparison of the engine's stored GRADIENT tensors with the teacher-forced
emulation (oracle16 with _GRADS capture) -- where does the residual ~0.5-1 % come from?"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ghnd_oracle16 as E, weights
from tests.test_gpu_distill import build_pair, criterion_config, targets_for, engine_forward_tensors, nchw_cpu
from tests.golden.make_golden import small_images
from hnd_ghnd_object_detectors_b200 import models, module_util
from hnd_ghnd_object_detectors_b200.tool import DistillationBox


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


t_sd, s_sd = weights.teacher_student(3, seed=0)
env = {"models": models, "module_util": module_util, "t_sd": t_sd, "s_sd": s_sd}
teacher, student = build_pair(env)
box = DistillationBox(teacher, student, criterion_config(), use_cuda_graph=False)
host = small_images()
images = [im.cuda() for im in host]
loss = box(images, targets_for(images))
torch.cuda.synchronize()
plan = list(box._plans.values())[0]
force, t_feats = engine_forward_tensors(plan)
E._GRADS = {}
res = E.distill_step16(t_sd, s_sd, host, force=force, teacher_feats=t_feats)
G = E._GRADS
E._GRADS = None
p = "backbone.body."
print("loss grads (engine loss_grads vs emulation gradient at the level outputs)")
for name in ("layer4", "layer3", "layer2"):
    r = plan.s_layers[name]
    for b in reversed(range(len(r.blocks))):
        blk = r.blocks[b]
        pre = "%s%s.%d" % (p, name, b)
        for key, t in ((".out", blk.g_out), (".a2", blk.g_a2), (".a1", blk.g_a1)):
            if pre + key in G:
                e = nchw_cpu(t)
                print("%-34s rel %.3e   |g| %.3e" % (pre + key, rel(e, G[pre + key]), float(G[pre + key].norm())))
l1 = plan.s_l1
e_, d_ = p + "layer1.encoder.encoder.", p + "layer1.decoder."
for key, u in ((d_ + "9", l1.d9), (d_ + "7", l1.d7), (d_ + "4", l1.d4), (e_ + "5", l1.e2), (e_ + "2", l1.e1), (e_ + "0", l1.e0)):
    if key + ".raw" in G:
        print("%-34s rel %.3e  (g_raw)" % (key + ".raw", rel(nchw_cpu(u.g_raw), G[key + ".raw"])))
    if key + ".out" in G:
        print("%-34s rel %.3e  (g_out)" % (key + ".out", rel(nchw_cpu(u.g_out), G[key + ".out"])))
if d_ + "2.raw" in G:
    print("%-34s rel %.3e  (g_raw3)" % (d_ + "2.raw", rel(nchw_cpu(l1.g_raw3), G[d_ + "2.raw"])))
