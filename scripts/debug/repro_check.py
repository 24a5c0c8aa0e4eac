"""Debug: run-to-run reproducibility of the forward features at 800x1333 (the teacher path has no
atomics: every run must be bit-identical) and distance to the fp32 oracle per level."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ghnd_oracle as O, weights
from tests.test_gpu_distill import criterion_config, targets_for
import tests.test_gpu_fullsize as F
from hnd_ghnd_object_detectors_b200 import models, module_util, ops
from hnd_ghnd_object_detectors_b200.tool import DistillationBox

t_sd, s_sd = weights.teacher_student(3, seed=0)
env = {"models": models, "module_util": module_util, "t_sd": t_sd, "s_sd": s_sd}
g = torch.Generator().manual_seed(5)
host = [torch.rand(3, 800, 1333, generator=g)]
images = [im.cuda() for im in host]
torch.set_num_threads(os.cpu_count())
x = O.transform_batch(host)
with torch.no_grad():
    ref = O.backbone_features(x, t_sd, student=False)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


first = None
for graph in (False, True, True):
    teacher, student = F.build_full_pair(env)
    box = DistillationBox(teacher, student, criterion_config(), use_cuda_graph=graph)
    for it in range(4):
        box(images, targets_for(images))
        torch.cuda.synchronize()
        plan = list(box._plans.values())[0]
        cur = {lv: plan.feat_t[lv].clone() for lv in plan.levels}
        if first is None:
            first = cur
        line = []
        for lv in plan.levels:
            same = torch.equal(cur[lv], first[lv])
            nd = int((cur[lv] != first[lv]).sum())
            line.append("%s %s(%d) rel %.2e" % (lv, "same" if same else "DIFF", nd, rel(ops.to_nchw_f32(cur[lv]), ref[lv])))
        print("graph=%s it=%d  " % (graph, it) + " | ".join(line), flush=True)
        if not all(torch.equal(cur[lv], first[lv]) for lv in plan.levels):
            lv = "layer1"
            d = (cur[lv].float() - first[lv].float()).abs()
            idx = d.reshape(-1).argmax().item()
            bad = (d > 0).nonzero()
            print("   layer1 first diffs at", bad[:5].tolist(), "max", float(d.max()), "rows", sorted(set(bad[:, 1].tolist()))[:10], "cols", sorted(set(bad[:, 2].tolist()))[:10], "chans", sorted(set(bad[:, 3].tolist()))[:16])
    del box
