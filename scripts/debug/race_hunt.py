"""Debug: which teacher tensor is first corrupted when the step runs as a CUDA graph with the side stream?"""
import os, sys
os.environ.setdefault("GHND_STEM_POOL", "0")  # this script looks at conv1's output, which only the un-fused stem stores
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import weights
from tests.test_gpu_distill import criterion_config, targets_for
import tests.test_gpu_fullsize as F
from hnd_ghnd_object_detectors_b200 import models, module_util
from hnd_ghnd_object_detectors_b200.tool import DistillationBox

t_sd, s_sd = weights.teacher_student(3, seed=0)
env = {"models": models, "module_util": module_util, "t_sd": t_sd, "s_sd": s_sd}
g = torch.Generator().manual_seed(5)
images = [torch.rand(3, 800, 1333, generator=g).cuda()]


def tensors(plan):
    out = {"stem.conv": plan.stem2.conv, "t_stem.out": plan.t_stem.out, "s_stem.out": plan.s_stem.out}
    for b, blk in enumerate(plan.t_layers["layer1"].blocks):
        for k in ("a1", "a2", "idn", "out"):
            t = getattr(blk, k)
            if t is not None:
                out["t.l1.%d.%s" % (b, k)] = t
    l1 = plan.s_l1
    for k, u in (("e0", l1.e0), ("e1", l1.e1), ("e2", l1.e2), ("d4", l1.d4), ("d7", l1.d7), ("d9", l1.d9)):
        out["s.%s.raw" % k] = u.raw
        out["s.%s.w" % k] = u.w
    out["s.z"] = l1.z
    out["stem.w"] = plan.stem2.w
    return {k: v.clone() for k, v in out.items()}


ref = None
for graph in (False, True):
    teacher, student = F.build_full_pair(env)
    box = DistillationBox(teacher, student, criterion_config(), use_cuda_graph=graph)
    for it in range(6):
        box(images, targets_for(images))
        torch.cuda.synchronize()
        plan = list(box._plans.values())[0]
        cur = tensors(plan)
        if ref is None:
            ref = cur
            continue
        bad = []
        for k in cur:
            nd = int((cur[k] != ref[k]).sum())
            if nd:
                d = (cur[k].float() - ref[k].float()).abs()
                nz = (d > 0).nonzero()
                bad.append("%s: %d differ, max %.3g, first %s" % (k, nd, float(d.max()), nz[0].tolist()))
        print("graph=%s it=%d: %s" % (graph, it, "all identical" if not bad else ""), flush=True)
        for b in bad:
            print("     ", b)
