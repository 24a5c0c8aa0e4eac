"""Host-side time of every segment of the public-API training step (no device synchronisation inside the loop):
how much Python / launch work one step costs the rank's CPU thread."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from hnd_ghnd_object_detectors_b200 import models, module_util
from hnd_ghnd_object_detectors_b200.optim import FusedAdam
from hnd_ghnd_object_detectors_b200.prefetch import AsyncScalarReader, DevicePrefetcher
from hnd_ghnd_object_detectors_b200.tool import DistillationBox

dev = torch.device("cuda", 0)
torch.manual_seed(0)
teacher = models.get_model(bench.model_config(False), dev)
student = models.get_model(bench.model_config(True), dev)
student.load_state_dict(teacher.state_dict(), strict=False)
module_util.freeze_module_params(teacher)
for path in bench.model_config(True)["frozen_modules"]:
    module_util.freeze_module_params(module_util.get_module(student, path))
teacher.eval(); student.train()
teacher.distill_backbone_only = student.distill_backbone_only = True
box = DistillationBox(teacher, student, bench.criterion_config())
flat = box.flatten_parameters()
opt = FusedAdam([p for p in student.parameters() if p.requires_grad], lr=1e-3, flat=flat)
host_images = [torch.rand(3, bench.IMG_H, bench.IMG_W).pin_memory() for _ in range(bench.PER_GPU_BATCH)]
targets = [{"boxes": torch.tensor([[10., 10., 100., 100.]], device=dev), "labels": torch.tensor([1], device=dev)}
           for _ in range(bench.PER_GPU_BATCH)]


class _Endless(object):
    def __iter__(self):
        while True:
            yield host_images, None

    def __len__(self):
        return 1 << 30


it, reader = iter(DevicePrefetcher(_Endless(), dev)), AsyncScalarReader()
seg = {k: 0.0 for k in ("prefetch", "forward", "zero_grad", "backward", "adam", "reader")}
N = 30
for step in range(N + 5):
    if step == 5:
        torch.cuda.synchronize()
        seg = {k: 0.0 for k in seg}
        t_all = time.perf_counter()
    t = time.perf_counter(); imgs, _ = next(it); seg["prefetch"] += time.perf_counter() - t
    t = time.perf_counter(); l = box(imgs, targets); seg["forward"] += time.perf_counter() - t
    t = time.perf_counter(); opt.zero_grad(); seg["zero_grad"] += time.perf_counter() - t
    t = time.perf_counter(); l.backward(); seg["backward"] += time.perf_counter() - t
    t = time.perf_counter(); opt.step(); seg["adam"] += time.perf_counter() - t
    t = time.perf_counter(); reader.push(l); seg["reader"] += time.perf_counter() - t
t_all = time.perf_counter() - t_all
torch.cuda.synchronize()
print("host time per step (us): " + ", ".join("%s %.0f" % (k, v / N * 1e6) for k, v in seg.items()),
      " | loop %.0f us/step (includes blocking on the in-flight bound)" % (t_all / N * 1e6))
