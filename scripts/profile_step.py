"""One eager (un-graphed) GHND step at the bench workload between cudaProfilerStart/Stop, for
`ncu --profile-from-start off`.  Also writes gpurun_out/step_ops.json: the tensor-core plan runs of
the step in launch order (description, launches, FLOPs) so the ncu launch list can be joined to
layers by scripts/join_launches.py.  Usage: python scripts/profile_step.py [batch]"""
import json, os, sys
os.environ.setdefault("GHND_SIDE_STREAM", "0")  # ncu serialises kernels anyway; keeps plan runs in one ordered stream
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hnd_ghnd_object_detectors_b200 import models, module_util, ops
from hnd_ghnd_object_detectors_b200.tool import DistillationBox

batch = int(sys.argv[1]) if len(sys.argv) > 1 else bench.PER_GPU_BATCH
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
box = DistillationBox(teacher, student, bench.criterion_config(), use_cuda_graph=False)
images = [torch.rand(3, bench.IMG_H, bench.IMG_W, device=dev) for _ in range(batch)]
targets = [{"boxes": torch.tensor([[10., 10., 100., 100.]], device=dev), "labels": torch.tensor([1], device=dev)}
           for _ in range(batch)]
box(images, targets)
box(images, targets)
torch.cuda.synchronize()

trace = []
selected = None  # None: profile the whole step; else the set of plan descs to profile (first run of each)
rt = torch.cuda.cudart()
def wrap(cls, kernel):
    orig = cls.run
    def run(self, stream=None):
        trace.append({"kernel": kernel, "desc": self.desc, "launches": getattr(self, "n_launches", 1),
                      "flops": self.flops})
        if selected is not None and self.desc in selected:
            selected.discard(self.desc)
            rt.cudaProfilerStart()
            try:
                return orig(self, stream)
            finally:
                rt.cudaProfilerStop()
        return orig(self, stream)
    cls.run = run
wrap(ops.ConvPlan, "conv_tc_kernel"); wrap(ops.StemPlan, "conv_tc_kernel"); wrap(ops.WgradPlan, "wgrad_tc_kernel")
top = int(os.environ.get("GHND_PROFILE_TOP", "0"))
want = os.environ.get("GHND_PROFILE_DESC")  # ';'-separated substrings of plan descriptions to profile (first run each)
if want:
    box(images, targets)
    torch.cuda.synchronize()
    selected = set()
    for w in want.split(";"):
        for t in trace:
            if w in t["desc"]:
                selected.add(t["desc"])
                break
    print("selected:", sorted(selected))
    del trace[:]
    box(images, targets)
    torch.cuda.synchronize()
elif top:
    # GHND_PROFILE_TOP=K: profile only the first run of the K distinct plan shapes with the most FLOPs
    # (for `ncu --set full`, which replays every profiled kernel ~40 times)
    box(images, targets)
    torch.cuda.synchronize()
    best = {}
    for t in trace:
        best[t["desc"]] = max(best.get(t["desc"], 0.0), t["flops"])
    selected = set(sorted(best, key=lambda d: -best[d])[:top])
    print("selected:", sorted(selected))
    del trace[:]
    box(images, targets)
    torch.cuda.synchronize()
else:
    rt.cudaProfilerStart()
    box(images, targets)
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/step_ops.json", "w") as f:
    json.dump(trace, f)
print("profiled one step, %d plan runs" % len(trace))
