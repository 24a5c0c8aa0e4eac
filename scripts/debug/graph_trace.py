"""In-graph timeline of the GHND step: torch.profiler (CUPTI) around a few CUDA-graph replays, then per kernel
name the time INSIDE the graph (with the side stream running), the busy / idle split of the step and the
largest idle gaps.  Complements the ncu launch list (serialised, cold caches)."""
import json
import os
import sys
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from hnd_ghnd_object_detectors_b200 import models, module_util
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
box = DistillationBox(teacher, student, bench.criterion_config())
images = [torch.rand(3, bench.IMG_H, bench.IMG_W, device=dev) for _ in range(batch)]
targets = [{"boxes": torch.tensor([[10., 10., 100., 100.]], device=dev), "labels": torch.tensor([1], device=dev)}
           for _ in range(batch)]
for _ in range(5):
    box(images, targets)
torch.cuda.synchronize()
plan = next(iter(box._plans.values()))
REPS = 4
from torch.profiler import profile, ProfilerActivity
E2E = len(sys.argv) > 2 and sys.argv[2] == "e2e"  # the public-API loop (host images, prefetcher, backward, Adam)
if E2E:
    from hnd_ghnd_object_detectors_b200.optim import FusedAdam
    from hnd_ghnd_object_detectors_b200.prefetch import AsyncScalarReader, DevicePrefetcher
    flat = box.flatten_parameters()
    opt = FusedAdam([p for p in student.parameters() if p.requires_grad], lr=1e-3, flat=flat)
    host_images = [im.cpu().pin_memory() for im in images]

    class _Endless(object):
        def __iter__(self):
            while True:
                yield host_images, None

        def __len__(self):
            return 1 << 30
    it, reader = iter(DevicePrefetcher(_Endless(), dev)), AsyncScalarReader()

    def one_step():
        imgs, _ = next(it)
        l = box(imgs, targets)
        opt.zero_grad()
        l.backward()
        opt.step()
        reader.push(l)
    for _ in range(5):
        one_step()
    torch.cuda.synchronize()
    REPS = 6
else:
    one_step = plan.step
import time
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    th = time.perf_counter()
    for _ in range(REPS):
        one_step()
    th = time.perf_counter() - th
    torch.cuda.synchronize()
print("host loop time %.1f us per step (returns before the GPU is done unless the in-flight bound blocks)" % (th / REPS * 1e6))
path = "gpurun_out/graph_trace.json"
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
print("%d device events over %d replays" % (len(ev), REPS))
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
by = defaultdict(lambda: [0, 0.0])
for e in ev:
    n = e["name"].split("(")[0].replace("void ", "").replace("ghnd::", "")
    by[n][0] += 1
    by[n][1] += e["dur"]
tot = sum(v[1] for v in by.values())
print("wall %.1f us per replay, sum of kernel durations %.1f us per replay" % ((t1 - t0) / REPS, tot / REPS))
for n, (c, d) in sorted(by.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%-60s %4d  %8.1f us  %5.1f%%" % (n[:60], c // REPS, d / REPS, 100 * d / tot))
# busy / idle: sweep
pts = []
for e in ev:
    pts.append((e["ts"], 1))
    pts.append((e["ts"] + e["dur"], -1))
pts.sort()
act, last, hist, gaps = 0, pts[0][0], defaultdict(float), []
for ts, d in pts:
    hist[min(act, 3)] += ts - last
    if act == 0 and ts - last > 0:
        gaps.append((ts - last, last))
    act += d
    last = ts
print("time with 0 / 1 / 2 / 3+ kernels in flight per replay: " + " / ".join("%.1f" % (hist[i] / REPS) for i in range(4)))
gaps.sort(reverse=True)
print("largest idle gaps (us): " + ", ".join("%.1f" % g[0] for g in gaps[:12]), " n_gaps %d" % len(gaps))
# what precedes the biggest gaps
for g, at in gaps[1:9]:
    prev = max((e for e in ev if e["ts"] + e["dur"] <= at + 0.01), key=lambda e: e["ts"] + e["dur"])
    nxt = min((e for e in ev if e["ts"] >= at + g - 0.01), key=lambda e: e["ts"])
    print("  gap %.1f us between %s and %s" % (g, prev["name"][:50], nxt["name"][:50]))
# chronological list of the LAST replay (stream, start offset, duration): the critical chain can be read off it
per = len(ev) // REPS
last = ev[-per:]
base = last[0]["ts"]
with open("gpurun_out/graph_timeline.txt", "w") as f:
    for e in last:
        a = e.get("args", {})
        n = e["name"].split("(")[0].replace("void ", "").replace("ghnd::", "")
        f.write("%9.1f %8.1f  s%-3s %s\n" % (e["ts"] - base, e["dur"], a.get("stream", "?"), n[:70]))
print("timeline of the last replay: gpurun_out/graph_timeline.txt (%d events)" % per)
