#!/usr/bin/env python
"""bench.py -- GHND distillation throughput on B200 (BASELINE.json configs[1]).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

Workload (N=1): Faster R-CNN ResNet-50-FPN b3ch GHND distillation step -- teacher forward, student
forward, 4-level SSE loss, student backward, gradient all-reduce (N>1), fused Adam -- on 4 synthetic
3x800x1333 images per GPU (config/ghnd/faster_rcnn-backbone_resnet50-b3ch.yaml: batch_size 4),
random-init weights.  One JSON line is printed by rank 0.  Beside the headline it carries the other
BASELINE.json configs: `config1` (head + 8-bit quantize: CPU reference leg and the CUDA path, batch 2),
`config4` (Keypoint R-CNN, batch 8 per GPU, all-800 and random-scale seed 0; N=1 and N=8 runs),
`encode` (config 5: batch 1..64 sweep), and an informational `stock_torch_cuda` leg (the oracle port
of the same step run by stock PyTorch/cuDNN on the same B200, fp32 and TF32).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG_H, IMG_W = 800, 1333
PER_GPU_BATCH = 4
METRIC = "ghnd_distill_images_per_sec"
WORKLOAD = ("Faster R-CNN ResNet-50-FPN b3ch GHND distillation step (teacher fwd + student fwd/bwd + "
            "4-level SSE loss + Adam), %d synthetic 3x800x1333 images per GPU, random init" % PER_GPU_BATCH)


def model_config(student, bch=3):
    cfg = {
        "name": "faster_rcnn",
        "backbone": {"name": "custom_resnet50" if student else "resnet50",
                     "params": {"pretrained": False, "freeze_layers": not student}},
        "params": {"num_classes": 91, "pretrained": False},
        "ckpt": "./resource/ckpt/none.pt",
    }
    if student:
        cfg["backbone"]["params"]["layer1"] = {"name": "Bottleneck4LargeResNet", "bottleneck_channel": bch}
        cfg["frozen_modules"] = ["backbone.body.layer2", "backbone.body.layer3", "backbone.body.layer4",
                                 "backbone.fpn", "rpn", "roi_heads"]
    return cfg


def criterion_config():
    terms = {lv: {"ts_modules": ["backbone.body." + lv, "backbone.body." + lv],
                  "criterion": {"type": "MSELoss", "params": {"reduction": "sum"}}, "factor": 1.0}
             for lv in ("layer1", "layer2", "layer3", "layer4")}
    return {"type": "general", "params": {"org_loss_factor": 0.0}, "terms": terms}


class ClockSampler(object):
    """SM clock / throttle reasons sampled DURING the timed region: NVML polled every 5 ms from a thread
    (the driver's 20-step region lasts ~0.1 s -- too short for `nvidia-smi -lms`); falls back to
    nvidia-smi when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml = index, [], None, None
        self.stop_flag = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
            except Exception:
                pass
            self.handle = None
            if uuid:
                for i in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(i)
                    u = pynvml.nvmlDeviceGetUUID(h)
                    u = u.decode() if isinstance(u, bytes) else u
                    if uuid in u:
                        self.handle = h
            if self.handle is None:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self.nvml = None

    def _poll(self):
        nv, h = self.nvml, self.handle
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(sm), [k for k, b in bits.items() if reasons & b]))
            except Exception:
                break
            time.sleep(0.005)

    def start(self):
        if self.nvml is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            sm = sorted(r[0] for r in self.rows)
            reasons = sorted({k for r in self.rows for k in r[1]})
            try:
                mx = float(self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
            except Exception:
                mx = None
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons,
                    "samples": len(sm), "source": "NVML polled every 5 ms during the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except Exception:
                continue
            for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6),
                              ("sw_power_cap", 7)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 50"}


def ncu_traffic():
    """DRAM bytes of the conv family from the committed ncu capture (profiles/), if present."""
    for name in ("r2_conv_dram_traffic.json", "r1_conv_dram_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.isfile(path):
            with open(path) as f:
                d = json.load(f)
            d["file"] = "profiles/" + name
            return d
    return {}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_tflops_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of the reference algorithm on the host cores
# ------------------------------------------------------------------------------------------------
CPU_SAMPLE_BATCH = 2  # the reference arm steps on 2 images


def cpu_step_fn(sample_batch, device="cpu", tf32=False):
    """One step of the reference algorithm (oracle port, stock torch ops): teacher forward, student
    forward + backward through autograd, 4-level MSE(sum) loss, Adam over the 25 trainable tensors
    (src/distillation/tool.py:40-61, src/mimic_runner.py:51-54).  device="cuda": the same Python on the
    GPU through cuDNN/ATen ("stock PyTorch on the same box")."""
    import torch
    from oracle import ghnd_oracle as O
    from oracle import weights
    if device == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    else:
        torch.backends.cudnn.allow_tf32 = bool(tf32)
        torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    t_sd, s_sd = weights.teacher_student(3, seed=0)
    t_sd = {k: v.to(device) for k, v in t_sd.items()}
    s_sd = {k: v.to(device) for k, v in s_sd.items()}
    g = torch.Generator().manual_seed(0)
    images = [torch.rand(3, IMG_H, IMG_W, generator=g).to(device) for _ in range(sample_batch)]
    state = {"step": 0, "m": {}, "v": {}}

    def step():
        res = O.distill_step(t_sd, s_sd, images)
        state["step"] += 1
        for n, gr in res["grads"].items():
            m = state["m"].get(n, torch.zeros_like(gr))
            v = state["v"].get(n, torch.zeros_like(gr))
            s_sd[n], state["m"][n], state["v"][n] = O.adam_step(s_sd[n], gr, m, v, state["step"])
        return float(res["loss"])
    return step


def cpu_encode_fn(sample_batch):
    """BASELINE config 1: Faster R-CNN b3ch head forward + 8-bit quantize on the host cores."""
    import torch
    from oracle import ghnd_oracle as O
    from oracle import weights
    torch.set_num_threads(os.cpu_count() or 1)
    _, s_sd = weights.teacher_student(3, seed=0)
    g = torch.Generator().manual_seed(0)
    images = [torch.rand(3, IMG_H, IMG_W, generator=g) for _ in range(sample_batch)]

    def step():
        return O.encode_head(images, s_sd, 8)[1]
    return step


def time_host(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return (time.perf_counter() - t0) / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_batch = CPU_SAMPLE_BATCH
    step = cpu_step_fn(sample_batch)
    dt = time_host(step, args.warmup, args.steps) * args.steps
    val = sample_batch * args.steps / dt
    cores = os.cpu_count() or 1
    sample = ("each step = oracle port (torch CPU fp32, oracle/ghnd_oracle.py distill_step + adam_step) of the "
              "same GHND step (teacher fwd, student fwd/bwd, 4-level loss, Adam) on %d images of 3x800x1333, "
              "%d torch threads" % (sample_batch, cores))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOAD, "sample_batch": sample_batch},
            "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def conv_plans(plan):
    """Every tcgen05 conv plan (forward / dgrad / stem) one GHND step runs, in no particular order."""
    out = [plan.stem2.plan] if plan.stem2 is not None else [plan.t_stem.plan, plan.s_stem.plan]
    for r in list(plan.t_layers.values()) + list(plan.s_layers.values()):
        for b in r.blocks:
            out += list(b.fwd) + list(b.bwd)
    l1 = plan.s_l1
    for u in (l1.e0, l1.e1, l1.e2, l1.d4, l1.d7, l1.d9):
        out.append(u.plan)
        if u.dgrad is not None:
            out.append(u.dgrad)
    return out


def conv_flops_per_step(plan):
    """Algorithmic FLOPs (2*MACs) of every conv_tc_kernel launch in one step, from the plans."""
    return float(sum(p.flops for p in conv_plans(plan)))


def time_entry_points(fn):
    """CUDA-event time of every C-ABI entry point called by fn() (one eager, un-graphed pass):
    {name: (seconds, calls)}.  Events are recorded on torch's current stream, which is the stream
    every kernel of the path is launched on."""
    import torch
    from hnd_ghnd_object_detectors_b200 import _lib, ops
    spans = []
    orig_ops, orig_lib = ops.call, _lib.call

    def timed_call(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig_lib(name, *a)
        e1.record()
        spans.append((name, e0, e1))
    ops.call = _lib.call = timed_call
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        ops.call, _lib.call = orig_ops, orig_lib
    out = {}
    for name, e0, e1 in spans:
        t, c = out.get(name, (0.0, 0))
        out[name] = (t + e0.elapsed_time(e1) * 1e-3, c + 1)
    return out


def flush_l2(dev, _buf={}):
    """Write a buffer larger than the 126 MB L2 so the next kernel starts cold."""
    import torch
    if "b" not in _buf:
        _buf["b"] = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    _buf["b"].zero_()


def time_cold(fns, dev, iters=5):
    """Median CUDA-event time (s) PER CALL of the closures in `fns`, launched back to back after an L2
    flush.  The closures do the same work on DISTINCT buffers whose total exceeds the L2, so every call
    streams from HBM while the ~5 us of launch + event overhead of a lone short kernel is amortised."""
    import torch
    if callable(fns):
        fns = [fns]
    ts = []
    for _ in range(iters + 1):
        flush_l2(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for fn in fns:
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3 / len(fns))
    ts = sorted(ts[1:])
    return ts[len(ts) // 2]


def encode_sweep(dev, batches, iters=10, model_name="keypoint_rcnn"):
    """Config 5: Keypoint R-CNN b3ch split-computing head (stem + layer1 encoder + 8-bit quantizer),
    images/s per batch size, inputs resident in HBM; also achieved HBM GB/s of the bottleneck-side
    kernels at the largest batch (cold L2): the encoder's last conv (64 -> bch, narrow_out), the
    one-pass quantizer fed by its min/max, the stand-alone quantizer and the dequantizer."""
    import torch
    from hnd_ghnd_object_detectors_b200 import models, ops
    from hnd_ghnd_object_detectors_b200.split_rcnn import split_rcnn_model
    cfg = model_config(True)
    cfg["name"] = model_name
    if model_name == "keypoint_rcnn":
        cfg["params"] = {"num_classes": 2, "pretrained": False, "num_keypoints": 17}
    torch.manual_seed(0)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        model = models.get_model(cfg, dev).eval()
    head, _ = split_rcnn_model(model, 8)
    out = {}
    g = torch.Generator().manual_seed(5)
    pool = [torch.rand(3, IMG_H, IMG_W, generator=g).to(dev) for _ in range(4)]
    kernels = {}
    for b in batches:
        images = [pool[i % len(pool)] for i in range(b)]
        head(images)  # builds + captures the fixed-shape plan
        plan = head.plan
        for _ in range(3):
            plan.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            plan.run()
        e1.record()
        torch.cuda.synchronize()
        out[str(b)] = b * iters / (e0.elapsed_time(e1) * 1e-3)
        if b == batches[-1]:
            z, l1 = plan.z, plan.l1
            nz = float(z.numel())
            wide = float(l1.e2.out.numel())
            # four buffer sets (4 x 66 MB > L2): back-to-back launches stay HBM-cold
            zs = [z] + [z.clone() for _ in range(3)]
            qs = [plan.q] + [torch.empty_like(plan.q) for _ in range(3)]
            qps = [plan.qparams] + [torch.empty_like(plan.qparams) for _ in range(3)]
            t = time_cold([lambda i=i: ops.quantize_u8_minmax(zs[i], plan.minmax, l1.z_pairs, 8, plan.scale_mode,
                                                              q=qs[i], qparams=qps[i]) for i in range(4)], dev)
            kernels["quant_apply (one-pass 8-bit quantizer of the encode path, min/max from the encoder's last conv; 5 B/elem, batch %d)" % b] = (5.0 * nz / t, 0.8)
            qws = ops.quantize_ws(z.numel(), dev)
            t = time_cold([lambda i=i: ops.quantize_u8(zs[i], 8, q=qs[i], qparams=qps[i], ws=qws) for i in range(4)], dev)
            kernels["quantize_u8 (stand-alone quantize_tensor: min/max + quantize in one persistent launch; 5 B/elem compulsory, batch %d)" % b] = (5.0 * nz / t, 0.8)
            q, qp = ops.quantize_u8(z, 8)
            qq = [q] + [q.clone() for _ in range(3)]
            t = time_cold([lambda i=i: ops.dequantize_u8(qq[i], qp, out=zs[i]) for i in range(4)], dev)
            kernels["dequantize_u8 (1 B in, 4 B out per elem, batch %d)" % b] = (5.0 * nz / t, 0.2)
            xs = [l1.e2.out] + [l1.e2.out.clone() for _ in range(3)]
            t = time_cold([lambda i=i: ops.conv_narrow_out(xs[i], l1.enc7.weight, 1, y=zs[i], ws=l1.nws,
                                                           minmax=plan.minmax) for i in range(4)], dev)
            kernels["narrow_out (encoder's last conv 64 -> bch k2 + min/max; 2 B/elem of the 64-ch input + 4 B/elem of z, batch %d)" % b] = ((2.0 * wide + 4.0 * nz) / t, 2.0 * wide / (2.0 * wide + 4.0 * nz))
            del zs, qs, qq, xs
            # The batch-64 tensor is 53 MB: a launch's fixed cost (~8 us of launch, prologue, ramp and tail; a
            # torch copy of the same size has ~4 us and reaches 0.7 of the copy peak) weighs as much as the
            # streaming.  The same kernels on a 16x larger tensor show the streaming rate by itself.
            big = torch.randn(16 * z.numel(), device=dev)
            bq, bqp = torch.empty(big.numel(), dtype=torch.uint8, device=dev), torch.empty_like(plan.qparams)
            mm = torch.stack([big.min(), big.max()]).contiguous()
            t = time_cold(lambda: ops.quantize_u8_minmax(big, mm, 1, 8, plan.scale_mode, q=bq, qparams=bqp), dev)
            kernels["quant_apply, streaming rate (same kernel, 16x the batch-%d tensor = %.0f MB of fp32)" % (b, big.numel() * 4e-6)] = (5.0 * big.numel() / t, 0.8)
            bws = ops.quantize_ws(big.numel(), dev)
            t = time_cold(lambda: ops.quantize_u8(big, 8, q=bq, qparams=bqp, ws=bws), dev)
            kernels["quantize_u8 stand-alone, streaming rate (two passes above 20 M elements: 9 B/elem of traffic for 5 B/elem compulsory; 16x tensor)"] = (5.0 * big.numel() / t, 0.8)
            t = time_cold(lambda: ops.dequantize_u8(bq, bqp, out=big), dev)
            kernels["dequantize_u8, streaming rate (16x tensor)"] = (5.0 * big.numel() / t, 0.2)
            del big, bq
        head.plan = None
    return out, kernels


def config4_bench(dev, world, steps=10, warmup=3):
    """BASELINE config 4: Keypoint R-CNN b3ch GHND distillation, 8 images per GPU.  The Keypoint path
    draws a per-image `fixed_sizes` from min_size = (640..800) (src/distillation/tool.py:44-49), so
    every step is a bilinear resize + a padded shape picked by the largest image of the batch; the
    box keeps one plan + CUDA graph per padded shape.  Two runs: all images at 800 (one shape) and
    random scales with random.seed(0).  Through the public API (DistillationBox -> backward -> flat
    all-reduce -> FusedAdam), images resident on the device."""
    import contextlib
    import random
    import torch
    import torch.distributed as dist
    from hnd_ghnd_object_detectors_b200 import models, module_util, parallel
    from hnd_ghnd_object_detectors_b200.optim import FusedAdam
    from hnd_ghnd_object_detectors_b200.tool import DistillationBox
    batch = 8
    rank = int(os.environ.get("RANK", "0"))

    def cfg(student):
        c = model_config(student)
        c["name"] = "keypoint_rcnn"
        c["params"] = {"num_classes": 2, "pretrained": False, "num_keypoints": 17}
        return c
    torch.manual_seed(0)
    with contextlib.redirect_stdout(sys.stderr):
        teacher = models.get_model(cfg(False), dev)
        student = models.get_model(cfg(True), dev)
    student.load_state_dict(teacher.state_dict(), strict=False)
    module_util.freeze_module_params(teacher)
    for path in cfg(True)["frozen_modules"]:
        module_util.freeze_module_params(module_util.get_module(student, path))
    teacher.eval()
    student.train()
    teacher.distill_backbone_only = student.distill_backbone_only = True
    g = torch.Generator().manual_seed(4000 + rank)
    images = [torch.rand(3, IMG_H, IMG_W, generator=g).to(dev) for _ in range(batch)]
    targets = [{"boxes": torch.tensor([[10., 10., 100., 100.]], device=dev), "labels": torch.tensor([1], device=dev)}
               for _ in range(batch)]
    all_sizes = tuple(teacher.transform.min_size)
    out = {"workload": "Keypoint R-CNN ResNet-50-FPN b3ch GHND distillation, %d synthetic 3x800x1333 images per GPU, "
                       "per-image fixed_sizes from %s, public API loop, images resident on the device" % (batch, list(all_sizes)),
           "per_gpu_batch": batch, "n_gpus": world}
    for mode, sizes in (("all_800", (800,)), ("random_scale_seed0", all_sizes)):
        teacher.transform.min_size = sizes
        box = DistillationBox(teacher, student, criterion_config())
        flat = box.flatten_parameters()
        opt = FusedAdam([p for p in student.parameters() if p.requires_grad], lr=1e-3, grad_scale=1.0 / world, flat=flat)
        random.seed(0)

        def step():
            loss = box(images, targets)
            opt.zero_grad()
            loss.backward()
            parallel.allreduce_flat_grad(flat)
            opt.step()
        # warm-up: visit (build + capture) every padded shape the seeded sequence will need
        state = random.getstate()
        for _ in range(warmup + steps):
            step()
        random.setstate(state)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        out[mode] = {"value": batch * world * steps / (ms * 1e-3), "unit": "images/s", "ms_per_step": ms / steps,
                     "steps": steps, "padded_shapes": sorted("%dx%d" % (k[1], k[2]) for k in box._plans)}
        del box, opt
        torch.cuda.empty_cache()
    teacher.transform.min_size = all_sizes
    return out


def stock_torch_cuda(dev):
    """Informational: the oracle port of the same GHND step (stock torch ops -> cuDNN/ATen, autograd,
    per-tensor Adam math) on the same B200, fp32 and TF32, 4 images per step."""
    import torch
    out = {}
    for name, tf32 in (("fp32", False), ("tf32", True)):
        step = cpu_step_fn(PER_GPU_BATCH, device=str(dev), tf32=tf32)
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 3
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        out[name] = {"value": PER_GPU_BATCH * n / (e0.elapsed_time(e1) * 1e-3), "unit": "images/s"}
        del step
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = True
    out["what"] = ("oracle/ghnd_oracle.py distill_step + adam_step with every tensor on cuda: stock PyTorch %s "
                   "(cuDNN/ATen kernels, autograd), %d images of 3x800x1333 per step, 1 warm-up + 3 timed steps; "
                   "not this repo's kernels" % (torch.__version__, PER_GPU_BATCH))
    return out


def run_cuda(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the GHND path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__
    __graft_entry__.build()
    from hnd_ghnd_object_detectors_b200 import models, module_util, ops
    from hnd_ghnd_object_detectors_b200.optim import FusedAdam
    from hnd_ghnd_object_detectors_b200.tool import DistillationBox

    torch.manual_seed(0)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):  # "ckpt file is not found" notices: keep stdout = one JSON line
        teacher = models.get_model(model_config(False), dev)
        student = models.get_model(model_config(True), dev)
    student.load_state_dict(teacher.state_dict(), strict=False)  # layer2-4 identical (pretrained flow)
    module_util.freeze_module_params(teacher)
    for path in model_config(True)["frozen_modules"]:
        module_util.freeze_module_params(module_util.get_module(student, path))
    teacher.eval()
    student.train()
    teacher.distill_backbone_only = student.distill_backbone_only = True
    box = DistillationBox(teacher, student, criterion_config(),
                          use_cuda_graph=os.environ.get("GHND_NO_GRAPH") is None)
    opt = FusedAdam([p for p in student.parameters() if p.requires_grad],
                    lr=float(os.environ.get("GHND_BENCH_LR", "1e-3")),
                    grad_scale=1.0 / world)

    g = torch.Generator().manual_seed(1000 + rank)
    host_images = [torch.rand(3, IMG_H, IMG_W, generator=g).pin_memory() for _ in range(PER_GPU_BATCH)]
    dev_images = [im.to(dev) for im in host_images]
    targets = [{"boxes": torch.tensor([[10., 10., 100., 100.]], device=dev),
                "labels": torch.tensor([1], device=dev)} for _ in range(PER_GPU_BATCH)]

    # first call builds + captures the plan
    loss = box(dev_images, targets)
    opt.attach(box.flat)
    plan = list(box._plans.values())[0]
    plan_shared, plan_side = bool(plan.shared), plan.side is not None
    torch.cuda.synchronize()
    from hnd_ghnd_object_detectors_b200 import parallel
    parallel.broadcast_flat_params(box.flat)  # every rank starts from rank 0's student (DDP semantics)
    parallel.broadcast_buffers(student)

    dbg = {"n": 0}

    def device_step():
        if os.environ.get("GHND_BENCH_TRACE"):
            dbg["n"] += 1
            if dbg["n"] % 20 == 0:
                torch.cuda.synchronize()
                print("step %d loss %s" % (dbg["n"], plan.loss_out.tolist()), file=sys.stderr, flush=True)
        plan.step()  # images already packed in HBM; graph replay of fwd + loss + bwd
        parallel.allreduce_flat_grad(box.flat)
        opt.step()

    # end-to-end: the loop a user writes with the package's public API (mimic_runner.distill_model):
    # every step uploads its 4 images from pinned host memory (DevicePrefetcher: batch i+1 is copied
    # on a side stream while step i runs) and reads one loss value back (AsyncScalarReader: the
    # previous step's value, so the host never stalls the GPU).
    from hnd_ghnd_object_detectors_b200.prefetch import AsyncScalarReader, DevicePrefetcher

    class _Endless(object):
        def __iter__(self):
            while True:
                yield host_images, None

        def __len__(self):
            return 1 << 30
    e2e_state = {"iter": iter(DevicePrefetcher(_Endless(), dev)), "reader": AsyncScalarReader(), "last": None}

    def e2e_step():
        imgs, _ = next(e2e_state["iter"])
        l = box(imgs, targets)
        opt.zero_grad()
        l.backward()                            # p.grad become views of the flat gradient buffer
        parallel.allreduce_flat_grad(box.flat)  # one NCCL all-reduce (N > 1)
        opt.step()
        v = e2e_state["reader"].push(l)  # device -> host read of a step's loss, one per step
        if v is not None:
            e2e_state["last"] = v

    def timed(fn, steps, warmup, sampler=None):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launches()
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        if world > 1:
            dist.barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), ops.launches() - l0, clocks

    sampler = ClockSampler(local) if rank == 0 and not os.environ.get("GHND_NO_CLOCKS") else None
    ms, launches, clocks = timed(device_step, args.steps, args.warmup, sampler)
    e2e_steps = max(3, min(args.steps, 30))
    ms_e2e, _, _ = timed(e2e_step, e2e_steps, max(3, min(args.warmup, 3)))
    images_per_step = PER_GPU_BATCH * world
    value = images_per_step * args.steps / (ms * 1e-3)
    e2e_value = images_per_step * e2e_steps / (ms_e2e * 1e-3)

    roof = cpu = hbm_kernels = encode = config1 = stock = None
    if rank == 0:
        peaks = measured_peaks()
        hbm = peaks["hbm_gbs"]

        def entry(gbs, read_share=None):
            """frac: against the measured COPY bandwidth (MEASURED_PEAKS.json).  HBM on this part is
            directional (profiles/r2_hbm_directional.txt: read-only 5777, write-only 3872, copy 6441 GB/s),
            so a kernel with read share r of its bytes cannot stream faster than
            1 / max(r/5777, (1-r)/3872, 1/6441) GB/s; frac_directional is measured against that."""
            e = {"achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm}
            if read_share is not None:
                lim = 1.0 / max(read_share / 5777.0, (1.0 - read_share) / 3872.0, 1.0 / 6441.0)
                e.update({"read_share": round(read_share, 3), "directional_peak": round(lim, 1),
                          "frac_directional": gbs / lim})
            return e
        # per-kernel timing pass: single stream (the side-stream branches would overlap the spans)
        side, plan.side = plan.side, None
        try:
            ep = time_entry_points(plan.forward_backward)
        finally:
            plan.side = side
        conv_s = sum(ep.get(k, (0.0, 0))[0] for k in ("ghnd_conv_plan_run", "ghnd_conv_plan_run_range",
                                                      "ghnd_stem_conv_plan_run", "ghnd_stem_pool_plan_run"))
        n_conv = sum(p.n_launches for p in conv_plans(plan))
        flops = conv_flops_per_step(plan)
        achieved = flops / conv_s / 1e12
        ncu = ncu_traffic()
        traffic = ncu.get("conv_tc_dram_bytes_per_step")
        note = ncu.get("note")
        if traffic is not None and ncu.get("launches") not in (None, n_conv):
            note = "STALE (%s launches captured, this build runs %d): %s" % (ncu.get("launches"), n_conv, note)
        roof = {"bound": "tensor",
                "kernel": "conv_tc_kernel + stem_pool_kernel (tcgen05 implicit-GEMM conv fwd/dgrad, conv1 fused with its max-pool; %d launches/step; "
                          "achieved = algorithmic FLOPs of all launches / summed CUDA-event time)" % n_conv,
                "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops_sustained"],
                "traffic": traffic, "traffic_note": note, "traffic_file": ncu.get("file"),
                "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                "conv_share_of_step": conv_s / sum(t for t, _ in ep.values()),
                "flops_per_step": flops,
                "whole_step_tflops": value / world * 701.8e9 / 1e12,
                "whole_step_frac": value / world * 701.8e9 / 1e12 / peaks["bf16_tflops_sustained"]}
        # memory-bound kernels of the path against the measured copy bandwidth
        sse_bytes = 6.0 * sum(t.numel() for t in plan.feat_s.values())
        hbm_kernels = {}
        if "ghnd_sse_fwd_bwd" in ep:
            hbm_kernels["sse_kernel (4-level loss fwd+bwd, 6 B/elem)"] = entry(sse_bytes / ep["ghnd_sse_fwd_bwd"][0] / 1e9, 4.0 / 6.0)
        # the other streaming kernels of the step: algorithmic (read, write) bytes from the plan's tensors
        l1 = plan.s_l1
        units = [l1.e0, l1.e1, l1.e2, l1.d4, l1.d7, l1.d9]
        raw_elems = float(sum(u.raw.numel() for u in units))
        raw3 = float(l1.raw3.numel())
        pool_elems = float(plan.s_stem.out.numel())
        conv_elems = 4.0 * pool_elems  # one model's 64-channel conv1 output (2x2 the pooled size)
        nz = float(l1.z.numel())
        wide_e2, wide_r3 = float(l1.e2.out.numel()), raw3
        streaming = {
            "bn_finalize_apply (x -> y f16 [+ bf16 copy where a tensor-core dW reads it], 4-6 B/elem)":
                ("ghnd_bn_finalize_apply", 2.0 * (raw_elems + raw3),
                 sum(u.raw.numel() * (4.0 if u.out_g is not None else 2.0) for u in units) + 4.0 * raw3),
            "bn_bwd_reduce (g, x -> sums, 4 B/elem)": ("ghnd_bn_bwd_reduce", 4.0 * (raw_elems + raw3) + 8.0 * nz, 0.0),
            "bn_bwd_apply (g, x -> dx, 6 B/elem)":
                ("ghnd_bn_bwd_apply", 4.0 * (raw_elems + raw3) + 8.0 * nz, 2.0 * (raw_elems + raw3) + 4.0 * nz),
            "maxpool 3x3 s2 fwd, teacher + student (2 B in, 2 B out, +1 B argmax)":
                ("ghnd_maxpool3x3s2_strided", 4.0 * conv_elems, 5.0 * pool_elems),
            "maxpool bwd + ReLU mask (dx 2 B/elem written; dy + argmax 3 B/pooled elem read)":
                ("ghnd_maxpool3x3s2_bwd_strided", 3.0 * pool_elems, 2.0 * conv_elems),
            # the bottleneck side (north_star (b)): wide 64-channel 16-bit tensor <-> planar fp32 bch tensor
            "narrow_out fwd (enc7: 64 -> bch k2 p1; 2 B/elem wide in + 4 B/elem z out)":
                ("ghnd_conv_narrow_out", 2.0 * wide_e2, 4.0 * nz),
            "narrow_in (dec2 fwd bch -> 64 with BN+ReLU prologue, and enc7 dgrad; 4 B/elem planar in + 2 B/elem wide out, 2 launches)":
                ("ghnd_conv_narrow_in", 8.0 * nz, 2.0 * wide_r3 + 2.0 * wide_e2),
            "narrow_out dgrad (dec2: 64 -> bch; 2 B/elem wide in + 4 B/elem out)":
                ("ghnd_conv_narrow_out_dgrad", 2.0 * wide_r3, 4.0 * nz),
            "wgrad_narrow (dec2 + enc7 dW; wide 2 B/elem + planar 4 B/elem, 2 launches)":
                ("ghnd_wgrad_narrow", 2.0 * wide_r3 + 4.0 * nz + 2.0 * wide_e2 + 4.0 * nz, 0.0),
        }
        for label, (name, rd, wr) in streaming.items():
            if name in ep and ep[name][0] > 0:
                hbm_kernels[label] = entry((rd + wr) / ep[name][0] / 1e9, rd / (rd + wr))
        breakdown = {k: round(v[0] * 1e3, 4) for k, v in sorted(ep.items(), key=lambda kv: -kv[1][0])}
        roof["entry_point_ms_per_step"] = breakdown
        if world == 1 and not args.no_encode:
            sweep, kern = encode_sweep(dev, [8, 64] if args.short_encode else [1, 2, 4, 8, 16, 32, 64])
            encode = {"metric": "head_quant_encode_images_per_sec", "unit": "images/s",
                      "workload": "BASELINE config 5: Keypoint R-CNN b3ch RcnnHead (stem + layer1 encoder + 8-bit "
                                  "quantize), synthetic 3x800x1333, inputs resident in HBM, one scale/zero-point "
                                  "per call", "by_batch": sweep}
            for label, (bps, rshare) in kern.items():
                hbm_kernels[label] = entry(bps / 1e9, rshare)
            # config 1 on the CUDA path: Faster R-CNN head + quantize at batch 2
            f_sweep, _ = encode_sweep(dev, [2], model_name="faster_rcnn")
            config1 = {"workload": "BASELINE config 1: Faster R-CNN b3ch head forward + 8-bit quantize, batch 2, "
                                   "synthetic 3x800x1333", "cuda": {"value": f_sweep["2"], "unit": "images/s"}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n = 2
        dt = time_host(cpu_step_fn(CPU_SAMPLE_BATCH), 1, n)
        cores = os.cpu_count() or 1
        cpu = {"value": CPU_SAMPLE_BATCH / dt, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": "oracle port (torch CPU fp32, autograd + Adam) of the same GHND step on %d images of "
                         "3x800x1333 per step, %d torch threads, 1 warm-up + %d timed steps" % (CPU_SAMPLE_BATCH, cores, n)}
        dt1 = time_host(cpu_encode_fn(2), 1, 3)
        if config1 is None:
            config1 = {"workload": "BASELINE config 1"}
        config1["cpu"] = {"value": 2.0 / dt1, "unit": "images/s", "cores": cores, "kind": "port",
                          "sample": "oracle port (torch CPU fp32 + numpy quantizer) of RcnnHead on 2 images of "
                                    "3x800x1333, 1 warm-up + 3 timed calls"}
        if not args.no_stock_torch:
            stock = stock_torch_cuda(dev)
    config4 = None
    if (world in (1, 8) and not args.no_config4) or args.config4:
        del box, plan
        torch.cuda.empty_cache()
        try:
            config4 = config4_bench(dev, world)
        except Exception as e:  # the headline line must survive a failure of the secondary workload
            config4 = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    if rank == 0:
        h2d = sum(h.numel() * 4 for h in host_images)
        line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": "f16 forward / bf16 gradients, fp32 accumulate (tcgen05 kind::f16)",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": images_per_step, "per_gpu_batch": PER_GPU_BATCH,
                           "parallelism": "dp%d" % world,
                           "l2": "per-step working set (several GB of activations) exceeds the 126 MB L2",
                           "shared_frozen_trunk": plan_shared,
                           "cuda_graph": True, "side_stream": plan_side,
                           "device_resident_value": "`value` replays the step on images already packed in HBM "
                                                    "(the stem-pack kernel of the 4 inputs is not in it); `e2e` "
                                                    "includes the H2D copy and the packing"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 4, "steps": e2e_steps,
                        "api": "DevicePrefetcher(pinned host images) -> DistillationBox -> backward -> flat all-reduce "
                               "-> FusedAdam -> AsyncScalarReader(loss), i.e. mimic_runner.distill_model's loop"},
                "gpu_launches": launches, "roofline": roof, "roofline_hbm_kernels": hbm_kernels,
                "encode": encode, "config1": config1, "config4": config4, "cpu_baseline": cpu,
                "stock_torch_cuda": stock,
                "loss": float(loss.item())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-encode", action="store_true", help="skip the split-computing head line")
    ap.add_argument("--short-encode", action="store_true", help="encode path at batch 8 and 64 only")
    ap.add_argument("--encode-sweep", action="store_true", help="(default now) encode path at batch 1..64")
    ap.add_argument("--no-config4", action="store_true", help="skip the Keypoint batch-8 line (config 4)")
    ap.add_argument("--config4", action="store_true", help="run config 4 at any N (default: N=1 and N=8)")
    ap.add_argument("--no-stock-torch", action="store_true", help="skip the stock-PyTorch-on-cuda leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
