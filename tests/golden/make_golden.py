"""Generates tests/golden/*.npz by executing the UNMODIFIED reference (/root/reference) on CPU.

Run in the authoring container only:  PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py
The fixtures are committed; the GPU box has no /root/reference and never runs this script.

What is recorded
  quantizer.npz : tensor_util.quantize_tensor / dequantize_tensor outputs (tensor_util.py:8-22) for
                  the known-answer vector of SURVEY.md 8(a10), seeded random bottleneck-shaped
                  tensors and edge cases (all-positive, all-negative, ties at .5, constant).
  distill_small.npz : one DistillationBox forward+backward (tool.py:40-61) of Faster R-CNN b3ch
                  GHND on two small ragged images with the deterministic weights of
                  oracle/weights.py: loss, per-level terms, strided samples + moments of the
                  teacher/student layer1..4 features, of all 25 gradients, BN running stats after
                  the step, and one Adam step of conv1.weight.
  encode_small.npz : RcnnHead (split_rcnn.py:23-37) quantized bottleneck bytes / scale / zero-point
                  for the same student in eval mode.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ref_loader, weights  # noqa: E402

SAMPLE_STRIDE = 7
GRAD_STRIDE = 5


def quantizer_cases():
    """name -> float32 array; shared with the tests (which regenerate the inputs)."""
    cases = {"kat": np.array([-1, -.5, 0, .5, 1, 2], dtype=np.float32)}
    for seed in range(6):
        rng = np.random.RandomState(seed)
        n = 1 + seed % 3
        cases["rand%d" % seed] = (rng.randn(n, 3, 20, 34) * (1 + seed)).astype(np.float32)
    rng = np.random.RandomState(100)
    cases["all_pos"] = (rng.rand(2, 3, 9, 11) + 0.5).astype(np.float32)
    cases["all_neg"] = (-rng.rand(1, 3, 9, 11) - 0.25).astype(np.float32)
    cases["ties"] = (np.arange(-64, 192, dtype=np.float32) * 0.5).reshape(1, 1, 16, 16)
    cases["const_pos"] = np.full((1, 3, 4, 4), 1.5, dtype=np.float32)
    cases["const_neg"] = np.full((1, 3, 4, 4), -2.0, dtype=np.float32)
    cases["b6ch"] = (rng.randn(2, 6, 12, 20) * 3).astype(np.float32)
    cases["odd_len"] = rng.randn(1, 3, 7, 13).astype(np.float32)
    return cases


def small_images():
    g = torch.Generator().manual_seed(1234)
    return [torch.rand(3, 96, 128, generator=g), torch.rand(3, 80, 128, generator=g)]


def sample(t, stride):
    flat = t.detach().reshape(-1).to(torch.float64)
    return {"sample": flat[::stride].to(torch.float32).numpy(), "sum": float(flat.sum()),
            "sumsq": float((flat * flat).sum()), "shape": np.array(t.shape)}


def main():
    ref_loader.load()
    from myutils.pytorch import tensor_util, module_util
    out = {}
    for name, x in quantizer_cases().items():
        qt = tensor_util.quantize_tensor(torch.from_numpy(x), num_bits=8)
        out[name + ".q"] = qt.tensor.numpy()
        out[name + ".scale"] = np.float32(qt.scale.item())
        out[name + ".zp"] = np.int32(qt.zero_point)
        out[name + ".deq"] = tensor_util.dequantize_tensor(qt).numpy()
    np.savez_compressed(os.path.join(HERE, "quantizer.npz"), **out)

    # ---- distillation step on the real reference classes ----
    from models.org import rcnn
    from distillation.tool import DistillationBox
    cfg = ref_loader.load_config("ghnd", "faster_rcnn", 3)
    for k in ("teacher_model", "student_model"):
        cfg[k]["params"]["min_size"] = 96
        cfg[k]["params"]["max_size"] = 128
    tc, sc = cfg["teacher_model"], cfg["student_model"]
    teacher = rcnn.get_model(tc["name"], backbone_config=tc["backbone"], **tc["params"])
    student = rcnn.get_model(sc["name"], backbone_config=sc["backbone"], **sc["params"])
    t_sd, s_sd = weights.teacher_student(3, seed=0)
    missing = teacher.load_state_dict(t_sd, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    missing = student.load_state_dict(s_sd, strict=False)
    assert not missing.unexpected_keys, missing.unexpected_keys
    module_util.freeze_module_params(teacher)
    for path in sc["frozen_modules"]:
        module_util.freeze_module_params(module_util.get_module(student, path))
    names = module_util.get_updatable_param_names(student)
    assert len(names) == 25, names
    teacher.eval()
    student.train()
    teacher.distill_backbone_only = True
    student.distill_backbone_only = True
    student.backbone.body.layer1.use_bottleneck_transformer = False
    box = DistillationBox(teacher, student, cfg["train"]["criterion"])
    images = small_images()
    targets = [{"boxes": torch.tensor([[1., 1., 20., 20.]]), "labels": torch.tensor([1])} for _ in images]
    opt = torch.optim.Adam([p for p in student.parameters() if p.requires_grad], lr=1e-3)
    loss = box(images, targets)
    opt.zero_grad()
    loss.backward()
    rec = {"loss": np.float64(loss.item())}
    for lv in ("layer1", "layer2", "layer3", "layer4"):
        t = module_util.get_module(teacher, "backbone.body." + lv).__dict__["distillation_box"]["output"]
        s = module_util.get_module(student, "backbone.body." + lv).__dict__["distillation_box"]["output"]
        rec[lv + ".term"] = np.float64(torch.nn.functional.mse_loss(t, s, reduction="sum").item())
        for tag, v in (("teacher", t), ("student", s)):
            for k, a in sample(v, SAMPLE_STRIDE).items():
                rec["%s.%s.%s" % (lv, tag, k)] = a
    for n, p in student.named_parameters():
        if p.requires_grad:
            for k, a in sample(p.grad, GRAD_STRIDE).items():
                rec["grad.%s.%s" % (n, k)] = a
    for n, b in student.backbone.body.layer1.named_buffers():
        rec["buf.backbone.body.layer1.%s" % n] = b.detach().numpy().copy()
    w_before = student.backbone.body.conv1.weight.detach().clone()
    opt.step()
    rec["adam.conv1.weight.delta"] = (student.backbone.body.conv1.weight.detach() - w_before).numpy()
    np.savez_compressed(os.path.join(HERE, "distill_small.npz"), **rec)

    # ---- encode path (fresh student: hooks break deepcopy, split deletes modules) ----
    from models.mimic.split_rcnn import split_rcnn_model
    student2 = rcnn.get_model(sc["name"], backbone_config=sc["backbone"], **sc["params"])
    student2.load_state_dict(s_sd, strict=False)
    student2.eval()
    head, _tail = split_rcnn_model(student2, 8)
    with torch.no_grad():
        qz, tshape, image_sizes, orig_sizes = head(images)
    enc = {"q": qz.tensor.numpy(), "scale": np.float32(qz.scale.item()), "zp": np.int32(qz.zero_point),
           "tensors_shape": np.array(tshape), "image_sizes": np.array(image_sizes)}
    np.savez_compressed(os.path.join(HERE, "encode_small.npz"), **enc)
    for f in ("quantizer.npz", "distill_small.npz", "encode_small.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
    print("loss", rec["loss"])


if __name__ == "__main__":
    main()
