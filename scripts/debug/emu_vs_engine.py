"""Debug: where does the engine's forward leave the storage-precision emulation (oracle16)?
Compares intermediate teacher tensors of a GhndPlan with the emulation, stage by stage."""
import os, sys
os.environ.setdefault("GHND_STEM_POOL", "0")  # this script looks at conv1's output, which only the un-fused stem stores
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ghnd_oracle as O, ghnd_oracle16 as E, weights
from tests.test_gpu_distill import build_pair, criterion_config, targets_for
from tests.golden.make_golden import small_images
from hnd_ghnd_object_detectors_b200 import models, module_util, ops
from hnd_ghnd_object_detectors_b200.tool import DistillationBox

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

def stats(name, got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    d = (got - ref).abs()
    ulp = (ref.abs().clamp_min(6e-8)) * 2 ** -10
    frac = float((d > 0).float().mean())
    big = float((d > 1.5 * ulp).float().mean())
    idx = int(d.argmax())
    pos = []
    for dim in reversed(got.shape):
        pos.append(idx % dim)
        idx //= dim
    print("%-28s rel %.3e  differing %.4f  >1.5ulp %.5f  max|d| %.3e at %s of %s" % (
        name, rel(got, ref), frac, big, float(d.max()), list(reversed(pos)), list(got.shape)))
    if big > 0:
        m = (d > 1.5 * ulp)
        # where do the large differences sit? histogram over rows / cols
        rows = m.sum(dim=(0, 1, 3)).nonzero().flatten().tolist()
        cols = m.sum(dim=(0, 1, 2)).nonzero().flatten().tolist()
        print("     rows with >1.5ulp: %d (first %s last %s)  cols: %d (first %s last %s)" % (
            len(rows), rows[:4], rows[-4:], len(cols), cols[:4], cols[-4:]))

t_sd, s_sd = weights.teacher_student(3, seed=0)
env = {"models": models, "module_util": module_util, "t_sd": t_sd, "s_sd": s_sd}
teacher, student = build_pair(env)
box = DistillationBox(teacher, student, criterion_config(), use_cuda_graph=False)
host = small_images()
if len(sys.argv) > 1 and sys.argv[1] == "full":
    g = torch.Generator().manual_seed(5)
    host = [torch.rand(3, 800, 1333, generator=g)]
    import tests.test_gpu_fullsize as F
    teacher, student = F.build_full_pair(env)
    box = DistillationBox(teacher, student, criterion_config(), use_cuda_graph=False)
images = [im.cuda() for im in host]
loss = box(images, targets_for(images))
torch.cuda.synchronize()
plan = list(box._plans.values())[0]
nchw = lambda t: t.float().permute(0, 3, 1, 2).contiguous().cpu()

x = O.transform_batch(host)
x16 = E.r16(x)
N, _, H, W = x.shape
packed = plan.packed[:, 3:3 + H, 3:3 + W, :3]
stats("packed image", nchw(packed), x16)
sd = t_sd
p = "backbone.body."
with torch.no_grad():
    c = E.frozen_conv(x16, sd, p + "conv1", p + "bn1", 2, 3, True)
    stats("conv1+bn+relu (teacher)", nchw(plan.stem2.conv[..., :64]), c)
    c32 = torch.relu(O.frozen_bn(torch.nn.functional.conv2d(x, sd[p + "conv1.weight"], None, 2, 3), sd, p + "bn1"))
    stats("  emulation vs fp32", c, c32)
    stats("  engine vs fp32", nchw(plan.stem2.conv[..., :64]), c32)
    pool = torch.nn.functional.max_pool2d(c, 3, 2, 1)
    stats("pooled", nchw(plan.t_stem.out), pool)
    b0 = plan.t_layers["layer1"].blocks[0]
    pre = p + "layer1.0"
    # feed the ENGINE's tensors into each emulated op so that every line isolates one kernel
    xin = nchw(plan.t_stem.out)
    a1 = E.frozen_conv(xin, sd, pre + ".conv1", pre + ".bn1")
    stats("l1.0 conv1 (1x1)", nchw(b0.a1), a1)
    a2 = E.frozen_conv(nchw(b0.a1), sd, pre + ".conv2", pre + ".bn2", 1, 1)
    stats("l1.0 conv2 (3x3)", nchw(b0.a2), a2)
    idn = E.frozen_conv(xin, sd, pre + ".downsample.0", pre + ".downsample.1", 1, 0, relu=False)
    stats("l1.0 downsample", nchw(b0.idn), idn)
    out = E.frozen_conv(nchw(b0.a2), sd, pre + ".conv3", pre + ".bn3", relu=True, residual=nchw(b0.idn))
    stats("l1.0 conv3 + res", nchw(b0.out), out)
    # student layer1 units (train): raw, out
    l1 = plan.s_l1
    xs = nchw(plan.s_stem.out)
    e = p + "layer1.encoder.encoder."
    raw = E.r16(torch.nn.functional.conv2d(xs, E.r16(s_sd[e + "0.weight"]), None, 1, 1))
    stats("student e0 raw", nchw(l1.e0.raw), raw)
    a = torch.nn.functional.batch_norm(nchw(l1.e0.raw), None, None, s_sd[e + "1.weight"], s_sd[e + "1.bias"], True, 0.0, 1e-5)
    stats("student e0 bn out", nchw(l1.e0.out), E.r16(a))
    stats("student e0 bn out_g", nchw(l1.e0.out_g), E.rbf(a))

    # chained: the emulation on its OWN tensors all the way (what the tests compare)
    tf, _ = E.backbone_features16(x16, t_sd, student=False)
    for lv in ("layer1", "layer2", "layer3", "layer4"):
        stats("chain teacher " + lv, nchw(plan.feat_t[lv]), tf[lv])
    sf, _ = E.backbone_features16(x16, s_sd, student=True, training=True)
    for lv in ("layer1", "layer2", "layer3", "layer4"):
        stats("chain student " + lv, nchw(plan.feat_s[lv]), sf[lv])
    # layer1 blocks chained from the engine's pooled tensor
    cur = xin
    for b in range(3):
        cur, _ = E.bottleneck16(cur, sd, p + "layer1.%d" % b, 1)
        stats("chain l1.%d out (from engine pool)" % b, nchw(plan.t_layers["layer1"].blocks[b].out), cur)
