"""Debug aid: compares intermediate gradient buffers of GhndPlan with oracle autograd (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import ghnd_oracle as O, weights
from tests.golden.make_golden import small_images
from tests.test_gpu_distill import build_pair, criterion_config, targets_for, rel, LEVELS
from hnd_ghnd_object_detectors_b200 import models, module_util, ops
from hnd_ghnd_object_detectors_b200.tool import DistillationBox

t_sd, s_sd = weights.teacher_student(3, 0)
env = {"models": models, "module_util": module_util, "t_sd": t_sd, "s_sd": s_sd}
for levels in (LEVELS, ("layer1",)):
    teacher, student = build_pair(env)
    box = DistillationBox(teacher, student, criterion_config(levels), use_cuda_graph=False)
    images = [im.cuda() for im in small_images()]
    loss = box(images, targets_for(images))
    plan = list(box._plans.values())[0]
    # oracle with retained intermediate grads
    x = O.transform_batch(small_images())
    with torch.no_grad():
        tf = O.backbone_features(x, t_sd, False)
    sd = dict(s_sd)
    names = O.trainable_names(s_sd)
    for n in names:
        sd[n] = s_sd[n].clone().requires_grad_(True)
    xs = O.stem_forward(x, sd)
    xs.retain_grad()
    e = "backbone.body.layer1.encoder.encoder."; d = "backbone.body.layer1.decoder."
    a0 = O._bn(F.conv2d(xs, sd[e + "0.weight"], None, 1, 1), sd, e + "1", True); a0.retain_grad()
    a1 = F.relu(O._bn(F.conv2d(a0, sd[e + "2.weight"], None, 1, 1), sd, e + "3", True)); a1.retain_grad()
    a2 = O._bn(F.conv2d(a1, sd[e + "5.weight"], None, 1, 1), sd, e + "6", True); a2.retain_grad()
    z = F.conv2d(a2, sd[e + "7.weight"], None, 1, 1); z.retain_grad()
    zz = F.relu(O._bn(z, sd, d + "0", True))
    r3 = F.conv2d(zz, sd[d + "2.weight"]); r3.retain_grad()
    a3 = O._bn(r3, sd, d + "3", True); a3.retain_grad()
    a4 = F.relu(O._bn(F.conv2d(a3, sd[d + "4.weight"]), sd, d + "5", True)); a4.retain_grad()
    a5 = O._bn(F.conv2d(a4, sd[d + "7.weight"]), sd, d + "8", True); a5.retain_grad()
    r6 = F.conv2d(a5, sd[d + "9.weight"]); r6.retain_grad()
    out1 = F.relu(O._bn(r6, sd, d + "10", True)); out1.retain_grad()
    f = {"layer1": out1}; h = out1
    for name in LEVELS[1:]:
        if name in levels:
            h = O.frozen_layer_forward(h, sd, name); h.retain_grad(); f[name] = h
    l, _ = O.ghnd_loss(tf, f, levels)
    l.backward()
    print("==== levels", levels, "loss", loss.item(), l.item())
    L1 = plan.s_l1
    def cmp(tag, buf, ref, mask=None):
        got = ops.to_nchw_f32(buf) if buf.dim() == 4 and buf.dtype != torch.float32 else buf.float()
        r = ref if mask is None else ref * mask
        print("  %-28s rel %.4f   |ref| %.4g" % (tag, rel(got, r), float(r.norm())))
    cmp("g layer1 out (masked)", L1.d9.g_out, out1.grad, (out1 > 0).float())
    cmp("g_raw6", L1.d9.g_raw, r6.grad)
    cmp("g act5 (d7 out)", L1.g_d7out, a5.grad)
    cmp("g act4 (d4 out)", L1.g_d4out, a4.grad)
    cmp("g act3", L1.g_act3, a3.grad)
    cmp("g raw3", L1.g_raw3, r3.grad)
    cmp("g z", L1.g_z, z.grad)
    cmp("g act2 (e2 out)", L1.g_e2out, a2.grad)
    cmp("g act1 (e1 out)", L1.g_e1out, a1.grad)
    cmp("g act0 (e0 out)", L1.g_e0out, a0.grad)
    cmp("g x (stem out)", L1.g_x, xs.grad)
    for name in LEVELS[1:]:
        if name in plan.s_layers:
            r = plan.s_layers[name]
            for i, b in enumerate(r.blocks):
                pass
    # weights of frozen dgrad chain: compare g at each layer boundary
    for name in reversed(LEVELS[1:]):
        if name in plan.s_layers:
            b0 = plan.s_layers[name].blocks[0]
            gx = b0.bwd[-1]._keep[2]  # dst of the last plan = g_x
            below = LEVELS[LEVELS.index(name) - 1]
            cmp("g %s out (masked)" % below, gx, f[below].grad, (f[below] > 0).float())
    loss.backward()
    params = dict(student.named_parameters())
    for n in names:
        print("  grad %-55s rel %.4f |ref| %.3g" % (n, rel(params[n].grad, sd[n].grad), float(sd[n].grad.norm())))
