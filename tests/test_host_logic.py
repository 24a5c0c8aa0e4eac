"""CPU tests of the host-side mirror of the reference interface (no GPU, no compute kernels):
names, registries, error behaviour, YAML/JSON config handling, state_dict compatibility, parameter
flattening and the world_size-2 gloo data-parallel plumbing."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _yaml(tmp_path, text):
    p = tmp_path / "c.yaml"
    p.write_text(text)
    return str(p)


def test_yaml_join_and_json_override(tmp_path):
    from hnd_ghnd_object_detectors_b200 import main_util, yaml_util
    cfg = yaml_util.load_yaml_file(_yaml(tmp_path, """
dataset:
    name: &n 'coco2017'
    root: &r !join ['./resource/dataset/', *n]
student_model:
    bch: &bch 3
    ckpt: !join ['./ckpt/', *n, '-b', *bch, 'ch.pt']
train:
    batch_size: 4
    optimizer: {type: 'Adam', params: {lr: 0.001}}
"""))
    assert cfg["dataset"]["root"] == "./resource/dataset/coco2017"
    assert cfg["student_model"]["ckpt"] == "./ckpt/coco2017-b3ch.pt"
    main_util.overwrite_config(cfg, json.dumps({"train": {"batch_size": 8, "optimizer": {"params": {"lr": 0.01}}}}))
    assert cfg["train"]["batch_size"] == 8 and cfg["train"]["optimizer"]["params"]["lr"] == 0.01
    assert cfg["train"]["optimizer"]["type"] == "Adam"


@pytest.mark.parametrize("model", ["faster_rcnn", "mask_rcnn", "keypoint_rcnn"])
def test_state_dict_keys_match_reference(model, golden_dir):
    """Released checkpoints must load with strict=True: same keys and shapes as the reference models
    (tests/golden/state_dict_keys.json was dumped from the real reference classes)."""
    from hnd_ghnd_object_detectors_b200 import rcnn
    gold = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))[model]
    params = {"num_classes": 2 if model == "keypoint_rcnn" else 91, "pretrained": False}
    t = rcnn.get_model(model, backbone_config={"name": "resnet50", "params": {"pretrained": False, "freeze_layers": True}}, **params)
    s = rcnn.get_model(model, backbone_config={"name": "custom_resnet50", "params": {
        "pretrained": False, "freeze_layers": False,
        "layer1": {"name": "Bottleneck4LargeResNet", "bottleneck_channel": 3}}}, **params)
    for tag, m in (("teacher", t), ("student", s)):
        got = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert got == gold[tag], (set(got) ^ set(gold[tag]))
    names = [n for n, p in s.backbone.body.named_parameters()]
    assert "layer1.encoder.encoder.7.weight" in names and "layer1.decoder.10.bias" in names


def test_trainable_set_is_the_25_tensors():
    from hnd_ghnd_object_detectors_b200 import module_util, rcnn
    s = rcnn.get_model("faster_rcnn", False, backbone_config={"name": "custom_resnet50", "params": {
        "pretrained": False, "freeze_layers": False,
        "layer1": {"name": "Bottleneck4LargeResNet", "bottleneck_channel": 3}}})
    for path in ["backbone.body.layer2", "backbone.body.layer3", "backbone.body.layer4", "backbone.fpn", "rpn", "roi_heads"]:
        module_util.freeze_module_params(module_util.get_module(s, path))
    names = module_util.get_updatable_param_names(s)
    assert len(names) == 25 and sum(dict(s.named_parameters())[n].numel() for n in names) == 586566
    assert module_util.get_module(s, "backbone.body.nope") is None  # prints and returns None


def test_error_behaviour_matches_reference():
    from hnd_ghnd_object_detectors_b200 import loss, models, rcnn, resnet_layer, transformer
    with pytest.raises(ValueError):
        loss.get_loss({"type": "nope", "params": {"org_loss_factor": 0}, "terms": {}})
    with pytest.raises(KeyError):
        transformer.get_bottleneck_transformer({"order": ["nope"], "components": {"nope": {"params": {}}}})
    with pytest.raises(ValueError):
        resnet_layer.get_mimic_layers("custom_resnet50", {"params": {"layer1": {"name": "X", "bottleneck_channel": 3}}})
    with pytest.raises(ValueError):
        models.get_model({"name": "yolo", "ckpt": "x", "params": {}}, "cpu")
    with pytest.raises(KeyError):
        rcnn.get_model_config("yolo")
    tr = transformer.get_bottleneck_transformer({"order": ["quantizer", "dequantizer"], "components": {
        "quantizer": {"params": {"num_bits": 8}}, "dequantizer": {"params": {"num_bits": 8}}}})
    assert [type(t).__name__ for t in tr.transforms] == ["Quantizer", "Dequantizer"]
    assert transformer.get_bottleneck_transformer({"order": [], "components": {}}) is None


def test_resize_decisions_match_reference_transform(golden_dir):
    """CustomRCNN._scaled_images decides each image's scale factor like CustomRCNNTransform.resize
    (src/models/org/rcnn.py:29-42): fixed_sizes > random.choice(min_size) in training > min_size[-1]
    in eval, capped so the long side stays <= max_size; identity scale returns the tensor itself,
    otherwise an ops.ScaledImage whose geometry equals what the reference transform produced
    (tests/golden/transform.npz, generated from the unmodified reference)."""
    import numpy as np
    import os
    import random
    from hnd_ghnd_object_detectors_b200 import models, ops
    from oracle import ghnd_oracle as O
    from tests.golden.make_transform_golden import case_images, transform_cases
    g = np.load(os.path.join(golden_dir, "transform.npz"))
    for name, (shapes, sizes, max_size, seed) in transform_cases().items():
        cfg = {"name": "keypoint_rcnn", "ckpt": "/nonexistent",
               "backbone": {"name": "resnet50", "params": {"pretrained": False, "freeze_layers": True}},
               "params": {"num_classes": 2, "pretrained": False, "num_keypoints": 17,
                          "min_size": tuple(sorted(set(sizes))), "max_size": max_size}}
        model = models.get_model(cfg, "cpu")
        imgs = case_images(shapes, seed)
        out = model._scaled_images(imgs, fixed_sizes=sizes)
        for im, o, size, (h, w) in zip(imgs, out, sizes, g[name + "/image_sizes"]):
            sc = O.resize_scale(im.shape[1], im.shape[2], size, max_size)
            assert tuple(o.shape[-2:]) == (int(h), int(w)), name
            if sc == 1.0:
                assert o is im
            else:
                assert isinstance(o, ops.ScaledImage) and o.src is im and abs(o.scale - sc) < 1e-12
        # eval: the largest configured scale (rcnn.py:38-40); training: one of the configured scales
        model.eval()
        ev = model._scaled_images(imgs, None)
        want = [O.resize_scale(im.shape[1], im.shape[2], max(sizes), max_size) for im in imgs]
        for o, im, sc in zip(ev, imgs, want):
            assert (o is im) if sc == 1.0 else abs(o.scale - sc) < 1e-12
        model.train()
        random.seed(3)
        tr = model._scaled_images(imgs, None)
        random.seed(3)
        picks = [random.choice(model.transform.min_size) for _ in imgs]
        for o, im, size in zip(tr, imgs, picks):
            sc = O.resize_scale(im.shape[1], im.shape[2], size, max_size)
            assert (o is im) if sc == 1.0 else abs(o.scale - sc) < 1e-12


def test_no_cpu_fallback():
    from hnd_ghnd_object_detectors_b200 import _lib, resnet_layer, tensor_util
    with pytest.raises(_lib.GhndError):
        tensor_util.quantize_tensor(torch.randn(1, 3, 4, 4))
    layer = resnet_layer.Bottleneck4LargeResNet(3, None, None)
    with pytest.raises(_lib.GhndError):
        layer(torch.randn(1, 64, 8, 8))


def test_flat_params_views():
    from hnd_ghnd_object_detectors_b200.engine import FlatParams
    ps = [("a", torch.nn.Parameter(torch.randn(3, 5))), ("b", torch.nn.Parameter(torch.randn(7))),
          ("c", torch.nn.Parameter(torch.randn(2, 2), requires_grad=False))]
    ref = {n: p.detach().clone() for n, p in ps}
    flat = FlatParams(ps)
    assert flat.names == ["a", "b"] and flat.total % 4 == 0
    for n in flat.names:
        assert torch.equal(flat.params[n].data, ref[n])
    flat.flat.add_(1.0)
    assert torch.equal(ps[0][1].data, ref["a"] + 1.0)  # parameters are views of the flat buffer
    flat.grads["b"].fill_(2.0)
    assert float(flat.grad.sum()) == 14.0


def test_shard_indices_cover_all_items():
    from hnd_ghnd_object_detectors_b200.parallel import shard_indices
    for n, world in ((10, 2), (7, 4), (8, 8)):
        seen = set()
        sizes = set()
        for r in range(world):
            idx = shard_indices(n, r, world)
            sizes.add(len(idx))
            seen.update(idx)
        assert seen == set(range(n)) and len(sizes) == 1


def _gloo_worker(rank, world, port, q):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "WORLD_SIZE": str(world)})
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from hnd_ghnd_object_detectors_b200 import parallel
    from hnd_ghnd_object_detectors_b200.engine import FlatParams
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(rank)
    flat = FlatParams([("w", torch.nn.Parameter(torch.randn(5, 3))), ("b", torch.nn.Parameter(torch.randn(6)))])
    parallel.broadcast_flat_params(flat, 0)
    flat.grad.copy_(torch.arange(flat.total, dtype=torch.float32) * (rank + 1))
    parallel.allreduce_flat_grad(flat)
    q.put((rank, flat.flat.clone(), flat.grad.clone(), parallel.world_size()))
    dist.destroy_process_group()


def test_data_parallel_flat_allreduce_gloo():
    """world_size 2 on CPU: after broadcast both ranks hold rank 0's parameters; the flat gradient
    all-reduce is the SUM over ranks (FusedAdam applies 1/world)."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    (r0, p0, g0, w0), (r1, p1, g1, w1) = res
    assert w0 == w1 == 2
    assert torch.equal(p0, p1)
    expect = torch.arange(p0.numel(), dtype=torch.float32) * 3
    assert torch.equal(g0, expect) and torch.equal(g1, expect)


# ------------------------------------------------------------------------------------------------
# data-parallel training semantics (the all-reduced gradient must be the one Adam uses)
# ------------------------------------------------------------------------------------------------
def _torch_adam_step(param, grad, m, v, lr, b1, b2, eps, wd, grad_scale, step):
    """Stand-in for the ghnd_adam_step kernel in CPU tests of the HOST logic (same update rule)."""
    g = grad * grad_scale
    if wd:
        g = g + wd * param
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    param.addcdiv_(m, (v.sqrt() / bc2 ** 0.5).add_(eps), value=-lr / bc1)


class _FakeBox(object):
    """What tool._GradInjector needs from a DistillationBox: the flat buffer the plan wrote into."""

    def __init__(self, flat):
        self.flat = flat


def _train_worker(rank, world, port, q):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank),
                       "WORLD_SIZE": str(world)})
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from hnd_ghnd_object_detectors_b200 import ops, parallel
    from hnd_ghnd_object_detectors_b200.engine import FlatParams
    from hnd_ghnd_object_detectors_b200.optim import FusedAdam
    from hnd_ghnd_object_detectors_b200.tool import _GradInjector
    ops.adam_step = _torch_adam_step
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)  # ranks start from DIFFERENT parameters, like un-seeded processes
    model = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.BatchNorm1d(3))
    model[1].running_mean.fill_(float(rank))
    flat = FlatParams(list(model.named_parameters()))
    parallel.broadcast_flat_params(flat)
    parallel.broadcast_buffers(model)
    opt = FusedAdam([p for p in model.parameters()], lr=1e-2, grad_scale=1.0 / world, flat=flat)
    box = _FakeBox(flat)
    start = flat.flat.clone()
    local = []
    for step in range(3):
        # the fused plan writes this rank's LOCAL gradient into the flat buffer during forward
        flat.grad.copy_(torch.randn(flat.total, generator=torch.Generator().manual_seed(10 * step + rank)))
        local.append(flat.grad.clone())
        loss = _GradInjector.apply(torch.tensor(1.0 + rank), box, *[flat.params[n] for n in flat.names])
        opt.zero_grad()       # mimic_runner.py:51-54 order
        loss.backward()
        for n in flat.names:  # p.grad is a VIEW of the flat gradient buffer
            assert flat.params[n].grad.data_ptr() == flat.grads[n].data_ptr()
        parallel.allreduce_flat_grad(flat)
        opt.step()
    # numpy arrays travel through the queue by value (torch tensors are passed as shared-memory handles that
    # die with this process: the parent then fails with FileNotFoundError if it reads after our exit)
    q.put((rank, start.numpy(), flat.flat.detach().clone().numpy(), model[1].running_mean.clone().numpy(),
           [g.numpy() for g in local]))
    dist.destroy_process_group()


def test_two_rank_training_keeps_replicas_identical():
    """world_size 2, gloo: ranks begin with different random parameters and see different data; after
    broadcast + 3 x (backward -> flat all-reduce -> fused Adam) both hold identical parameters, and
    those equal single-process Adam on the rank-averaged gradient."""
    import queue
    import socket
    import torch.multiprocessing as mp
    res = None
    for attempt in range(3):  # the rendezvous port is picked, released and re-bound: retry a lost race
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
        for p in procs:
            p.start()
        got, waited = [], 0
        while len(got) < 2 and waited < 180:
            try:
                got.append(q.get(timeout=5))
            except queue.Empty:
                waited += 5
                if not any(p.is_alive() for p in procs):  # a worker died without reporting
                    break
        res = sorted(got, key=lambda t: t[0]) if len(got) == 2 else None
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
        if res is not None:
            break
    assert res is not None, "2-rank gloo workers did not report"
    def to_t(r):
        return (r[0], torch.from_numpy(r[1]), torch.from_numpy(r[2]), torch.from_numpy(r[3]),
                [torch.from_numpy(g) for g in r[4]])
    (_, s0, p0, rm0, g0), (_, s1, p1, rm1, g1) = [to_t(r) for r in res]
    assert torch.equal(s0, s1) and torch.equal(rm0, rm1)      # broadcast at start
    assert torch.equal(p0, p1) and not torch.equal(p0, s0)    # replicas stay identical and did move
    ref, m, v = s0.clone(), torch.zeros_like(s0), torch.zeros_like(s0)
    for step in range(3):
        _torch_adam_step(ref, (g0[step] + g1[step]), m, v, 1e-2, 0.9, 0.999, 1e-8, 0, 0.5, step + 1)
    assert torch.allclose(p0, ref, atol=1e-6)


def test_grad_injector_scales_once_and_zero_grad_keeps_new_gradients():
    from hnd_ghnd_object_detectors_b200.engine import FlatParams
    from hnd_ghnd_object_detectors_b200.optim import FusedAdam
    from hnd_ghnd_object_detectors_b200.tool import _GradInjector
    w = torch.nn.Parameter(torch.randn(4, 3))
    flat = FlatParams([("w", w)])
    opt = FusedAdam([w], flat=flat)
    box = _FakeBox(flat)
    flat.grad.fill_(1.5)
    loss = _GradInjector.apply(torch.tensor(2.0), box, w)
    (3.0 * loss).backward()
    assert w.grad.data_ptr() == flat.grads["w"].data_ptr() and torch.all(w.grad == 4.5)
    # next step: the plan rewrites the buffer during forward, THEN the reference loop calls zero_grad
    flat.grad.fill_(2.0)
    loss = _GradInjector.apply(torch.tensor(2.0), box, w)
    opt.zero_grad(set_to_none=False)
    assert torch.all(flat.grad[:12] == 2.0)  # never wiped in place
    loss.backward()
    assert torch.all(w.grad == 2.0)


def test_fused_adam_state_dict_is_torch_adam_format():
    """Checkpoint resume (src/mimic_runner.py:75): moments and step count survive a save/load round
    trip and are interchangeable with torch.optim.Adam."""
    from hnd_ghnd_object_detectors_b200 import ops
    from hnd_ghnd_object_detectors_b200.engine import FlatParams
    from hnd_ghnd_object_detectors_b200.optim import FusedAdam
    saved = ops.adam_step
    ops.adam_step = _torch_adam_step
    try:
        torch.manual_seed(0)
        ps = [torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7))]
        ref_ps = [torch.nn.Parameter(p.detach().clone()) for p in ps]
        flat = FlatParams([("a", ps[0]), ("b", ps[1])])
        opt = FusedAdam(ps, lr=1e-2, flat=flat)
        ref = torch.optim.Adam(ref_ps, lr=1e-2)
        grads = [[torch.randn_like(p) for p in ps] for _ in range(4)]
        for k in range(2):
            for p, r, g in zip(ps, ref_ps, grads[k]):
                flat.grads["a" if p is ps[0] else "b"].copy_(g)
                r.grad = g.clone()
            opt.step()
            ref.step()
        sd = opt.state_dict()
        assert set(sd["state"].keys()) == {0, 1} and int(sd["state"][0]["step"]) == 2
        assert torch.allclose(sd["state"][0]["exp_avg"], ref.state_dict()["state"][0]["exp_avg"], atol=1e-7)
        # resume into a fresh FusedAdam (flat) and into torch.optim.Adam: both continue identically
        ps2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
        flat2 = FlatParams([("a", ps2[0]), ("b", ps2[1])])
        opt2 = FusedAdam(ps2, lr=1e-2, flat=flat2)
        opt2.load_state_dict(sd)
        ps3 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
        opt3 = torch.optim.Adam(ps3, lr=1e-2)
        opt3.load_state_dict({k: v for k, v in sd.items() if k != "fused_step"})
        for k in range(2, 4):
            for i, n in enumerate(("a", "b")):
                flat.grads[n].copy_(grads[k][i])
                flat2.grads[n].copy_(grads[k][i])
                ps3[i].grad = grads[k][i].clone()
            opt.step()
            opt2.step()
            opt3.step()
        for a, b, c in zip(ps, ps2, ps3):
            assert torch.allclose(a, b, atol=1e-7) and torch.allclose(a, c, atol=1e-6)
    finally:
        ops.adam_step = saved


def test_wire_format_sizes():
    """file_util.get_binary_object_size (src/myutils/common/file_util.py:52-53) = KB of the pickle: the
    8-bit QuantizedTensor of a b3ch bottleneck is ~1/4 of the fp32 tensor and ~1/2 of the fp16 one."""
    import pickle
    from hnd_ghnd_object_detectors_b200 import file_util
    from hnd_ghnd_object_detectors_b200.tensor_util import QuantizedTensor
    z = torch.randn(1, 3, 204, 340)
    q = QuantizedTensor(tensor=torch.zeros(1, 3, 204, 340, dtype=torch.uint8), scale=torch.tensor(0.1), zero_point=7)
    s32, s16, s8 = (file_util.get_binary_object_size(x) for x in (z, z.short(), q))
    assert s32 == sys.getsizeof(pickle.dumps(z)) / 1024
    n = z.numel()
    assert 4 * n / 1024 < s32 < 4 * n / 1024 + 2 and 2 * n / 1024 < s16 < 2 * n / 1024 + 2
    assert n / 1024 < s8 < n / 1024 + 2
    assert file_util.get_binary_object_size(z, unit_size=1) == sys.getsizeof(pickle.dumps(z))
