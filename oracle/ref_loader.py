"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference from /root/reference.

Used in the authoring container to (a) validate the oracle restatement in
``oracle/ghnd_oracle.py`` and (b) generate the golden fixtures committed under
``tests/golden/`` (see ``tests/golden/make_golden.py``).  /root/reference does
not exist on the GPU box, so nothing that runs there may import this module.

Recipe follows SURVEY.md Appendix B: two in-memory shims, no file under
/root/reference is touched.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("GHND_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "src"))


def load():
    """Make the reference importable (``import models, distillation, ...``)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    sys.dont_write_bytecode = True
    import torch
    import torchvision
    from torchvision.models import resnet as tv_resnet
    if "torchvision.models.utils" not in sys.modules:
        shim = types.ModuleType("torchvision.models.utils")
        shim.load_state_dict_from_url = torch.hub.load_state_dict_from_url
        sys.modules["torchvision.models.utils"] = shim
        torchvision.models.utils = shim
    if not hasattr(tv_resnet, "model_urls"):
        tv_resnet.model_urls = {}
    src = os.path.join(REF_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)


def load_config(kind="ghnd", model="faster_rcnn", bch=3):
    load()
    from myutils.common import yaml_util
    path = os.path.join(REF_ROOT, "config", kind,
                        "%s-backbone_resnet50-b%dch.yaml" % (model, bch))
    cfg = yaml_util.load_yaml_file(path)
    for k in ("teacher_model", "student_model"):
        cfg[k]["params"]["pretrained"] = False
        cfg[k]["backbone"]["params"]["pretrained"] = False
    return cfg


def build_pair(cfg, seed=0):
    """teacher/student exactly as mimic_runner.main does (src/mimic_runner.py:131-135)."""
    load()
    import torch
    from models.org import rcnn
    from myutils.pytorch import module_util
    torch.manual_seed(seed)
    tc, sc = cfg["teacher_model"], cfg["student_model"]
    teacher = rcnn.get_model(tc["name"], backbone_config=tc["backbone"], **tc["params"])
    student = rcnn.get_model(sc["name"], backbone_config=sc["backbone"], **sc["params"])
    student.load_state_dict(teacher.state_dict(), strict=False)
    module_util.freeze_module_params(teacher)
    for path in sc["frozen_modules"]:
        module_util.freeze_module_params(module_util.get_module(student, path))
    teacher.eval()
    student.train()
    teacher.distill_backbone_only = True
    student.distill_backbone_only = True
    return teacher, student
