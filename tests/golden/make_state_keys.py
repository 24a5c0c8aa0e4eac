"""Generates tests/golden/state_dict_keys.json from the UNMODIFIED reference models (run in the
authoring container only): state_dict key -> shape for the Faster/Mask/Keypoint R-CNN teacher and
b3ch student, so the drop-in modules can be checked for checkpoint compatibility offline."""
import json
import os
import sys
import warnings

warnings.filterwarnings('ignore')
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader  # noqa: E402

ref_loader.load()
from models.org import rcnn  # noqa: E402

out = {}
for model in ('faster_rcnn', 'mask_rcnn', 'keypoint_rcnn'):
    cfg = ref_loader.load_config('ghnd', model, 3)
    tc, sc = cfg['teacher_model'], cfg['student_model']
    t = rcnn.get_model(tc['name'], backbone_config=tc['backbone'], **tc['params'])
    s = rcnn.get_model(sc['name'], backbone_config=sc['backbone'], **sc['params'])
    out[model] = {'teacher': {k: list(v.shape) for k, v in t.state_dict().items()},
                  'student': {k: list(v.shape) for k, v in s.state_dict().items()}}
json.dump(out, open(os.path.join(HERE, 'state_dict_keys.json'), 'w'))
