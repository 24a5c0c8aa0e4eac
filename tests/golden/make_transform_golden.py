"""Generates tests/golden/transform.npz by running the UNMODIFIED reference's input transform
(CustomRCNNTransform.forward: normalize -> resize -> batch_images, src/models/org/rcnn.py:25-82)
on CPU.  Authoring container only:
    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_transform_golden.py

Cases: ragged small images with per-image `fixed_sizes` (the Keypoint multi-scale path of
DistillationBox, src/distillation/tool.py:44-49), min_size / max_size scaled down so the fixture
stays small; one case where the max-side cap decides the scale; one identity case.
Recorded per case: the padded batch (fp32) and the per-image sizes after resizing.
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

from oracle import ref_loader  # noqa: E402

IMAGE_MEAN = [0.485, 0.456, 0.406]
IMAGE_STD = [0.229, 0.224, 0.225]


def transform_cases():
    """name -> (image shapes, fixed_sizes, max_size, seed); shared with the tests."""
    return {
        "multi_scale": ([(3, 60, 100), (3, 75, 90), (3, 64, 64)], [48, 64, 80], 133, 11),
        "max_side_cap": ([(3, 40, 120), (3, 50, 70)], [64, 64], 133, 12),
        "identity": ([(3, 64, 96), (3, 64, 80)], [64, 64], 133, 13),
        "upsample": ([(3, 30, 44)], [75], 200, 14),
    }


def case_images(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(*s, generator=g) for s in shapes]


def main():
    ref_loader.load()
    from models.org.rcnn import CustomRCNNTransform
    rec = {}
    for name, (shapes, sizes, max_size, seed) in transform_cases().items():
        tr = CustomRCNNTransform(min(sizes), max_size, IMAGE_MEAN, IMAGE_STD)
        tr.eval()
        image_list, _ = tr(case_images(shapes, seed), None, fixed_sizes=sizes)
        rec[name + "/batch"] = image_list.tensors.numpy().astype(np.float32)
        rec[name + "/image_sizes"] = np.array([list(s) for s in image_list.image_sizes], dtype=np.int64)
        print(name, tuple(image_list.tensors.shape), image_list.image_sizes)
    np.savez_compressed(os.path.join(HERE, "transform.npz"), **rec)


if __name__ == "__main__":
    main()
