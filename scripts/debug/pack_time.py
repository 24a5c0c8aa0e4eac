"""Time of the image pack kernel (normalise + pad + 4-channel interleave [+ bilinear resize]) per image."""
import os
import sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from hnd_ghnd_object_detectors_b200 import ops
from oracle import ghnd_oracle as O

imgs = [torch.rand(3, 800, 1333, device="cuda") for _ in range(4)]
small = [torch.rand(3, 480, 640, device="cuda") for _ in range(4)]
packed = torch.zeros((4, 806, 1352, 4), dtype=torch.float16, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    ts = []
    for _ in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[3]


def plain():
    for i, im in enumerate(imgs):
        ops.stem_pack_image(im, packed, i, 800, 1344, O.IMAGE_MEAN, O.IMAGE_STD)


def resized():
    for i, im in enumerate(small):
        ops.stem_pack_image(ops.ScaledImage(im, 800 / 480), packed, i, 800, 1344, O.IMAGE_MEAN, O.IMAGE_STD)


print("pack 4 x 3x800x1333: %.1f us   pack+resize 4 x 3x480x640 -> 800x1067: %.1f us" % (timed(plain), timed(resized)))
