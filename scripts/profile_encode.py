"""One eager (un-graphed) pass of the split-computing encode path (Keypoint R-CNN b3ch RcnnHead:
stem -> layer1 encoder -> 8-bit quantizer) at batch N between cudaProfilerStart/Stop, for
`ncu --profile-from-start off`.  Usage: python scripts/profile_encode.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hnd_ghnd_object_detectors_b200 import models
from hnd_ghnd_object_detectors_b200.split_rcnn import split_rcnn_model

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
cfg = bench.model_config(True)
cfg["name"] = "keypoint_rcnn"
cfg["params"] = {"num_classes": 2, "pretrained": False, "num_keypoints": 17}
torch.manual_seed(0)
model = models.get_model(cfg, dev).eval()
head, _ = split_rcnn_model(model, 8)
head.use_cuda_graph = False
pool = [torch.rand(3, bench.IMG_H, bench.IMG_W, device=dev) for _ in range(4)]
images = [pool[i % 4] for i in range(batch)]
head(images)
plan = head.plan
plan.graph = None
plan.run()
torch.cuda.synchronize()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
plan.run()
torch.cuda.synchronize()
rt.cudaProfilerStop()
print("profiled one encode pass, batch", batch)
