"""Head / tail split for split computing -- mirror of src/models/mimic/split_rcnn.py
(RcnnHead :13-37, RcnnTail :162-212, split_rcnn_model :215-221).

RcnnHead = transform -> conv1/bn1/relu/maxpool -> layer1.encoder -> Quantizer, executed as the
fixed-shape EncodePlan (engine.py): tcgen05 stem + encoder convs with eval-mode BN folded, then the
fused 8-bit quantizer.  It returns the reference's tuple
(QuantizedTensor, tensors.shape, image_sizes, original_image_sizes).
RcnnTail = Dequantizer -> layer1.decoder -> layer2-4 on the CUDA kernels, then torchvision
FPN / RPN / RoI heads (outside the hot path)."""
from collections import OrderedDict

import torch
from torch import nn
from torchvision.models.detection.image_list import ImageList

from . import _lib, ops, tensor_util
from .engine import EncodePlan, FrozenLayerRunner, StudentLayer1Runner, _empty, LEVELS
from .rcnn import round_up
from .transformer import Compose, Dequantizer, Quantizer


class RcnnHead(nn.Module):
    def __init__(self, rcnn_model, bottleneck_transformer=None):
        super().__init__()
        self.model = rcnn_model
        self.transform = rcnn_model.transform
        self.bottleneck_transformer = bottleneck_transformer
        self.num_bits = None
        if bottleneck_transformer is not None:
            qs = [t for t in bottleneck_transformer.transforms if isinstance(t, Quantizer)]
            if len(qs) != 1 or len(bottleneck_transformer.transforms) != 1:
                raise ValueError("RcnnHead supports a single Quantizer as encoder-side transformer")
            self.num_bits = qs[0].num_bits
        self.plan = None
        self.use_cuda_graph = True

    def forward(self, images, targets=None):
        if not images[0].is_cuda:
            raise _lib.GhndError("RcnnHead runs on CUDA only (no CPU fallback)")
        original_image_sizes = [img.shape[-2:] for img in images]
        was_training = self.model.training
        imgs = self.model._scaled_images(images, None)
        image_sizes = [tuple(i.shape[-2:]) for i in imgs]
        hp = round_up(max(i.shape[1] for i in imgs), 32)
        wp = round_up(max(i.shape[2] for i in imgs), 32)
        key = (len(imgs), hp, wp)
        if self.plan is None or self.plan_key != key:
            if was_training:
                raise _lib.GhndError("RcnnHead is an inference path: call model.eval() first")
            self.plan = EncodePlan(self.model.backbone.body, len(imgs), hp, wp,
                                   num_bits=self.num_bits if self.num_bits else 16,
                                   image_mean=self.transform.image_mean, image_std=self.transform.image_std,
                                   scale_mode=tensor_util._scale_mode())
            self.plan_key = key
            if self.use_cuda_graph and not getattr(self.model.backbone.body.layer1, "uses_ext_encoder", False):
                self.plan.capture()
        layer1 = self.model.backbone.body.layer1
        if getattr(layer1, "uses_ext_encoder", False):
            res, _ext_z = self.plan.run_filtered(imgs, layer1.encoder)
            if res is None:
                # Stop inference since it is decided that there is no object we are interested in
                return None
            q, qp = res
        else:
            q, qp = self.plan.run(imgs)
        tshape = torch.Size((len(imgs), 3, hp, wp))
        if self.num_bits is None:
            return self.plan.z, tshape, image_sizes, original_image_sizes
        if self.num_bits == 16:
            return self.plan.z.half(), tshape, image_sizes, original_image_sizes
        host = qp.cpu()  # python-int zero point of the reference API: one D2H read
        zp = int(host[1])
        tensor_util.check_zero_point(zp)
        qz = tensor_util.QuantizedTensor(tensor=q, scale=qp[0:1].view(torch.float32).reshape(()), zero_point=zp)
        return qz, tshape, image_sizes, original_image_sizes


class RcnnTail(nn.Module):
    def __init__(self, rcnn_model, bottleneck_transformer=None):
        super().__init__()
        self.model = rcnn_model
        self.bottleneck_transformer = bottleneck_transformer
        self._plans = {}

    def _plan(self, z):
        n, c, hz, wz = z.shape
        key = (n, hz, wz)
        if key not in self._plans:
            body = self.model.backbone.body
            h, w = hz - 4, wz - 4
            x = _empty((n, h, w, 64), torch.float16, z.device)  # unused encoder input placeholder
            l1 = StudentLayer1Runner(body.layer1, x, n, h, w, torch.float16, torch.bfloat16, False)
            layers, cur, ch, cw = [], l1.out, h, w
            for name in LEVELS[1:]:
                r = FrozenLayerRunner(getattr(body, name), cur, n, ch, cw, torch.float16, torch.bfloat16, False)
                layers.append(r)
                cur, ch, cw = r.out, r.Ho, r.Wo
            from .engine import FpnPlan
            fpn = FpnPlan(self.model.backbone.fpn, [l1.out] + [r.out for r in layers])
            self._plans = {key: (l1, layers, fpn)}
        return self._plans[key]

    def backbone_features(self, z, targets=None):
        """Server half of split computing up to the body outputs (split_rcnn.py:186-189):
        Dequantizer -> layer1.decoder (eval BN folded) -> layer2-4 on the CUDA kernels; NCHW fp32
        tensors keyed '0'..'3' like IntermediateLayerGetter's return_layers."""
        if self.bottleneck_transformer is not None:
            z, _ = self.bottleneck_transformer(z, targets)
        z = z.float().contiguous()
        l1, layers, _ = self._run_body(z)
        feats = OrderedDict()
        feats['0'] = ops.to_nchw_f32(l1.out)
        for i, r in enumerate(layers):
            feats[str(i + 1)] = ops.to_nchw_f32(r.out)
        return feats

    def _run_body(self, z):
        z = z.float().contiguous()
        plan = self._plan(z)
        plan[0].forward_decoder(z)
        for r in plan[1]:
            r.forward()
        return plan

    def forward(self, z, tensors_shape, image_sizes, original_image_sizes, targets=None):
        """split_rcnn.py:186-212: dequantize -> decoder -> layer2-4 -> FPN on the CUDA kernels, then the
        torchvision RPN / RoI heads."""
        if self.bottleneck_transformer is not None:
            z, _ = self.bottleneck_transformer(z, targets)
        _, _, fpn = self._run_body(z)
        names = [str(i) for i in range(len(fpn.out))]
        results = [ops.to_nchw_f32(t) for t in fpn.run()]
        extra = self.model.backbone.fpn.extra_blocks
        if extra is not None:
            results, names = extra(results, [None] * len(results), names)
        features = OrderedDict(zip(names, results))
        image_list = ImageList(torch.empty(tuple(tensors_shape), device=results[0].device),
                               [tuple(s) for s in image_sizes])
        proposals, proposal_losses = self.model.rpn(image_list, features, targets)
        detections, detector_losses = self.model.roi_heads(features, proposals, image_list.image_sizes, targets)
        detections = self.model.transform.postprocess(detections, image_list.image_sizes, original_image_sizes)
        if self.training:
            loss_dict = dict()
            loss_dict.update(detector_losses)
            loss_dict.update(proposal_losses)
            return loss_dict
        return detections


def split_rcnn_model(model, quantization):
    encoder_transformer = None if quantization is None else Compose([Quantizer(num_bits=quantization)])
    decoder_transformer = None if quantization is None else Compose([Dequantizer(num_bits=quantization)])
    head_model = RcnnHead(model, bottleneck_transformer=encoder_transformer)
    tail_model = RcnnTail(model, bottleneck_transformer=decoder_transformer)
    return head_model, tail_model
