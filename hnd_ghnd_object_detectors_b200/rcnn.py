"""R-CNN wrappers with the distillation short-circuit -- drop-in mirror of the reference's
src/models/org/rcnn.py (get_model / get_base_backbone / get_fpn_backbone / MODEL_CLASS_DICT) and
src/models/custom/resnet.py (custom_resnet50 with an injectable layer1).

Module containers (ResNet-50 body, FPN, RPN, RoI heads, transform) come from torchvision exactly as
in the reference, so state_dict keys match the released checkpoints.  What changes is execution:
with `distill_backbone_only` (rcnn.py:107-110) or through `backbone_features`, the stem, the student
bottleneck layer, the frozen Bottleneck stacks and the loss run as hand-written sm_100a kernels
(engine.py); FPN/RPN/RoI heads -- never executed in distillation, SURVEY.md section 2 "out of
scope" -- stay torchvision modules fed by those features.
"""
from collections import OrderedDict

import torch
from torch import nn
from torchvision.models import resnet as tv_resnet
from torchvision.models.detection import FasterRCNN as TVFasterRCNN
from torchvision.models.detection import KeypointRCNN as TVKeypointRCNN
from torchvision.models.detection import MaskRCNN as TVMaskRCNN
from torchvision.models.detection.backbone_utils import BackboneWithFPN
from torchvision.models.detection.image_list import ImageList
from torchvision.ops import misc as misc_nn_ops

from . import _lib, ops
from .resnet_layer import get_mimic_layers

MODEL_URL_DICT = {
    'fasterrcnn_resnet50_fpn_coco': 'https://download.pytorch.org/models/fasterrcnn_resnet50_fpn_coco-258fb6c6.pth',
    'maskrcnn_resnet50_fpn_coco': 'https://download.pytorch.org/models/maskrcnn_resnet50_fpn_coco-bf2d0c1e.pth',
    'keypointrcnn_resnet50_fpn_coco': 'https://download.pytorch.org/models/keypointrcnn_resnet50_fpn_coco-fc266e95.pth',
}


def round_up(v, m):
    return (v + m - 1) // m * m


class CustomRCNN(nn.Module):
    """rcnn.py:85-134.  forward(images, targets=None, fixed_sizes=None)."""

    def __init__(self, tv_model):
        super().__init__()
        self.transform = tv_model.transform
        self.backbone = tv_model.backbone
        self.rpn = tv_model.rpn
        self.roi_heads = tv_model.roi_heads
        self.ext_training = False
        self.distill_backbone_only = False
        self.act_dtype = torch.float16
        self._feature_plans = {}

    # ---- CUDA feature path ----------------------------------------------------------------
    def _scaled_images(self, images, fixed_sizes=None):
        """normalize, resize and zero-pad all happen inside the stem pack kernel; this only decides
        each image's scale factor like CustomRCNNTransform.resize (rcnn.py:29-45): identity when the
        image is already at network scale, otherwise an ops.ScaledImage (source + output geometry)
        that ghnd_stem_pack_image_resized resamples bilinearly while packing."""
        import random
        out = []
        tr = self.transform
        for i, img in enumerate(images):
            if img.dim() != 3:
                raise ValueError("images is expected to be a list of 3d tensors of shape [C, H, W], "
                                 "got {}".format(img.shape))
            h, w = img.shape[-2:]
            mn, mx = float(min(h, w)), float(max(h, w))
            if fixed_sizes is not None:
                size = fixed_sizes[i]
            elif self.training:
                size = random.choice(tr.min_size)
            else:
                size = tr.min_size[-1]
            scale = size / mn
            if mx * scale > tr.max_size:
                scale = tr.max_size / mx
            if scale != 1.0:
                img = ops.ScaledImage(img, scale)  # resampled inside the stem pack kernel
            out.append(img)
        return out

    def _run_body(self, images, fixed_sizes=None):
        """Transform + body (stem -> layer1..4) on the CUDA kernels -> (BodyPlan holding the NHWC
        16-bit features, image sizes after resizing, padded batch shape)."""
        from .engine import BodyPlan
        imgs = self._scaled_images(images, fixed_sizes)
        hp = round_up(max(i.shape[1] for i in imgs), 32)
        wp = round_up(max(i.shape[2] for i in imgs), 32)
        # A BodyPlan bakes in layer1's train/eval mode and packs the frozen weights once: key it on the
        # mode and on a version of the backbone's tensors so train()/eval()/load_state_dict() rebuild it.
        body = self.backbone.body
        l1_train = bool(getattr(body.layer1, "training", False))
        version = sum(int(t._version) for t in self.backbone.state_dict(keep_vars=True).values())
        key = (len(imgs), hp, wp, l1_train, version)
        plan = self._feature_plans.get(key)
        if plan is None:
            plan = BodyPlan(body, len(imgs), hp, wp, act_dtype=self.act_dtype,
                            image_mean=self.transform.image_mean, image_std=self.transform.image_std)
            self._feature_plans = {key: plan}
        plan.run(imgs)
        image_sizes = [tuple(i.shape[-2:]) for i in imgs]
        return plan, image_sizes, (len(imgs), 3, hp, wp)

    def backbone_features(self, images, fixed_sizes=None, levels=("layer1", "layer2", "layer3", "layer4")):
        """Forward-only body features (eval-mode BN) as NCHW fp32 tensors, keyed '0'..'3' like
        IntermediateLayerGetter's return_layers (rcnn.py:405)."""
        plan, image_sizes, tshape = self._run_body(images, fixed_sizes)
        feats = plan.feats
        return OrderedDict((str(i), ops.to_nchw_f32(feats[l])) for i, l in enumerate(levels) if l in feats), \
            image_sizes, tshape

    def fpn_features(self, plan):
        """BackboneWithFPN.fpn on the body plan's NHWC features through engine.FpnPlan (tcgen05 convs);
        returns the reference's OrderedDict {'0','1','2','3','pool'} of NCHW fp32 maps for the
        torchvision RPN / RoI heads."""
        from .engine import FpnPlan, LEVELS
        fp = getattr(plan, "fpn_plan", None)
        if fp is None:
            fp = plan.fpn_plan = FpnPlan(self.backbone.fpn, [plan.feats[l] for l in LEVELS], act_dtype=self.act_dtype)
        names = [str(i) for i in range(len(fp.out))]
        results = [ops.to_nchw_f32(t) for t in fp.run()]
        extra = self.backbone.fpn.extra_blocks
        if extra is not None:  # LastLevelMaxPool: a stride-2 subsample of the coarsest map
            results, names = extra(results, [None] * len(results), names)
        return OrderedDict(zip(names, results))

    @staticmethod
    def _resize_targets(targets, original_sizes, new_sizes):
        """What GeneralizedRCNNTransform.resize does to the targets next to the image
        (rcnn.py:46-62): boxes / keypoints scaled by the per-axis ratios, masks resampled."""
        from torchvision.models.detection.transform import resize_boxes, resize_keypoints
        out = []
        for t, (h, w), (nh, nw) in zip(targets, original_sizes, new_sizes):
            if (h, w) == (nh, nw):
                out.append(t)
                continue
            t = dict(t)
            t["boxes"] = resize_boxes(t["boxes"], (h, w), (nh, nw))
            if "keypoints" in t:
                t["keypoints"] = resize_keypoints(t["keypoints"], (h, w), (nh, nw))
            if "masks" in t:
                t["masks"] = torch.nn.functional.interpolate(t["masks"][None].float(), size=(nh, nw))[0].byte()
            out.append(t)
        return out

    def train_ext(self):
        raise _lib.GhndError("training the neural filter (ext_runner.py) is outside the B200 hot path")

    def get_ext_classifier(self):
        """rcnn.py:99-100."""
        layer1 = self.backbone.body.layer1
        return layer1.get_ext_classifier() if hasattr(layer1, "get_ext_classifier") else None

    def forward(self, images, targets=None, fixed_sizes=None):
        if self.training and targets is None:
            raise ValueError("In training mode, targets should be passed")
        if not images[0].is_cuda:
            raise _lib.GhndError("CustomRCNN runs its backbone on CUDA only (no CPU fallback)")
        original_image_sizes = [img.shape[-2:] for img in images]
        plan, image_sizes, tshape = self._run_body(images, fixed_sizes)
        if plan.skipped:  # rcnn.py:115-124: the neural filter found nothing of interest (eval, batch 1)
            _, ch, height, width = tshape
            pred_dict = {'boxes': torch.empty(0, 4), 'labels': torch.empty(0, dtype=torch.int64),
                         'scores': torch.empty(0), 'masks': torch.zeros(100, ch, height, width),
                         'keypoints': torch.empty(0, 17, 3), 'keypoints_scores': torch.empty(0, 17)}
            return [pred_dict]
        if self.distill_backbone_only:
            # rcnn.py:109-110 (the reference also runs a discarded FPN here)
            return OrderedDict((str(i), ops.to_nchw_f32(f)) for i, f in enumerate(plan.feats.values()))
        features = self.fpn_features(plan)
        image_list = ImageList(torch.empty(tshape, device=images[0].device), image_sizes)
        if targets is not None and self.training:
            targets = self._resize_targets(targets, original_image_sizes, image_sizes)
        proposals, proposal_losses = self.rpn(image_list, features, targets)
        detections, detector_losses = self.roi_heads(features, proposals, image_list.image_sizes, targets)
        detections = self.transform.postprocess(detections, image_list.image_sizes, original_image_sizes)
        if self.training:
            loss_dict = dict()
            loss_dict.update(detector_losses)
            loss_dict.update(proposal_losses)
            return loss_dict
        return detections


class FasterRCNN(CustomRCNN):
    def __init__(self, backbone, num_classes=None, min_size=800, max_size=1333, **kwargs):
        super().__init__(TVFasterRCNN(backbone, num_classes, min_size=min_size, max_size=max_size, **kwargs))


class MaskRCNN(CustomRCNN):
    def __init__(self, backbone, num_classes=None, min_size=800, max_size=1333, **kwargs):
        super().__init__(TVMaskRCNN(backbone, num_classes, min_size=min_size, max_size=max_size, **kwargs))


class KeypointRCNN(CustomRCNN):
    def __init__(self, backbone, num_classes=None, min_size=None, max_size=1333, num_keypoints=17, **kwargs):
        if min_size is None:
            min_size = (640, 672, 704, 736, 768, 800)  # rcnn.py:325-326
        super().__init__(TVKeypointRCNN(backbone, num_classes, min_size=min_size, max_size=max_size,
                                        num_keypoints=num_keypoints, **kwargs))


MODEL_CLASS_DICT = {
    'faster_rcnn': (FasterRCNN, 'fasterrcnn_resnet50_fpn_coco'),
    'mask_rcnn': (MaskRCNN, 'maskrcnn_resnet50_fpn_coco'),
    'keypoint_rcnn': (KeypointRCNN, 'keypointrcnn_resnet50_fpn_coco')
}


def custom_resnet50(layer1=None, layer2=None, layer3=None, layer4=None, norm_layer=None, **kwargs):
    """custom/resnet.py:137-144: torchvision ResNet-50 with injectable layers."""
    model = tv_resnet.resnet50(weights=None, norm_layer=norm_layer)
    for name, layer in (("layer1", layer1), ("layer2", layer2), ("layer3", layer3), ("layer4", layer4)):
        if layer is not None:
            setattr(model, name, layer)
    return model


def get_base_backbone(backbone_name, backbone_config, bottleneck_transformer=None):
    """rcnn.py:388-396."""
    pretrained = backbone_config['params']['pretrained']
    if pretrained:
        raise ValueError("ImageNet-pretrained backbones need a download; pass a checkpoint instead "
                         "(set backbone.params.pretrained: False)")
    if backbone_name.startswith('resne') or backbone_name.startswith('wide_resne'):
        if backbone_name != 'resnet50':
            raise ValueError('backbone_name `{}` is not expected'.format(backbone_name))
        return tv_resnet.resnet50(weights=None, norm_layer=misc_nn_ops.FrozenBatchNorm2d)
    elif backbone_name.startswith('custom_resne') or backbone_name.startswith('custom_wide_resne'):
        if backbone_name != 'custom_resnet50':
            raise ValueError('backbone_name `{}` is not expected'.format(backbone_name))
        layer1, layer2, layer3, layer4 = get_mimic_layers(backbone_name, backbone_config, bottleneck_transformer)
        return custom_resnet50(norm_layer=misc_nn_ops.FrozenBatchNorm2d, layer1=layer1, layer2=layer2,
                               layer3=layer3, layer4=layer4)
    raise ValueError('backbone_name `{}` is not expected'.format(backbone_name))


def get_fpn_backbone(backbone, freeze_layers):
    """rcnn.py:399-414."""
    if freeze_layers:
        for name, parameter in backbone.named_parameters():
            if 'layer2' not in name and 'layer3' not in name and 'layer4' not in name:
                parameter.requires_grad_(False)
    return_layers = {'layer1': '0', 'layer2': '1', 'layer3': '2', 'layer4': '3'}
    return BackboneWithFPN(backbone, return_layers, [256, 512, 1024, 2048], 256)


def get_model_config(model_name):
    if model_name in MODEL_CLASS_DICT:
        return MODEL_CLASS_DICT[model_name]
    raise KeyError('model_name `{}` is not expected'.format(model_name))


def get_model(model_name, pretrained, num_classes=91, backbone_config=None, custom_backbone=None,
              strict=True, progress=True, bottleneck_transformer=None, **kwargs):
    """rcnn.py:423-451."""
    backbone_name = backbone_config['name']
    backbone_params_config = backbone_config['params']
    if pretrained:
        backbone_params_config['pretrained'] = False
    if custom_backbone is None:
        base_backbone = get_base_backbone(backbone_name, backbone_config, bottleneck_transformer)
        # ext_config (neural filter, rcnn.py:432-435): the reference swaps in ExtBackboneWithFPN; here the
        # filter hangs on layer1.encoder.ext_classifier (same state_dict keys) and CustomRCNN.forward
        # consults it on the CUDA path
        backbone = get_fpn_backbone(base_backbone, backbone_params_config['freeze_layers'])
    else:
        backbone = custom_backbone
    model_class, pretrained_key = get_model_config(model_name)
    model = model_class(backbone, num_classes, **kwargs)
    if pretrained and backbone_name.endswith('resnet50'):
        print('Loading pretrained state dict of {}'.format(backbone_name))
        if backbone_name != 'resnet50':
            strict = False
        state_dict = torch.hub.load_state_dict_from_url(MODEL_URL_DICT[pretrained_key], progress=progress)
        model.load_state_dict(state_dict, strict=strict)
    return model
