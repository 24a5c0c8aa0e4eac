"""Fixed-shape execution plans of the GHND hot path on one B200.

A plan owns the HBM-resident NHWC 16-bit activation / gradient buffers (allocated once through
torch, the allocator is plumbing) and the C-ABI launch plans bound to them, so that one
distillation step is a fixed sequence of hand-written kernels that can be captured in a CUDA graph:

  teacher : stem -> layer1..4 (frozen Bottlenecks, FrozenBN folded)            forward
  student : stem -> layer1 (encoder -> bch bottleneck -> decoder, batch-stat BN) -> layer2..4
  loss    : fused multi-level SSE forward + dL/ds
  backward: layer4..2 dgrad -> layer1 BN-bwd / dgrad / wgrad -> pool+ReLU bwd -> conv1 wgrad

Numeric types: forward tensors and weights fp16, gradient tensors bf16, all accumulation fp32
(BN statistics fp64).  tcgen05 kind::f16 needs both MMA operands in ONE 16-bit format, so the
student's layer1 activations are additionally kept as bf16 copies for the weight-gradient GEMMs.

Reference mapping: src/distillation/tool.py:40-61 (step), src/models/org/rcnn.py:102-110
(distill_backbone_only short-circuit), src/models/mimic/resnet_layer.py:40-70 (student layer1),
torchvision Bottleneck / FrozenBatchNorm2d (frozen layers), src/mimic_runner.py:48-59 (loop).
"""
import torch

from . import _lib, ops
from ._lib import CONV_DGRAD, CONV_FWD

import os

# GHND_FUSE_BN_REDUCE=1: take the BN-backward reductions inside the producing dgrad's epilogue
# (ConvPlan stats_mode 1) instead of a separate bn_bwd_reduce pass.  Off by default: an A/B on the
# B200 gave the same step time (655.5 vs 655.6 img/s) -- the small streaming reduce kernels already
# run concurrently with the persistent tensor-core kernels of the second stream -- and the fused
# form divides by gamma (undefined for a channel whose gamma is exactly 0).
FUSE_BN_REDUCE = os.environ.get("GHND_FUSE_BN_REDUCE", "0") == "1"
# A/B switch: the bottleneck-side dW launches on the side stream (default) or in the data-gradient chain
# A/B switch: the student's weight repacking as ONE launch on the MAIN stream in front of the stem kernel (default).
# As twelve launches on the side stream it started only when the stem kernel (every SM, all shared memory) let go, and
# the first layer1 conv, which waits for it, started 37 us after the stem had finished (in-graph timeline).
PREPACK_BATCHED = os.environ.get("GHND_PREPACK_BATCHED", "1") != "0"
# A/B switch: one memset per step over all BatchNorm reduction buffers (SumsPool) instead of one memset node per use
SUMS_POOL = os.environ.get("GHND_SUMS_POOL", "1") != "0"
NARROW_DW_SIDE = int(os.environ.get("GHND_NARROW_DW_SIDE", "1"))  # 0 chain, 1 side stream at once, 2 side stream, deferred

LEVELS = ("layer1", "layer2", "layer3", "layer4")
PLANES = {"layer1": 64, "layer2": 128, "layer3": 256, "layer4": 512}
IMAGE_MEAN = (0.485, 0.456, 0.406)
IMAGE_STD = (0.229, 0.224, 0.225)


def _empty(shape, dtype, device):
    return torch.empty(shape, dtype=dtype, device=device)


class SideStream(object):
    """A second CUDA stream for the independent branches of a step (teacher forward next to the
    student's stem/layer1 forward; weight repacking; the weight-gradient GEMMs next to the data-
    gradient chain).  Every conv kernel is a persistent one-CTA-per-SM grid, so two of them never
    share an SM -- what the second stream buys is the tail: when the CTAs of one kernel run out of
    tiles, CTAs of the independent kernel start on the freed SMs instead of idling until the slowest
    CTA is done.  Inside CUDA-graph capture the fork/join events become parallel graph branches.
    Only allocation-free work may run here (all buffers are preallocated by the plans)."""

    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.dirty = False

    def fork(self):
        """Everything enqueued on the current stream so far happens-before later side work."""
        self.stream.wait_stream(torch.cuda.current_stream())

    def run(self, fn):
        with torch.cuda.stream(self.stream):
            fn()
        self.dirty = True

    def mark(self):
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return ev

    def join(self):
        if self.dirty:
            torch.cuda.current_stream().wait_stream(self.stream)
            self.dirty = False


def fold_frozen_bn(bn):
    """FrozenBatchNorm2d -> per-channel (scale, shift) fp32 (tiny host-side tensor math)."""
    eps = float(getattr(bn, "eps", 1e-5))
    w, b = bn.weight.detach().float(), bn.bias.detach().float()
    rm, rv = bn.running_mean.detach().float(), bn.running_var.detach().float()
    scale = w * torch.rsqrt(rv + eps)
    return scale.contiguous(), (b - rm * scale).contiguous()


class FrozenConv(object):
    """Packed weights of one frozen conv + FrozenBN: forward [K][R][S][C] (act dtype) and, when a
    data gradient is needed, transposed [C][R][S][K] (grad dtype); bias = BN shift."""

    def __init__(self, conv, bn, act_dtype, grad_dtype, need_dgrad):
        scale, shift = fold_frozen_bn(bn)
        self.K, self.C, self.R, self.S = conv.weight.shape
        self.stride = conv.stride[0]
        self.pad = conv.padding[0]
        self.w = ops.pack_weight(conv.weight, scale, False, act_dtype)
        self.wt = ops.pack_weight(conv.weight, scale, True, grad_dtype) if need_dgrad else None
        self.bias = shift


class BottleneckRunner(object):
    """One torchvision Bottleneck with FrozenBN: forward, and dgrad when need_bwd."""

    def __init__(self, block, x, N, H, W, act_dtype, grad_dtype, need_bwd, out=None, bwd_n=None):
        dev = x.device
        self.x, self.N, self.H, self.W = x, N, H, W
        self.need_bwd = need_bwd
        # the backward pass may cover only the LAST bwd_n images of the batch (shared frozen trunk:
        # teacher images first, student images last; only the student half needs gradients)
        self.bwd_n = N if bwd_n is None else bwd_n
        self.c1 = FrozenConv(block.conv1, block.bn1, act_dtype, grad_dtype, need_bwd)
        self.c2 = FrozenConv(block.conv2, block.bn2, act_dtype, grad_dtype, need_bwd)
        self.c3 = FrozenConv(block.conv3, block.bn3, act_dtype, grad_dtype, need_bwd)
        self.cd = None
        if block.downsample is not None:
            self.cd = FrozenConv(block.downsample[0], block.downsample[1], act_dtype, grad_dtype, need_bwd)
        cin, p = self.c1.C, self.c1.K
        s = self.c2.stride
        self.stride = s
        self.Ho, self.Wo = (H + 2 - 3) // s + 1, (W + 2 - 3) // s + 1
        self.cin, self.planes, self.cout = cin, p, self.c3.K
        self.a1 = _empty((N, H, W, p), act_dtype, dev)
        self.a2 = _empty((N, self.Ho, self.Wo, p), act_dtype, dev)
        self.out = out if out is not None else _empty((N, self.Ho, self.Wo, self.cout), act_dtype, dev)
        assert tuple(self.out.shape) == (N, self.Ho, self.Wo, self.cout)
        self.idn = _empty((N, self.Ho, self.Wo, self.cout), act_dtype, dev) if self.cd else None
        self.fwd = [
            ops.ConvPlan(CONV_FWD, N, H, W, cin, p, 1, 1, 1, 0, x, self.c1.w, self.a1, bias=self.c1.bias,
                         relu=True),
            ops.ConvPlan(CONV_FWD, N, H, W, p, p, 3, 3, s, 1, self.a1, self.c2.w, self.a2,
                         bias=self.c2.bias, relu=True),
        ]
        if self.cd:
            self.fwd.append(ops.ConvPlan(CONV_FWD, N, H, W, cin, self.cout, 1, 1, s, 0, x, self.cd.w,
                                         self.idn, bias=self.cd.bias))
        self.fwd.append(ops.ConvPlan(CONV_FWD, N, self.Ho, self.Wo, p, self.cout, 1, 1, 1, 0, self.a2,
                                     self.c3.w, self.out, bias=self.c3.bias,
                                     residual=self.idn if self.cd else x, relu=True))
        self.bwd = []

    def plan_backward(self, g_out, g_x, loss_grad_x, grad_dtype):
        """g_out: gradient w.r.t. this block's pre-ReLU output (already masked).  Writes g_x =
        mask(x>0) * (dL/dx [+ loss_grad_x]) i.e. the same convention for the producer of x.
        All gradient tensors have batch bwd_n; the saved activations are sliced to those images."""
        dev = g_out.device
        self.g_out, self.g_x = g_out, g_x  # kept for inspection (tests / debugging)
        N, H, W, Ho, Wo = self.bwd_n, self.H, self.W, self.Ho, self.Wo
        p, cin, cout, s = self.planes, self.cin, self.cout, self.stride
        o = self.N - N
        x_b, a1_b, a2_b = self.x[o:], self.a1[o:], self.a2[o:]
        self.g_a2 = _empty((N, Ho, Wo, p), grad_dtype, dev)
        self.g_a1 = _empty((N, H, W, p), grad_dtype, dev)
        self.bwd = [
            ops.ConvPlan(CONV_DGRAD, N, Ho, Wo, p, cout, 1, 1, 1, 0, g_out, self.c3.wt, self.g_a2,
                         mask=a2_b),
            ops.ConvPlan(CONV_DGRAD, N, H, W, p, p, 3, 3, s, 1, self.g_a2, self.c2.wt, self.g_a1,
                         mask=a1_b),
        ]
        if self.cd is None:
            assert loss_grad_x is None
            self.bwd.append(ops.ConvPlan(CONV_DGRAD, N, H, W, cin, p, 1, 1, 1, 0, self.g_a1, self.c1.wt,
                                         g_x, residual=g_out, mask=x_b))
        else:
            # the stride-2 downsample dgrad reaches only the even lattice of g_x, so it has to be the
            # accumulating launch and stays behind the main branch
            self.bwd.append(ops.ConvPlan(CONV_DGRAD, N, H, W, cin, p, 1, 1, 1, 0, self.g_a1, self.c1.wt,
                                         g_x, residual=loss_grad_x, mask=x_b))
            self.bwd.append(ops.ConvPlan(CONV_DGRAD, N, H, W, cin, cout, 1, 1, s, 0, g_out, self.cd.wt,
                                         g_x, mask=x_b, accumulate=True))

    def forward(self, side=None):
        if self.cd is not None and side is not None:
            # [conv1, conv2, downsample, conv3]: the downsample conv only needs x
            side.fork()
            side.run(self.fwd[2].run)
            self.fwd[0].run()
            self.fwd[1].run()
            side.join()
            self.fwd[3].run()
            return
        for p in self.fwd:
            p.run()

    def backward(self, side=None):
        for p in self.bwd:
            n = p.n_launches
            if side is not None and n > 1 and p is self.bwd[1] and os.environ.get("GHND_S2_SIDE", "0") == "1":
                # stride-2 3x3 dgrad = one small launch per output parity class, disjoint outputs:
                # half of them on the second stream
                side.fork()
                side.run(lambda: p.run_range(n // 2, n - n // 2))
                p.run_range(0, n // 2)
                side.join()
            else:
                p.run()


class FrozenLayerRunner(object):
    """nn.Sequential of Bottlenecks (teacher layer1..4, student layer2..4).  `out`: preallocated
    buffer for the last block's output; `bwd_n`: backward covers only the last bwd_n images."""

    def __init__(self, layer, x, N, H, W, act_dtype, grad_dtype, need_bwd, out=None, bwd_n=None):
        self.blocks = []
        n_blocks = len(layer)
        self.bwd_n = N if bwd_n is None else bwd_n
        for i, block in enumerate(layer):
            r = BottleneckRunner(block, x, N, H, W, act_dtype, grad_dtype, need_bwd,
                                 out=out if i == n_blocks - 1 else None, bwd_n=bwd_n)
            self.blocks.append(r)
            x, H, W = r.out, r.Ho, r.Wo
        self.out, self.Ho, self.Wo = x, H, W

    def forward(self, side=None):
        for b in self.blocks:
            b.forward(side)

    def plan_backward(self, g_out, g_x, loss_grad_x, grad_dtype):
        """g_out: masked gradient at this layer's output; g_x: buffer for the layer input's."""
        dev = g_out.device
        g = g_out
        for i in range(len(self.blocks) - 1, -1, -1):
            b = self.blocks[i]
            if i == 0:
                b.plan_backward(g, g_x, loss_grad_x, grad_dtype)
            else:
                gx = _empty((b.bwd_n, b.H, b.W, b.cin), grad_dtype, dev)
                b.plan_backward(g, gx, None, grad_dtype)
                g = gx

    def backward(self, side=None):
        for b in reversed(self.blocks):
            b.backward(side)


def _stem_pool_fused():
    """conv1 and the max-pool as one kernel (the conv output never reaches HBM); GHND_STEM_POOL=0 keeps
    the two-kernel path (conv GEMM, then pool) for A/B runs and as the bitwise reference in the tests."""
    return os.environ.get("GHND_STEM_POOL", "1") != "0"


class DualStemConv(object):
    """The teacher's and the student's conv1 over the shared packed image (teacher = channels 0-63,
    student = 64-127 of the packed weights).  Fused mode: ONE stem+pool kernel writes both pooled maps
    (and the student's argmax codes).  Otherwise one K=128 GEMM writes both conv outputs and the two
    StemRunners pool their halves."""

    def __init__(self, teacher_body, student_body, packed, N, Hp, Wp, act_dtype):
        dev = packed.device
        self.student_body = student_body
        t_scale, t_shift = fold_frozen_bn(teacher_body.bn1)
        self.s_scale, s_shift = fold_frozen_bn(student_body.bn1)
        self.w = _empty((128, 7, 32), act_dtype, dev)
        self.bias = torch.cat([t_shift, s_shift]).contiguous()
        self.fused = _stem_pool_fused()
        if self.fused:
            Ho, Wo = (Hp // 2 + 1) // 2, (Wp // 2 + 1) // 2
            self.conv = None
            self.outs = [_empty((N, Ho, Wo, 64), act_dtype, dev) for _ in range(2)]
            self.argmax = [None, _empty((N, Ho, Wo, 64), torch.uint8, dev)]
            self.plan = ops.StemPoolPlan(packed, self.w, self.bias, self.outs, self.argmax, N, Hp, Wp)
        else:
            self.conv = _empty((N, Hp // 2, Wp // 2, 128), act_dtype, dev)
            self.plan = ops.StemPlan(packed, self.w, self.bias, self.conv, N, Hp, Wp)
        ops.stem_pack_weight(teacher_body.conv1.weight, t_scale, out=self.w[:64])  # frozen: once
        self.refresh_student()

    def refresh_student(self):
        ops.stem_pack_weight(self.student_body.conv1.weight, self.s_scale, out=self.w[64:])

    def forward(self):
        self.refresh_student()  # conv1.weight is trainable
        self.plan.run()


class StemRunner(object):
    """conv1 7x7 s2 + FrozenBN + ReLU -> MaxPool 3x3 s2 (tcgen05 implicit GEMM; one fused kernel unless
    GHND_STEM_POOL=0).  shared = (DualStemConv, channel offset): the conv (fused mode: and the pool) is
    done by the two-stem kernel and this runner only owns the backward pass of its 64-channel half."""

    def __init__(self, body, packed, N, Hp, Wp, act_dtype, grad_dtype, trainable, shared=None):
        dev = packed.device
        self.body, self.packed = body, packed
        self.N, self.Hp, self.Wp = N, Hp, Wp
        self.trainable = trainable
        self.shared = shared
        self.scale, self.shift = fold_frozen_bn(body.bn1)
        self.Ho, self.Wo = (Hp // 2 + 1) // 2, (Wp // 2 + 1) // 2
        self.fused = shared[0].fused if shared is not None else _stem_pool_fused()
        if shared is not None and self.fused:
            self.out, self.argmax = shared[0].outs[shared[1] // 64], shared[0].argmax[shared[1] // 64]
            assert (self.argmax is not None) == bool(trainable)
        else:
            self.out = _empty((N, self.Ho, self.Wo, 64), act_dtype, dev)
            self.argmax = _empty((N, self.Ho, self.Wo, 64), torch.uint8, dev) if trainable else None
        if shared is not None:
            self.conv, self.c_off, self.plan, self.w = shared[0].conv, shared[1], None, None
        else:
            self.c_off = 0
            self.w = _empty((64, 7, 32), act_dtype, dev)
            if self.fused:
                self.conv = None
                self.plan = ops.StemPoolPlan(packed, self.w, self.shift, [self.out], [self.argmax], N, Hp, Wp)
            else:
                self.conv = _empty((N, Hp // 2, Wp // 2, 64), act_dtype, dev)
                self.plan = ops.StemPlan(packed, self.w, self.shift, self.conv, N, Hp, Wp)
            self.refresh_weights()
        if trainable:
            self.g_conv = _empty((N, Hp // 2, Wp // 2, 64), grad_dtype, dev)
            self.ws = _empty((_lib.load().ghnd_stem_wgrad_workspace_bytes(),), torch.uint8, dev)
            # the tensor-core dW needs image and gradient in ONE 16-bit format: keep a copy of the
            # packed image in the gradient dtype (refreshed every step, 2 x 35 MB of traffic)
            self.packed_g = packed if packed.dtype == grad_dtype else torch.zeros_like(packed, dtype=grad_dtype)
            self.wgrad = None

    def refresh_weights(self):
        if self.shared is None:
            ops.stem_pack_weight(self.body.conv1.weight, self.scale, out=self.w)

    def forward(self):
        if self.shared is None:
            if self.trainable:
                self.refresh_weights()
            self.plan.run()
        if not self.fused:
            ops.maxpool3x3s2(self.conv, self.out, self.argmax, channels=64, channel_offset=self.c_off)

    def convert_image(self):
        """The packed image in the gradient format (operand of the tensor-core dW).  Depends on the image only:
        the plan runs it on the side stream at the start of the step instead of in the backward tail."""
        if self.trainable and self.packed_g is not self.packed:
            ops.convert16(self.packed, self.packed_g)
        self.image_converted = True

    def backward(self, g_out, dw):
        """g_out: gradient w.r.t. the pooled output; dw: fp32 OIHW view for conv1.weight.grad."""
        ops.maxpool3x3s2_bwd(self.conv, self.argmax, g_out, self.g_conv, channel_offset=self.c_off)
        if self.packed_g is not self.packed and not getattr(self, "image_converted", False):
            ops.convert16(self.packed, self.packed_g)
        self.image_converted = False
        if self.wgrad is None or self.wgrad_dw is not dw:
            self.wgrad = ops.StemWgradPlan(self.packed_g, self.g_conv, self.scale, dw, self.N, self.Hp, self.Wp,
                                           ws=self.ws)
            self.wgrad_dw = dw
        self.wgrad.run()


class SumsPool(object):
    """One fp64 allocation for every per-step BatchNorm reduction buffer of a plan (forward statistics and backward
    sums of each BN separately), zeroed by ONE memset at the start of the step.  The reductions then only accumulate
    (GHND_SUMS_ZEROED) instead of each heading its kernel with a memset node of its own: inside the CUDA graph a
    kernel -> memset -> kernel hop costs ~9 us against ~2 us for kernel -> kernel (in-graph timeline), and the
    student's layer1 has fifteen of them on its critical chain."""
    current = None  # the pool new _BN objects draw from (set by GhndPlan while it builds the student's layer1)

    def __init__(self, device, n_doubles=16384):
        self.buf = torch.zeros(n_doubles, dtype=torch.float64, device=device)
        self.used = 0

    def take(self, n):
        n = (n + 15) // 16 * 16
        assert self.used + n <= self.buf.numel(), "SumsPool exhausted"
        v = self.buf[self.used:self.used + n]
        self.used += n
        return v

    def zero(self):
        if self.used:
            self.buf[:self.used].zero_()


class _BN(object):
    """Training-mode nn.BatchNorm2d state for one NHWC tensor (or the planar bottleneck)."""

    def __init__(self, bn, C, count, dev):
        self.bn, self.C, self.count = bn, C, count
        pool = SumsPool.current
        self.pooled = pool is not None  # forward / backward sums are zeroed by the plan, once per step
        if self.pooled:
            self.sums, self.sums_b = pool.take(2 * C)[:2 * C], pool.take(2 * C)[:2 * C]
        else:
            self.sums = self.sums_b = _empty((2 * C,), torch.float64, dev)
        self.scale_shift = _empty((2 * C,), torch.float32, dev)
        self.mean_invstd = _empty((2 * C,), torch.float32, dev)

    def finalize(self):
        bn = self.bn
        ops.bn_finalize(self.sums, self.count, self.C, bn.weight, bn.bias, bn.eps,
                        0.1 if bn.momentum is None else bn.momentum, bn.running_mean, bn.running_var,
                        bn.num_batches_tracked, self.scale_shift, self.mean_invstd)

    def finalize_apply(self, x, y, relu, y2=None):
        """finalize() fused into the normalising pass over x (one launch)."""
        bn = self.bn
        ops.bn_finalize_apply(x, y, relu, self.sums, self.count, bn.weight, bn.bias, bn.eps,
                              0.1 if bn.momentum is None else bn.momentum, bn.running_mean, bn.running_var,
                              bn.num_batches_tracked, self.scale_shift, self.mean_invstd, y2=y2)

    def eval_params(self):
        bn = self.bn
        ops.bn_eval_params(self.C, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps,
                           self.scale_shift)


class _WideUnit(object):
    """conv(k2) -> BatchNorm(batch stats) [-> ReLU] of the student's layer1, forward + backward."""

    def __init__(self, conv, bn, relu, x, x_g, N, H, W, act_dtype, grad_dtype, train, out=None,
                 need_out_g=True):
        dev = x.device
        K, C, R, S = conv.weight.shape
        pad = conv.padding[0]
        self.conv, self.relu, self.x, self.x_g = conv, relu, x, x_g
        self.prepacked = False
        self.reduce_fused = False
        self.N, self.H, self.W, self.C, self.K, self.pad = N, H, W, C, K, pad
        self.Ho, self.Wo = H + 2 * pad - R + 1, W + 2 * pad - S + 1
        self.train = train
        self.w = _empty((K, R, S, C), act_dtype, dev)
        self.bn = _BN(bn, K, N * self.Ho * self.Wo, dev)
        self.out = out if out is not None else _empty((N, self.Ho, self.Wo, K), act_dtype, dev)
        assert tuple(self.out.shape) == (N, self.Ho, self.Wo, K)
        if train:
            self.raw = _empty((N, self.Ho, self.Wo, K), act_dtype, dev)
            # bf16 copy of the activation for the NEXT unit's weight-gradient GEMM (skipped when no
            # tensor-core dW reads it: the layer's last unit, and e2 whose consumer is the narrow dW)
            self.out_g = _empty((N, self.Ho, self.Wo, K), grad_dtype, dev) if need_out_g else None
            # batch statistics of the stored conv output are accumulated by the conv epilogue
            self.plan = ops.ConvPlan(CONV_FWD, N, H, W, C, K, R, S, 1, pad, x, self.w, self.raw,
                                     stats=self.bn.sums, stats_zeroed=self.bn.pooled)
        else:
            # eval: BN folded into the conv (scale into the weights, shift as bias, ReLU in epilogue)
            self.raw = self.out_g = None
            self.plan = ops.ConvPlan(CONV_FWD, N, H, W, C, K, R, S, 1, pad, x, self.w, self.out,
                                     bias=self.bn.scale_shift[K:], relu=relu)

    def prepack(self):
        """Repack this step's weights (forward [K][R][S][C] and transposed for the dgrad): depends on
        the parameters only, so the plan runs it off the critical path at the start of the step."""
        if not self.train:
            return  # eval folds the BN scale into the weights inside forward()
        ops.pack_weight(self.conv.weight, None, False, out=self.w)
        if getattr(self, "wt", None) is not None:
            ops.pack_weight(self.conv.weight, None, True, out=self.wt)
        self.prepacked = True

    def forward(self):
        if self.train:
            if not self.prepacked:
                ops.pack_weight(self.conv.weight, None, False, out=self.w)
            self.plan.run()
            self.bn.finalize_apply(self.raw, self.out, self.relu, y2=self.out_g)
        else:
            self.bn.eval_params()
            ops.pack_weight(self.conv.weight, self.bn.scale_shift[:self.K], False, out=self.w)
            self.plan.run()

    def plan_backward(self, g_out, g_x, grads, names, grad_dtype, below=None):
        """g_out: gradient w.r.t. self.out; g_x: gradient buffer for the input (None: not needed);
        grads: dict name -> fp32 grad view; names = (conv.weight, bn.weight, bn.bias).
        below = (activation, relu, sums): the BatchNorm whose output is this unit's input.  Its
        backward reductions (sum g', sum g'*activation) are then taken in this unit's dgrad epilogue
        (the activation tile doubles as the ReLU mask), and `below` skips its bn_bwd_reduce pass."""
        dev = g_out.device
        N, H, W, C, K = self.N, self.H, self.W, self.C, self.K
        R, S = self.conv.weight.shape[2:]
        self.g_out, self.g_x = g_out, g_x
        self.g_raw = _empty((N, self.Ho, self.Wo, K), grad_dtype, dev)
        self.dw_packed = _empty((K, R, S, C), torch.float32, dev)
        self.dw, self.dgamma, self.dbeta = grads[names[0]], grads[names[1]], grads[names[2]]
        self.wgrad = ops.WgradPlan(N, H, W, C, K, R, S, self.pad, self.x_g, self.g_raw, self.dw_packed)
        self.wt = None
        self.dgrad = None
        if g_x is not None:
            self.wt = _empty((C, R, S, K), grad_dtype, dev)
            if below is not None and FUSE_BN_REDUCE:
                act, relu, sums, pooled = below
                self.dgrad = ops.ConvPlan(CONV_DGRAD, N, H, W, C, K, R, S, 1, self.pad, self.g_raw, self.wt,
                                          g_x, mask=act, stats=sums, stats_mode=1, mask_stats_only=not relu,
                                          stats_zeroed=pooled)
            else:
                self.dgrad = ops.ConvPlan(CONV_DGRAD, N, H, W, C, K, R, S, 1, self.pad, self.g_raw, self.wt, g_x)

    def _wgrad(self):
        self.wgrad.run()
        ops.unpack_wgrad(self.dw_packed, self.dw)

    def backward(self, side=None):
        bn = self.bn
        if not self.reduce_fused:  # else: the dgrad that produced g_out already left the sums
            ops.bn_bwd_reduce(self.g_out, self.raw, bn.scale_shift, bn.mean_invstd, self.relu, bn.sums_b,
                              zeroed=bn.pooled)
        ops.bn_bwd_apply(self.g_out, self.raw, self.g_raw, bn.bn.weight, bn.scale_shift, bn.mean_invstd,
                         self.relu, bn.sums_b, self.dgamma, self.dbeta, fused_sums=self.reduce_fused)
        if side is not None:  # dW next to the dgrad chain (reads g_raw / x_g only)
            side.fork()
            side.run(self._wgrad)
        else:
            self._wgrad()
        if self.dgrad is not None:
            if not self.prepacked:
                ops.pack_weight(self.conv.weight, None, True, out=self.wt)
            self.dgrad.run()
        self.prepacked = False


class StudentLayer1Runner(object):
    """Bottleneck4LargeResNet (resnet_layer.py:40-70): three wide encoder convs, the narrow
    64->bch / bch->64 pair around the planar fp32 bottleneck z, three wide decoder convs."""

    def __init__(self, layer1, x, N, H, W, act_dtype, grad_dtype, train, x_g=None, out=None):
        dev = x.device
        enc, dec = layer1.encoder.encoder, layer1.decoder
        self.layer1, self.train = layer1, train
        self.N, self.H, self.W = N, H, W
        self.act_dtype, self.grad_dtype = act_dtype, grad_dtype
        self.x = x
        if train and x_g is None:
            x_g = _empty(tuple(x.shape), grad_dtype, dev)
            self._own_xg = True
        else:
            self._own_xg = False
        self.x_g = x_g
        mk = lambda conv, bn, relu, xin, xin_g, h, w, og=True: _WideUnit(conv, bn, relu, xin, xin_g, N, h, w,
                                                                       act_dtype, grad_dtype, train,
                                                                       need_out_g=og)
        self.e0 = mk(enc[0], enc[1], False, x, x_g, H, W)
        self.e1 = mk(enc[2], enc[3], True, self.e0.out, self.e0.out_g, self.e0.Ho, self.e0.Wo)
        self.e2 = mk(enc[5], enc[6], False, self.e1.out, self.e1.out_g, self.e1.Ho, self.e1.Wo, False)
        self.enc7, self.bn0, self.dec2 = enc[7], dec[0], dec[2]
        self.bch = enc[7].weight.shape[0]
        self.Hz, self.Wz = self.e2.Ho + 1, self.e2.Wo + 1
        self.z = _empty((N, self.bch, self.Hz, self.Wz), torch.float32, dev)
        self.bnz = _BN(dec[0], self.bch, N * self.Hz * self.Wz, dev)
        self.H3, self.W3 = self.Hz - 1, self.Wz - 1
        self.nws = ops._narrow_ws(self.bch, 64, 2, 2, dev)
        # dec2 output + its BN (dec[3], no ReLU)
        self.raw3 = _empty((N, self.H3, self.W3, 64), act_dtype, dev)
        self.act3 = _empty((N, self.H3, self.W3, 64), act_dtype, dev)
        self.act3_g = _empty((N, self.H3, self.W3, 64), grad_dtype, dev) if train else None
        self.bn3 = _BN(dec[3], 64, N * self.H3 * self.W3, dev)
        self.d4 = mk(dec[4], dec[5], True, self.act3, self.act3_g, self.H3, self.W3)
        self.d7 = mk(dec[7], dec[8], False, self.d4.out, self.d4.out_g, self.d4.Ho, self.d4.Wo)
        self.d9 = _WideUnit(dec[9], dec[10], True, self.d7.out, self.d7.out_g, N, self.d7.Ho, self.d7.Wo,
                            act_dtype, grad_dtype, train, out=out, need_out_g=False)
        self.out = self.d9.out
        assert (self.d9.Ho, self.d9.Wo) == (H, W)
        self.q = None  # set by forward_encoder_quantized

    # ---- forward ----
    def forward_encoder(self, minmax=None):
        """minmax: fp32 buffer for the (min, max) partial pairs of z (the quantizer's first pass,
        fused into the last encoder conv); the number of pairs is left in self.z_pairs."""
        if self.train and self._own_xg and not getattr(self, "xg_converted", False):
            ops.convert16(self.x, self.x_g)
        self.xg_converted = False
        self.e0.forward()
        self.e1.forward()
        self.e2.forward()
        if minmax is not None:
            _, self.z_pairs = ops.conv_narrow_out(self.e2.out, self.enc7.weight, 1, y=self.z, ws=self.nws,
                                                  minmax=minmax)
        else:
            ops.conv_narrow_out(self.e2.out, self.enc7.weight, 1, y=self.z, ws=self.nws)
        return self.z

    def forward_decoder(self, z=None):
        z = self.z if z is None else z
        if self.train:
            bz = self.bnz  # statistics of z and the finalize step in one launch
            ops.bn_stats_finalize(z, bz.sums, bz.bn.weight, bz.bn.bias, bz.bn.eps,
                                  0.1 if bz.bn.momentum is None else bz.bn.momentum, bz.bn.running_mean,
                                  bz.bn.running_var, bz.bn.num_batches_tracked, bz.scale_shift, bz.mean_invstd,
                                  planar=True, zeroed=bz.pooled)
        else:
            self.bnz.eval_params()
        ops.conv_narrow_in(z, self.dec2.weight, 0, pre=self.bnz.scale_shift, pre_relu=True, y=self.raw3,
                           ws=self.nws)
        if self.train:
            ops.bn_stats(self.raw3, self.bn3.sums, zeroed=self.bn3.pooled)
            self.bn3.finalize_apply(self.raw3, self.act3, False, y2=self.act3_g)
        else:
            self.bn3.eval_params()
            ops.bn_apply(self.raw3, self.act3, self.bn3.scale_shift, False, y2=self.act3_g)
        self.d4.forward()
        self.d7.forward()
        self.d9.forward()
        return self.out

    def forward(self):
        self.forward_encoder()
        return self.forward_decoder()

    # ---- backward ----
    def plan_backward(self, g_out, grads, prefix, need_gx=True):
        """g_out: gradient w.r.t. layer1's output (bf16 NHWC).  grads: name -> fp32 view."""
        dev = g_out.device
        N, gd = self.N, self.grad_dtype
        e, d = prefix + "encoder.encoder.", prefix + "decoder."
        g = lambda u: _empty((N, u.H, u.W, u.C), gd, dev)
        self.g_d7out, self.g_d4out, self.g_act3 = g(self.d9), g(self.d7), g(self.d4)
        def below(u):  # the BN under a unit's input, for the fused backward reductions
            return (u.out, u.relu, u.bn.sums_b, u.bn.pooled)
        fuse = FUSE_BN_REDUCE
        self.d9.plan_backward(g_out, self.g_d7out, grads, (d + "9.weight", d + "10.weight", d + "10.bias"), gd,
                              below=below(self.d7))
        self.d7.plan_backward(self.g_d7out, self.g_d4out, grads, (d + "7.weight", d + "8.weight", d + "8.bias"), gd,
                              below=below(self.d4))
        self.d4.plan_backward(self.g_d4out, self.g_act3, grads, (d + "4.weight", d + "5.weight", d + "5.bias"), gd,
                              below=(self.act3, False, self.bn3.sums_b, self.bn3.pooled))
        self.d7.reduce_fused = self.d4.reduce_fused = self.bn3_reduce_fused = fuse
        self.g_raw3 = _empty((N, self.H3, self.W3, 64), gd, dev)
        self.g_zact = _empty(tuple(self.z.shape), torch.float32, dev)  # grad wrt relu(bn0(z))
        self.g_z = _empty(tuple(self.z.shape), torch.float32, dev)
        self.wws = _empty((_lib.load().ghnd_wgrad_narrow_workspace_bytes(self.bch, 64, 2, 2),), torch.uint8, dev)
        self.gr = {k: grads[k] for k in (d + "3.weight", d + "3.bias", d + "2.weight", d + "0.weight",
                                         d + "0.bias", e + "7.weight")}
        self.names = (e, d)
        self.g_e2out, self.g_e1out, self.g_e0out = g_like(self.e2, gd, dev), g_like(self.e1, gd, dev), g_like(self.e0, gd, dev)
        self.g_x = _empty((N, self.H, self.W, 64), gd, dev) if need_gx else None
        self.e2.plan_backward(self.g_e2out, self.g_e1out, grads, (e + "5.weight", e + "6.weight", e + "6.bias"), gd,
                              below=below(self.e1))
        self.e1.plan_backward(self.g_e1out, self.g_e0out, grads, (e + "2.weight", e + "3.weight", e + "3.bias"), gd,
                              below=below(self.e0))
        self.e0.plan_backward(self.g_e0out, self.g_x, grads, (e + "0.weight", e + "1.weight", e + "1.bias"), gd)
        self.e1.reduce_fused = self.e0.reduce_fused = fuse

    def wide_units(self):
        return (self.e0, self.e1, self.e2, self.d4, self.d7, self.d9)

    def prepack(self):
        """This step's packed weights of the six wide units (forward + transposed).  PREPACK_BATCHED: ONE launch
        (ops.pack_weights) instead of twelve."""
        units = [u for u in self.wide_units() if u.train]
        if PREPACK_BATCHED and units:
            items = []
            for u in units:
                items.append((u.conv.weight, None, False, u.w))
                if getattr(u, "wt", None) is not None:
                    items.append((u.conv.weight, None, True, u.wt))
            ops.pack_weights(items)
            for u in units:
                u.prepacked = True
            return
        for u in self.wide_units():
            u.prepack()

    def convert_xg(self):
        """The layer input in the gradient format (operand of enc.0's tensor-core dW, needed in the backward
        pass only): the plan runs it on the side stream once the stem has produced the input."""
        if self.train and self._own_xg:
            ops.convert16(self.x, self.x_g)
        self.xg_converted = True

    def backward(self, side=None):
        e, d = self.names
        self.d9.backward(side)
        self.d7.backward(side)
        self.d4.backward(side)
        # BN dec[3] (no ReLU) on raw3
        b3 = self.bn3
        if not self.bn3_reduce_fused:
            ops.bn_bwd_reduce(self.g_act3, self.raw3, b3.scale_shift, b3.mean_invstd, False, b3.sums_b,
                              zeroed=b3.pooled)
        ops.bn_bwd_apply(self.g_act3, self.raw3, self.g_raw3, b3.bn.weight, b3.scale_shift, b3.mean_invstd,
                         False, b3.sums_b, self.gr[d + "3.weight"], self.gr[d + "3.bias"],
                         fused_sums=self.bn3_reduce_fused)
        # dec2 (narrow-in conv on relu(bn0(z)))
        bz = self.bnz
        # the two bottleneck-side dW launches feed nothing on the data-gradient chain: like the wide dW GEMMs they
        # run on the side stream (one after the other there, so they may share the workspace)
        narrow_side = side if NARROW_DW_SIDE else None
        # mode 2: both launches wait until the end of the layer's backward pass (forked right here they fill every SM
        # for 22 us each and the small kernels of the chain below queue behind them for ~11 us twice).  Measured
        # WORSE than mode 1 (same box, 150 steps x 3: 770.3 / 768.4 -> 766.2 / 766.5 / 765.0 img/s): at the end they
        # hold the SMs the conv1 dW GEMM is waiting for.
        defer = narrow_side is not None and NARROW_DW_SIDE == 2

        def dw_dec2():
            ops.wgrad_narrow(self.z, self.g_raw3, self.gr[d + "2.weight"], False, 2, 2, 0, pre=bz.scale_shift,
                             pre_relu=True, ws=self.wws)

        def dw_enc7():
            ops.wgrad_narrow(self.g_z, self.e2.out, self.gr[e + "7.weight"], True, 2, 2, 1, ws=self.wws)

        if defer:
            pass
        elif narrow_side is not None:
            narrow_side.fork()
            narrow_side.run(dw_dec2)
        else:
            dw_dec2()
        ops.conv_narrow_out_dgrad(self.g_raw3, self.dec2.weight, 0, self.Hz, self.Wz, dx=self.g_zact,
                                  ws=self.nws)
        # BN dec[0] + ReLU on the planar bottleneck
        ops.bn_bwd_reduce(self.g_zact, self.z, bz.scale_shift, bz.mean_invstd, True, bz.sums_b, planar=True,
                          zeroed=bz.pooled)
        ops.bn_bwd_apply(self.g_zact, self.z, self.g_z, bz.bn.weight, bz.scale_shift, bz.mean_invstd, True,
                         bz.sums_b, self.gr[d + "0.weight"], self.gr[d + "0.bias"], planar=True)
        # enc7 (narrow-out conv)
        if defer:
            pass
        elif narrow_side is not None:
            narrow_side.fork()
            narrow_side.run(dw_enc7)
        else:
            dw_enc7()
        ops.conv_narrow_in(self.g_z, self.enc7.weight, 1, flip=True, y=self.g_e2out, ws=self.nws)
        self.e2.backward(side)
        self.e1.backward(side)
        self.e0.backward(side)
        if defer:
            narrow_side.fork()
            narrow_side.run(dw_dec2)
            narrow_side.run(dw_enc7)


def _same_frozen_layers(teacher_body, student_body, names):
    """True iff every tensor of the named frozen layers is identical in both bodies."""
    for name in names:
        t_sd, s_sd = getattr(teacher_body, name).state_dict(), getattr(student_body, name).state_dict()
        if t_sd.keys() != s_sd.keys():
            return False
        for k, v in t_sd.items():
            w = s_sd[k]
            if v.shape != w.shape or v.dtype != w.dtype or not torch.equal(v, w):
                return False
    return True


def g_like(unit, dtype, dev):
    return _empty((unit.N, unit.Ho, unit.Wo, unit.K), dtype, dev)


class FlatParams(object):
    """All trainable student tensors as views of ONE flat fp32 buffer (+ flat grad / Adam moments):
    a single all-reduce and a single fused Adam kernel cover the 25 tensors."""

    def __init__(self, named_params):
        named_params = [(n, p) for n, p in named_params if p.requires_grad]
        assert named_params, "no trainable parameters"
        dev = named_params[0][1].device
        sizes = [((p.numel() + 3) // 4) * 4 for _, p in named_params]  # keep 16-byte alignment
        total = sum(sizes)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.names, self.params, self.grads = [], {}, {}
        off = 0
        for (n, p), sz in zip(named_params, sizes):
            view = self.flat[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            self.grads[n] = self.grad[off:off + p.numel()].view(p.shape)
            self.params[n] = p
            self.names.append(n)
            off += sz
        self.total = total


class GhndPlan(object):
    """One fixed-shape GHND / HND distillation step for a (teacher body, student body) pair."""

    def __init__(self, teacher_body, student_body, N, Hp, Wp, levels=LEVELS, factors=None,
                 act_dtype=torch.float16, grad_dtype=torch.bfloat16, device=None, flat=None,
                 image_mean=IMAGE_MEAN, image_std=IMAGE_STD, share_frozen_trunk=None):
        _lib.check(_lib.load().ghnd_device_check(), "ghnd_device_check")
        dev = torch.device(device) if device is not None else next(student_body.parameters()).device
        self.device, self.N, self.Hp, self.Wp = dev, N, Hp, Wp
        self.levels = tuple(l for l in LEVELS if l in levels)
        assert self.levels, "no loss levels"
        self.factors = {l: 1.0 for l in self.levels}
        if factors:
            self.factors.update(factors)
        self.act_dtype, self.grad_dtype = act_dtype, grad_dtype
        self.mean, self.std = tuple(image_mean), tuple(image_std)
        top = max(LEVELS.index(l) for l in self.levels)
        self.top = LEVELS[top]
        self.packed = torch.zeros((N, Hp + 6, Wp + 8, 4), dtype=act_dtype, device=dev)
        # Frozen layers 2..top usually hold IDENTICAL tensors in teacher and student (the student is
        # the teacher with layer1 replaced, config/ghnd/*.yaml) -> run them once on a 2N batch
        # (teacher images first): half the launches, twice the tiles per launch.
        upper = LEVELS[1:top + 1]
        if share_frozen_trunk is None:
            share_frozen_trunk = bool(upper) and _same_frozen_layers(teacher_body, student_body, upper)
        self.shared = bool(share_frozen_trunk) and bool(upper)
        # ---- teacher (forward only) ----
        # both conv1's as one K=128 GEMM over the shared image (GHND_DUAL_STEM=0: two K=64 GEMMs)
        self.stem2 = None
        if os.environ.get("GHND_DUAL_STEM", "1") != "0":
            self.stem2 = DualStemConv(teacher_body, student_body, self.packed, N, Hp, Wp, act_dtype)
        self.t_stem = StemRunner(teacher_body, self.packed, N, Hp, Wp, act_dtype, grad_dtype, False,
                                 shared=(self.stem2, 0) if self.stem2 else None)
        H1, W1 = self.t_stem.Ho, self.t_stem.Wo
        self.trunk_in = _empty((2 * N, H1, W1, 256), act_dtype, dev) if self.shared else None
        self.t_layers = {}
        self.t_layers["layer1"] = FrozenLayerRunner(teacher_body.layer1, self.t_stem.out, N, H1, W1, act_dtype,
                                                    grad_dtype, False,
                                                    out=self.trunk_in[:N] if self.shared else None)
        if not self.shared:
            x, H, W = self.t_layers["layer1"].out, H1, W1
            for name in upper:
                r = FrozenLayerRunner(getattr(teacher_body, name), x, N, H, W, act_dtype, grad_dtype, False)
                self.t_layers[name] = r
                x, H, W = r.out, r.Ho, r.Wo
        # ---- student ----
        self.flat = flat if flat is not None else FlatParams(
            [("backbone.body." + n, p) for n, p in student_body.named_parameters()])
        grads = self.flat.grads
        self.s_stem = StemRunner(student_body, self.packed, N, Hp, Wp, act_dtype, grad_dtype, True,
                                 shared=(self.stem2, 64) if self.stem2 else None)
        # every BatchNorm reduction buffer of the student's layer1 in one allocation, zeroed once per step
        self.sums_pool = SumsPool(dev) if SUMS_POOL else None
        SumsPool.current = self.sums_pool
        try:
            self.s_l1 = StudentLayer1Runner(student_body.layer1, self.s_stem.out, N, self.s_stem.Ho,
                                            self.s_stem.Wo, act_dtype, grad_dtype, True,
                                            out=self.trunk_in[N:] if self.shared else None)
        finally:
            SumsPool.current = None
        self.s_layers = {}
        if self.shared:
            x, H, W = self.trunk_in, H1, W1
            for name in upper:
                r = FrozenLayerRunner(getattr(student_body, name), x, 2 * N, H, W, act_dtype, grad_dtype, True,
                                      bwd_n=N)
                self.s_layers[name] = r
                x, H, W = r.out, r.Ho, r.Wo
        else:
            x, H, W = self.s_l1.out, H1, W1
            for name in upper:
                r = FrozenLayerRunner(getattr(student_body, name), x, N, H, W, act_dtype, grad_dtype, True)
                self.s_layers[name] = r
                x, H, W = r.out, r.Ho, r.Wo
        # ---- loss ----
        self.feat_t = {"layer1": self.t_layers["layer1"].out}
        self.feat_s = {"layer1": self.s_l1.out}
        for name in upper:
            if self.shared:
                self.feat_t[name] = self.s_layers[name].out[:N]
                self.feat_s[name] = self.s_layers[name].out[N:]
            else:
                self.feat_t[name] = self.t_layers[name].out
                self.feat_s[name] = self.s_layers[name].out
        self.loss_grads = {l: torch.empty_like(self.feat_s[l], dtype=grad_dtype) for l in self.levels}
        self.loss_out = torch.zeros(1 + len(self.levels), dtype=torch.float32, device=dev)
        self.sse_ws = _empty((_lib.load().ghnd_sse_workspace_bytes(),), torch.uint8, dev)
        # ---- backward plans (top level down) ----
        g = self.loss_grads[self.top]  # masked by the SSE kernel (relu_mask on the top level)
        for name in reversed(LEVELS[1:top + 1]):
            r = self.s_layers[name]
            below = LEVELS[LEVELS.index(name) - 1]
            b0 = r.blocks[0]
            g_x = _empty((N, b0.H, b0.W, b0.cin), grad_dtype, dev)
            r.plan_backward(g, g_x, self.loss_grads.get(below), grad_dtype)
            g = g_x
        self.s_l1.plan_backward(g, grads, "backbone.body.layer1.")
        self.graph = None
        self.step_count = 0
        self.side = SideStream(dev) if os.environ.get("GHND_SIDE_STREAM", "1") != "0" else None

    # ------------------------------------------------------------------------------------------
    def load_images(self, images):
        """images: list of N fp32 [3,H,W] CUDA tensors in [0,1] (already at network scale)."""
        assert len(images) == self.N, "plan was built for batch %d" % self.N
        ops.stem_pack_images(images, self.packed, self.Hp, self.Wp, self.mean, self.std)

    def forward_backward(self):
        """Enqueue teacher fwd, student fwd, loss and student bwd on the current stream."""
        side = self.side
        if side is not None:
            # branch 1 (side): this step's weight repacking, then the whole teacher forward;
            # branch 2 (main): student stem + layer1.  They meet at the (shared) frozen trunk.
            if PREPACK_BATCHED:
                self.s_l1.prepack()  # one small launch in front of the stem kernel
                side.fork()
            else:
                side.fork()
                side.run(self.s_l1.prepack)
            packed = None if PREPACK_BATCHED else side.mark()
            side.run(self.s_stem.convert_image)  # bf16 copy of the packed image: off the backward tail
            if self.stem2 is not None:
                self.stem2.forward()  # both conv1's; the two pools + layer1's then run side by side
                side.fork()
                if self.stem2.fused:  # the student's pooled output exists: its bf16 copy leaves the main chain
                    side.run(self.s_l1.convert_xg)
            if os.environ.get("GHND_TEACHER_SIDE", "1") != "0":
                side.run(self._teacher_forward)
            else:
                self._teacher_forward()
            self.s_stem.forward()
            if packed is not None:
                torch.cuda.current_stream().wait_event(packed)
            self.s_l1.forward()
            side.join()
        else:
            if self.stem2 is not None:
                self.stem2.forward()
            self._teacher_forward()
            self.s_stem.forward()
            self.s_l1.forward()
        for name in LEVELS[1:]:
            if name in self.s_layers:
                # both models' images when the frozen trunk is shared (the teacher has been joined by
                # then, so the side stream is free for the downsample convs)
                self.s_layers[name].forward(side if self.shared else None)
        lv = [(self.feat_t[l], self.feat_s[l], self.loss_grads[l], self.factors[l], l == self.top)
              for l in self.levels]
        ops.sse_fwd_bwd(lv, self.grad_dtype, self.loss_out, self.sse_ws)
        for name in reversed(LEVELS[1:]):
            if name in self.s_layers:
                self.s_layers[name].backward(side)
        self.s_l1.backward(side)
        self.s_stem.backward(self.s_l1.g_x, self.flat.grads["backbone.body.conv1.weight"])
        if side is not None:
            side.join()
        if self.sums_pool is not None:
            # every reduction buffer has been consumed: leave the pool zeroed for the next step (it is created
            # zeroed; as the first node of the step the fill kernel sat 11 us in front of the next launch)
            self.sums_pool.zero()
        return self.loss_out

    def _teacher_forward(self):
        self.t_stem.forward()
        for name in LEVELS:
            if name in self.t_layers:
                self.t_layers[name].forward()

    def capture(self):
        """Capture forward_backward() into a CUDA graph (buffers and plans are static)."""
        torch.cuda.synchronize()
        # the warm-up run must not count as a training step: keep the BN running statistics
        bufs = list(self.s_l1.layer1.buffers())
        saved = [b.clone() for b in bufs]
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.forward_backward()  # warm-up outside capture (lazy module loading)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        for b, v in zip(bufs, saved):
            b.copy_(v)
        before = ops.launches()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.forward_backward()
        self.launches_per_step = ops.launches() - before
        self.graph = g
        self._inflight = [torch.cuda.Event() for _ in range(3)]
        return g

    def step(self, images=None):
        if images is not None:
            self.load_images(images)
        if self.graph is not None:
            # Bound the number of graph launches in flight: a loop that never reads anything back
            # (bench.py's device-resident arm) otherwise queues hundreds of 200-node launches, which
            # crashed inside cudaGraphLaunch on this driver (580.159) after a few hundred replays.
            ev = self._inflight[self.step_count % len(self._inflight)]
            if self.step_count >= len(self._inflight):
                ev.synchronize()
            self.graph.replay()
            ev.record()
            ops._count(self.launches_per_step)
        else:
            self.forward_backward()
        self.step_count += 1
        return self.loss_out


class EncodePlan(object):
    """Split-computing head (RcnnHead, split_rcnn.py:23-37) for a fixed shape: stem -> layer1
    encoder with eval-mode BN folded into the convs -> fused 8-bit quantizer."""

    def __init__(self, student_body, N, Hp, Wp, num_bits=8, act_dtype=torch.float16, device=None,
                 image_mean=IMAGE_MEAN, image_std=IMAGE_STD, scale_mode=_lib.QSCALE_DIV):
        _lib.check(_lib.load().ghnd_device_check(), "ghnd_device_check")
        dev = torch.device(device) if device is not None else next(student_body.parameters()).device
        self.N, self.Hp, self.Wp, self.num_bits, self.scale_mode = N, Hp, Wp, num_bits, scale_mode
        self.mean, self.std = tuple(image_mean), tuple(image_std)
        self.packed = torch.zeros((N, Hp + 6, Wp + 8, 4), dtype=act_dtype, device=dev)
        self.stem = StemRunner(student_body, self.packed, N, Hp, Wp, act_dtype, torch.bfloat16, False)
        self.l1 = StudentLayer1Runner(student_body.layer1, self.stem.out, N, self.stem.Ho, self.stem.Wo,
                                      act_dtype, torch.bfloat16, False)
        self.z = self.l1.z
        self.q = torch.empty(tuple(self.z.shape), dtype=torch.uint8, device=dev)
        self.qparams = torch.zeros(4, dtype=torch.int32, device=dev)
        # the encoder's last conv publishes min / max of z, so the quantizer is one streaming pass
        self.minmax = torch.zeros((2 * ops.MINMAX_CAPACITY,), dtype=torch.float32, device=dev)
        self.graph = None

    def load_images(self, images):
        assert len(images) == self.N
        ops.stem_pack_images(images, self.packed, self.Hp, self.Wp, self.mean, self.std)

    def forward(self):
        self.stem.refresh_weights()
        self.stem.forward()
        if self.num_bits == 16:
            self.l1.forward_encoder()
            return self.z
        self.l1.forward_encoder(minmax=self.minmax)
        ops.quantize_u8_minmax(self.z, self.minmax, self.l1.z_pairs, self.num_bits, self.scale_mode,
                               q=self.q, qparams=self.qparams)
        return self.q

    def capture(self):
        torch.cuda.synchronize()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.forward()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        before = ops.launches()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.forward()
        self.launches_per_step = ops.launches() - before
        self.graph = g
        return g

    def run(self, images=None):
        if images is not None:
            self.load_images(images)
        if self.graph is not None:
            self.graph.replay()
            ops._count(self.launches_per_step)
        else:
            self.forward()
        return self.q, self.qparams

    def run_filtered(self, images, ext_encoder):
        """RcnnHead with a neural filter (split_rcnn.py:30-35): stem -> filter -> [stop | encoder +
        quantizer].  The decision is a host-side branch (the reference returns None from Python), so
        this path runs eagerly: one D2H read of the 2 filter probabilities."""
        self.load_images(images)
        self.stem.refresh_weights()
        self.stem.forward()
        skip, ext_z = ext_encoder.filter_decision(self.stem.out)
        if skip:
            return None, ext_z
        if self.num_bits == 16:
            self.l1.forward_encoder()
            return (self.z, None), ext_z
        self.l1.forward_encoder(minmax=self.minmax)
        ops.quantize_u8_minmax(self.z, self.minmax, self.l1.z_pairs, self.num_bits, self.scale_mode,
                               q=self.q, qparams=self.qparams)
        return (self.q, self.qparams), ext_z


class BodyPlan(object):
    """Forward-only backbone body (stem -> layer1..4) for a fixed shape: the teacher's frozen
    ResNet-50 or the student's bottleneck-injected one (BN in eval mode unless layer1.training).
    Used by CustomRCNN.backbone_features / distill_backbone_only (rcnn.py:102-110)."""

    def __init__(self, body, N, Hp, Wp, act_dtype=torch.float16, image_mean=IMAGE_MEAN,
                 image_std=IMAGE_STD, device=None):
        from .resnet_layer import BottleneckBase4Ext
        _lib.check(_lib.load().ghnd_device_check(), "ghnd_device_check")
        dev = torch.device(device) if device is not None else body.conv1.weight.device
        self.body, self.N, self.Hp, self.Wp = body, N, Hp, Wp
        self.mean, self.std = tuple(image_mean), tuple(image_std)
        self.packed = torch.zeros((N, Hp + 6, Wp + 8, 4), dtype=act_dtype, device=dev)
        self.stem = StemRunner(body, self.packed, N, Hp, Wp, act_dtype, torch.bfloat16, False)
        x, H, W = self.stem.out, self.stem.Ho, self.stem.Wo
        self.student = isinstance(body.layer1, BottleneckBase4Ext)
        self.feats = {}
        self.skipped, self.ext_z = False, None
        if self.student:
            self.l1_train = bool(body.layer1.training)
            self.l1 = StudentLayer1Runner(body.layer1, x, N, H, W, act_dtype, torch.bfloat16, self.l1_train)
            x = self.l1.out
        else:
            self.l1 = FrozenLayerRunner(body.layer1, x, N, H, W, act_dtype, torch.bfloat16, False)
            x, H, W = self.l1.out, self.l1.Ho, self.l1.Wo
        self.feats["layer1"] = x
        self.layers = []
        for name in LEVELS[1:]:
            r = FrozenLayerRunner(getattr(body, name), x, N, H, W, act_dtype, torch.bfloat16, False)
            self.layers.append(r)
            x, H, W = r.out, r.Ho, r.Wo
            self.feats[name] = x

    def run(self, images):
        assert len(images) == self.N
        ops.stem_pack_images(images, self.packed, self.Hp, self.Wp, self.mean, self.std)
        self.stem.refresh_weights()
        self.stem.forward()
        self.skipped, self.ext_z = False, None
        if self.student:
            l1 = self.body.layer1
            if getattr(l1, "uses_ext_encoder", False) and not l1.training:
                # neural filter on the stem output (base.py:13-16): may stop inference here
                self.skipped, self.ext_z = l1.encoder.filter_decision(self.stem.out)
                if self.skipped:
                    return self.feats
            z = self.l1.forward_encoder()
            if (not l1.training) and l1.bottleneck_transformer is not None and l1.use_bottleneck_transformer:
                from .resnet_layer import apply_bottleneck_transformer
                z = apply_bottleneck_transformer(l1, z).contiguous()  # base.py:55-57
            self.l1.forward_decoder(z)
        else:
            self.l1.forward()
        for r in self.layers:
            r.forward()
        return self.feats


class FpnPlan(object):
    """BackboneWithFPN.fpn forward on the conv kernels: torchvision
    FeaturePyramidNetwork.forward reached through src/models/org/rcnn.py:399-414 (called at :108 and
    :127) -- 1x1 lateral convs (+bias), top-down nearest-upsample + add, 3x3 output convs (+bias), all
    on the tcgen05 implicit-GEMM ConvPlan; the upsample-add is one streaming kernel per level.
    `feats`: the body's NHWC 16-bit outputs, finest first (layer1..layer4).  Outputs stay NHWC 16-bit
    in self.out (finest first); extra blocks (LastLevelMaxPool = a stride-2 subsample) are left to the
    caller."""

    def __init__(self, fpn, feats, act_dtype=torch.float16):
        from torch import nn
        _lib.check(_lib.load().ghnd_device_check(), "ghnd_device_check")

        def conv_of(blk):  # torchvision >= 0.13 wraps each conv in Conv2dNormActivation
            return blk[0] if isinstance(blk, nn.Sequential) else blk
        assert len(feats) == len(fpn.inner_blocks) == len(fpn.layer_blocks)
        self.feats = list(feats)
        self.lat, self.out, self.inner, self.outer = [], [], [], []
        self._keep = []
        for i, x in enumerate(self.feats):
            n, h, w, c = x.shape
            ci, co = conv_of(fpn.inner_blocks[i]), conv_of(fpn.layer_blocks[i])
            K = ci.weight.shape[0]
            if ci.weight.shape[2:] != (1, 1) or co.weight.shape[2:] != (3, 3) or co.padding[0] != 1:
                raise ValueError("FpnPlan expects 1x1 lateral and 3x3/p1 output convolutions")
            wi = ops.pack_weight(ci.weight, None, False, act_dtype)
            wo = ops.pack_weight(co.weight, None, False, act_dtype)
            bi = ci.bias.detach().float().contiguous() if ci.bias is not None else None
            bo = co.bias.detach().float().contiguous() if co.bias is not None else None
            lat = _empty((n, h, w, K), act_dtype, x.device)
            po = _empty((n, h, w, K), act_dtype, x.device)
            self.inner.append(ops.ConvPlan(CONV_FWD, n, h, w, c, K, 1, 1, 1, 0, x, wi, lat, bias=bi))
            self.outer.append(ops.ConvPlan(CONV_FWD, n, h, w, K, K, 3, 3, 1, 1, lat, wo, po, bias=bo))
            self.lat.append(lat)
            self.out.append(po)
            self._keep += [wi, wo, bi, bo]
        self.flops = sum(p.flops for p in self.inner + self.outer)

    def run(self):
        top = len(self.feats) - 1
        for i in range(top, -1, -1):
            self.inner[i].run()
            if i < top:  # last_inner = inner_lateral + interpolate(last_inner, nearest)
                ops.upsample_add(self.lat[i], self.lat[i + 1], out=self.lat[i])
            self.outer[i].run()
        return self.out
