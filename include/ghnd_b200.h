/*
 * ghnd_b200.h -- C ABI of the B200-native GHND hot path (libghnd_b200.so).
 *
 * The reference (yoshitomo-matsubara/hnd-ghnd-object-detectors) is pure Python and has no FFI;
 * each entry point below replaces the implicit torch/cuDNN library call made at the cited
 * reference location (paths relative to the reference root).  Rules of the boundary:
 *   - extern "C", plain pointers and sizes, a cudaStream_t passed as void*; no torch types.
 *   - every pointer is DEVICE memory owned by the caller unless the comment says "host".
 *   - no allocation of tensor memory, no ownership transfer, no hidden synchronisation;
 *     work is enqueued on the given stream (CUDA-graph capturable).
 *   - return value: GHND_OK or an error code; ghnd_last_error() gives the message (thread local).
 *
 * Tensors on the 16-bit activation path are NHWC ("pixel-major": [N][H][W][C], C contiguous);
 * element format is a run-time flag: GHND_F16 or GHND_BF16.
 */
#ifndef GHND_B200_H
#define GHND_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GHND_OK 0
#define GHND_ERR_INVALID 1
#define GHND_ERR_CUDA 2
#define GHND_ERR_UNSUPPORTED 3

#define GHND_F16 0  /* IEEE half   (matches the tcgen05 kind::f16 a/b format code) */
#define GHND_BF16 1 /* bfloat16 */

#define GHND_ABI_VERSION 1

const char* ghnd_last_error(void);
int ghnd_abi_version(void);
/* GHND_OK iff the current CUDA device is compute capability 10.x (the only target). */
int ghnd_device_check(void);

/* ------------------------------------------------------------------------------------------
 * 8-bit affine quantizer of the bottleneck tensor.
 * Replaces myutils.pytorch.tensor_util.quantize_tensor / dequantize_tensor
 * (src/myutils/pytorch/tensor_util.py:8-22), reached through structure.transformer.Quantizer /
 * Dequantizer (src/structure/transformer.py:131-153).
 *   scale = (max-min)/(2^bits-1); zp = trunc(clamp(0 - min/scale, 0, qmax));
 *   q = u8(rint(clamp(zp + x/scale, 0, qmax)))          (one scale for the whole tensor)
 * scale_mode selects how torch evaluates "tensor / python_float" for `scale`:
 *   GHND_QSCALE_DIV   IEEE division            (reference executed on CPU; oracle and goldens)
 *   GHND_QSCALE_RECIP multiply by fl(1/255)    (reference executed by torch on a CUDA device)
 * qparams (device, 16 bytes): [0] float scale, [1] int32 zero_point, [2] float min, [3] float max
 *   zero_point INT_MIN     : the reference would raise (int(NaN))
 *   zero_point INT_MIN + 1 : the grid barrier of the single-launch kernel timed out (see below)
 * workspace: ghnd_quantize_u8_workspace_bytes(n) bytes that MUST BE ZERO on first use; the call
 *   leaves them zeroed again (per-CTA min/max partials followed by the words of a grid-wide
 *   barrier: the 16-byte-aligned path is ONE persistent launch, one CTA per SM, that keeps its
 *   slice of x in shared memory between the min/max pass and the quantize pass).  One workspace
 *   per stream: concurrent calls must not share it.
 * ------------------------------------------------------------------------------------------ */
#define GHND_QSCALE_DIV 0
#define GHND_QSCALE_RECIP 1
size_t ghnd_quantize_u8_workspace_bytes(int64_t n);
int ghnd_quantize_u8(const float* x, int64_t n, int num_bits, int scale_mode, uint8_t* q,
                     void* qparams, void* workspace, size_t workspace_bytes, void* stream);
/* The same quantizer as ONE streaming pass (4 B read + 1 B write per element) when the tensor's
 * min / max are already known as `n_partial` (min, max) pairs -- written by the kernel that
 * produced x (ghnd_conv_narrow_out_minmax: the encoder's last conv, src/models/mimic/
 * split_rcnn.py:31-35 runs encoder then Quantizer back to back).  Folding the pairs, deriving
 * scale / zero-point and quantizing follow the same operation order; results are bit-identical to
 * ghnd_quantize_u8 on the same tensor.  No workspace. */
int ghnd_quantize_u8_minmax(const float* x, int64_t n, int num_bits, int scale_mode,
                            const float* minmax_partial, int n_partial, uint8_t* q, void* qparams,
                            void* stream);
/* out = scale * (float(q) - zero_point)   (tensor_util.py:21-22) */
int ghnd_dequantize_u8(const uint8_t* q, int64_t n, const void* qparams, float* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * HND / GHND feature-mimicking loss, forward + backward fused.
 * Replaces GeneralizedCustomLoss.forward with MSELoss(reduction='sum') terms
 * (src/distillation/loss.py:25-34, src/myutils/pytorch/func_util.py:9-13) and its autograd
 * backward:  L = sum_l factor_l * sum (t_l - s_l)^2 ;  dL/ds_l = 2 * factor_l * (s_l - t_l).
 * loss_out (device): float[1 + n_levels] = { L, factor_0*SSE_0, ... }.
 * `levels` is a HOST array. grad may be NULL (forward only).
 * ------------------------------------------------------------------------------------------ */
#define GHND_SSE_MAX_LEVELS 8
typedef struct ghnd_sse_level {
  const void* teacher;
  const void* student;
  void* grad; /* same shape as student, grad_fmt; may be NULL */
  int64_t n;  /* elements, multiple of 8 */
  float factor;
  int relu_mask; /* 1: grad = student>0 ? 2f(s-t) : 0  (student is a post-ReLU output whose
                    gradient is consumed directly as the pre-activation gradient) */
} ghnd_sse_level_t;
size_t ghnd_sse_workspace_bytes(void);
int ghnd_sse_fwd_bwd(const ghnd_sse_level_t* levels, int n_levels, int in_fmt, int grad_fmt,
                     float* loss_out, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Layout / dtype boundary kernels (module boundary: reference tensors are NCHW fp32).
 * ------------------------------------------------------------------------------------------ */
int ghnd_nchw_f32_to_nhwc16(const float* src, void* dst, int dst_fmt, int N, int C, int H, int W,
                            void* stream);
int ghnd_nhwc16_to_nchw_f32(const void* src, int src_fmt, float* dst, int N, int C, int H, int W,
                            void* stream);

/* Input transform of GeneralizedRCNNTransform.normalize + batch_images for images that need no
 * resize (src/models/org/rcnn.py:73-80; torchvision transform.py normalize/batch_images):
 * dst[n][h+3][w+3][0..3] = ((img-mean)/std, 0) as 4-channel pixels inside a zero frame of 3 px
 * (the 7x7 stem's padding), dst dims [N][Hp+6][Wp+8][4], dst_fmt 16-bit.  One image per call.
 * mean/std: host float[3]. */
int ghnd_stem_pack_image(const float* img_chw, int H, int W, const float* mean, const float* std_,
                         void* dst, int dst_fmt, int n_index, int Hp, int Wp, void* stream);

/* The same with GeneralizedRCNNTransform.resize fused in (src/models/org/rcnn.py:29-45: normalize ->
 * interpolate(scale_factor, bilinear, align_corners=False) -> batch_images zero pad): the H x W
 * source is resampled to Ho x Wo (= floor(H*scale), floor(W*scale)) while packing.  rscale_* is the
 * source step per output pixel, (float)(1.0/scale_factor) as ATen derives it when scale_factor is
 * given; tap indices/weights follow ATen's upsample_bilinear2d.  One image per call. */
int ghnd_stem_pack_image_resized(const float* img_chw, int H, int W, int Ho, int Wo, float rscale_h,
                                 float rscale_w, const float* mean, const float* std_, void* dst,
                                 int dst_fmt, int n_index, int Hp, int Wp, void* stream);
/* The whole batch in ONE launch (the per-image calls above cost a launch each on the end-to-end path):
 * image k = imgs_chw[k] of H[k] x W[k], written at H'[k] = Ho[k] x Wo[k] into slot n_index0 + k; rscale[k] = 0
 * means no resize (then Ho = H, Wo = W), else the ATen source-index scale 1/scale_factor of
 * ghnd_stem_pack_image_resized.  All arrays are HOST arrays of n entries. */
int ghnd_stem_pack_images(const float* const* imgs_chw, const int* H, const int* W, const int* Ho, const int* Wo,
                          const float* rscale, int n, const float* mean, const float* std_, void* dst, int dst_fmt,
                          int n_index0, int Hp, int Wp, void* stream);

/* ------------------------------------------------------------------------------------------
 * Wide convolutions: implicit GEMM on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 * Replaces nn.Conv2d (+FrozenBatchNorm2d +ReLU +residual) forward and its autograd dgrad:
 *   student layer1 k=2 convs   src/models/mimic/resnet_layer.py:42-65
 *   teacher layer1, layer2-4   torchvision Bottleneck via src/models/org/rcnn.py:388-396
 * A plan binds geometry + device pointers (TMA descriptors embed them); run enqueues it.
 *   FWD  : src = x  [N,H,W,C]   weights = [K][R][S][C]    dst = y  [N,Ho,Wo,K]
 *   DGRAD: src = dy [N,Ho,Wo,K] weights = [C][R][S][K]    dst = dx [N,H,W,C]
 *          (weights for DGRAD are the forward weights transposed, NOT flipped)
 * Epilogue (all indexed like dst): v = acc + bias[ch] + residual; relu; v = mask>0 ? v : 0;
 * accumulate: dst += v.
 * ------------------------------------------------------------------------------------------ */
#define GHND_CONV_FWD 0
#define GHND_CONV_DGRAD 1
typedef struct ghnd_conv_desc {
  int kind;
  int N, H, W, C; /* forward-input geometry */
  int K;          /* forward-output channels */
  int R, S, stride, pad;
  const void* src;
  int src_fmt;
  const void* weights;
  int w_fmt;
  void* dst;
  int dst_fmt;
  const float* bias;
  const void* residual;
  int res_fmt;
  int relu;
  const void* mask;
  int mask_fmt;
  int accumulate;
  double* stats; /* optional: [2*channels] sum / sum-of-squares of the stored (rounded) output */
  /* stats_mode 1 (needs `mask`): stats = sum of dst, sum of dst * mask-operand per channel -- the two
   * reductions of the BatchNorm backward of the layer that produced the mask operand (its
   * post-BN/ReLU activation `a`): sum g' and sum g'*a, from which bn_bwd_apply (sums_mode 1) recovers
   * sum g'*xhat = (sum g'*a - beta * sum g') / gamma.  Lets a dgrad launch absorb the separate
   * bn_bwd_reduce pass over (g, x) of the previous layer. */
  int stats_mode;
  /* 1: the mask operand only feeds the statistics (layer without ReLU); 0: it also masks dst */
  int mask_stats_only;
} ghnd_conv_desc_t;
typedef struct ghnd_conv_plan ghnd_conv_plan_t;
int ghnd_conv_plan_create(const ghnd_conv_desc_t* desc, ghnd_conv_plan_t** plan);
int ghnd_conv_plan_run(const ghnd_conv_plan_t* plan, void* stream);
/* Enqueue only launches [first, first+count) of the plan.  The launches of one plan write disjoint
 * parts of dst (the parity classes of a stride-2 dgrad), so a caller may spread them over streams. */
int ghnd_conv_plan_run_range(const ghnd_conv_plan_t* plan, int first, int count, void* stream);
void ghnd_conv_plan_destroy(ghnd_conv_plan_t* plan);
/* number of kernel launches one run of the plan enqueues (for gpu_launches accounting) */
int ghnd_conv_plan_launches(const ghnd_conv_plan_t* plan);

/* Weight gradient of a wide conv on tcgen05 (MN-major operands straight from NHWC tensors):
 *   dw[K][R][S][C] (fp32) = sum_{n,ho,wo} dy[n,ho,wo,k] * x[n, ho*stride+r-pad, wo*stride+s-pad, c]
 * stride must be 1 (only the student's layer1 convs are trainable, all stride 1). */
typedef struct ghnd_wgrad_desc {
  int N, H, W, C, K, R, S, pad;
  const void* x;
  int x_fmt;
  const void* dy;
  int dy_fmt;
  float* dw; /* [K][R][S][C] fp32, overwritten */
} ghnd_wgrad_desc_t;
typedef struct ghnd_wgrad_plan ghnd_wgrad_plan_t;
int ghnd_wgrad_plan_create(const ghnd_wgrad_desc_t* desc, ghnd_wgrad_plan_t** plan);
int ghnd_wgrad_plan_run(const ghnd_wgrad_plan_t* plan, void* stream);
void ghnd_wgrad_plan_destroy(ghnd_wgrad_plan_t* plan);

/* Weight repack: OIHW fp32 (nn.Conv2d.weight) * optional per-O scale -> [O][R][S][I] 16-bit
 * (transpose=0, forward layout) or [I][R][S][O] (transpose=1, dgrad layout).  I is zero-padded to
 * I_pad / O to O_pad in the packed tensor. */
int ghnd_pack_weight(const float* w_oihw, const float* scale_o, int O, int I, int R, int S,
                     int transpose, void* dst, int dst_fmt, void* stream);
/* Several repacks in ONE launch (the student's layer1 repacks 12 tensors at the start of every step: as 12 small
 * launches they queued behind the stem kernel and held the first layer1 conv back).  descs: HOST array of n <=
 * GHND_PACK_MAX entries with the arguments of ghnd_pack_weight. */
#define GHND_PACK_MAX 16
typedef struct {
  const float* w_oihw;
  const float* scale_o; /* may be NULL */
  int O, I, R, S, transpose;
  void* dst;
  int dst_fmt;
} ghnd_pack_weight_desc_t;
int ghnd_pack_weights(const ghnd_pack_weight_desc_t* descs, int n, void* stream);
/* inverse for gradients: dw [O][R][S][I] fp32 -> OIHW fp32 (scaled by alpha) */
int ghnd_unpack_wgrad(const float* dw_orsi, float* dst_oihw, int O, int I, int R, int S,
                      float alpha, void* stream);

/* ------------------------------------------------------------------------------------------
 * Narrow (bottleneck-side) convolutions, k=2 stride 1: HBM-bound SIMT kernels.
 *   enc7: Conv2d(64 -> bch, k2, p1)  resnet_layer.py:50     (narrow OUTPUT, planar fp32 NCHW z)
 *   dec2: Conv2d(bch -> 64, k2, p0)  resnet_layer.py:55     (narrow INPUT,  planar fp32 NCHW z)
 * wide side is NHWC 16-bit, narrow side is NCHW fp32 (it is the quantizer's tensor).
 * ------------------------------------------------------------------------------------------ */
/* y[n][k][ho][wo] = sum_{r,s,c} w[k][c][r][s] * x[n][ho+r-pad][wo+s-pad][c];  w: OIHW fp32 */
size_t ghnd_conv_narrow_workspace_bytes(int C, int K, int R, int S);
int ghnd_conv_narrow_out(const void* x, int x_fmt, const float* w, float* y, int N, int H, int W,
                         int C, int K, int R, int S, int pad, void* workspace,
                         size_t workspace_bytes, void* stream);
/* The same conv that also publishes per-CTA (min, max) pairs of the stored fp32 outputs
 * (NaN-propagating like torch.min / torch.max): minmax_partial[2*i], [2*i+1] for i < *n_partial,
 * *n_partial <= partial_capacity (2 * SM count pairs always suffice).  64-channel 16-bit input only.
 * Feeds ghnd_quantize_u8_minmax. */
int ghnd_conv_narrow_out_minmax(const void* x, int x_fmt, const float* w, float* y, int N, int H,
                                int W, int C, int K, int R, int S, int pad, void* workspace,
                                size_t workspace_bytes, float* minmax_partial, int partial_capacity,
                                int* n_partial, void* stream);
/* flip=0: y[n][ho][wo][k] = sum_{r,s,c} w[k][c][r][s] * f(x[n][c][ho+r-pad][wo+s-pad]),
 *   f(v) = pre_scale_shift ? (v*scale[c]+shift[c], then max(.,0) if pre_relu) : v
 *   (decoder BN0+ReLU fused; zero padding is applied AFTER f).  w: OIHW [K][C][R][S].
 * flip=1: data-gradient of a narrow-OUT conv (wide K ch -> narrow C ch, weight OIHW [C][K][R][S],
 *   padding pad): x = dy of that conv, planar [N][C][H][W]; y = dx, NHWC [N][H-2pad+R-1][..][K]. */
int ghnd_conv_narrow_in(const float* x, const float* pre_scale_shift, int pre_relu, const float* w,
                        int flip, void* y, int y_fmt, int N, int H, int W, int C, int K, int R,
                        int S, int pad, void* workspace, size_t workspace_bytes, void* stream);
/* dgrad of a narrow-IN conv: dx[n][c][h][w] = sum_{r,s,k} w[k][c][r][s]*dy[n][h+pad-r][w+pad-s][k]
 * (dy NHWC 16-bit, dx NCHW fp32) -- same arithmetic as narrow_out with flipped taps. */
int ghnd_conv_narrow_out_dgrad(const void* dy, int dy_fmt, const float* w, float* dx, int N, int H,
                               int W, int C, int K, int R, int S, int pad, void* workspace,
                               size_t workspace_bytes, void* stream);
/* weight gradient between a planar fp32 narrow tensor a[N][Ca][Ha][Wa] and an NHWC 16-bit wide
 * tensor b[N][Hb][Wb][Cb]:  out[ca][cb][r][s] = sum a[n][ca][i][j] * b[n][i+r-pad][j+s-pad][cb]
 * (a_is_output=1: a = dy of narrow-out conv, b = its input  -> dw[k=ca][c=cb][r][s])
 * (a_is_output=0: a = x of narrow-in conv, b = its dy      -> dw[k=cb][c=ca][r][s], taps mirrored)
 * dw is written in OIHW fp32.  pre_* as in narrow_in (applied to a when a_is_output=0). */
int ghnd_wgrad_narrow(const float* a, const float* pre_scale_shift, int pre_relu, const void* b,
                      int b_fmt, float* dw_oihw, int a_is_output, int N, int Ha, int Wa, int Ca,
                      int Hb, int Wb, int Cb, int R, int S, int pad, void* workspace,
                      size_t workspace_bytes, void* stream);
size_t ghnd_wgrad_narrow_workspace_bytes(int Ca, int Cb, int R, int S);

/* ------------------------------------------------------------------------------------------
 * Stem: conv1 7x7 s2 p3 (3->64) + FrozenBatchNorm2d + ReLU, then MaxPool 3x3 s2 p1
 * (src/models/custom/resnet.py:26-30,96-99).  Input = ghnd_stem_pack_image output.
 * Implicit GEMM on tcgen05: K = 7 rows x 32 (7 px x 4 ch + 4 zero) = 224.
 *   weights: [64][7][32] 16-bit from ghnd_stem_pack_weight (FrozenBN scale folded), bias fp32[64]
 * ------------------------------------------------------------------------------------------ */
int ghnd_stem_pack_weight(const float* w_oihw /*[64][3][7][7]*/, const float* scale_o, void* dst,
                          int dst_fmt, void* stream);
typedef struct ghnd_stem_plan ghnd_stem_plan_t;
/* y = relu(conv(x)+bias) : [N][Hp/2][Wp/2][64] */
int ghnd_stem_conv_plan_create(const void* x_packed, int x_fmt, const void* w_packed, int w_fmt,
                               const float* bias, void* y, int y_fmt, int N, int Hp, int Wp,
                               ghnd_stem_plan_t** plan);
/* The same for K = 64 * m output channels: m stems that read the SAME packed image as one GEMM
 * (weights [K][7][32], bias[K], y [N,Hp/2,Wp/2,K]).  The distillation step runs the teacher's and
 * the student's conv1 this way (channels 0-63 / 64-127): the im2col traffic, which bounds the stem,
 * is paid once. */
int ghnd_stem_conv_plan_create_k(const void* x_packed, int x_fmt, const void* w_packed, int w_fmt,
                                 const float* bias, void* y, int y_fmt, int N, int Hp, int Wp, int K,
                                 ghnd_stem_plan_t** plan);
int ghnd_stem_conv_plan_run(const ghnd_stem_plan_t* plan, void* stream);
void ghnd_stem_plan_destroy(ghnd_stem_plan_t* plan);
/* conv1 + FrozenBN + ReLU + MaxPool 3x3 s2 p1 as ONE kernel (src/models/custom/resnet.py:26-30,96-99 --
 * `self.conv1 .. self.maxpool` of the backbone body): conv1's [N][Hp/2][Wp/2][64] output is pooled in
 * shared memory and never written.  n_models (1 or 2) stems read the SAME packed image (weights
 * [64*n_models][7][32], bias[64*n_models]; the distillation step runs teacher + student this way);
 * y[m] receives model m's pooled map [N][Ho][Wo][64] (Ho = (Hp/2+1)/2), argmax[m] (array or entries
 * nullable) the window codes of ghnd_maxpool3x3s2 (0xff where the maximum is not > 0).  Bit-identical
 * to ghnd_stem_conv_plan_* followed by ghnd_maxpool3x3s2_strided. */
typedef struct ghnd_stem_pool_plan ghnd_stem_pool_plan_t;
int ghnd_stem_pool_plan_create(const void* x_packed, int x_fmt, const void* w_packed, int w_fmt,
                               const float* bias, int n_models, void* const* y, void* const* argmax,
                               int y_fmt, int N, int Hp, int Wp, ghnd_stem_pool_plan_t** plan);
int ghnd_stem_pool_plan_run(const ghnd_stem_pool_plan_t* plan, void* stream);
void ghnd_stem_pool_plan_destroy(ghnd_stem_pool_plan_t* plan);
/* maxpool 3x3 s2 p1 on NHWC 16-bit: y[N][Ho][Wo][C], Ho=(H+1)/2; argmax (nullable) receives the
 * window position 0..8 of the first maximum, one byte per output element -- or 0xff when that maximum is
 * not > 0: the ReLU mask of x's producer, folded in so that the backward pass need not re-read x. */
int ghnd_maxpool3x3s2(const void* x, void* y, void* argmax, int fmt, int N, int H, int W, int C,
                      void* stream);
/* backward of relu+maxpool: dx[n][h][w][c] = sum of dy over the pool windows whose argmax is
 * (h,w), zero where x (the post-ReLU conv output) is not > 0.  The mask arrives through the argmax
 * codes (0xff); x / x_fmt are kept in the signature for geometry checks only and are not read. */
int ghnd_maxpool3x3s2_bwd(const void* x, int x_fmt, const void* argmax, const void* dy, int dy_fmt,
                          void* dx, int dx_fmt, int N, int H, int W, int C, void* stream);
/* Strided-input variants: x holds x_channels channels per pixel and the pool works on the C channels
 * starting at x_channel_offset (multiples of 8) -- one half of the two-stem conv output. */
int ghnd_maxpool3x3s2_strided(const void* x, int x_channels, int x_channel_offset, void* y, void* argmax,
                              int fmt, int N, int H, int W, int C, void* stream);
int ghnd_maxpool3x3s2_bwd_strided(const void* x, int x_fmt, int x_channels, int x_channel_offset,
                                  const void* argmax, const void* dy, int dy_fmt, void* dx, int dx_fmt,
                                  int N, int H, int W, int C, void* stream);
/* dW of conv1: dw[k][c][r][s] = scale[k] * sum g[n][ho][wo][k] * xpacked[n][2ho+r][2wo+s][c]
 * (fp32 OIHW [64][3][7][7]); SIMT register-tiled, split over pixels, fp32 atomics into workspace. */
size_t ghnd_stem_wgrad_workspace_bytes(void);
int ghnd_stem_wgrad(const void* x_packed, int x_fmt, const void* g, int g_fmt,
                    const float* scale_o, float* dw_oihw, int N, int Hp, int Wp, void* workspace,
                    size_t workspace_bytes, void* stream);

/* Same dW on tcgen05 tensor cores: GEMM M = (filter row, 8 px x 4 ch window) = 7 x 32, N = 64,
 * K = output pixels, both operands MN-major straight from the NHWC tensors by TMA.  x_packed and g
 * must share ONE 16-bit format (tcgen05 kind::f16 restriction), so the caller keeps a copy of the
 * packed image in the gradient format.  workspace as for ghnd_stem_wgrad. */
typedef struct ghnd_stem_wgrad_plan ghnd_stem_wgrad_plan_t;
int ghnd_stem_wgrad_plan_create(const void* x_packed, const void* g, int fmt, const float* scale_o,
                                float* dw_oihw, int N, int Hp, int Wp, void* workspace,
                                size_t workspace_bytes, ghnd_stem_wgrad_plan_t** plan);
int ghnd_stem_wgrad_plan_run(const ghnd_stem_wgrad_plan_t* plan, void* stream);
void ghnd_stem_wgrad_plan_destroy(ghnd_stem_wgrad_plan_t* plan);

/* ------------------------------------------------------------------------------------------
 * Training-mode BatchNorm2d around the student's layer1 convs (nn.BatchNorm2d forward/backward,
 * resnet_layer.py:43-64; batch statistics, biased var for normalisation, unbiased for running).
 * x is NHWC 16-bit with npix = N*H*W pixels, or planar fp32 (planar=1: [N][C][HW]).
 * ------------------------------------------------------------------------------------------ */
/* The reductions below (and a conv launch with `stats`) zero their [2C] double buffer with a memset node of their
 * own before accumulating.  Inside a CUDA graph every such node costs ~6 us on a dependent chain (kernel -> memset
 * -> kernel instead of kernel -> kernel); a caller that keeps all these buffers in one allocation and zeroes it ONCE
 * per step ORs GHND_SUMS_ZEROED into `planar` (ghnd_bn_stats, ghnd_bn_bwd_reduce) or `stats_mode`
 * (ghnd_conv_desc_t): the call then only accumulates. */
#define GHND_SUMS_ZEROED 0x100
/* sums[2C] (double) = {sum x, sum x^2}; zeroed by the call itself before accumulation. */
int ghnd_bn_stats(const void* x, int fmt, int planar, int N, int64_t hw, int C, double* sums,
                  void* stream);
/* ghnd_bn_stats followed by ghnd_bn_finalize (count = N*hw) as ONE launch for planar tensors (the last block to finish
 * finalizes); NHWC tensors take the two launches.  `planar` accepts GHND_SUMS_ZEROED like ghnd_bn_stats. */
int ghnd_bn_stats_finalize(const void* x, int fmt, int planar, int N, int64_t hw, int C, double* sums,
                           const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                           float* running_var, int64_t* num_batches_tracked, float* scale_shift,
                           float* mean_invstd, void* stream);
/* From sums: scale_shift[2C] = {gamma*invstd, beta-mean*gamma*invstd}, mean_invstd[2C];
 * running stats updated in place (momentum), num_batches_tracked += 1 (int64, nullable). */
int ghnd_bn_finalize(const double* sums, int64_t count, int C, const float* gamma,
                     const float* beta, float eps, float momentum, float* running_mean,
                     float* running_var, int64_t* num_batches_tracked, float* scale_shift,
                     float* mean_invstd, void* stream);
/* eval mode: scale_shift from running stats */
int ghnd_bn_eval_params(int C, const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, float eps, float* scale_shift, void* stream);
/* y = x*scale+shift (optionally relu); NHWC 16-bit in/out.  y2 (nullable) receives the same values
 * in a second 16-bit format: forward tensors are fp16, but the weight-gradient MMA needs its two
 * operands in one format, so the student's activations are also kept as bf16 next to the bf16
 * gradients. */
int ghnd_bn_apply(const void* x, int x_fmt, void* y, int y_fmt, void* y2, int y2_fmt, int64_t npix,
                  int C, const float* scale_shift, int relu, void* stream);
/* ghnd_bn_finalize + ghnd_bn_apply as ONE launch (training forward of nn.BatchNorm2d,
 * src/models/mimic/resnet_layer.py:43-64): every thread derives scale / shift of its channels from
 * the batch sums; block 0 also writes scale_shift / mean_invstd (for the backward kernels) and
 * updates running_mean / running_var / num_batches_tracked exactly like ghnd_bn_finalize. */
int ghnd_bn_finalize_apply(const void* x, int x_fmt, void* y, int y_fmt, void* y2, int y2_fmt,
                           int64_t npix, int C, int relu, const double* sums, int64_t count,
                           const float* gamma, const float* beta, float eps, float momentum,
                           float* running_mean, float* running_var, int64_t* num_batches_tracked,
                           float* scale_shift, float* mean_invstd, void* stream);
/* FPN top-down step: y = fine + nearest_upsample(coarse) on NHWC 16-bit tensors (fine/y [N,H,W,C],
 * coarse [N,Hc,Wc,C]; source index floor(dst*in/out)).  Replaces F.interpolate(mode="nearest") + add in
 * torchvision's FeaturePyramidNetwork.forward, reached from the reference through
 * BackboneWithFPN.fpn (src/models/org/rcnn.py:399-414, called at :108,:127). */
int ghnd_upsample_add(const void* fine, const void* coarse, void* y, int fmt, int N, int H, int W, int Hc,
                      int Wc, int C, void* stream);
/* 16-bit format conversion (fp16 <-> bf16), n elements (multiple of 8) */
int ghnd_convert16(const void* x, int x_fmt, void* y, int y_fmt, int64_t n, void* stream);
/* backward, pass 1: sums[2C] = { sum g', sum g'*xhat } with g' = dy * (relu ? (x*scale+shift>0):1) */
int ghnd_bn_bwd_reduce(const void* dy, int dy_fmt, const void* x, int x_fmt, int planar, int N,
                       int64_t hw, int C, const float* scale_shift, const float* mean_invstd,
                       int relu, double* sums, void* stream);
/* backward, pass 2: dx = gamma*invstd*(g' - sum_g/M - xhat*sum_gx/M); also dgamma=sum_gx,
 * dbeta=sum_g written (fp32[C] each, nullable). */
int ghnd_bn_bwd_apply(const void* dy, int dy_fmt, const void* x, int x_fmt, void* dx, int dx_fmt,
                      int planar, int N, int64_t hw, int C, const float* gamma,
                      const float* scale_shift, const float* mean_invstd, int relu,
                      const double* sums, float* dgamma, float* dbeta, void* stream);
/* The same when `sums` were produced by a conv launch with stats_mode 1 (sum g', sum g'*a with a =
 * the layer's post-BN activation): sum g'*xhat is recovered as (is/sc) * (sum g'*a - (sf + mu*sc) * sum g')
 * in fp64.  NHWC f16 activations / bf16 gradients only.  Undefined for gamma == 0 (a carries no
 * information about xhat then). */
int ghnd_bn_bwd_apply_fused_sums(const void* dy, int dy_fmt, const void* x, int x_fmt, void* dx,
                                 int dx_fmt, int N, int64_t hw, int C, const float* gamma,
                                 const float* scale_shift, const float* mean_invstd, int relu,
                                 const double* sums, float* dgamma, float* dbeta, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused multi-tensor Adam over one flat fp32 parameter/gradient buffer
 * (torch.optim.Adam step, src/mimic_runner.py:52-54; lr schedule stays on the host).
 * grad_scale multiplies the gradient first (1/world_size after the all-reduce SUM).
 * step is the 1-based step count used for bias correction.
 * ------------------------------------------------------------------------------------------ */
int ghnd_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                   double lr, double beta1, double beta2, double eps, double weight_decay,
                   double grad_scale, int step, void* stream);

/* ---- neural-filter head Ext4ResNet (src/models/ext/classifier.py:16-37), inference ----------------
 * ghnd_adaptive_avgpool_nhwc16: nn.AdaptiveAvgPool2d((OH,OW)) of an NHWC 16-bit tensor [N,H,W,C] into
 *   NHWC fp32 [N,OH,OW,C] (bin i = [floor(i*H/OH), ceil((i+1)*H/OH)), classifier.py:20).
 * ghnd_small_conv_f32: nn.Conv2d (no padding) + folded eval BatchNorm + optional ReLU on NHWC fp32
 *   tensors, weights repacked [R][S][C][K]; y = act(scale[k]*conv + shift[k]) (classifier.py:21-29).
 * ghnd_avgpool_linear: AdaptiveAvgPool2d((OH,OW)) -> flatten(1) in NCHW order -> nn.Linear -> optional
 *   softmax(dim=1) (classifier.py:30-37); lw is the Linear weight [n_out][C*OH*OW], out [N][n_out]. */
int ghnd_adaptive_avgpool_nhwc16(const void* x, int fmt, int N, int H, int W, int C, float* y, int OH, int OW,
                                 void* stream);
int ghnd_small_conv_f32(const float* x, const float* w_rsck, const float* scale, const float* shift, int relu,
                        float* y, int N, int H, int W, int C, int K, int R, int S, int stride, void* stream);
int ghnd_avgpool_linear(const float* x, int N, int H, int W, int C, int OH, int OW, const float* lw, const float* lb,
                        int n_out, int softmax, float* out, void* stream);

/* ---- data-parallel exchange (src/mimic_runner.py:141-143: DistributedDataParallel; src/utils/main_util.py:43-62)
 * One process per GPU.  Rank 0 makes a 128-byte NCCL unique id (ghnd_comm_unique_id), ships it to the other
 * ranks by any side channel (the Python runner uses torch.distributed's store), every rank calls
 * ghnd_comm_init_from_unique_id on its own device.  ghnd_comm_allreduce_flat: in-place SUM over ranks of
 * the flat fp32 gradient buffer (the 1/world factor is FusedAdam's grad_scale); ghnd_comm_broadcast_flat:
 * rank `root`'s flat parameter buffer to everyone (what DDP's constructor does).  NCCL over NVLink 5 /
 * NVSwitch, enqueued on `stream`; libnccl.so.2 is resolved with dlopen at first use. */
typedef struct ghnd_comm ghnd_comm_t;
int ghnd_comm_unique_id(void* id128);
int ghnd_comm_init_from_unique_id(const void* id128, int world_size, int rank, ghnd_comm_t** comm);
int ghnd_comm_allreduce_flat(ghnd_comm_t* comm, float* buf, int64_t n, void* stream);
int ghnd_comm_broadcast_flat(ghnd_comm_t* comm, float* buf, int64_t n, int root, void* stream);
void ghnd_comm_destroy(ghnd_comm_t* comm);

#ifdef __cplusplus
}
#endif
#endif /* GHND_B200_H */
