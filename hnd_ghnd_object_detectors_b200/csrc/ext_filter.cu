// Neural-filter head `Ext4ResNet` (src/models/ext/classifier.py:16-37), inference only:
//   AdaptiveAvgPool2d(64,64) over the stem output -> Conv 4x4 s2 (+BN, ReLU) -> Conv 3x3 s2 (+BN, ReLU)
//   -> Conv 2x2 (+BN, ReLU) -> AdaptiveAvgPool2d(8,8) -> Linear(1024, 2) -> softmax.
// The first pool reads the whole 64-channel stem output once (HBM-bound, coalesced NHWC rows); everything
// after it works on a 64x64x64 fp32 tensor per image (a few hundred KB) with plain fp32 SIMT kernels.
#include "common.cuh"

namespace ghnd {

// x NHWC 16-bit [N,H,W,C] -> y NHWC fp32 [N,OH,OW,C]; bin i = [floor(i*H/OH), ceil((i+1)*H/OH))
// grid (OH, N), block 256: thread -> (8-channel group, output column) pairs
template <int FMT>
__global__ void __launch_bounds__(256)
    adaptive_avgpool_nhwc16_kernel(const uint4* __restrict__ x, float* __restrict__ y, int H, int W, int C8,
                                   int OH, int OW) {
  const int oh = blockIdx.x, n = blockIdx.y;
  const int h0 = (oh * H) / OH, h1 = ((oh + 1) * H + OH - 1) / OH;
  for (int v = threadIdx.x; v < OW * C8; v += blockDim.x) {
    const int cg = v % C8, ow = v / C8;
    const int w0 = (ow * W) / OW, w1 = ((ow + 1) * W + OW - 1) / OW;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int h = h0; h < h1; ++h)
      for (int w = w0; w < w1; ++w) {
        const uint4 q = ld_stream(x + (((int64_t)n * H + h) * W + w) * C8 + cg);
        const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack2_t<FMT>(u[e]);
          acc[2 * e] += f.x;
          acc[2 * e + 1] += f.y;
        }
      }
    const float inv = 1.f / (float)((h1 - h0) * (w1 - w0));
    float* o = y + ((((int64_t)n * OH + oh) * OW + ow) * C8 + cg) * 8;
    *reinterpret_cast<float4*>(o) = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
    *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
  }
}

// Direct convolution, no padding, fp32 NHWC: y[n,ho,wo,k] = act(scale[k] * sum_{r,s,c} x[n,ho*st+r,wo*st+s,c]
// * w[r][s][c][k] + shift[k]).  grid (Ho, N); the R input rows of an output row are staged in shared
// memory; a warp covers 32 consecutive k of one output column (x reads are broadcasts, w reads coalesced).
__global__ void __launch_bounds__(256)
    small_conv_f32_kernel(const float* __restrict__ x, const float* __restrict__ w,
                          const float* __restrict__ scale, const float* __restrict__ shift, int relu,
                          float* __restrict__ y, int H, int W, int C, int K, int R, int S, int stride, int Ho,
                          int Wo) {
  extern __shared__ float rows[];  // [R][W][C]
  const int ho = blockIdx.x, n = blockIdx.y;
  const int row_elems = W * C;
  const float* src = x + ((int64_t)n * H + (int64_t)ho * stride) * row_elems;
  for (int i = threadIdx.x * 4; i < R * row_elems; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(rows + i) = *reinterpret_cast<const float4*>(src + i);
  __syncthreads();
  for (int idx = threadIdx.x; idx < Wo * K; idx += blockDim.x) {
    const int k = idx % K, wo = idx / K;
    float acc = 0.f;
    for (int r = 0; r < R; ++r)
      for (int s = 0; s < S; ++s) {
        const float* xr = rows + (r * W + wo * stride + s) * C;
        const float* wr = w + ((int64_t)(r * S + s) * C) * K + k;
        for (int c = 0; c < C; c += 4) {
          const float4 xv = *reinterpret_cast<const float4*>(xr + c);
          acc = fmaf(xv.x, __ldg(wr + (int64_t)c * K), acc);
          acc = fmaf(xv.y, __ldg(wr + (int64_t)(c + 1) * K), acc);
          acc = fmaf(xv.z, __ldg(wr + (int64_t)(c + 2) * K), acc);
          acc = fmaf(xv.w, __ldg(wr + (int64_t)(c + 3) * K), acc);
        }
      }
    float v = fmaf(acc, scale ? __ldg(scale + k) : 1.f, shift ? __ldg(shift + k) : 0.f);
    if (relu) v = fmaxf(v, 0.f);
    y[(((int64_t)n * Ho + ho) * Wo + wo) * K + k] = v;
  }
}

// x NHWC fp32 [N,H,W,C] -> AdaptiveAvgPool2d(OH,OW) -> flatten in NCHW order (c*OH*OW + oh*OW + ow) ->
// Linear(C*OH*OW, n_out) -> optional softmax.  One block per image.
__global__ void __launch_bounds__(256)
    avgpool_linear_kernel(const float* __restrict__ x, int H, int W, int C, int OH, int OW,
                          const float* __restrict__ lw, const float* __restrict__ lb, int n_out, int softmax,
                          float* __restrict__ out) {
  extern __shared__ float pooled[];  // [C*OH*OW] in NCHW-flatten order, then [n_out] logits
  const int n = blockIdx.x;
  const int feat = C * OH * OW;
  for (int i = threadIdx.x; i < feat; i += blockDim.x) {
    const int ow = i % OW, oh = (i / OW) % OH, c = i / (OW * OH);
    const int h0 = (oh * H) / OH, h1 = ((oh + 1) * H + OH - 1) / OH;
    const int w0 = (ow * W) / OW, w1 = ((ow + 1) * W + OW - 1) / OW;
    float acc = 0.f;
    for (int h = h0; h < h1; ++h)
      for (int w = w0; w < w1; ++w) acc += x[(((int64_t)n * H + h) * W + w) * C + c];
    pooled[i] = acc / (float)((h1 - h0) * (w1 - w0));
  }
  __syncthreads();
  float* logits = pooled + feat;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < n_out; o += blockDim.x >> 5) {
    float acc = 0.f;
    for (int i = lane; i < feat; i += 32) acc = fmaf(pooled[i], __ldg(lw + (int64_t)o * feat + i), acc);
    acc = warp_sum(acc);
    if (lane == 0) logits[o] = acc + (lb ? lb[o] : 0.f);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (softmax) {
      float m = logits[0];
      for (int o = 1; o < n_out; ++o) m = fmaxf(m, logits[o]);
      float sum = 0.f;
      for (int o = 0; o < n_out; ++o) sum += expf(logits[o] - m);
      for (int o = 0; o < n_out; ++o) out[(int64_t)n * n_out + o] = expf(logits[o] - m) / sum;
    } else {
      for (int o = 0; o < n_out; ++o) out[(int64_t)n * n_out + o] = logits[o];
    }
  }
}

}  // namespace ghnd

extern "C" {
using namespace ghnd;

int ghnd_adaptive_avgpool_nhwc16(const void* x, int fmt, int N, int H, int W, int C, float* y, int OH, int OW,
                                 void* stream) {
  GHND_CHECK_ARG(x && y && (fmt == GHND_F16 || fmt == GHND_BF16), "adaptive_avgpool: bad argument");
  // OH > H (the pool up-samples a small map by repeating bins) is legal in nn.AdaptiveAvgPool2d too
  GHND_CHECK_ARG(N > 0 && N <= 65535 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && OH > 0 && OW > 0,
                 "adaptive_avgpool: bad geometry N=%d %dx%d C=%d -> %dx%d", N, H, W, C, OH, OW);
  dim3 grid((unsigned)OH, (unsigned)N);
  if (fmt == GHND_F16)
    adaptive_avgpool_nhwc16_kernel<GHND_F16><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, y, H, W, C / 8, OH, OW);
  else
    adaptive_avgpool_nhwc16_kernel<GHND_BF16><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, y, H, W, C / 8, OH, OW);
  GHND_LAUNCH_CHECK("adaptive_avgpool_nhwc16_kernel");
  return GHND_OK;
}

int ghnd_small_conv_f32(const float* x, const float* w_rsck, const float* scale, const float* shift, int relu,
                        float* y, int N, int H, int W, int C, int K, int R, int S, int stride, void* stream) {
  GHND_CHECK_ARG(x && w_rsck && y, "small_conv_f32: null argument");
  GHND_CHECK_ARG(N > 0 && N <= 65535 && C > 0 && C % 4 == 0 && K > 0 && R > 0 && S > 0 && stride > 0 && H >= R &&
                     W >= S,
                 "small_conv_f32: bad geometry N=%d %dx%d C=%d K=%d %dx%d s%d", N, H, W, C, K, R, S, stride);
  const int Ho = (H - R) / stride + 1, Wo = (W - S) / stride + 1;
  const size_t smem = (size_t)R * W * C * sizeof(float);
  GHND_CHECK_ARG(smem <= 200 * 1024, "small_conv_f32: %zu bytes of input rows exceed shared memory", smem);
  if (smem > 48 * 1024) {
    static size_t attr = 0;
    if (smem > attr) {
      GHND_CUDA(cudaFuncSetAttribute(small_conv_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr = 200 * 1024;
    }
  }
  dim3 grid((unsigned)Ho, (unsigned)N);
  small_conv_f32_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, w_rsck, scale, shift, relu, y, H, W, C, K, R, S,
                                                                 stride, Ho, Wo);
  GHND_LAUNCH_CHECK("small_conv_f32_kernel");
  return GHND_OK;
}

int ghnd_avgpool_linear(const float* x, int N, int H, int W, int C, int OH, int OW, const float* lw, const float* lb,
                        int n_out, int softmax, float* out, void* stream) {
  GHND_CHECK_ARG(x && lw && out, "avgpool_linear: null argument");
  GHND_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && OH > 0 && OW > 0 && n_out > 0 && n_out <= 64,
                 "avgpool_linear: bad geometry");
  const size_t smem = ((size_t)C * OH * OW + n_out) * sizeof(float);
  GHND_CHECK_ARG(smem <= 48 * 1024, "avgpool_linear: %zu bytes of pooled features exceed shared memory", smem);
  avgpool_linear_kernel<<<N, 256, smem, (cudaStream_t)stream>>>(x, H, W, C, OH, OW, lw, lb, n_out, softmax, out);
  GHND_LAUNCH_CHECK("avgpool_linear_kernel");
  return GHND_OK;
}
}
