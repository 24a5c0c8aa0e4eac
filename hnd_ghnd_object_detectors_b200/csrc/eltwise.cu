// HBM-bound helper kernels of the GHND path: layout boundary, stem input packing, max-pool,
// training-mode BatchNorm (stats / apply / backward), weight repacking, fused Adam.
#include <cstdlib>
#include "common.cuh"

namespace ghnd {

// developer knob for on-GPU sweeps (scripts/bench_kernels.py); unset = the tuned default
static inline int tune_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}
static unsigned* red_counter_slot();  // {claim, done} counters of a launch in flight (defined with the BN kernels)
static inline int grid_for(int64_t work_items, int threads, int per_sm = 8) {
  int64_t b = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)num_sms() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// The streaming BatchNorm kernels ask for the same (maximum) shared-memory split as the tensor-core kernels they run
// next to, so that a CTA never has to wait for an SM to change its L1 / shared-memory split.  Measured neutral on the
// B200 (scripts/debug/coresidency.py: 165 vs 169 us for dW GEMM + bn_bwd_reduce with and without) -- what decides
// whether a streaming kernel shares the SMs with a GEMM is the size of its CTAs (bn_bwd_apply_light_kernel below).
// GHND_STREAM_CARVEOUT=0: the driver's choice.
template <auto kernel>  // the kernel is a template argument: one `done` flag per instantiation
static inline void share_sm_with_gemm() {
  static const bool on = tune_int("GHND_STREAM_CARVEOUT", 1) != 0;
  static bool done = false;  // one flag per kernel instantiation
  if (on && !done) {
    cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         (int)cudaSharedmemCarveoutMaxShared);
    done = true;
  }
}

// ------------------------------------------------------------------------------------------------
// NCHW fp32 <-> NHWC 16-bit (module boundary only; 32x32 smem transpose)
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst,
                                    int fmt, int C, int64_t HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const int64_t p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? src[((int64_t)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t p = p0 + i;
    const int c = c0 + threadIdx.x;
    if (c < C && p < HW) dst[((int64_t)n * HW + p) * C + c] = float_to_h16(tile[threadIdx.x][i], fmt);
  }
}

__global__ void nhwc_to_nchw_kernel(const uint16_t* __restrict__ src, int fmt,
                                    float* __restrict__ dst, int C, int64_t HW) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t p = p0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? h16_to_float(src[((int64_t)n * HW + p) * C + c], fmt) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const int64_t p = p0 + threadIdx.x;
    if (c < C && p < HW) dst[((int64_t)n * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}

// ------------------------------------------------------------------------------------------------
// stem input: normalise + zero-pad + 4-channel pixel interleave, inside a 3-px zero frame
// ------------------------------------------------------------------------------------------------
__global__ void stem_pack_kernel(const float* __restrict__ img, int H, int W, float m0, float m1,
                                 float m2, float s0, float s1, float s2, uint2* __restrict__ dst,
                                 int fmt, int rows, int cols) {
  // grid (column blocks, rows): no per-pixel division (the flat-index version spent ~100 instructions per
  // pixel on a 64-bit i / cols and ran at 1.1 TB/s)
  const int r = blockIdx.y;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x) {
    const int64_t i = (int64_t)r * cols + c;
    const int h = r - 3, w = c - 3;
    uint2 o = make_uint2(0u, 0u);
    if (h >= 0 && h < H && w >= 0 && w < W) {
      const int64_t hw = (int64_t)H * W, at = (int64_t)h * W + w;
      // (image - mean) / std, true division like torchvision's normalize
      const float a = __fdiv_rn(__fsub_rn(img[at], m0), s0);
      const float b = __fdiv_rn(__fsub_rn(img[hw + at], m1), s1);
      const float d = __fdiv_rn(__fsub_rn(img[2 * hw + at], m2), s2);
      o.x = pack2(a, b, fmt);
      o.y = pack2(d, 0.f, fmt);
    }
    dst[i] = o;
  }
}

// Same packing with the transform's bilinear resize fused in: the reference
// normalises, then calls interpolate(scale_factor, mode='bilinear', align_corners=False)
// (src/models/org/rcnn.py:29-45), then zero-pads in batch_images.  Source index and tap weights
// follow ATen's upsample_bilinear2d (area_pixel_compute_source_index: src = rscale*(dst+0.5)-0.5,
// clamped at 0; second tap = first + (first < size-1)); the four taps are normalised first, like
// the reference's order of operations.
__global__ void stem_pack_resize_kernel(const float* __restrict__ img, int H, int W, int Ho, int Wo,
                                        float rh, float rw, float m0, float m1, float m2, float s0,
                                        float s1, float s2, uint2* __restrict__ dst, int fmt,
                                        int rows, int cols) {
  const int64_t hw = (int64_t)H * W;
  const int r = blockIdx.y;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x) {
    const int64_t i = (int64_t)r * cols + c;
    const int h = r - 3, w = c - 3;
    uint2 o = make_uint2(0u, 0u);
    if (h >= 0 && h < Ho && w >= 0 && w < Wo) {
      float hr = __fsub_rn(__fmul_rn(rh, (float)h + 0.5f), 0.5f);
      float wr = __fsub_rn(__fmul_rn(rw, (float)w + 0.5f), 0.5f);
      hr = hr < 0.f ? 0.f : hr;
      wr = wr < 0.f ? 0.f : wr;
      int h1 = (int)hr, w1 = (int)wr;
      h1 = h1 > H - 1 ? H - 1 : h1;
      w1 = w1 > W - 1 ? W - 1 : w1;
      const int hp = (h1 < H - 1) ? 1 : 0, wp = (w1 < W - 1) ? 1 : 0;
      const float h1l = hr - (float)h1, h0l = 1.f - h1l;
      const float w1l = wr - (float)w1, w0l = 1.f - w1l;
      const int64_t a00 = (int64_t)h1 * W + w1, a01 = a00 + wp, a10 = a00 + (int64_t)hp * W,
                    a11 = a10 + wp;
      const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
      float v[3];
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        const float* p = img + ch * hw;
        const float x00 = __fdiv_rn(__fsub_rn(p[a00], mean[ch]), sd[ch]);
        const float x01 = __fdiv_rn(__fsub_rn(p[a01], mean[ch]), sd[ch]);
        const float x10 = __fdiv_rn(__fsub_rn(p[a10], mean[ch]), sd[ch]);
        const float x11 = __fdiv_rn(__fsub_rn(p[a11], mean[ch]), sd[ch]);
        v[ch] = h0l * (w0l * x00 + w1l * x01) + h1l * (w0l * x10 + w1l * x11);
      }
      o.x = pack2(v[0], v[1], fmt);
      o.y = pack2(v[2], 0.f, fmt);
    }
    dst[i] = o;
  }
}

// The whole batch in ONE launch (grid z = image): per-image source pointer, size and resize factor travel in
// the kernel parameters.  rscale == 0: no resize.  Same arithmetic as the two single-image kernels above.
static constexpr int kPackMaxImages = 16;
struct PackBatch {
  const float* img[kPackMaxImages];
  int H[kPackMaxImages], W[kPackMaxImages], Ho[kPackMaxImages], Wo[kPackMaxImages];
  float rh[kPackMaxImages], rw[kPackMaxImages];
};
__global__ void stem_pack_batch_kernel(const __grid_constant__ PackBatch b, float m0, float m1, float m2, float s0,
                                       float s1, float s2, uint2* __restrict__ dst, int fmt, int rows, int cols,
                                       int n_index0) {
  const int z = blockIdx.z, r = blockIdx.y;
  const float* __restrict__ img = b.img[z];
  const int H = b.H[z], W = b.W[z], Ho = b.Ho[z], Wo = b.Wo[z];
  const float rh = b.rh[z], rw = b.rw[z];
  const bool resize = rh != 0.f;
  const int64_t hw = (int64_t)H * W;
  uint2* out = dst + (int64_t)(n_index0 + z) * rows * cols;
  const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cols; c += gridDim.x * blockDim.x) {
    const int h = r - 3, w = c - 3;
    uint2 o = make_uint2(0u, 0u);
    if (h >= 0 && h < Ho && w >= 0 && w < Wo) {
      float v[3];
      if (!resize) {
        const int64_t at = (int64_t)h * W + w;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) v[ch] = __fdiv_rn(__fsub_rn(img[ch * hw + at], mean[ch]), sd[ch]);
      } else {
        float hr = __fsub_rn(__fmul_rn(rh, (float)h + 0.5f), 0.5f);
        float wr = __fsub_rn(__fmul_rn(rw, (float)w + 0.5f), 0.5f);
        hr = hr < 0.f ? 0.f : hr;
        wr = wr < 0.f ? 0.f : wr;
        int h1 = (int)hr, w1 = (int)wr;
        h1 = h1 > H - 1 ? H - 1 : h1;
        w1 = w1 > W - 1 ? W - 1 : w1;
        const int hp = (h1 < H - 1) ? 1 : 0, wp = (w1 < W - 1) ? 1 : 0;
        const float h1l = hr - (float)h1, h0l = 1.f - h1l;
        const float w1l = wr - (float)w1, w0l = 1.f - w1l;
        const int64_t a00 = (int64_t)h1 * W + w1, a01 = a00 + wp, a10 = a00 + (int64_t)hp * W, a11 = a10 + wp;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          const float* p = img + ch * hw;
          const float x00 = __fdiv_rn(__fsub_rn(p[a00], mean[ch]), sd[ch]);
          const float x01 = __fdiv_rn(__fsub_rn(p[a01], mean[ch]), sd[ch]);
          const float x10 = __fdiv_rn(__fsub_rn(p[a10], mean[ch]), sd[ch]);
          const float x11 = __fdiv_rn(__fsub_rn(p[a11], mean[ch]), sd[ch]);
          v[ch] = h0l * (w0l * x00 + w1l * x01) + h1l * (w0l * x10 + w1l * x11);
        }
      }
      o.x = pack2(v[0], v[1], fmt);
      o.y = pack2(v[2], 0.f, fmt);
    }
    out[(int64_t)r * cols + c] = o;
  }
}

// ------------------------------------------------------------------------------------------------
// max-pool 3x3 s2 p1 (NHWC, 8 channels per thread) + backward fused with the ReLU mask
// ------------------------------------------------------------------------------------------------
// grid = (ceil(Wo*C8/256), Ho, N): no per-element division (C8 is a power of two: shift).
// The 3x3 scan stays in packed 16-bit pairs: one max + one compare-mask + one select per pair and
// tap (the first version unpacked to fp32 and was instruction-issue bound: ~900 SASS instructions
// per 8 channels, 55 us for a 172 MB stream).  Strict '>' keeps the first maximum in scan order,
// like torch's max_pool2d; padding taps are skipped.
template <int FMT>
__device__ __forceinline__ void pool_tap(uint32_t& best, uint32_t& idx, uint32_t v, uint32_t code) {
  if (FMT == GHND_F16) {
    const __half2 b = *reinterpret_cast<const __half2*>(&best);
    const __half2 x = *reinterpret_cast<const __half2*>(&v);
    const uint32_t m = __hgt2_mask(x, b);
    const __half2 r = __hmax2(b, x);
    best = *reinterpret_cast<const uint32_t*>(&r);
    idx = (idx & ~m) | (code & m);
  } else {
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&best);
    const __nv_bfloat162 x = *reinterpret_cast<const __nv_bfloat162*>(&v);
    const uint32_t m = __hgt2_mask(x, b);
    const __nv_bfloat162 r = __hmax2(b, x);
    best = *reinterpret_cast<const uint32_t*>(&r);
    idx = (idx & ~m) | (code & m);
  }
}
template <int FMT, bool kArg>
__global__ void __launch_bounds__(256)
    maxpool_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, uint2* __restrict__ argmax,
                   int N, int H, int W, int C8, int c8_shift, int Ho, int Wo, int XC8, int xoff8) {
  const int v = blockIdx.x * 256 + threadIdx.x;
  if (v >= Wo * C8) return;
  const int cg = v & (C8 - 1);
  const int wo = v >> c8_shift;
  const int ho = blockIdx.y;
  const int n = blockIdx.z;
  const int64_t i = (((int64_t)n * Ho + ho) * Wo + wo) * C8 + cg;
  const uint32_t ninf = FMT == GHND_F16 ? 0xfc00fc00u : 0xff80ff80u;
  uint32_t best[4] = {ninf, ninf, ninf, ninf};
  uint32_t idx[4] = {0u, 0u, 0u, 0u};  // two 16-bit tap codes per word
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const int h = 2 * ho - 1 + r;
    if (h < 0 || h >= H) continue;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int w = 2 * wo - 1 + s;
      if (w < 0 || w >= W) continue;
      const uint4 q = __ldg(x + (((int64_t)n * H + h) * W + w) * XC8 + xoff8 + cg);
      const uint32_t code = (uint32_t)(r * 3 + s) * 0x00010001u;
      pool_tap<FMT>(best[0], idx[0], q.x, code);
      pool_tap<FMT>(best[1], idx[1], q.y, code);
      pool_tap<FMT>(best[2], idx[2], q.z, code);
      pool_tap<FMT>(best[3], idx[3], q.w, code);
    }
  }
  y[i] = make_uint4(best[0], best[1], best[2], best[3]);
  if (kArg) {
    // The backward pass routes dy to the argmax tap AND applies the ReLU mask of the pooled tensor's
    // producer; at the argmax the producer's value IS the pooled value, so the mask is folded in here:
    // a non-positive maximum gets tap code 0xff ("no tap") and the backward kernel never has to re-read
    // the (4x larger) pre-pool tensor.
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      uint32_t pos;
      if (FMT == GHND_F16) pos = __hgt2_mask(*reinterpret_cast<const __half2*>(&best[e]), __float2half2_rn(0.f));
      else pos = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&best[e]), __float2bfloat162_rn(0.f));
      idx[e] = (idx[e] & pos) | (0x00ff00ffu & ~pos);
    }
    uint2 a;  // one byte per channel, channel order
    a.x = __byte_perm(idx[0], idx[1], 0x6420);
    a.y = __byte_perm(idx[2], idx[3], 0x6420);
    argmax[i] = a;
  }
}

// Backward of the pool fused with the ReLU mask of the stem output (carried by the argmax codes, see
// maxpool_kernel).  A thread owns the input pixel
// pair (h, 2m), (h, 2m+1) for one 8-channel group: the pair shares its windows (wo = m, m+1; ho =
// h/2 and, for odd h, h/2+1 -- uniform per block), so every window vector is loaded once and each
// (pixel, window) combination that can hold the pixel is tested exactly once: 1.5 (even rows) or 3
// (odd rows) tests per pixel, all loads issued before the first use.  (The per-pixel version
// tested 4 windows per pixel and was instruction-issue bound at ~40 % of HBM.)
template <int DYF>
__device__ __forceinline__ void pool_bwd_acc(float (&g)[8], const uint2 a, const uint4 d, uint32_t code) {
  const uint32_t du[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 f = unpack2_t<DYF>(du[e]);
    const uint32_t word = e < 2 ? a.x : a.y;
    const uint32_t a0 = (word >> (16 * (e & 1))) & 0xffu;
    const uint32_t a1 = (word >> (16 * (e & 1) + 8)) & 0xffu;
    g[2 * e] += a0 == code ? f.x : 0.f;
    g[2 * e + 1] += a1 == code ? f.y : 0.f;
  }
}
template <int DXF>
__device__ __forceinline__ uint4 pool_bwd_pack(const float (&g)[8]) {
  return make_uint4(pack2_t<DXF>(g[0], g[1]), pack2_t<DXF>(g[2], g[3]), pack2_t<DXF>(g[4], g[5]),
                    pack2_t<DXF>(g[6], g[7]));
}
template <int XF, int DYF, int DXF>
__global__ void __launch_bounds__(256)
    maxpool_bwd_kernel(const uint4* __restrict__ x, const uint2* __restrict__ argmax,
                       const uint4* __restrict__ dy, uint4* __restrict__ dx, int N, int H, int W,
                       int C8, int c8_shift, int Ho, int Wo, int XC8, int xoff8) {
  const int v = blockIdx.x * 256 + threadIdx.x;
  const int m = v >> c8_shift;  // pixel pair index along w
  const int w0 = 2 * m;
  if (w0 >= W) return;
  const int cg = v & (C8 - 1);
  const int h = blockIdx.y;
  const int n = blockIdx.z;
  const bool has1 = w0 + 1 < W;
  const int64_t i0 = (((int64_t)n * H + h) * W + w0) * C8 + cg;
  // x (the pre-pool tensor) is not read: its ReLU mask travels inside the argmax codes (0xff = masked)
  const int k = h >> 1;
  const bool odd = h & 1;  // uniform per block
  // window (row q, col c): q = 0 -> ho = k, q = 1 -> ho = k + 1 (odd rows only); c = 0 -> wo = m,
  // c = 1 -> wo = m + 1 (only the odd pixel of the pair lies in it)
  uint2 a[2][2];
  uint4 d[2][2];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int ho = k + q, wo = m + c;
      const bool ok = (q == 0 || odd) && ho < Ho && wo < Wo && (c == 0 || has1);
      const int64_t o = (((int64_t)n * Ho + ho) * Wo + wo) * C8 + cg;
      a[q][c] = ok ? __ldg(argmax + o) : make_uint2(0xffffffffu, 0xffffffffu);
      d[q][c] = ok ? __ldg(dy + o) : make_uint4(0u, 0u, 0u, 0u);
    }
  float g0[8], g1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) g0[j] = g1[j] = 0.f;
  // tap code inside a window = r*3 + s with r = h - (2ho-1), s = w - (2wo-1)
  const uint32_t r0 = odd ? 2u : 1u;  // row of this pixel inside window row q = 0
  pool_bwd_acc<DYF>(g0, a[0][0], d[0][0], r0 * 3 + 1);  // (w0 in window m: s = 1)
  pool_bwd_acc<DYF>(g1, a[0][0], d[0][0], r0 * 3 + 2);  // (w1 in window m: s = 2)
  pool_bwd_acc<DYF>(g1, a[0][1], d[0][1], r0 * 3 + 0);  // (w1 in window m+1: s = 0)
  if (odd) {                                            // second window row: r = 0
    pool_bwd_acc<DYF>(g0, a[1][0], d[1][0], 1);
    pool_bwd_acc<DYF>(g1, a[1][0], d[1][0], 2);
    pool_bwd_acc<DYF>(g1, a[1][1], d[1][1], 0);
  }
  dx[i0] = pool_bwd_pack<DXF>(g0);
  if (has1) dx[i0 + C8] = pool_bwd_pack<DXF>(g1);
}

// ------------------------------------------------------------------------------------------------
// BatchNorm (training): statistics, finalize, apply, backward
// ------------------------------------------------------------------------------------------------
// Block-level fold of per-thread channel partials (a[8] = sums, b[8] = second sums of the 8 channels of
// group cg = tid % C8) into `sums`, for C8 a power of two <= 32 and 256 threads: lanes of a warp that
// share a group are combined by xor-shuffles, every warp writes its totals into its OWN row of sh
// (plain stores), the block adds the 8 rows and issues ONE fp64 atomic per channel.  The previous
// version did 16 shared-memory float atomicAdds per thread: those are compare-and-swap spin loops in
// SASS (ATOMS.CAST.SPIN) with 8-32 threads per address -- the fixed ~15 us tail of the reduce kernels.
// sh: [8 warps][2*C] floats.  Requires C8 in {1, 2, 4, 8, 16, 32} (C <= 256).
__device__ __forceinline__ void block_channel_fold(float (&a)[8], float (&b)[8], int C8, float* sh,
                                                   double* __restrict__ sums) {
  const int C = C8 * 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = C8; o < 32; o <<= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
      b[j] += __shfl_xor_sync(0xffffffffu, b[j], o);
    }
  }
  float* row = sh + (size_t)warp * 2 * C;
  if (lane < C8) {
    const int cg = threadIdx.x % C8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      row[cg * 8 + j] = a[j];
      row[C + cg * 8 + j] = b[j];
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 2 * C; k += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sh[(size_t)w * 2 * C + k];
    atomicAdd(&sums[k], (double)t);
  }
}
// NHWC 16-bit: thread owns one 8-channel group and strides over pixels
template <bool kBwd>
__global__ void __launch_bounds__(256)
    bn_reduce_nhwc_kernel(const uint4* __restrict__ x, int x_fmt, const uint4* __restrict__ dy,
                          int dy_fmt, int64_t npix, int C, const float* __restrict__ scale_shift,
                          const float* __restrict__ mean_invstd, int relu,
                          double* __restrict__ sums) {
  extern __shared__ float sh[];  // [2*C], or [8 warps][2*C] on the fold path
  const int C8 = C >> 3;
  const int lanes = blockDim.x / C8;  // pixel lanes per block
  const int cg = threadIdx.x % C8;
  const int lane = threadIdx.x / C8;
  const bool fold = C8 <= 32 && (C8 & (C8 - 1)) == 0 && blockDim.x == 256;  // host sizes sh to match
  if (!fold) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
  }
  float a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
  float sc[8], shf[8], mu[8], is[8];
  if (kBwd) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = scale_shift[cg * 8 + j];
      shf[j] = scale_shift[C + cg * 8 + j];
      mu[j] = mean_invstd[cg * 8 + j];
      is[j] = mean_invstd[C + cg * 8 + j];
    }
  }
  if (lane < lanes) {
    for (int64_t p = (int64_t)blockIdx.x * lanes + lane; p < npix; p += (int64_t)gridDim.x * lanes) {
      const uint4 v = ld_stream(x + p * C8 + cg);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
      if (!kBwd) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack2(u[e], x_fmt);
          a[2 * e] += f.x;
          a[2 * e + 1] += f.y;
          b[2 * e] = fmaf(f.x, f.x, b[2 * e]);
          b[2 * e + 1] = fmaf(f.y, f.y, b[2 * e + 1]);
        }
      } else {
        const uint4 dv = ld_stream(dy + p * C8 + cg);
        const uint32_t du[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack2(u[e], x_fmt);
          float2 g = unpack2(du[e], dy_fmt);
          if (relu) {
            if (!(fmaf(f.x, sc[2 * e], shf[2 * e]) > 0.f)) g.x = 0.f;
            if (!(fmaf(f.y, sc[2 * e + 1], shf[2 * e + 1]) > 0.f)) g.y = 0.f;
          }
          a[2 * e] += g.x;
          a[2 * e + 1] += g.y;
          b[2 * e] = fmaf(g.x, (f.x - mu[2 * e]) * is[2 * e], b[2 * e]);
          b[2 * e + 1] = fmaf(g.y, (f.y - mu[2 * e + 1]) * is[2 * e + 1], b[2 * e + 1]);
        }
      }
    }
    if (!fold) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&sh[cg * 8 + j], a[j]);
        atomicAdd(&sh[C + cg * 8 + j], b[j]);
      }
    }
  }
  if (fold) {  // all 256 threads are pixel lanes here (256 % C8 == 0)
    block_channel_fold(a, b, C8, sh, sums);
    return;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&sums[i], (double)sh[i]);
}

// planar fp32 [N][C][HW]: grid (chunks, C, N)
template <bool kBwd>
__global__ void __launch_bounds__(256)
    bn_reduce_planar_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t hw,
                            int C, const float* __restrict__ scale_shift,
                            const float* __restrict__ mean_invstd, int relu,
                            double* __restrict__ sums) {
  const int c = blockIdx.y, n = blockIdx.z;
  const float* xp = x + ((int64_t)n * C + c) * hw;
  const float* dp = kBwd ? dy + ((int64_t)n * C + c) * hw : nullptr;
  float sc = 0, shf = 0, mu = 0, is = 0;
  if (kBwd) {
    sc = scale_shift[c];
    shf = scale_shift[C + c];
    mu = mean_invstd[c];
    is = mean_invstd[C + c];
  }
  double a = 0.0, b = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float v = xp[i];
    if (!kBwd) {
      a += v;
      b += (double)v * v;
    } else {
      float g = dp[i];
      if (relu && !(fmaf(v, sc, shf) > 0.f)) g = 0.f;
      a += g;
      b += (double)g * ((v - mu) * is);
    }
  }
  __shared__ double ra[8], rb[8];
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) {
    ra[threadIdx.x >> 5] = a;
    rb[threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0, tb = 0;
    for (int w = 0; w < 8; ++w) {
      ta += ra[w];
      tb += rb[w];
    }
    atomicAdd(&sums[c], ta);
    atomicAdd(&sums[C + c], tb);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float eps, float momentum, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int64_t* __restrict__ nbt,
                                   float* __restrict__ scale_shift, float* __restrict__ mean_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt != nullptr) *nbt += 1;
  if (c >= C) return;
  const double mean = sums[c] / count;
  double var = sums[C + c] / count - mean * mean;  // biased
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float sc = g * invstd;
  scale_shift[c] = sc;
  scale_shift[C + c] = b - (float)mean * sc;
  mean_invstd[c] = (float)mean;
  mean_invstd[C + c] = invstd;
  if (running_mean != nullptr) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// Forward statistics of a planar tensor AND the finalize step in one launch: every block adds its partial sums, the
// block that takes the last ticket (ctr[1], reset for the next launch) reads the totals back through L2 and does what
// bn_finalize_kernel does.  Saves one launch on the student's chain in the bottleneck section, where every launch
// queues behind a teacher conv that holds all SMs (in-graph timeline: bn_finalize started 34 us after the reduce
// had finished and the narrow conv after it another 41 us later).
__global__ void __launch_bounds__(256)
    bn_stats_finalize_planar_kernel(const float* __restrict__ x, int64_t hw, int C, double* __restrict__ sums,
                                    unsigned* __restrict__ ctr, double count, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, float eps, float momentum,
                                    float* __restrict__ running_mean, float* __restrict__ running_var,
                                    int64_t* __restrict__ nbt, float* __restrict__ scale_shift,
                                    float* __restrict__ mean_invstd) {
  const int c = blockIdx.y, n = blockIdx.z;
  const float* xp = x + ((int64_t)n * C + c) * hw;
  double a = 0.0, b = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = xp[i];
    a += v;
    b += (double)v * v;
  }
  __shared__ double ra[8], rb[8];
  __shared__ unsigned s_last;
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) {
    ra[threadIdx.x >> 5] = a;
    rb[threadIdx.x >> 5] = b;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0, tb = 0;
    for (int w = 0; w < 8; ++w) {
      ta += ra[w];
      tb += rb[w];
    }
    atomicAdd(&sums[c], ta);
    atomicAdd(&sums[C + c], tb);
    __threadfence();
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    s_last = atomicAdd(&ctr[1], 1u) == total - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (s_last == 0u) return;
  __threadfence();
  if (threadIdx.x == 0) {
    ctr[1] = 0u;
    if (nbt != nullptr) *nbt += 1;
  }
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    const double mean = __ldcg(&sums[k]) / count;
    double var = __ldcg(&sums[C + k]) / count - mean * mean;  // biased
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[k] : 1.f, bb = beta ? beta[k] : 0.f;
    const float sc = g * invstd;
    scale_shift[k] = sc;
    scale_shift[C + k] = bb - (float)mean * sc;
    mean_invstd[k] = (float)mean;
    mean_invstd[C + k] = invstd;
    if (running_mean != nullptr) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[k] = (1.f - momentum) * running_mean[k] + momentum * (float)mean;
      running_var[k] = (1.f - momentum) * running_var[k] + momentum * (float)unbiased;
    }
  }
}

__global__ void bn_eval_params_kernel(int C, const float* __restrict__ gamma,
                                      const float* __restrict__ beta, const float* __restrict__ rm,
                                      const float* __restrict__ rv, float eps,
                                      float* __restrict__ scale_shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.0f / sqrtf(rv[c] + eps);
  const float sc = (gamma ? gamma[c] : 1.f) * invstd;
  scale_shift[c] = sc;
  scale_shift[C + c] = (beta ? beta[c] : 0.f) - rm[c] * sc;
}

__global__ void __launch_bounds__(256)
    bn_apply_kernel(const uint4* __restrict__ x, int x_fmt, uint4* __restrict__ y, int y_fmt,
                    uint4* __restrict__ y2, int y2_fmt, int64_t nvec, int C8,
                    const float* __restrict__ scale_shift, int relu) {
  const int C = C8 * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % C8);
    const uint4 v = ld_stream(x + i);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4], o2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack2(u[e], x_fmt);
      const int c = cg * 8 + 2 * e;
      float a = fmaf(f.x, __ldg(scale_shift + c), __ldg(scale_shift + C + c));
      float b = fmaf(f.y, __ldg(scale_shift + c + 1), __ldg(scale_shift + C + c + 1));
      if (relu) {
        a = fmaxf(a, 0.f);
        b = fmaxf(b, 0.f);
      }
      o[e] = pack2(a, b, y_fmt);
      o2[e] = pack2(a, b, y2_fmt);
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
    if (y2 != nullptr) y2[i] = make_uint4(o2[0], o2[1], o2[2], o2[3]);
  }
}

__global__ void __launch_bounds__(256)
    bn_bwd_apply_nhwc_kernel(const uint4* __restrict__ dy, int dy_fmt, const uint4* __restrict__ x,
                             int x_fmt, uint4* __restrict__ dx, int dx_fmt, int64_t nvec, int C8,
                             double inv_count, const float* __restrict__ gamma,
                             const float* __restrict__ scale_shift,
                             const float* __restrict__ mean_invstd, int relu,
                             const double* __restrict__ sums) {
  const int C = C8 * 8;
  // per-channel constants staged once per block (the fp64 -> fp32 means are computed here, not per
  // element: the fp64 pipe of B200 is slow)
  extern __shared__ float s_c[];  // [6][C]: scale, shift, mean, invstd, mean_g, mean_gx
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_c[c] = scale_shift[c];
    s_c[C + c] = scale_shift[C + c];
    s_c[2 * C + c] = mean_invstd[c];
    s_c[3 * C + c] = mean_invstd[C + c];
    s_c[4 * C + c] = (float)(sums[c] * inv_count);
    s_c[5 * C + c] = (float)(sums[C + c] * inv_count);
  }
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % C8);
    const uint4 xv = ld_stream(x + i), dv = ld_stream(dy + i);
    const uint32_t xu[4] = {xv.x, xv.y, xv.z, xv.w}, du[4] = {dv.x, dv.y, dv.z, dv.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack2(xu[e], x_fmt);
      const float2 g = unpack2(du[e], dy_fmt);
      float r[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = cg * 8 + 2 * e + h;
        const float xv1 = h ? f.y : f.x;
        float gv = h ? g.y : g.x;
        const float sc = s_c[c], sf = s_c[C + c];
        if (relu && !(fmaf(xv1, sc, sf) > 0.f)) gv = 0.f;
        const float xhat = (xv1 - s_c[2 * C + c]) * s_c[3 * C + c];
        r[h] = sc * (gv - s_c[4 * C + c] - xhat * s_c[5 * C + c]);  // sc = gamma*invstd
      }
      o[e] = pack2(r[0], r[1], dx_fmt);
    }
    dx[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  (void)gamma;
}

__global__ void __launch_bounds__(256)
    bn_bwd_apply_planar_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                               float* __restrict__ dx, int64_t hw, int C, double inv_count,
                               const float* __restrict__ scale_shift,
                               const float* __restrict__ mean_invstd, int relu,
                               const double* __restrict__ sums) {
  const int c = blockIdx.y, n = blockIdx.z;
  const int64_t base = ((int64_t)n * C + c) * hw;
  const float sc = scale_shift[c], sf = scale_shift[C + c];
  const float mu = mean_invstd[c], is = mean_invstd[C + c];
  const float mg = (float)(sums[c] * inv_count), mgx = (float)(sums[C + c] * inv_count);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hw;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[base + i];
    float g = dy[base + i];
    if (relu && !(fmaf(v, sc, sf) > 0.f)) g = 0.f;
    dx[base + i] = sc * (g - mg - (v - mu) * is * mgx);
  }
}

__global__ void bn_param_grads_kernel(const double* __restrict__ sums, int C,
                                      float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dbeta) dbeta[c] = (float)sums[c];
  if (dgamma) dgamma[c] = (float)sums[C + c];
}


// ------------------------------------------------------------------------------------------------
// Fast paths of the three streaming BatchNorm kernels (C/8 divides 256): a thread always meets the
// same 8-channel group, so its per-channel constants live in registers, the loop carries no
// division, formats are compile-time, and several 16-byte loads are in flight per thread.
// ------------------------------------------------------------------------------------------------
// Optional "finalize" job of the apply kernel (training forward): derive scale / shift from the
// batch sums in the prologue -- every thread for its own 8 channels, same fp64 formulas as
// bn_finalize_kernel -- while block 0 also publishes scale_shift / mean_invstd for the backward pass
// and updates the running statistics.  Saves one tiny launch per BatchNorm on the critical path.
struct BnFin {
  const double* sums;  // nullptr: plain apply with the given scale_shift
  double count;
  const float* gamma;
  const float* beta;
  float eps, momentum;
  float* running_mean;
  float* running_var;
  int64_t* nbt;
  float* scale_shift_out;
  float* mean_invstd_out;
};
template <int XF, int YF, int Y2F, bool RELU>
__global__ void __launch_bounds__(256)
    bn_apply_fast_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, uint4* __restrict__ y2,
                         int64_t nvec, int C8, const float* __restrict__ scale_shift, const BnFin fin) {
  constexpr int U = 4;
  const int cg = threadIdx.x % C8;
  const int C = C8 * 8;
  float sc[8], sh[8];
  if (fin.sums != nullptr) {
    // one thread per channel does the fp64 math, the block shares the result through shared memory
    __shared__ float s_ss[2 * 2048];
    const bool publish = blockIdx.x == 0;
    if (publish && threadIdx.x == 0 && fin.nbt != nullptr) *fin.nbt += 1;
    for (int c = threadIdx.x; c < C; c += 256) {
      const double mean = fin.sums[c] / fin.count;
      double var = fin.sums[C + c] / fin.count - mean * mean;  // biased
      if (var < 0.0) var = 0.0;
      const float invstd = (float)(1.0 / sqrt(var + (double)fin.eps));
      const float g = fin.gamma ? __ldg(fin.gamma + c) : 1.f, b = fin.beta ? __ldg(fin.beta + c) : 0.f;
      const float scv = g * invstd, shv = b - (float)mean * scv;
      s_ss[c] = scv;
      s_ss[C + c] = shv;
      if (publish) {
        fin.scale_shift_out[c] = scv;
        fin.scale_shift_out[C + c] = shv;
        fin.mean_invstd_out[c] = (float)mean;
        fin.mean_invstd_out[C + c] = invstd;
        if (fin.running_mean != nullptr) {
          const double unbiased = fin.count > 1.0 ? var * fin.count / (fin.count - 1.0) : var;
          fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * (float)mean;
          fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * (float)unbiased;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = s_ss[cg * 8 + j];
      sh[j] = s_ss[C + cg * 8 + j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = __ldg(scale_shift + cg * 8 + j);
      sh[j] = __ldg(scale_shift + C + cg * 8 + j);
    }
  }
  const int64_t stride = (int64_t)gridDim.x * 256;
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  auto one = [&](const uint4 v, int64_t at) {
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4], o2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack2_t<XF>(u[e]);
      float a = fmaf(f.x, sc[2 * e], sh[2 * e]);
      float b = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]);
      if (RELU) {
        a = fmaxf(a, 0.f);
        b = fmaxf(b, 0.f);
      }
      o[e] = pack2_t<YF>(a, b);
      if (Y2F >= 0) o2[e] = pack2_t<(Y2F >= 0 ? Y2F : 0)>(a, b);
    }
    y[at] = make_uint4(o[0], o[1], o[2], o[3]);
    if (Y2F >= 0) y2[at] = make_uint4(o2[0], o2[1], o2[2], o2[3]);
  };
  for (; i + (U - 1) * stride < nvec; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) v[k] = ld_stream(x + i + k * stride);
#pragma unroll
    for (int k = 0; k < U; ++k) one(v[k], i + k * stride);
  }
  for (; i < nvec; i += stride) one(ld_stream(x + i), i);
}

// bn_apply_fast_kernel in the library-kernel shape (see bn_bwd_apply_light_kernel): 128-thread CTAs, 1024 vectors
// each; the finalize prologue (fp64, one or two channels per thread) runs in every CTA, block 0 publishes.
template <int XF, int YF, int Y2F, bool RELU>
__global__ void __launch_bounds__(128, 8)
    bn_apply_light_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, uint4* __restrict__ y2,
                          int64_t nvec, int C8, const float* __restrict__ scale_shift, const BnFin fin) {
  constexpr int U = 4, ROUNDS = 2;
  const int C = C8 * 8;  // <= 256 (host)
  __shared__ __align__(16) float s_ss[2][256];
  const bool publish = blockIdx.x == 0;
  if (fin.sums != nullptr) {
    if (publish && threadIdx.x == 0 && fin.nbt != nullptr) *fin.nbt += 1;
    for (int c = threadIdx.x; c < C; c += 128) {
      const double mean = fin.sums[c] / fin.count;
      double var = fin.sums[C + c] / fin.count - mean * mean;  // biased
      if (var < 0.0) var = 0.0;
      const float invstd = (float)(1.0 / sqrt(var + (double)fin.eps));
      const float g = fin.gamma ? __ldg(fin.gamma + c) : 1.f, b = fin.beta ? __ldg(fin.beta + c) : 0.f;
      const float scv = g * invstd, shv = b - (float)mean * scv;
      s_ss[0][c] = scv;
      s_ss[1][c] = shv;
      if (publish) {
        fin.scale_shift_out[c] = scv;
        fin.scale_shift_out[C + c] = shv;
        fin.mean_invstd_out[c] = (float)mean;
        fin.mean_invstd_out[C + c] = invstd;
        if (fin.running_mean != nullptr) {
          const double unbiased = fin.count > 1.0 ? var * fin.count / (fin.count - 1.0) : var;
          fin.running_mean[c] = (1.f - fin.momentum) * fin.running_mean[c] + fin.momentum * (float)mean;
          fin.running_var[c] = (1.f - fin.momentum) * fin.running_var[c] + fin.momentum * (float)unbiased;
        }
      }
    }
  } else {
    for (int c = threadIdx.x; c < C; c += 128) {
      s_ss[0][c] = __ldg(scale_shift + c);
      s_ss[1][c] = __ldg(scale_shift + C + c);
    }
  }
  __syncthreads();
  const int cg = threadIdx.x % C8;
  float sc[8], sh[8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float4 a = *reinterpret_cast<const float4*>(&s_ss[0][cg * 8 + 4 * h]);
    const float4 b = *reinterpret_cast<const float4*>(&s_ss[1][cg * 8 + 4 * h]);
    sc[4 * h] = a.x, sc[4 * h + 1] = a.y, sc[4 * h + 2] = a.z, sc[4 * h + 3] = a.w;
    sh[4 * h] = b.x, sh[4 * h + 1] = b.y, sh[4 * h + 2] = b.z, sh[4 * h + 3] = b.w;
  }
  auto one = [&](const uint4 v, int64_t at) {
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4], o2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack2_t<XF>(u[e]);
      float a = fmaf(f.x, sc[2 * e], sh[2 * e]);
      float b = fmaf(f.y, sc[2 * e + 1], sh[2 * e + 1]);
      if (RELU) {
        a = fmaxf(a, 0.f);
        b = fmaxf(b, 0.f);
      }
      o[e] = pack2_t<YF>(a, b);
      if (Y2F >= 0) o2[e] = pack2_t<(Y2F >= 0 ? Y2F : 0)>(a, b);
    }
    y[at] = make_uint4(o[0], o[1], o[2], o[3]);
    if (Y2F >= 0) y2[at] = make_uint4(o2[0], o2[1], o2[2], o2[3]);
  };
  int64_t i = (int64_t)blockIdx.x * (128 * U * ROUNDS) + threadIdx.x;
#pragma unroll 1
  for (int r = 0; r < ROUNDS; ++r, i += 128 * U) {
    if (i + (U - 1) * 128 < nvec) {
      uint4 v[U];
#pragma unroll
      for (int k = 0; k < U; ++k) v[k] = ld_stream(x + i + k * 128);
#pragma unroll
      for (int k = 0; k < U; ++k) one(v[k], i + k * 128);
    } else {
      for (int k = 0; k < U; ++k)
        if (i + k * 128 < nvec) one(ld_stream(x + i + k * 128), i + k * 128);
    }
  }
}

template <int XF, int GF, bool RELU>
__global__ void __launch_bounds__(256)
    bn_bwd_apply_fast_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x,
                             uint4* __restrict__ dx, int64_t nvec, int C8, float inv_count,
                             const float* __restrict__ scale_shift,
                             const float* __restrict__ mean_invstd,
                             const double* __restrict__ sums, float* __restrict__ dgamma,
                             float* __restrict__ dbeta, int sums_mode) {
  constexpr int U = 2;
  const int cg = threadIdx.x % C8;
  const int C = C8 * 8;
  // sum g'*xhat of channel c.  sums_mode 0: sums[C+c] is that sum.  sums_mode 1 (reductions fused
  // into the producing dgrad, which only sees the post-BN activation a = sc*x + sf): sums[C+c] =
  // sum g'*a, and (x - mu)*is = (a - sf - mu*sc) * is/sc, all in fp64 (the two terms cancel).
  auto sum_gxhat = [&](int c) -> double {
    const double s2 = sums[C + c];
    if (sums_mode == 0) return s2;
    const double scd = (double)__ldg(scale_shift + c), sfd = (double)__ldg(scale_shift + C + c);
    const double mud = (double)__ldg(mean_invstd + c), isd = (double)__ldg(mean_invstd + C + c);
    return (isd / scd) * (s2 - (sfd + mud * scd) * sums[c]);
  };
  // parameter gradients are the two sums themselves: block 0 copies them out (was a separate launch)
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < C; c += 256) {
      if (dbeta != nullptr) dbeta[c] = (float)sums[c];
      if (dgamma != nullptr) dgamma[c] = (float)sum_gxhat(c);
    }
  // dx = sc*(g' - mg - xhat*mgx) with xhat = (x-mu)*is  ==  sc*g' + x*k1 + k0
  float sc[8], sf[8], k1[8], k0[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cg * 8 + j;
    sc[j] = __ldg(scale_shift + c);
    sf[j] = __ldg(scale_shift + C + c);
    const float mu = __ldg(mean_invstd + c), is = __ldg(mean_invstd + C + c);
    const float mg = (float)sums[c] * inv_count, mgx = (float)sum_gxhat(c) * inv_count;
    k1[j] = -sc[j] * is * mgx;
    k0[j] = sc[j] * (mu * is * mgx - mg);
  }
  const int64_t stride = (int64_t)gridDim.x * 256;
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  auto one = [&](const uint4 xv, const uint4 dv, int64_t at) {
    const uint32_t xu[4] = {xv.x, xv.y, xv.z, xv.w}, du[4] = {dv.x, dv.y, dv.z, dv.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack2_t<XF>(xu[e]);
      float2 g = unpack2_t<GF>(du[e]);
      if (RELU) {
        if (!(fmaf(f.x, sc[2 * e], sf[2 * e]) > 0.f)) g.x = 0.f;
        if (!(fmaf(f.y, sc[2 * e + 1], sf[2 * e + 1]) > 0.f)) g.y = 0.f;
      }
      const float r0 = fmaf(sc[2 * e], g.x, fmaf(f.x, k1[2 * e], k0[2 * e]));
      const float r1 = fmaf(sc[2 * e + 1], g.y, fmaf(f.y, k1[2 * e + 1], k0[2 * e + 1]));
      o[e] = pack2_t<GF>(r0, r1);
    }
    dx[at] = make_uint4(o[0], o[1], o[2], o[3]);
  };
  for (; i + (U - 1) * stride < nvec; i += U * stride) {
    uint4 xv[U], dv[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      xv[k] = ld_stream(x + i + k * stride);
      dv[k] = ld_stream(dy + i + k * stride);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) one(xv[k], dv[k], i + k * stride);
  }
  for (; i < nvec; i += stride) one(ld_stream(x + i), ld_stream(dy + i), i);
}


// The same arithmetic as bn_bwd_apply_fast_kernel in the shape of a library copy kernel: MANY short-lived 128-thread
// CTAs (1024 vectors each, <= 64 registers) instead of a persistent grid.  scripts/debug/coresidency.py: next to the
// dW GEMM of the other stream (one 192-thread CTA per SM holding ~200 KB of shared memory) torch's copy kernel is
// absorbed completely (GEMM + copy take as long as the GEMM alone), while the persistent 256-thread / 80-register
// CTAs of the fast kernel mostly wait for the GEMM to finish: the block scheduler places small CTAs wherever
// registers are left and late CTAs of a persistent grid still own their static share of the tensor.  The
// per-channel constants are computed once per CTA (one or two channels per thread) and shared through 4 KB of smem.
template <int XF, int GF, bool RELU>
__global__ void __launch_bounds__(128, 8)
    bn_bwd_apply_light_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ x,
                              uint4* __restrict__ dx, int64_t nvec, int C8, float inv_count,
                              const float* __restrict__ scale_shift,
                              const float* __restrict__ mean_invstd,
                              const double* __restrict__ sums, float* __restrict__ dgamma,
                              float* __restrict__ dbeta, int sums_mode) {
  constexpr int U = 4, ROUNDS = 2;
  const int C = C8 * 8;  // <= 256 (host)
  __shared__ __align__(16) float s_k[4][256];  // sc, sf, k1, k0
  for (int c = threadIdx.x; c < C; c += 128) {
    const float scv = __ldg(scale_shift + c), sfv = __ldg(scale_shift + C + c);
    const float mu = __ldg(mean_invstd + c), is = __ldg(mean_invstd + C + c);
    const double s1 = sums[c];
    double s2 = sums[C + c];
    if (sums_mode != 0) s2 = ((double)is / (double)scv) * (s2 - ((double)sfv + (double)mu * (double)scv) * s1);
    if (blockIdx.x == 0) {  // parameter gradients are the two sums themselves
      if (dbeta != nullptr) dbeta[c] = (float)s1;
      if (dgamma != nullptr) dgamma[c] = (float)s2;
    }
    const float mg = (float)s1 * inv_count, mgx = (float)s2 * inv_count;
    s_k[0][c] = scv;
    s_k[1][c] = sfv;
    s_k[2][c] = -scv * is * mgx;
    s_k[3][c] = scv * (mu * is * mgx - mg);
  }
  __syncthreads();
  const int cg = threadIdx.x % C8;
  float sc[8], sf[8], k1[8], k0[8];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const float4 a = *reinterpret_cast<const float4*>(&s_k[0][cg * 8 + 4 * h]);
    const float4 b = *reinterpret_cast<const float4*>(&s_k[1][cg * 8 + 4 * h]);
    const float4 c = *reinterpret_cast<const float4*>(&s_k[2][cg * 8 + 4 * h]);
    const float4 d = *reinterpret_cast<const float4*>(&s_k[3][cg * 8 + 4 * h]);
    sc[4 * h] = a.x, sc[4 * h + 1] = a.y, sc[4 * h + 2] = a.z, sc[4 * h + 3] = a.w;
    sf[4 * h] = b.x, sf[4 * h + 1] = b.y, sf[4 * h + 2] = b.z, sf[4 * h + 3] = b.w;
    k1[4 * h] = c.x, k1[4 * h + 1] = c.y, k1[4 * h + 2] = c.z, k1[4 * h + 3] = c.w;
    k0[4 * h] = d.x, k0[4 * h + 1] = d.y, k0[4 * h + 2] = d.z, k0[4 * h + 3] = d.w;
  }
  auto one = [&](const uint4 xv, const uint4 dv, int64_t at) {
    const uint32_t xu[4] = {xv.x, xv.y, xv.z, xv.w}, du[4] = {dv.x, dv.y, dv.z, dv.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack2_t<XF>(xu[e]);
      float2 g = unpack2_t<GF>(du[e]);
      if (RELU) {
        if (!(fmaf(f.x, sc[2 * e], sf[2 * e]) > 0.f)) g.x = 0.f;
        if (!(fmaf(f.y, sc[2 * e + 1], sf[2 * e + 1]) > 0.f)) g.y = 0.f;
      }
      const float r0 = fmaf(sc[2 * e], g.x, fmaf(f.x, k1[2 * e], k0[2 * e]));
      const float r1 = fmaf(sc[2 * e + 1], g.y, fmaf(f.y, k1[2 * e + 1], k0[2 * e + 1]));
      o[e] = pack2_t<GF>(r0, r1);
    }
    dx[at] = make_uint4(o[0], o[1], o[2], o[3]);
  };
  int64_t i = (int64_t)blockIdx.x * (128 * U * ROUNDS) + threadIdx.x;
#pragma unroll 1
  for (int r = 0; r < ROUNDS; ++r, i += 128 * U) {
    if (i + (U - 1) * 128 < nvec) {
      uint4 xv[U], dv[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        xv[k] = ld_stream(x + i + k * 128);
        dv[k] = ld_stream(dy + i + k * 128);
      }
#pragma unroll
      for (int k = 0; k < U; ++k) one(xv[k], dv[k], i + k * 128);
    } else {
      for (int k = 0; k < U; ++k)
        if (i + k * 128 < nvec) one(ld_stream(x + i + k * 128), ld_stream(dy + i + k * 128), i + k * 128);
    }
  }
}

// backward reduction: sums[c] += sum g', sums[C+c] += sum g'*xhat
template <int XF, int GF, bool RELU>
__global__ void __launch_bounds__(256)
    bn_bwd_reduce_fast_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, int64_t nvec,
                              int C8, const float* __restrict__ scale_shift,
                              const float* __restrict__ mean_invstd, double* __restrict__ sums) {
  constexpr int U = 4;
  extern __shared__ float sh[];  // [8 warps][2*C]
  const int C = C8 * 8;
  const int cg = threadIdx.x % C8;
  float sc[8], sf[8], mu[8], is[8], a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cg * 8 + j;
    sc[j] = __ldg(scale_shift + c);
    sf[j] = __ldg(scale_shift + C + c);
    mu[j] = __ldg(mean_invstd + c);
    is[j] = __ldg(mean_invstd + C + c);
    a[j] = b[j] = 0.f;
  }
  const int64_t stride = (int64_t)gridDim.x * 256;
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  auto one = [&](const uint4 xv, const uint4 dv) {
    const uint32_t xu[4] = {xv.x, xv.y, xv.z, xv.w}, du[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack2_t<XF>(xu[e]);
      float2 g = unpack2_t<GF>(du[e]);
      if (RELU) {
        if (!(fmaf(f.x, sc[2 * e], sf[2 * e]) > 0.f)) g.x = 0.f;
        if (!(fmaf(f.y, sc[2 * e + 1], sf[2 * e + 1]) > 0.f)) g.y = 0.f;
      }
      a[2 * e] += g.x;
      a[2 * e + 1] += g.y;
      b[2 * e] = fmaf(g.x, (f.x - mu[2 * e]) * is[2 * e], b[2 * e]);
      b[2 * e + 1] = fmaf(g.y, (f.y - mu[2 * e + 1]) * is[2 * e + 1], b[2 * e + 1]);
    }
  };
  for (; i + (U - 1) * stride < nvec; i += U * stride) {
    uint4 xv[U], dv[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      xv[k] = ld_stream(x + i + k * stride);
      dv[k] = ld_stream(dy + i + k * stride);
    }
#pragma unroll
    for (int k = 0; k < U; ++k) one(xv[k], dv[k]);
  }
  for (; i < nvec; i += stride) one(ld_stream(x + i), ld_stream(dy + i));
  block_channel_fold(a, b, C8, sh, sums);
}

// bn_bwd_reduce for sharing the SMs with a GEMM: 128-thread CTAs that CLAIM chunks of 2048 vectors from a global
// counter until the tensor is exhausted.  Whatever subset of the grid is resident does all the work -- alone that is
// the whole grid (5 CTAs per SM, 8 vectors in flight per thread), next to the dW GEMM of the other stream the two or
// three CTAs per SM that fit beside it; CTAs that start late find nothing left and leave (a static share per CTA made
// the kernel wait for the GEMM: 143 us against 52 alone).  (f - mu) is accumulated and invstd applied once at the end,
// which keeps the per-thread constants to sc / sf / mu.  The next claim is in flight while a chunk is folded.
// ctr[0] = next chunk, ctr[1] = CTAs done; the last CTA to leave resets both for the next launch that uses the slot.
// sh: [4 warps][2*C] floats.
static constexpr int kRedChunk = 2048;  // vectors per claim: 128 threads x 4 x 4
template <int XF, int GF, bool RELU>
__global__ void __launch_bounds__(128, 5)
    bn_bwd_reduce_light_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, int64_t nvec,
                               int C8, unsigned n_chunks, unsigned* __restrict__ ctr,
                               const float* __restrict__ scale_shift,
                               const float* __restrict__ mean_invstd, double* __restrict__ sums) {
  extern __shared__ float sh[];
  __shared__ unsigned s_claim[2];
  const int C = C8 * 8;
  const int cg = threadIdx.x % C8;
  if (threadIdx.x == 0) s_claim[0] = atomicAdd(&ctr[0], 1u);
  float sc[8], sf[8], mu[8], a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cg * 8 + j;
    if (RELU) {
      sc[j] = __ldg(scale_shift + c);
      sf[j] = __ldg(scale_shift + C + c);
    }
    mu[j] = __ldg(mean_invstd + c);
    a[j] = b[j] = 0.f;
  }
  auto one = [&](const uint4 xv, const uint4 dv) {
    const uint32_t xu[4] = {xv.x, xv.y, xv.z, xv.w}, du[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack2_t<XF>(xu[e]);
      float2 g = unpack2_t<GF>(du[e]);
      if (RELU) {
        if (!(fmaf(f.x, sc[2 * e], sf[2 * e]) > 0.f)) g.x = 0.f;
        if (!(fmaf(f.y, sc[2 * e + 1], sf[2 * e + 1]) > 0.f)) g.y = 0.f;
      }
      a[2 * e] += g.x;
      a[2 * e + 1] += g.y;
      b[2 * e] = fmaf(g.x, f.x - mu[2 * e], b[2 * e]);
      b[2 * e + 1] = fmaf(g.y, f.y - mu[2 * e + 1], b[2 * e + 1]);
    }
  };
  __syncthreads();
  unsigned chunk = s_claim[0];
  bool any = false;
  int par = 1;
  while (chunk < n_chunks) {
    any = true;
    if (threadIdx.x == 0) s_claim[par] = atomicAdd(&ctr[0], 1u);  // next claim, read after the barrier below
    const int64_t base = (int64_t)chunk * kRedChunk + threadIdx.x;
    if (base - threadIdx.x + kRedChunk <= nvec) {
#pragma unroll 1
      for (int r = 0; r < 4; ++r) {
        const int64_t i = base + r * 512;
        uint4 xv[4], dv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          xv[k] = ld_stream(x + i + k * 128);
          dv[k] = ld_stream(dy + i + k * 128);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) one(xv[k], dv[k]);
      }
    } else {
      for (int64_t i = base; i < nvec; i += 128) one(ld_stream(x + i), ld_stream(dy + i));
    }
    __syncthreads();
    chunk = s_claim[par];
    par ^= 1;
  }
  if (any) {  // uniform per CTA
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] *= __ldg(mean_invstd + C + cg * 8 + j);
    // fold: lanes of one channel group, then the four warp rows, then one fp64 atomic per sum
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = C8; o < 32; o <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[j] += __shfl_xor_sync(0xffffffffu, a[j], o);
        b[j] += __shfl_xor_sync(0xffffffffu, b[j], o);
      }
    }
    float* row = sh + (size_t)warp * 2 * C;
    if (lane < C8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        row[cg * 8 + j] = a[j];
        row[C + cg * 8 + j] = b[j];
      }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 2 * C; k += 128) {
      const float t = (sh[k] + sh[2 * C + k]) + (sh[4 * C + k] + sh[6 * C + k]);
      atomicAdd(&sums[k], (double)t);
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctr[1], 1u) == gridDim.x - 1) {  // every CTA has made its last claim
      ctr[0] = 0u;
      ctr[1] = 0u;
    }
  }
}
// claim counters of the launches in flight: a launch takes the next slot (zero outside a launch)
static unsigned* red_counter_slot() {
  static unsigned* pool = nullptr;
  static int next = 0;
  constexpr int kSlots = 256;
  if (pool == nullptr) {
    if (cudaMalloc(&pool, kSlots * 2 * sizeof(unsigned)) != cudaSuccess) return nullptr;
    cudaMemset(pool, 0, kSlots * 2 * sizeof(unsigned));
  }
  unsigned* p = pool + 2 * next;
  next = (next + 1) % kSlots;
  return p;
}

static size_t bn_reduce_smem(int C) {  // bn_reduce_nhwc_kernel: [8 warps][2*C] on its fold path
  const int c8 = C / 8;
  return (size_t)((c8 <= 32 && (c8 & (c8 - 1)) == 0) ? 16 : 2) * C * sizeof(float);
}
static bool fast_c8(int C) { return C >= 8 && C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0; }

template <int XF, int YF, int Y2F>
static void launch_bn_apply_fast3(const uint4* x, uint4* y, uint4* y2, int64_t nvec, int C8,
                                  const float* ss, int relu, int grid, const BnFin& fin, cudaStream_t st) {
  static const int light = tune_int("GHND_BN_LIGHT", 3);  // A/B switch: 0 = the persistent kernels
  if ((light & 4) && C8 <= 32 && 128 % C8 == 0) {
    const unsigned blocks = (unsigned)((nvec + 1023) / 1024);
    share_sm_with_gemm<bn_apply_light_kernel<XF, YF, Y2F, true>>();
    share_sm_with_gemm<bn_apply_light_kernel<XF, YF, Y2F, false>>();
    if (relu) bn_apply_light_kernel<XF, YF, Y2F, true><<<blocks, 128, 0, st>>>(x, y, y2, nvec, C8, ss, fin);
    else bn_apply_light_kernel<XF, YF, Y2F, false><<<blocks, 128, 0, st>>>(x, y, y2, nvec, C8, ss, fin);
    return;
  }
  share_sm_with_gemm<bn_apply_fast_kernel<XF, YF, Y2F, true>>();
  share_sm_with_gemm<bn_apply_fast_kernel<XF, YF, Y2F, false>>();
  if (relu) bn_apply_fast_kernel<XF, YF, Y2F, true><<<grid, 256, 0, st>>>(x, y, y2, nvec, C8, ss, fin);
  else bn_apply_fast_kernel<XF, YF, Y2F, false><<<grid, 256, 0, st>>>(x, y, y2, nvec, C8, ss, fin);
}
template <int XF, int YF>
static void launch_bn_apply_fast2(int y2_fmt, const uint4* x, uint4* y, uint4* y2, int64_t nvec, int C8,
                                  const float* ss, int relu, int grid, const BnFin& fin, cudaStream_t st) {
  if (y2 == nullptr) launch_bn_apply_fast3<XF, YF, -1>(x, y, y2, nvec, C8, ss, relu, grid, fin, st);
  else if (y2_fmt == GHND_F16) launch_bn_apply_fast3<XF, YF, GHND_F16>(x, y, y2, nvec, C8, ss, relu, grid, fin, st);
  else launch_bn_apply_fast3<XF, YF, GHND_BF16>(x, y, y2, nvec, C8, ss, relu, grid, fin, st);
}
static void launch_bn_apply_fast(int x_fmt, int y_fmt, int y2_fmt, const uint4* x, uint4* y, uint4* y2,
                                 int64_t nvec, int C8, const float* ss, int relu, const BnFin& fin,
                                 cudaStream_t st) {
  const int grid = grid_for(nvec, 256 * 4, tune_int("GHND_BNAPP_PER_SM", 8));
  if (x_fmt == GHND_F16) {
    if (y_fmt == GHND_F16) launch_bn_apply_fast2<GHND_F16, GHND_F16>(y2_fmt, x, y, y2, nvec, C8, ss, relu, grid, fin, st);
    else launch_bn_apply_fast2<GHND_F16, GHND_BF16>(y2_fmt, x, y, y2, nvec, C8, ss, relu, grid, fin, st);
  } else {
    if (y_fmt == GHND_F16) launch_bn_apply_fast2<GHND_BF16, GHND_F16>(y2_fmt, x, y, y2, nvec, C8, ss, relu, grid, fin, st);
    else launch_bn_apply_fast2<GHND_BF16, GHND_BF16>(y2_fmt, x, y, y2, nvec, C8, ss, relu, grid, fin, st);
  }
}

__global__ void __launch_bounds__(256)
    convert16_kernel(const uint4* __restrict__ x, int x_fmt, uint4* __restrict__ y, int y_fmt,
                     int64_t nvec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = ld_stream(x + i);
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack2(u[e], x_fmt);
      o[e] = pack2(f.x, f.y, y_fmt);
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// FPN top-down step: y = lateral + nearest-upsample(coarse), NHWC 16-bit, 8 channels per
// thread.  torchvision FeaturePyramidNetwork: F.interpolate(last_inner, size=lateral.shape[-2:],
// mode="nearest") -> source index floor(dst * in / out) (= dst/2 for the exact 2x of /32-padded maps).
// ------------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256)
    upsample_add_kernel(const uint4* __restrict__ fine, const uint4* __restrict__ coarse,
                        uint4* __restrict__ y, int H, int W, int Hc, int Wc, int C8, int64_t nvec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t t = i;
    const int cg = (int)(t % C8);
    t /= C8;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const int n = (int)(t / H);
    const int hc = (int)(((int64_t)h * Hc) / H), wc = (int)(((int64_t)w * Wc) / W);
    const uint4 a = ld_stream(fine + i);
    const uint4 b = __ldg(coarse + (((int64_t)n * Hc + hc) * Wc + wc) * C8 + cg);
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, bu[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack2_t<FMT>(au[e]), g = unpack2_t<FMT>(bu[e]);
      o[e] = pack2_t<FMT>(f.x + g.x, f.y + g.y);
    }
    y[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// weight repack / gradient unpack
// ------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale, int O,
                                   int I, int R, int S, int transpose, uint16_t* __restrict__ dst,
                                   int fmt) {
  const int64_t total = (int64_t)O * I * R * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes dst
    int64_t t = i;
    int o, c, r, s;
    if (!transpose) {  // [O][R][S][I]
      c = (int)(t % I);
      t /= I;
      s = (int)(t % S);
      t /= S;
      r = (int)(t % R);
      o = (int)(t / R);
    } else {  // [I][R][S][O]
      o = (int)(t % O);
      t /= O;
      s = (int)(t % S);
      t /= S;
      r = (int)(t % R);
      c = (int)(t / R);
    }
    float v = w[(((int64_t)o * I + c) * R + r) * S + s];
    if (scale != nullptr) v *= scale[o];
    dst[i] = float_to_h16(v, fmt);
  }
}

struct WeightPackBatch {
  ghnd_pack_weight_desc_t d[GHND_PACK_MAX];
};
// blockIdx.y = tensor; the index arithmetic of pack_weight_kernel in 32 bits (a conv weight has < 2^31 elements;
// the 64-bit divisions made the first version take 20 us for 1.2 M elements)
__global__ void __launch_bounds__(256) pack_weights_kernel(const __grid_constant__ WeightPackBatch b) {
  const ghnd_pack_weight_desc_t& d = b.d[blockIdx.y];
  const unsigned O = (unsigned)d.O, I = (unsigned)d.I, R = (unsigned)d.R, S = (unsigned)d.S;
  const unsigned total = O * I * R * S;
  const unsigned RS = R * S;
  uint16_t* dst = static_cast<uint16_t*>(d.dst);
  const float* __restrict__ w = d.w_oihw;
  const float* __restrict__ scale = d.scale_o;
  const int fmt = d.dst_fmt;
  const bool transpose = d.transpose != 0;
  const unsigned inner = transpose ? O : I;  // fastest dst axis
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned t = i / inner, a = i - t * inner;  // a = c (forward layout) or o (transposed)
    const unsigned outer = t / RS, rs = t - outer * RS;   // outer = o (forward) or c (transposed)
    const unsigned o = transpose ? a : outer, c = transpose ? outer : a;
    float v = w[(o * I + c) * RS + rs];
    if (scale != nullptr) v *= scale[o];
    dst[i] = float_to_h16(v, fmt);
  }
}

__global__ void unpack_wgrad_kernel(const float* __restrict__ dw, float* __restrict__ dst, int O, int I,
                                    int R, int S, float alpha) {
  const int64_t total = (int64_t)O * I * R * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes dst (OIHW)
    int64_t t = i;
    const int s = (int)(t % S);
    t /= S;
    const int r = (int)(t % R);
    t /= R;
    const int c = (int)(t % I);
    const int o = (int)(t / I);
    dst[i] = alpha * dw[(((int64_t)o * R + r) * S + s) * I + c];
  }
}

// ------------------------------------------------------------------------------------------------
// fused Adam over a flat buffer (torch.optim.Adam, amsgrad=False)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                float* __restrict__ v, int64_t n, float beta1, float beta2, float omb1, float omb2,
                float eps, float wd, float gscale, float step_size, float inv_sqrt_bc2) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float grad = g[i] * gscale;
    const float pv = p[i];
    if (wd != 0.f) grad = fmaf(wd, pv, grad);
    const float mi = fmaf(beta1, m[i], omb1 * grad);
    const float vi = fmaf(beta2, v[i], omb2 * grad * grad);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] = pv - step_size * (mi / denom);
  }
}

}  // namespace ghnd

extern "C" {
using namespace ghnd;

static bool fmt16(int f) { return f == GHND_F16 || f == GHND_BF16; }

int ghnd_nchw_f32_to_nhwc16(const float* src, void* dst, int dst_fmt, int N, int C, int H, int W,
                            void* stream) {
  GHND_CHECK_ARG(src && dst && fmt16(dst_fmt) && N > 0 && C > 0 && H > 0 && W > 0,
                 "nchw_f32_to_nhwc16: bad argument");
  const int64_t hw = (int64_t)H * W;
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)N), block(32, 8);
  nchw_to_nhwc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(src, (uint16_t*)dst, dst_fmt, C, hw);
  GHND_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return GHND_OK;
}

int ghnd_nhwc16_to_nchw_f32(const void* src, int src_fmt, float* dst, int N, int C, int H, int W,
                            void* stream) {
  GHND_CHECK_ARG(src && dst && fmt16(src_fmt) && N > 0 && C > 0 && H > 0 && W > 0,
                 "nhwc16_to_nchw_f32: bad argument");
  const int64_t hw = (int64_t)H * W;
  dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)N), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const uint16_t*)src, src_fmt, dst, C,
                                                                hw);
  GHND_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return GHND_OK;
}

int ghnd_stem_pack_image(const float* img_chw, int H, int W, const float* mean, const float* std_,
                         void* dst, int dst_fmt, int n_index, int Hp, int Wp, void* stream) {
  GHND_CHECK_ARG(img_chw && mean && std_ && dst && fmt16(dst_fmt), "stem_pack_image: bad argument");
  GHND_CHECK_ARG(H > 0 && W > 0 && H <= Hp && W <= Wp && Hp % 2 == 0 && Wp % 8 == 0 && n_index >= 0,
                 "stem_pack_image: bad geometry H=%d W=%d Hp=%d Wp=%d", H, W, Hp, Wp);
  const int rows = Hp + 6, cols = Wp + 8;
  uint2* d = (uint2*)dst + (int64_t)n_index * rows * cols;
  GHND_CHECK_ARG(rows <= 65535, "stem_pack_image: %d rows exceed the grid", rows);
  stem_pack_kernel<<<dim3((unsigned)((cols + 255) / 256), (unsigned)rows), 256, 0, (cudaStream_t)stream>>>(
      img_chw, H, W, mean[0], mean[1], mean[2], std_[0], std_[1], std_[2], d, dst_fmt, rows, cols);
  GHND_LAUNCH_CHECK("stem_pack_kernel");
  return GHND_OK;
}

int ghnd_stem_pack_image_resized(const float* img_chw, int H, int W, int Ho, int Wo, float rscale_h,
                                 float rscale_w, const float* mean, const float* std_, void* dst,
                                 int dst_fmt, int n_index, int Hp, int Wp, void* stream) {
  GHND_CHECK_ARG(img_chw && mean && std_ && dst && fmt16(dst_fmt),
                 "stem_pack_image_resized: bad argument");
  GHND_CHECK_ARG(H > 0 && W > 0 && Ho > 0 && Wo > 0 && Ho <= Hp && Wo <= Wp && Hp % 2 == 0 &&
                     Wp % 8 == 0 && n_index >= 0 && rscale_h > 0.f && rscale_w > 0.f,
                 "stem_pack_image_resized: bad geometry H=%d W=%d Ho=%d Wo=%d Hp=%d Wp=%d", H, W, Ho,
                 Wo, Hp, Wp);
  const int rows = Hp + 6, cols = Wp + 8;
  uint2* d = (uint2*)dst + (int64_t)n_index * rows * cols;
  GHND_CHECK_ARG(rows <= 65535, "stem_pack_image_resized: %d rows exceed the grid", rows);
  stem_pack_resize_kernel<<<dim3((unsigned)((cols + 255) / 256), (unsigned)rows), 256, 0, (cudaStream_t)stream>>>(
      img_chw, H, W, Ho, Wo, rscale_h, rscale_w, mean[0], mean[1], mean[2], std_[0], std_[1], std_[2],
      d, dst_fmt, rows, cols);
  GHND_LAUNCH_CHECK("stem_pack_resize_kernel");
  return GHND_OK;
}

int ghnd_stem_pack_images(const float* const* imgs_chw, const int* H, const int* W, const int* Ho, const int* Wo,
                          const float* rscale, int n, const float* mean, const float* std_, void* dst, int dst_fmt,
                          int n_index0, int Hp, int Wp, void* stream) {
  GHND_CHECK_ARG(imgs_chw && H && W && Ho && Wo && rscale && mean && std_ && dst && fmt16(dst_fmt) && n > 0 &&
                     n_index0 >= 0,
                 "stem_pack_images: bad argument");
  GHND_CHECK_ARG(Hp > 0 && Wp > 0 && Hp % 2 == 0 && Wp % 8 == 0 && Hp + 6 <= 65535, "stem_pack_images: bad slot %dx%d",
                 Hp, Wp);
  const int rows = Hp + 6, cols = Wp + 8;
  for (int i0 = 0; i0 < n; i0 += kPackMaxImages) {
    const int m = n - i0 < kPackMaxImages ? n - i0 : kPackMaxImages;
    PackBatch b;
    memset(&b, 0, sizeof(b));
    for (int i = 0; i < m; ++i) {
      const int k = i0 + i;
      GHND_CHECK_ARG(imgs_chw[k] && H[k] > 0 && W[k] > 0 && Ho[k] > 0 && Wo[k] > 0 && Ho[k] <= Hp && Wo[k] <= Wp &&
                         rscale[k] >= 0.f && (rscale[k] != 0.f || (Ho[k] == H[k] && Wo[k] == W[k])),
                     "stem_pack_images: bad geometry of image %d (H=%d W=%d Ho=%d Wo=%d Hp=%d Wp=%d)", k, H[k], W[k],
                     Ho[k], Wo[k], Hp, Wp);
      b.img[i] = imgs_chw[k];
      b.H[i] = H[k];
      b.W[i] = W[k];
      b.Ho[i] = Ho[k];
      b.Wo[i] = Wo[k];
      b.rh[i] = b.rw[i] = rscale[k];
    }
    stem_pack_batch_kernel<<<dim3((unsigned)((cols + 255) / 256), (unsigned)rows, (unsigned)m), 256, 0,
                             (cudaStream_t)stream>>>(b, mean[0], mean[1], mean[2], std_[0], std_[1], std_[2],
                                                     (uint2*)dst, dst_fmt, rows, cols, n_index0 + i0);
    GHND_LAUNCH_CHECK("stem_pack_batch_kernel");
  }
  return GHND_OK;
}

int ghnd_maxpool3x3s2(const void* x, void* y, void* argmax, int fmt, int N, int H, int W, int C,
                      void* stream) {
  return ghnd_maxpool3x3s2_strided(x, C, 0, y, argmax, fmt, N, H, W, C, stream);
}

int ghnd_maxpool3x3s2_strided(const void* x, int x_channels, int x_channel_offset, void* y, void* argmax,
                              int fmt, int N, int H, int W, int C, void* stream) {
  GHND_CHECK_ARG(x && y && fmt16(fmt) && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0,
                 "maxpool3x3s2: bad argument");
  GHND_CHECK_ARG(x_channels % 8 == 0 && x_channel_offset % 8 == 0 && x_channel_offset >= 0 &&
                     x_channel_offset + C <= x_channels,
                 "maxpool3x3s2: channels [%d, %d) outside the %d-channel input", x_channel_offset,
                 x_channel_offset + C, x_channels);
  const int XC8 = x_channels / 8, xoff8 = x_channel_offset / 8;
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int C8 = C / 8;
  GHND_CHECK_ARG((C8 & (C8 - 1)) == 0 && Ho <= 65535 && N <= 65535,
                 "maxpool3x3s2: C/8 must be a power of two (C=%d)", C);
  int shift = 0;
  while ((1 << shift) < C8) ++shift;
  dim3 grid((unsigned)((Wo * C8 + 255) / 256), (unsigned)Ho, (unsigned)N);
  cudaStream_t st = (cudaStream_t)stream;
  const uint4* xp = (const uint4*)x;
  uint4* yp = (uint4*)y;
  uint2* ap = (uint2*)argmax;
  if (fmt == GHND_F16) {
    if (ap) maxpool_kernel<GHND_F16, true><<<grid, 256, 0, st>>>(xp, yp, ap, N, H, W, C8, shift, Ho, Wo, XC8, xoff8);
    else maxpool_kernel<GHND_F16, false><<<grid, 256, 0, st>>>(xp, yp, ap, N, H, W, C8, shift, Ho, Wo, XC8, xoff8);
  } else {
    if (ap) maxpool_kernel<GHND_BF16, true><<<grid, 256, 0, st>>>(xp, yp, ap, N, H, W, C8, shift, Ho, Wo, XC8, xoff8);
    else maxpool_kernel<GHND_BF16, false><<<grid, 256, 0, st>>>(xp, yp, ap, N, H, W, C8, shift, Ho, Wo, XC8, xoff8);
  }
  GHND_LAUNCH_CHECK("maxpool_kernel");
  return GHND_OK;
}

int ghnd_maxpool3x3s2_bwd(const void* x, int x_fmt, const void* argmax, const void* dy, int dy_fmt,
                          void* dx, int dx_fmt, int N, int H, int W, int C, void* stream) {
  return ghnd_maxpool3x3s2_bwd_strided(x, x_fmt, C, 0, argmax, dy, dy_fmt, dx, dx_fmt, N, H, W, C, stream);
}

int ghnd_maxpool3x3s2_bwd_strided(const void* x, int x_fmt, int x_channels, int x_channel_offset,
                                  const void* argmax, const void* dy, int dy_fmt, void* dx, int dx_fmt,
                                  int N, int H, int W, int C, void* stream) {
  GHND_CHECK_ARG(x && argmax && dy && dx && fmt16(x_fmt) && fmt16(dy_fmt) && fmt16(dx_fmt),
                 "maxpool3x3s2_bwd: bad argument");
  GHND_CHECK_ARG(x_channels % 8 == 0 && x_channel_offset % 8 == 0 && x_channel_offset >= 0 &&
                     x_channel_offset + C <= x_channels,
                 "maxpool3x3s2_bwd: channels [%d, %d) outside the %d-channel input", x_channel_offset,
                 x_channel_offset + C, x_channels);
  const int XC8 = x_channels / 8, xoff8 = x_channel_offset / 8;
  GHND_CHECK_ARG(N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "maxpool3x3s2_bwd: bad geometry");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int C8 = C / 8;
  GHND_CHECK_ARG((C8 & (C8 - 1)) == 0 && H <= 65535 && N <= 65535,
                 "maxpool3x3s2_bwd: C/8 must be a power of two (C=%d)", C);
  int shift = 0;
  while ((1 << shift) < C8) ++shift;
  dim3 grid((unsigned)((((W + 1) / 2) * C8 + 255) / 256), (unsigned)H, (unsigned)N);
  cudaStream_t st = (cudaStream_t)stream;
#define GHND_POOL_BWD(XF, DYF, DXF)                                                                \
  maxpool_bwd_kernel<XF, DYF, DXF><<<grid, 256, 0, st>>>((const uint4*)x, (const uint2*)argmax,    \
                                                         (const uint4*)dy, (uint4*)dx, N, H, W, C8, \
                                                         shift, Ho, Wo, XC8, xoff8)
  const int combo = (x_fmt == GHND_F16 ? 4 : 0) | (dy_fmt == GHND_F16 ? 2 : 0) | (dx_fmt == GHND_F16 ? 1 : 0);
  switch (combo) {
    case 0: GHND_POOL_BWD(GHND_BF16, GHND_BF16, GHND_BF16); break;
    case 1: GHND_POOL_BWD(GHND_BF16, GHND_BF16, GHND_F16); break;
    case 2: GHND_POOL_BWD(GHND_BF16, GHND_F16, GHND_BF16); break;
    case 3: GHND_POOL_BWD(GHND_BF16, GHND_F16, GHND_F16); break;
    case 4: GHND_POOL_BWD(GHND_F16, GHND_BF16, GHND_BF16); break;
    case 5: GHND_POOL_BWD(GHND_F16, GHND_BF16, GHND_F16); break;
    case 6: GHND_POOL_BWD(GHND_F16, GHND_F16, GHND_BF16); break;
    default: GHND_POOL_BWD(GHND_F16, GHND_F16, GHND_F16); break;
  }
#undef GHND_POOL_BWD
  GHND_LAUNCH_CHECK("maxpool_bwd_kernel");
  return GHND_OK;
}

static int bn_geom_ok(int planar, int C) {
  if (planar) return C > 0;
  // NHWC: one thread per 8-channel group, 256 threads must be a multiple of the group count
  return C >= 8 && C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0;
}

int ghnd_bn_stats(const void* x, int fmt, int planar, int N, int64_t hw, int C, double* sums,
                  void* stream) {
  GHND_CHECK_ARG(x && sums && N > 0 && hw > 0, "bn_stats: bad argument");
  const bool zeroed = (planar & GHND_SUMS_ZEROED) != 0;  // the caller zeroed `sums` (see the header)
  planar &= ~GHND_SUMS_ZEROED;
  GHND_CHECK_ARG(bn_geom_ok(planar, C), "bn_stats: unsupported channel count %d", C);
  cudaStream_t st = (cudaStream_t)stream;
  if (!zeroed) GHND_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
  if (planar) {
    dim3 grid((unsigned)grid_for(hw, 256 * 8, 2), (unsigned)C, (unsigned)N);
    bn_reduce_planar_kernel<false><<<grid, 256, 0, st>>>((const float*)x, nullptr, hw, C, nullptr,
                                                         nullptr, 0, sums);
  } else {
    GHND_CHECK_ARG(fmt16(fmt), "bn_stats: bad format");
    const int64_t npix = (int64_t)N * hw;
    const int lanes = 256 / (C / 8);
    bn_reduce_nhwc_kernel<false><<<grid_for(npix, lanes * 8, 4), 256, bn_reduce_smem(C), st>>>(
        (const uint4*)x, fmt, nullptr, 0, npix, C, nullptr, nullptr, 0, sums);
  }
  GHND_LAUNCH_CHECK("bn_reduce_kernel");
  return GHND_OK;
}

int ghnd_bn_stats_finalize(const void* x, int fmt, int planar, int N, int64_t hw, int C, double* sums,
                           const float* gamma, const float* beta, float eps, float momentum, float* running_mean,
                           float* running_var, int64_t* num_batches_tracked, float* scale_shift,
                           float* mean_invstd, void* stream) {
  GHND_CHECK_ARG(x && sums && scale_shift && mean_invstd && N > 0 && hw > 0 && C > 0,
                 "bn_stats_finalize: bad argument");
  GHND_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr),
                 "bn_stats_finalize: running_mean/var must both be given or both be null");
  const bool zeroed = (planar & GHND_SUMS_ZEROED) != 0;
  unsigned* ctr = (planar & ~GHND_SUMS_ZEROED) ? red_counter_slot() : nullptr;
  if (ctr == nullptr) {  // NHWC (or no counter pool): the two launches
    int rc = ghnd_bn_stats(x, fmt, planar, N, hw, C, sums, stream);
    if (rc != GHND_OK) return rc;
    return ghnd_bn_finalize(sums, (int64_t)N * hw, C, gamma, beta, eps, momentum, running_mean, running_var,
                            num_batches_tracked, scale_shift, mean_invstd, stream);
  }
  GHND_CHECK_ARG(bn_geom_ok(1, C), "bn_stats_finalize: unsupported channel count %d", C);
  cudaStream_t st = (cudaStream_t)stream;
  if (!zeroed) GHND_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
  dim3 grid((unsigned)grid_for(hw, 256 * 8, 2), (unsigned)C, (unsigned)N);
  bn_stats_finalize_planar_kernel<<<grid, 256, 0, st>>>(
      (const float*)x, hw, C, sums, ctr, (double)N * (double)hw, gamma, beta, eps, momentum, running_mean,
      running_var, num_batches_tracked, scale_shift, mean_invstd);
  GHND_LAUNCH_CHECK("bn_stats_finalize_planar_kernel");
  return GHND_OK;
}

int ghnd_bn_finalize(const double* sums, int64_t count, int C, const float* gamma,
                     const float* beta, float eps, float momentum, float* running_mean,
                     float* running_var, int64_t* num_batches_tracked, float* scale_shift,
                     float* mean_invstd, void* stream) {
  GHND_CHECK_ARG(sums && scale_shift && mean_invstd && count > 0 && C > 0,
                 "bn_finalize: bad argument");
  GHND_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr),
                 "bn_finalize: running_mean/var must both be given or both be null");
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      sums, (double)count, C, gamma, beta, eps, momentum, running_mean, running_var,
      num_batches_tracked, scale_shift, mean_invstd);
  GHND_LAUNCH_CHECK("bn_finalize_kernel");
  return GHND_OK;
}

int ghnd_bn_eval_params(int C, const float* gamma, const float* beta, const float* running_mean,
                        const float* running_var, float eps, float* scale_shift, void* stream) {
  GHND_CHECK_ARG(C > 0 && running_mean && running_var && scale_shift, "bn_eval_params: bad argument");
  bn_eval_params_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
      C, gamma, beta, running_mean, running_var, eps, scale_shift);
  GHND_LAUNCH_CHECK("bn_eval_params_kernel");
  return GHND_OK;
}

int ghnd_bn_apply(const void* x, int x_fmt, void* y, int y_fmt, void* y2, int y2_fmt, int64_t npix,
                  int C, const float* scale_shift, int relu, void* stream) {
  GHND_CHECK_ARG(x && y && scale_shift && fmt16(x_fmt) && fmt16(y_fmt) && npix > 0 && C > 0 &&
                     C % 8 == 0 && (y2 == nullptr || fmt16(y2_fmt)),
                 "bn_apply: bad argument");
  const int64_t nvec = npix * (C / 8);
  if (fast_c8(C)) {
    BnFin none;
    memset(&none, 0, sizeof(none));
    launch_bn_apply_fast(x_fmt, y_fmt, y2_fmt, (const uint4*)x, (uint4*)y, (uint4*)y2, nvec, C / 8,
                         scale_shift, relu, none, (cudaStream_t)stream);
  } else
    bn_apply_kernel<<<grid_for(nvec, 256 * 2), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)x, x_fmt, (uint4*)y, y_fmt, (uint4*)y2, y2_fmt, nvec, C / 8, scale_shift, relu);
  GHND_LAUNCH_CHECK("bn_apply_kernel");
  return GHND_OK;
}

int ghnd_bn_finalize_apply(const void* x, int x_fmt, void* y, int y_fmt, void* y2, int y2_fmt,
                           int64_t npix, int C, int relu, const double* sums, int64_t count,
                           const float* gamma, const float* beta, float eps, float momentum,
                           float* running_mean, float* running_var, int64_t* num_batches_tracked,
                           float* scale_shift, float* mean_invstd, void* stream) {
  GHND_CHECK_ARG(sums && scale_shift && mean_invstd && count > 0, "bn_finalize_apply: bad argument");
  GHND_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr),
                 "bn_finalize_apply: running_mean/var must both be given or both be null");
  if (!fast_c8(C)) {  // generic widths: the two separate kernels
    int rc = ghnd_bn_finalize(sums, count, C, gamma, beta, eps, momentum, running_mean, running_var,
                              num_batches_tracked, scale_shift, mean_invstd, stream);
    if (rc != GHND_OK) return rc;
    return ghnd_bn_apply(x, x_fmt, y, y_fmt, y2, y2_fmt, npix, C, scale_shift, relu, stream);
  }
  GHND_CHECK_ARG(x && y && fmt16(x_fmt) && fmt16(y_fmt) && npix > 0 && (y2 == nullptr || fmt16(y2_fmt)),
                 "bn_finalize_apply: bad tensor argument");
  BnFin fin;
  fin.sums = sums;
  fin.count = (double)count;
  fin.gamma = gamma;
  fin.beta = beta;
  fin.eps = eps;
  fin.momentum = momentum;
  fin.running_mean = running_mean;
  fin.running_var = running_var;
  fin.nbt = num_batches_tracked;
  fin.scale_shift_out = scale_shift;
  fin.mean_invstd_out = mean_invstd;
  launch_bn_apply_fast(x_fmt, y_fmt, y2_fmt, (const uint4*)x, (uint4*)y, (uint4*)y2, npix * (C / 8), C / 8,
                       nullptr, relu, fin, (cudaStream_t)stream);
  GHND_LAUNCH_CHECK("bn_apply_fast_kernel(finalize)");
  return GHND_OK;
}

int ghnd_upsample_add(const void* fine, const void* coarse, void* y, int fmt, int N, int H, int W, int Hc,
                      int Wc, int C, void* stream) {
  GHND_CHECK_ARG(fine && coarse && y && fmt16(fmt), "upsample_add: bad argument");
  GHND_CHECK_ARG(N > 0 && H > 0 && W > 0 && Hc > 0 && Wc > 0 && Hc <= H && Wc <= W && C > 0 && C % 8 == 0,
                 "upsample_add: bad geometry N=%d %dx%d <- %dx%d C=%d", N, H, W, Hc, Wc, C);
  const int64_t nvec = (int64_t)N * H * W * (C / 8);
  const int grid = grid_for(nvec, 256, 16);
  if (fmt == GHND_F16)
    upsample_add_kernel<GHND_F16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)fine, (const uint4*)coarse, (uint4*)y, H, W, Hc, Wc, C / 8, nvec);
  else
    upsample_add_kernel<GHND_BF16><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)fine, (const uint4*)coarse, (uint4*)y, H, W, Hc, Wc, C / 8, nvec);
  GHND_LAUNCH_CHECK("upsample_add_kernel");
  return GHND_OK;
}

int ghnd_convert16(const void* x, int x_fmt, void* y, int y_fmt, int64_t n, void* stream) {
  GHND_CHECK_ARG(x && y && fmt16(x_fmt) && fmt16(y_fmt) && n > 0 && n % 8 == 0,
                 "convert16: bad argument (n must be a positive multiple of 8)");
  convert16_kernel<<<grid_for(n / 8, 256 * 2), 256, 0, (cudaStream_t)stream>>>(
      (const uint4*)x, x_fmt, (uint4*)y, y_fmt, n / 8);
  GHND_LAUNCH_CHECK("convert16_kernel");
  return GHND_OK;
}

int ghnd_bn_bwd_reduce(const void* dy, int dy_fmt, const void* x, int x_fmt, int planar, int N,
                       int64_t hw, int C, const float* scale_shift, const float* mean_invstd,
                       int relu, double* sums, void* stream) {
  GHND_CHECK_ARG(dy && x && scale_shift && mean_invstd && sums && N > 0 && hw > 0,
                 "bn_bwd_reduce: bad argument");
  const bool zeroed = (planar & GHND_SUMS_ZEROED) != 0;  // the caller zeroed `sums` (see the header)
  planar &= ~GHND_SUMS_ZEROED;
  GHND_CHECK_ARG(bn_geom_ok(planar, C), "bn_bwd_reduce: unsupported channel count %d", C);
  cudaStream_t st = (cudaStream_t)stream;
  if (!zeroed) GHND_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
  if (planar) {
    dim3 grid((unsigned)grid_for(hw, 256 * 8, 2), (unsigned)C, (unsigned)N);
    bn_reduce_planar_kernel<true><<<grid, 256, 0, st>>>((const float*)x, (const float*)dy, hw, C,
                                                        scale_shift, mean_invstd, relu, sums);
  } else {
    GHND_CHECK_ARG(fmt16(dy_fmt) && fmt16(x_fmt), "bn_bwd_reduce: bad format");
    const int64_t npix = (int64_t)N * hw;
    const int lanes = 256 / (C / 8);
    const int c8 = C / 8;
    if (x_fmt == GHND_F16 && dy_fmt == GHND_BF16 && c8 <= 32 && (c8 & (c8 - 1)) == 0) {  // the engine's combination: fast path
      const int64_t nvec = npix * (C / 8);
      // 3 CTAs/SM x 8 warps x 8 x 16-byte loads in flight; fewer CTAs also means fewer same-address
      // fp64 atomics in the tail (2*C per CTA)
      static const int light = tune_int("GHND_BN_LIGHT", 3);  // A/B switch: 0 = the persistent kernels
      unsigned* ctr = (light & 2) ? red_counter_slot() : nullptr;
      if (ctr != nullptr) {
        const unsigned n_chunks = (unsigned)((nvec + kRedChunk - 1) / kRedChunk);
        int64_t blocks = (int64_t)num_sms() * 5;
        if (blocks > n_chunks) blocks = n_chunks;
        const size_t smem = (size_t)8 * C * sizeof(float);
        share_sm_with_gemm<bn_bwd_reduce_light_kernel<GHND_F16, GHND_BF16, true>>();
        share_sm_with_gemm<bn_bwd_reduce_light_kernel<GHND_F16, GHND_BF16, false>>();
        if (relu)
          bn_bwd_reduce_light_kernel<GHND_F16, GHND_BF16, true><<<(unsigned)blocks, 128, smem, st>>>(
              (const uint4*)x, (const uint4*)dy, nvec, C / 8, n_chunks, ctr, scale_shift, mean_invstd, sums);
        else
          bn_bwd_reduce_light_kernel<GHND_F16, GHND_BF16, false><<<(unsigned)blocks, 128, smem, st>>>(
              (const uint4*)x, (const uint4*)dy, nvec, C / 8, n_chunks, ctr, scale_shift, mean_invstd, sums);
        GHND_LAUNCH_CHECK("bn_bwd_reduce_light_kernel");
        return GHND_OK;
      }
      const int grid = grid_for(nvec, 256 * 8, tune_int("GHND_BNRED_PER_SM", 3));
      share_sm_with_gemm<bn_bwd_reduce_fast_kernel<GHND_F16, GHND_BF16, true>>();
      share_sm_with_gemm<bn_bwd_reduce_fast_kernel<GHND_F16, GHND_BF16, false>>();
      if (relu)
        bn_bwd_reduce_fast_kernel<GHND_F16, GHND_BF16, true><<<grid, 256, 16 * C * sizeof(float), st>>>(
            (const uint4*)x, (const uint4*)dy, nvec, C / 8, scale_shift, mean_invstd, sums);
      else
        bn_bwd_reduce_fast_kernel<GHND_F16, GHND_BF16, false><<<grid, 256, 16 * C * sizeof(float), st>>>(
            (const uint4*)x, (const uint4*)dy, nvec, C / 8, scale_shift, mean_invstd, sums);
    } else {
      bn_reduce_nhwc_kernel<true><<<grid_for(npix, lanes * 8, 4), 256, bn_reduce_smem(C), st>>>(
          (const uint4*)x, x_fmt, (const uint4*)dy, dy_fmt, npix, C, scale_shift, mean_invstd, relu,
          sums);
    }
  }
  GHND_LAUNCH_CHECK("bn_bwd_reduce_kernel");
  return GHND_OK;
}

static int bn_bwd_apply_impl(const void* dy, int dy_fmt, const void* x, int x_fmt, void* dx, int dx_fmt,
                             int planar, int N, int64_t hw, int C, const float* gamma,
                             const float* scale_shift, const float* mean_invstd, int relu,
                             const double* sums, float* dgamma, float* dbeta, int sums_mode,
                             void* stream) {
  GHND_CHECK_ARG(dy && x && dx && scale_shift && mean_invstd && sums && N > 0 && hw > 0,
                 "bn_bwd_apply: bad argument");
  GHND_CHECK_ARG(sums_mode == 0 || (sums_mode == 1 && !planar && x_fmt == GHND_F16 &&
                                    dy_fmt == GHND_BF16 && dx_fmt == GHND_BF16),
                 "bn_bwd_apply: sums_mode 1 needs the NHWC f16-activation / bf16-gradient path");
  GHND_CHECK_ARG(bn_geom_ok(planar, C), "bn_bwd_apply: unsupported channel count %d", C);
  cudaStream_t st = (cudaStream_t)stream;
  const double inv_count = 1.0 / ((double)N * (double)hw);
  if (planar) {
    dim3 grid((unsigned)grid_for(hw, 256 * 4, 2), (unsigned)C, (unsigned)N);
    bn_bwd_apply_planar_kernel<<<grid, 256, 0, st>>>((const float*)dy, (const float*)x, (float*)dx, hw,
                                                     C, inv_count, scale_shift, mean_invstd, relu,
                                                     sums);
  } else {
    GHND_CHECK_ARG(fmt16(dy_fmt) && fmt16(x_fmt) && fmt16(dx_fmt), "bn_bwd_apply: bad format");
    const int64_t nvec = (int64_t)N * hw * (C / 8);
    if (x_fmt == GHND_F16 && dy_fmt == GHND_BF16 && dx_fmt == GHND_BF16) {  // engine's combination
      static const int light = tune_int("GHND_BN_LIGHT", 3);  // A/B switch: 0 = the persistent kernels
      if ((light & 1) && C <= 256 && 128 % (C / 8) == 0) {
        const int64_t blocks = (nvec + 1023) / 1024;
        share_sm_with_gemm<bn_bwd_apply_light_kernel<GHND_F16, GHND_BF16, true>>();
        share_sm_with_gemm<bn_bwd_apply_light_kernel<GHND_F16, GHND_BF16, false>>();
        if (relu)
          bn_bwd_apply_light_kernel<GHND_F16, GHND_BF16, true><<<(unsigned)blocks, 128, 0, st>>>(
              (const uint4*)dy, (const uint4*)x, (uint4*)dx, nvec, C / 8, (float)inv_count, scale_shift,
              mean_invstd, sums, dgamma, dbeta, sums_mode);
        else
          bn_bwd_apply_light_kernel<GHND_F16, GHND_BF16, false><<<(unsigned)blocks, 128, 0, st>>>(
              (const uint4*)dy, (const uint4*)x, (uint4*)dx, nvec, C / 8, (float)inv_count, scale_shift,
              mean_invstd, sums, dgamma, dbeta, sums_mode);
        GHND_LAUNCH_CHECK("bn_bwd_apply_light_kernel");
        return GHND_OK;
      }
      const int grid = grid_for(nvec, 256 * 4, tune_int("GHND_BNBWD_PER_SM", 6));
      share_sm_with_gemm<bn_bwd_apply_fast_kernel<GHND_F16, GHND_BF16, true>>();
      share_sm_with_gemm<bn_bwd_apply_fast_kernel<GHND_F16, GHND_BF16, false>>();
      if (relu)
        bn_bwd_apply_fast_kernel<GHND_F16, GHND_BF16, true><<<grid, 256, 0, st>>>(
            (const uint4*)dy, (const uint4*)x, (uint4*)dx, nvec, C / 8, (float)inv_count, scale_shift,
            mean_invstd, sums, dgamma, dbeta, sums_mode);
      else
        bn_bwd_apply_fast_kernel<GHND_F16, GHND_BF16, false><<<grid, 256, 0, st>>>(
            (const uint4*)dy, (const uint4*)x, (uint4*)dx, nvec, C / 8, (float)inv_count, scale_shift,
            mean_invstd, sums, dgamma, dbeta, sums_mode);
      GHND_LAUNCH_CHECK("bn_bwd_apply_fast_kernel");
      return GHND_OK;  // dgamma / dbeta written by the kernel's block 0
    } else {
      bn_bwd_apply_nhwc_kernel<<<grid_for(nvec, 256 * 2), 256, 6 * C * sizeof(float), st>>>(
          (const uint4*)dy, dy_fmt, (const uint4*)x, x_fmt, (uint4*)dx, dx_fmt, nvec, C / 8, inv_count,
          gamma, scale_shift, mean_invstd, relu, sums);
    }
  }
  GHND_LAUNCH_CHECK("bn_bwd_apply_kernel");
  if (dgamma != nullptr || dbeta != nullptr) {
    bn_param_grads_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, C, dgamma, dbeta);
    GHND_LAUNCH_CHECK("bn_param_grads_kernel");
  }
  return GHND_OK;
}

int ghnd_bn_bwd_apply(const void* dy, int dy_fmt, const void* x, int x_fmt, void* dx, int dx_fmt,
                      int planar, int N, int64_t hw, int C, const float* gamma,
                      const float* scale_shift, const float* mean_invstd, int relu,
                      const double* sums, float* dgamma, float* dbeta, void* stream) {
  return bn_bwd_apply_impl(dy, dy_fmt, x, x_fmt, dx, dx_fmt, planar, N, hw, C, gamma, scale_shift,
                           mean_invstd, relu, sums, dgamma, dbeta, 0, stream);
}

int ghnd_bn_bwd_apply_fused_sums(const void* dy, int dy_fmt, const void* x, int x_fmt, void* dx,
                                 int dx_fmt, int N, int64_t hw, int C, const float* gamma,
                                 const float* scale_shift, const float* mean_invstd, int relu,
                                 const double* sums, float* dgamma, float* dbeta, void* stream) {
  GHND_CHECK_ARG(fast_c8(C), "bn_bwd_apply_fused_sums: unsupported channel count %d", C);
  return bn_bwd_apply_impl(dy, dy_fmt, x, x_fmt, dx, dx_fmt, 0, N, hw, C, gamma, scale_shift,
                           mean_invstd, relu, sums, dgamma, dbeta, 1, stream);
}

int ghnd_pack_weight(const float* w_oihw, const float* scale_o, int O, int I, int R, int S,
                     int transpose, void* dst, int dst_fmt, void* stream) {
  GHND_CHECK_ARG(w_oihw && dst && fmt16(dst_fmt) && O > 0 && I > 0 && R > 0 && S > 0,
                 "pack_weight: bad argument");
  const int64_t total = (int64_t)O * I * R * S;
  pack_weight_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      w_oihw, scale_o, O, I, R, S, transpose, (uint16_t*)dst, dst_fmt);
  GHND_LAUNCH_CHECK("pack_weight_kernel");
  return GHND_OK;
}

int ghnd_pack_weights(const ghnd_pack_weight_desc_t* descs, int n, void* stream) {
  GHND_CHECK_ARG(descs && n > 0 && n <= GHND_PACK_MAX, "pack_weights: 1..%d tensors per launch (got %d)",
                 GHND_PACK_MAX, n);
  WeightPackBatch b;
  memset(&b, 0, sizeof(b));
  int64_t largest = 0;
  for (int k = 0; k < n; ++k) {
    const ghnd_pack_weight_desc_t& d = descs[k];
    GHND_CHECK_ARG(d.w_oihw && d.dst && fmt16(d.dst_fmt) && d.O > 0 && d.I > 0 && d.R > 0 && d.S > 0,
                   "pack_weights: bad argument in entry %d", k);
    b.d[k] = d;
    const int64_t total = (int64_t)d.O * d.I * d.R * d.S;
    GHND_CHECK_ARG(total < ((int64_t)1 << 31), "pack_weights: entry %d has too many elements", k);
    if (total > largest) largest = total;
  }
  int gx = (int)((largest + 255) / 256);
  if (gx > 2 * num_sms()) gx = 2 * num_sms();
  pack_weights_kernel<<<dim3((unsigned)gx, (unsigned)n), 256, 0, (cudaStream_t)stream>>>(b);
  GHND_LAUNCH_CHECK("pack_weights_kernel");
  return GHND_OK;
}

int ghnd_unpack_wgrad(const float* dw_orsi, float* dst_oihw, int O, int I, int R, int S,
                      float alpha, void* stream) {
  GHND_CHECK_ARG(dw_orsi && dst_oihw && O > 0 && I > 0 && R > 0 && S > 0, "unpack_wgrad: bad argument");
  const int64_t total = (int64_t)O * I * R * S;
  unpack_wgrad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dw_orsi, dst_oihw, O, I,
                                                                              R, S, alpha);
  GHND_LAUNCH_CHECK("unpack_wgrad_kernel");
  return GHND_OK;
}

int ghnd_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                   double lr, double beta1, double beta2, double eps, double weight_decay,
                   double grad_scale, int step, void* stream) {
  GHND_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n > 0 && step >= 1,
                 "adam_step: bad argument");
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2 = 1.0 - pow(beta2, (double)step);
  const float step_size = (float)(lr / bc1);
  const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  adam_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(
      param, grad, exp_avg, exp_avg_sq, n, (float)beta1, (float)beta2, (float)(1.0 - beta1),
      (float)(1.0 - beta2), (float)eps, (float)weight_decay, (float)grad_scale, step_size,
      inv_sqrt_bc2);
  GHND_LAUNCH_CHECK("adam_kernel");
  return GHND_OK;
}
}
