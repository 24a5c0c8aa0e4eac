// 8-bit affine quantizer / dequantizer of the bottleneck tensor (HBM-bound, bit-exact).
// Arithmetic follows src/myutils/pytorch/tensor_util.py:8-22 of the reference operation by
// operation with explicit round-to-nearest intrinsics (no FMA contraction, no fast division).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace ghnd {

static constexpr int kQThreads = 256;
static constexpr int kQMaxBlocks = 1024;

// torch.min/max propagate NaN: min.NaN / max.NaN do exactly that in ONE instruction (FMNMX.NAN; the
// compare-and-select form was 4-5 instructions per element and made the min/max passes issue-bound)
__device__ __forceinline__ float nan_min(float a, float b) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float nan_max(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

__device__ __forceinline__ void block_minmax(float& mn, float& mx) {
  __shared__ float s_mn[32], s_mx[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) {
    s_mn[w] = mn;
    s_mx[w] = mx;
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  mn = (l < nw) ? s_mn[l] : s_mn[0];
  mx = (l < nw) ? s_mx[l] : s_mx[0];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = nan_min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = nan_max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  __syncthreads();
}

// pass 1: per-block min / max partials (no atomics -> deterministic, no init needed)
__global__ void __launch_bounds__(kQThreads)
    quant_minmax_kernel(const float* __restrict__ x, int64_t n, int vec_ok,
                        float* __restrict__ partial) {
  float mn = INFINITY, mx = -INFINITY;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (vec_ok) {
    const int64_t n4 = n >> 2;
    const uint4* x4 = reinterpret_cast<const uint4*>(x);
    for (int64_t i = tid; i < n4; i += nthreads) {
      uint4 v = ld_stream(x4 + i);
      float a = __uint_as_float(v.x), b = __uint_as_float(v.y), c = __uint_as_float(v.z),
            d = __uint_as_float(v.w);
      mn = nan_min(nan_min(mn, a), nan_min(b, nan_min(c, d)));
      mx = nan_max(nan_max(mx, a), nan_max(b, nan_max(c, d)));
    }
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += nthreads) {
    float a = x[i];
    mn = nan_min(mn, a);
    mx = nan_max(mx, a);
  }
  block_minmax(mn, mx);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = mn;
    partial[2 * blockIdx.x + 1] = mx;
  }
}

struct QParams {
  float scale;
  int zero_point;
  float mn, mx;
};

// scale / zero-point from the tensor's min and max, operation by operation as the reference
__device__ __forceinline__ void derive_qparams(float mn, float mx, float qmax, float inv_range,
                                               int scale_mode, float& scale, int& zp) {
  // scale = (max - min) / (qmax - qmin)        tensor_util.py:12
  const float range = __fsub_rn(mx, mn);
  scale = scale_mode == GHND_QSCALE_RECIP ? __fmul_rn(range, inv_range) : __fdiv_rn(range, qmax);
  // initial_zero_point = qmin - min / scale    tensor_util.py:13
  const float izp = __fsub_rn(0.0f, __fdiv_rn(mn, scale));
  // clamp to [qmin, qmax] then int() truncation   tensor_util.py:14-15 ; NaN -> marker
  if (izp < 0.0f)
    zp = 0;
  else if (izp > qmax)
    zp = (int)qmax;
  else if (izp != izp)
    zp = INT_MIN;
  else
    zp = (int)izp;  // cvt.rzi
}
// qx = zero_point + x / scale ; clamp ; round-half-even ; byte   tensor_util.py:16-17
__device__ __forceinline__ uint32_t quant1(float v, float zpf, float scale, float qmax) {
  float t = __fadd_rn(zpf, __fdiv_rn(v, scale));
  t = t < 0.0f ? 0.0f : (t > qmax ? qmax : t);
  return (uint32_t)(int)rintf(t) & 0xffu;
}
// Exact-result fast path.  The IEEE division per element (~20 issue slots with its slow path) kept the
// quantizer at 30-40 % of HBM.  Here t' = RN(zp + x * RN(1/scale)) (one FFMA) replaces the reference's
// t = RN(zp + RN(x / scale)).  For a quotient inside [-256, 512]: |x*r - x/scale| <= 512 * 2^-24 = 3.1e-5,
// the reference's quotient rounding <= 2^-16, the two final roundings <= 2^-16 each -> |t' - t| < 7.7e-5;
// quotients outside clamp to the same end.  The byte is round-half-even(clamp(t)); it can only differ from
// the reference's when t' lies within that error of a rounding boundary k + 0.5, so an element whose t'
// comes within kQuantGuard = 2^-12 (3x the bound) of a boundary -- 0.05 % of elements -- is redone with the
// reference's exact sequence.  Rounding and the float -> byte conversion use the 1.5 * 2^23 magic constant
// (round-half-even in the FADD itself, the integer sits in the low mantissa byte): no F2I / FRND on the
// conversion pipe.  ~34 instructions per 4 elements: 4 x (FFMA, 2 FMNMX, 3 FADD) + |d| maximum, one
// compare and one branch per vector, 3 PRMT.
static constexpr float kQuantMagic = 12582912.0f;  // 1.5 * 2^23
static constexpr float kQuantGuard = 0.5f - 0.000244140625f;
struct QFast {
  float zpf, scale, rscale, qmax;
  bool ok;  // scale is a normal number with a finite reciprocal: the error bound above holds
};
__device__ __forceinline__ QFast make_qfast(float zpf, float scale, float qmax) {
  QFast f;
  f.zpf = zpf;
  f.scale = scale;
  f.qmax = qmax;
  f.rscale = __frcp_rn(scale);
  f.ok = scale >= 1e-30f && scale <= 1e30f;
  return f;
}
__device__ __forceinline__ uint32_t quant4(uint4 v, const QFast& f) {
  const float x[4] = {__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
  uint32_t m[4];
  float d[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float t = __fmaf_rn(x[e], f.rscale, f.zpf);
    t = fminf(fmaxf(t, 0.0f), f.qmax);            // NaN -> 0, like (int)rintf(NaN) & 0xff
    const float r = __fadd_rn(t, kQuantMagic);    // low mantissa byte = round-half-even(t)
    m[e] = __float_as_uint(r);
    d[e] = fabsf(__fsub_rn(t, __fsub_rn(r, kQuantMagic)));  // distance to the nearest integer, exact
  }
  if (!f.ok || fmaxf(fmaxf(d[0], d[1]), fmaxf(d[2], d[3])) > kQuantGuard) {  // rare: redo the doubtful elements
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (!f.ok || d[e] > kQuantGuard) m[e] = quant1(x[e], f.zpf, f.scale, f.qmax);
  }
  return __byte_perm(__byte_perm(m[0], m[1], 0x0040), __byte_perm(m[2], m[3], 0x0040), 0x5410);
}

// grid-stride quantization of 4-element vectors [first, n4), four loads in flight per thread
template <typename I>
__device__ __forceinline__ void quant_stream(const uint4* __restrict__ x4, uint32_t* __restrict__ q4, I i,
                                             const I nthreads, const I n4, const QFast& qf) {
  for (; i + 3 * nthreads < n4; i += 4 * nthreads) {
    const uint4 a = ld_stream(x4 + i), b = ld_stream(x4 + i + nthreads), c = ld_stream(x4 + i + 2 * nthreads),
                d = ld_stream(x4 + i + 3 * nthreads);
    q4[i] = quant4(a, qf);
    q4[i + nthreads] = quant4(b, qf);
    q4[i + 2 * nthreads] = quant4(c, qf);
    q4[i + 3 * nthreads] = quant4(d, qf);
  }
  for (; i < n4; i += nthreads) q4[i] = quant4(ld_stream(x4 + i), qf);
}

// pass 2: every block folds the partials (L2 resident), derives scale / zero-point, quantizes.
__global__ void __launch_bounds__(kQThreads)
    quant_apply_kernel(const float* __restrict__ x, int64_t n, int vec_ok,
                       const float* __restrict__ partial, int n_partial, float qmax,
                       float inv_range, int scale_mode, uint8_t* __restrict__ q,
                       QParams* __restrict__ qp) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  // first batch of loads in flight while the block folds the min/max partials (the prologue is ~10 % of
  // this 15-20 us kernel: every block starts with it)
  const int64_t n4 = vec_ok ? (n >> 2) : 0;
  const uint4* x4 = reinterpret_cast<const uint4*>(x);
  constexpr int kPre = 4;  // 64 B per thread in flight across the prologue (8 measured no faster, 60 registers)
  uint4 pv[kPre];
  const bool pre = tid + (kPre - 1) * nthreads < n4;
  if (pre) {
#pragma unroll
    for (int k = 0; k < kPre; ++k) pv[k] = ld_stream(x4 + tid + k * nthreads);
  }
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < n_partial; i += blockDim.x) {
    mn = nan_min(mn, partial[2 * i]);
    mx = nan_max(mx, partial[2 * i + 1]);
  }
  block_minmax(mn, mx);
  float scale;
  int zp;
  derive_qparams(mn, mx, qmax, inv_range, scale_mode, scale, zp);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    qp->scale = scale;
    qp->zero_point = zp;
    qp->mn = mn;
    qp->mx = mx;
  }
  const float zpf = (float)zp;
  const QFast qf = make_qfast(zpf, scale, qmax);
  int64_t done = 0;
  if (vec_ok) {
    // lane-contiguous accesses: a warp reads 512 contiguous bytes per load instruction and writes 128
    // contiguous bytes per store (the first version gave every lane its own 64-byte run: each 32-byte
    // sector was requested by two different instructions and, with L1 allocation off, fetched twice)
    uint32_t* q4 = reinterpret_cast<uint32_t*>(q);
    if (pre) {
#pragma unroll
      for (int k = 0; k < kPre; ++k) q4[tid + k * nthreads] = quant4(pv[k], qf);
    }
    const int64_t first = pre ? tid + kPre * nthreads : tid;
    // 32-bit indices whenever they fit (the 64-bit index arithmetic was ~10 % of the issued instructions)
    if (n4 + 4 * nthreads < (int64_t)0x7fffffff) quant_stream<int>(x4, q4, (int)first, (int)nthreads, (int)n4, qf);
    else quant_stream<int64_t>(x4, q4, first, nthreads, n4, qf);
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += nthreads) q[i] = (uint8_t)quant1(x[i], zpf, scale, qmax);
}

// ------------------------------------------------------------------------------------------------
// Single-launch quantizer: one persistent CTA per SM.  Phase 1 streams the CTA's contiguous slice
// of x from HBM ONCE into shared memory (up to kQFSmemBytes of it; the rest of a very large slice
// is only min/max-reduced) and publishes the CTA's min/max; a grid-wide barrier follows; phase 2
// folds the partials, derives scale / zero-point and quantizes the slice out of shared memory, so
// HBM sees the compulsory 4 B read + 1 B write per element.  For the split-computing tensor
// (bch x 204 x 340 per image) up to 36 images stay fully on chip.
// Barrier state: counters[0] (arrivals) and counters[1] (departures) must be zero on entry; the last
// departing CTA zeroes both again.  A CTA that waits longer than ~2 s gives up and flags the
// result (zero_point = INT_MIN + 1) instead of hanging the device.
// ------------------------------------------------------------------------------------------------
static constexpr int kQFThreads = 1024;
static constexpr int kQFSmemBytes = 200 * 1024;
static constexpr int kQFResidentVec = kQFSmemBytes / 16;  // uint4 (4-element) vectors kept on chip

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(kQFThreads, 1)
    quant_fused_kernel(const float* __restrict__ x, int64_t n, float* partial, unsigned* counters,
                       float qmax, float inv_range, int scale_mode, uint8_t* __restrict__ q,
                       QParams* __restrict__ qp) {
  extern __shared__ uint4 s_x[];
  __shared__ int s_timeout;
  const int64_t nvec = n >> 2;  // 4-element vectors (x and q are 16-byte aligned on this path)
  const int64_t per = (nvec + gridDim.x - 1) / gridDim.x;
  int64_t v0 = per * blockIdx.x;
  if (v0 > nvec) v0 = nvec;
  int64_t v1 = v0 + per;
  if (v1 > nvec) v1 = nvec;
  const int len = (int)(v1 - v0);
  const int res = len < kQFResidentVec ? len : kQFResidentVec;
  const uint4* x4 = reinterpret_cast<const uint4*>(x) + v0;
  uint32_t* q4 = reinterpret_cast<uint32_t*>(q) + v0;
  float mn = INFINITY, mx = -INFINITY;
  if (threadIdx.x == 0) s_timeout = 0;
#pragma unroll 4
  for (int i = threadIdx.x; i < res; i += kQFThreads) {
    const uint4 v = ld_stream(x4 + i);
    s_x[i] = v;
    const float a = __uint_as_float(v.x), b = __uint_as_float(v.y), c = __uint_as_float(v.z),
                d = __uint_as_float(v.w);
    mn = nan_min(nan_min(mn, a), nan_min(b, nan_min(c, d)));
    mx = nan_max(nan_max(mx, a), nan_max(b, nan_max(c, d)));
  }
#pragma unroll 4
  for (int i = res + threadIdx.x; i < len; i += kQFThreads) {
    const uint4 v = ld_stream(x4 + i);  // re-read in phase 2 (L2 hit at these sizes)
    const float a = __uint_as_float(v.x), b = __uint_as_float(v.y), c = __uint_as_float(v.z),
                d = __uint_as_float(v.w);
    mn = nan_min(nan_min(mn, a), nan_min(b, nan_min(c, d)));
    mx = nan_max(nan_max(mx, a), nan_max(b, nan_max(c, d)));
  }
  if (blockIdx.x == 0)
    for (int64_t i = (nvec << 2) + threadIdx.x; i < n; i += kQFThreads) {
      mn = nan_min(mn, x[i]);
      mx = nan_max(mx, x[i]);
    }
  block_minmax(mn, mx);
  // ---- grid barrier ----
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = mn;
    partial[2 * blockIdx.x + 1] = mx;
    __threadfence();
    atomicAdd(&counters[0], 1u);
    const long long t0 = clock64();
    while (ld_acquire_u32(&counters[0]) < gridDim.x) {
      if (clock64() - t0 > 4000000000ll) {
        s_timeout = 1;
        break;
      }
    }
  }
  __syncthreads();
  mn = INFINITY;
  mx = -INFINITY;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += kQFThreads) {
    mn = nan_min(mn, __ldcg(partial + 2 * i));
    mx = nan_max(mx, __ldcg(partial + 2 * i + 1));
  }
  block_minmax(mn, mx);
  float scale;
  int zp;
  derive_qparams(mn, mx, qmax, inv_range, scale_mode, scale, zp);
  const float zpf = (float)zp;
  const QFast qf = make_qfast(zpf, scale, qmax);
#pragma unroll 4
  for (int i = threadIdx.x; i < res; i += kQFThreads) q4[i] = quant4(s_x[i], qf);
#pragma unroll 4
  for (int i = res + threadIdx.x; i < len; i += kQFThreads) q4[i] = quant4(ld_stream(x4 + i), qf);
  if (blockIdx.x == 0)
    for (int64_t i = (nvec << 2) + threadIdx.x; i < n; i += kQFThreads)
      q[i] = (uint8_t)quant1(x[i], zpf, scale, qmax);
  // ---- departure: publish the parameters, re-arm the barrier for the next launch ----
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {  // a barrier that timed out did so for every waiting CTA, this one included
      qp->scale = scale;
      qp->zero_point = s_timeout ? INT_MIN + 1 : zp;
      qp->mn = mn;
      qp->mx = mx;
    }
    __threadfence();
    const unsigned d = atomicAdd(&counters[1], 1u);
    if (d == gridDim.x - 1) {
      counters[0] = 0;
      counters[1] = 0;
      __threadfence();
    }
  }
}

__global__ void __launch_bounds__(kQThreads)
    dequant_kernel(const uint8_t* __restrict__ q, int64_t n, int vec_ok,
                   const QParams* __restrict__ qp, float* __restrict__ out) {
  const float scale = qp->scale;
  const float zpf = (float)qp->zero_point;
  // scale * (q.float() - zero_point)   tensor_util.py:22
  // (float)b - zp without I2F: bits 0x4b000000 | b are the float 2^23 + b; 2^23 + zp is exact, so is the
  // difference (both integers below 2^24) -- identical to __fsub_rn((float)b, zpf)
  const float zp_magic = 8388608.0f + zpf;
  auto dq = [&](uint32_t b) -> float {
    return __fmul_rn(scale, __fsub_rn(__uint_as_float(0x4b000000u | b), zp_magic));
  };
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  int64_t done = 0;
  if (vec_ok) {
    // lane-contiguous: a warp reads 128 contiguous bytes and writes 512 contiguous bytes per instruction
    const int64_t n4 = n >> 2;
    const uint32_t* q4 = reinterpret_cast<const uint32_t*>(q);
    uint4* o4 = reinterpret_cast<uint4*>(out);
    auto dq4 = [&](uint32_t w) -> uint4 {
      return make_uint4(__float_as_uint(dq(w & 0xff)), __float_as_uint(dq((w >> 8) & 0xff)),
                        __float_as_uint(dq((w >> 16) & 0xff)), __float_as_uint(dq(w >> 24)));
    };
    int64_t i = tid;
    for (; i + 3 * nthreads < n4; i += 4 * nthreads) {
      const uint32_t a = __ldg(q4 + i), b = __ldg(q4 + i + nthreads), c = __ldg(q4 + i + 2 * nthreads),
                     d = __ldg(q4 + i + 3 * nthreads);
      st_stream(o4 + i, dq4(a));
      st_stream(o4 + i + nthreads, dq4(b));
      st_stream(o4 + i + 2 * nthreads, dq4(c));
      st_stream(o4 + i + 3 * nthreads, dq4(d));
    }
    for (; i < n4; i += nthreads) st_stream(o4 + i, dq4(__ldg(q4 + i)));
    done = n4 << 2;
  }
  for (int64_t i = done + tid; i < n; i += nthreads) out[i] = dq(q[i]);
}

static int quant_blocks(int64_t n) {
  int64_t per_block = (int64_t)kQThreads * 16;
  int64_t b = (n + per_block - 1) / per_block;
  int64_t cap = (int64_t)num_sms() * 4;
  if (cap > kQMaxBlocks) cap = kQMaxBlocks;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace ghnd

extern "C" {

size_t ghnd_quantize_u8_workspace_bytes(int64_t n) {
  (void)n;
  return (size_t)ghnd::kQMaxBlocks * 2 * sizeof(float) + 4 * sizeof(unsigned);
}

int ghnd_quantize_u8(const float* x, int64_t n, int num_bits, int scale_mode, uint8_t* q,
                     void* qparams, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace ghnd;
  GHND_CHECK_ARG(x && q && qparams && workspace, "quantize_u8: null pointer");
  GHND_CHECK_ARG(n > 0, "quantize_u8: empty tensor (reference: min() of an empty tensor raises)");
  GHND_CHECK_ARG(num_bits >= 1 && num_bits <= 8, "quantize_u8: num_bits %d not in [1,8]", num_bits);
  GHND_CHECK_ARG(scale_mode == GHND_QSCALE_DIV || scale_mode == GHND_QSCALE_RECIP,
                 "quantize_u8: bad scale_mode %d", scale_mode);
  GHND_CHECK_ARG(workspace_bytes >= ghnd_quantize_u8_workspace_bytes(n),
                 "quantize_u8: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int vec_ok = ((((uintptr_t)x) & 15) == 0 && (((uintptr_t)q) & 15) == 0) ? 1 : 0;
  const int blocks = quant_blocks(n);
  const float qmax = (float)((1 << num_bits) - 1);
  const float inv_range = 1.0f / qmax;  // torch CUDA: a / cpu_scalar == a * (1/scalar) in fp32
  // Above ~20 M elements a CTA's slice no longer fits its shared memory (the rest is re-read) and two
  // full-occupancy passes are faster than one launch of 148 CTAs (B200: 89 vs 115 us at 53 M elements,
  // 310 vs 453 us at 213 M; 33.5 vs 31.6 us at 13 M).  GHND_QUANT_TWO_PASS=1|0 forces either.
  static const int two_pass_env = [] {
    const char* e = getenv("GHND_QUANT_TWO_PASS");
    return e == nullptr ? -1 : atoi(e);
  }();
  const bool two_pass = two_pass_env >= 0 ? two_pass_env != 0 : n > (int64_t)20 << 20;
  if (vec_ok && n >= 4096 && !two_pass) {
    // single persistent launch; workspace = [2 * kQMaxBlocks floats of partials][4 barrier words]
    static bool attr_set = false;
    if (!attr_set) {
      GHND_CUDA(cudaFuncSetAttribute(quant_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kQFSmemBytes));
      attr_set = true;
    }
    int grid = num_sms();
    if (grid > kQMaxBlocks) grid = kQMaxBlocks;
    const int64_t want = ((n >> 2) + kQFThreads - 1) / kQFThreads;  // no CTA without a vector
    if (grid > want) grid = (int)want;
    unsigned* counters = (unsigned*)((float*)workspace + 2 * kQMaxBlocks);
    quant_fused_kernel<<<grid, kQFThreads, kQFSmemBytes, st>>>(x, n, (float*)workspace, counters, qmax,
                                                             inv_range, scale_mode, q,
                                                             (QParams*)qparams);
    GHND_LAUNCH_CHECK("quant_fused_kernel");
    return GHND_OK;
  }
  quant_minmax_kernel<<<blocks, kQThreads, 0, st>>>(x, n, vec_ok, (float*)workspace);
  GHND_LAUNCH_CHECK("quant_minmax_kernel");
  quant_apply_kernel<<<blocks, kQThreads, 0, st>>>(x, n, vec_ok, (const float*)workspace, blocks,
                                                   qmax, inv_range, scale_mode, q,
                                                   (QParams*)qparams);
  GHND_LAUNCH_CHECK("quant_apply_kernel");
  return GHND_OK;
}

int ghnd_quantize_u8_minmax(const float* x, int64_t n, int num_bits, int scale_mode,
                            const float* minmax_partial, int n_partial, uint8_t* q, void* qparams,
                            void* stream) {
  using namespace ghnd;
  GHND_CHECK_ARG(x && q && qparams && minmax_partial, "quantize_u8_minmax: null pointer");
  GHND_CHECK_ARG(n > 0 && n_partial > 0, "quantize_u8_minmax: empty tensor");
  GHND_CHECK_ARG(num_bits >= 1 && num_bits <= 8, "quantize_u8_minmax: num_bits %d not in [1,8]", num_bits);
  GHND_CHECK_ARG(scale_mode == GHND_QSCALE_DIV || scale_mode == GHND_QSCALE_RECIP,
                 "quantize_u8_minmax: bad scale_mode %d", scale_mode);
  const int vec_ok = ((((uintptr_t)x) & 15) == 0 && (((uintptr_t)q) & 15) == 0) ? 1 : 0;
  const float qmax = (float)((1 << num_bits) - 1);
  // one streaming pass: 4 B read + 1 B write per element; 16 elements per thread and iteration
  int64_t blocks = (n / 16 + kQThreads - 1) / kQThreads;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  quant_apply_kernel<<<(int)blocks, kQThreads, 0, (cudaStream_t)stream>>>(
      x, n, vec_ok, minmax_partial, n_partial, qmax, 1.0f / qmax, scale_mode, q, (QParams*)qparams);
  GHND_LAUNCH_CHECK("quant_apply_kernel");
  return GHND_OK;
}

int ghnd_dequantize_u8(const uint8_t* q, int64_t n, const void* qparams, float* out,
                       void* stream) {
  using namespace ghnd;
  GHND_CHECK_ARG(q && qparams && out, "dequantize_u8: null pointer");
  GHND_CHECK_ARG(n >= 0, "dequantize_u8: negative size");
  if (n == 0) return GHND_OK;
  const int vec_ok = ((((uintptr_t)out) & 15) == 0 && (((uintptr_t)q) & 15) == 0) ? 1 : 0;
  dequant_kernel<<<quant_blocks(n), kQThreads, 0, (cudaStream_t)stream>>>(
      q, n, vec_ok, (const QParams*)qparams, out);
  GHND_LAUNCH_CHECK("dequant_kernel");
  return GHND_OK;
}
}
