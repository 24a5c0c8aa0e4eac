// Stem fused with its max-pool: conv1 7x7 s2 p3 (3->64) + FrozenBatchNorm + ReLU + MaxPool 3x3 s2 p1
// (src/models/custom/resnet.py:26-30,96-99) WITHOUT materialising conv1's output.
//
// The un-fused path writes the [N][Hp/2][Wp/2][64] conv output (69 MB per image and model at 800x1344)
// at the 3.87 TB/s write-only rate of HBM and reads it back for the pool; nothing else ever needs it:
// the pool's backward works from the argmax codes (which carry the ReLU mask), conv1's dW from the
// pooled gradient.  Here a CTA job computes a 16-row x 32-column patch of conv outputs for ONE model
// on tcgen05 and pools it in shared memory:
//
//   * the implicit GEMM is the one of conv_tc.cu's stem halo mode: output columns in four classes
//     (wo = 4j + q) so that consecutive GEMM rows are 64 B apart in the packed image; per class ONE TMA
//     box of 37 image rows x 8 column slots x 64 B, the 7 filter rows are MMA descriptors at row shifts
//     of that tile (SWIZZLE_64B, SBO = two image rows); weights [64][7][32] resident in shared memory;
//     M = 128 (16 rows x 8 slots), N = 64, K = 7 x 32.
//   * a job = (model, patch): the four classes of a patch go to four 64-column TMEM accumulators; two
//     such sets (2 x 256 columns) alternate between consecutive jobs, so the MMAs of job i+1 run under
//     the epilogue of job i.
//   * epilogue: 16 warps move the accumulators (bias + ReLU in packed half2, as the un-fused epilogue)
//     into a swizzled [16][32][64 ch] 16-bit tile in shared memory, then pool 7 x 14 outputs from it
//     (packed half2 max / compare-mask / select, first maximum in scan order like torch) and store
//     pooled rows (+ argmax bytes, 0xff where the maximum is not > 0) with 128-byte-per-pixel coalescing.
//   * patches overlap by two conv rows / four conv columns (the 3x3 window's halo is recomputed: 16x32
//     conv outputs give 7x14 pooled ones, 77 % efficiency) -- the GEMM is 7 % of conv1's old cost centre,
//     the removed HBM round trip was all of it.
//
// Bit-identical to ghnd_stem_conv_plan_* + ghnd_maxpool3x3s2* (same MMA order, same epilogue rounding).
#include <stdlib.h>
#include <vector>

#include "common.cuh"

namespace ghnd {

static constexpr int kSpThreads = 576;      // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue + pool
static constexpr int kSpEpiThreads = 512;   // 16 warps = 4 TMEM lane quarters x 4 column classes
static constexpr int kSpMaxStages = 8;
static constexpr int kSpRows = 16, kSpCols = 32;          // conv patch
static constexpr int kSpPoolRows = 7, kSpPoolCols = 14;   // pooled outputs per patch
static constexpr int kSpHaloRows = 2 * kSpRows + 5;       // 37 image rows per class box
static constexpr int kSpBoxBytes = kSpHaloRows * 8 * 64;  // 18944
static constexpr int kSpStageBytes = (kSpBoxBytes + 1023) / 1024 * 1024;  // 19456
static constexpr int kSpWBytes = 64 * 64;                 // one filter row of one model: 64 k x 32 elements
static constexpr int kSpTileBytes = kSpRows * kSpCols * 128;  // 65536

struct StemPoolParams {
  CUtensorMap tmap_a[4];  // packed image, one per column class
  CUtensorMap tmap_w;     // weights [64*n_models][7*32], box {32, 64}
  const float* bias;      // [64*n_models]
  void* y[2];             // pooled output per model [N][Ho][Wo][64]
  uint8_t* argmax[2];     // nullable
  int n_models, n_img, Hc, Wc, Ho, Wo;
  int tiles_x, tiles_per_img, total_jobs;
  FastDiv fd_models, fd_tiles_per_img, fd_tiles_x;
  int n_stages, fmt;
  uint32_t idesc;
};

template <int FMT>
__device__ __forceinline__ uint32_t sp_max2(uint32_t a, uint32_t b) {
  if (FMT == GHND_F16) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  const __nv_bfloat162 r =
      __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <int FMT>
__device__ __forceinline__ void sp_pool_tap(uint32_t& best, uint32_t& idx, uint32_t v, uint32_t code) {
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
// Bias + ReLU + conversion of one channel pair, with the rounding of conv_tc.cu's epilogue for the same output
// format: f16 -> packed (h = cvt(acc); h += bias16; max(h, 0)); bf16 -> fp32 (cvt(max(acc + bias, 0))).
// `bias` points at the pair's bias: one packed f16 word, or two floats.
template <int FMT>
__device__ __forceinline__ uint32_t sp_bias_relu(float a, float b, const uint32_t* bias) {
  if (FMT == GHND_F16) {
    const uint32_t h = pack2_t<FMT>(a, b);
    __half2 r = __hadd2(*reinterpret_cast<const __half2*>(&h), *reinterpret_cast<const __half2*>(bias));
    r = __hmax2(r, __float2half2_rn(0.f));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  const float2 bf = *reinterpret_cast<const float2*>(bias);
  return pack2_t<FMT>(fmaxf(a + bf.x, 0.f), fmaxf(b + bf.y, 0.f));
}

// Two K=16 steps with separate high words for A and B (see conv_tc.cu).
__device__ __forceinline__ void sp_umma2(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                         uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 a, b;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.eq.b32 q, 0, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "add.u32 a, %1, 2;\n\t"
      "add.u32 b, %3, 2;\n\t"
      "mov.b64 da, {a, %2};\n\t"
      "mov.b64 db, {b, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int FMT>
__global__ void __launch_bounds__(kSpThreads, 1)
    stem_pool_kernel(const __grid_constant__ StemPoolParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();
  uint8_t* wres = smem + (size_t)p.n_stages * kSpStageBytes;      // [n_models][7][64 x 64 B]
  uint8_t* tile = wres + (size_t)p.n_models * 7 * kSpWBytes;      // [16][32][128 B], swizzled
  // bias per channel pair: one packed f16 word, or two floats (bf16 output: fp32 epilogue) -- see sp_bias_relu
  constexpr int kBw = FMT == GHND_F16 ? 1 : 2;
  uint32_t* sbias = reinterpret_cast<uint32_t*>(tile + kSpTileBytes);  // [n_models][32 pairs][kBw]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sbias + 128);
  uint64_t* afull = bars;                     // [kSpMaxStages]
  uint64_t* aempty = afull + kSpMaxStages;    // [kSpMaxStages]
  uint64_t* tfull = aempty + kSpMaxStages;    // [2]
  uint64_t* tempty = tfull + 2;               // [2]
  uint64_t* wfull = tempty + 2;               // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) prefetch_tmap(&p.tmap_a[i]);
    prefetch_tmap(&p.tmap_w);
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&afull[i], 1);
      mbar_init(&aempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kSpEpiThreads / 32);  // one arrival per epilogue warp
    }
    mbar_init(wfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull, (uint32_t)(p.n_models * 7 * kSpWBytes));
      for (int m = 0; m < p.n_models; ++m)
        for (int r = 0; r < 7; ++r)
          tma_load_2d(wres + (size_t)(m * 7 + r) * kSpWBytes, &p.tmap_w, wfull, r * 32, m * 64);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int job = blockIdx.x; job < p.total_jobs; job += gridDim.x) {
      int t, m, img, rem, ty, tx;
      fd_divmod(p.fd_models, job, t, m);
      fd_divmod(p.fd_tiles_per_img, t, img, rem);
      fd_divmod(p.fd_tiles_x, rem, ty, tx);
      const int r_base = 2 * kSpPoolRows * ty - 1;  // first conv row of the patch
      const int J0 = (kSpPoolCols / 2) * tx;        // first conv column = 4*J0 - 1
      for (int q = 0; q < 4; ++q) {
        mbar_wait(&aempty[stage], phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&afull[stage], (uint32_t)kSpBoxBytes);
          tma_load_4d(smem + (size_t)stage * kSpStageBytes, &p.tmap_a[q], &afull[stage], 0,
                      J0 - (q == 3 ? 1 : 0), 2 * r_base, img);
        }
        __syncwarp();
        if (++stage == p.n_stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0, it = 0;
    uint32_t phase = 0;
    const uint32_t smem_base = smem_u32(smem), w_base = smem_u32(wres);
    const uint32_t row_pitch = 8u * 64u;  // 8 column slots x 64 B per image row of a class tile
    const uint32_t a_hi = ((2u * row_pitch) >> 4) | (1u << 14) | ((uint32_t)UMMA_SW64 << 29);
    const uint32_t b_hi = (512u >> 4) | (1u << 14) | ((uint32_t)UMMA_SW64 << 29);
    mbar_wait(wfull, 0);
    tc_fence_after();
    for (int job = blockIdx.x; job < p.total_jobs; job += gridDim.x, ++it) {
      const int buf = it & 1;
      const int m = job - fd_div(p.fd_models, job) * (int)p.fd_models.d;
      mbar_wait(&tempty[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      for (int q = 0; q < 4; ++q) {
        mbar_wait(&afull[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256 + q * 64);
          const uint32_t sa = smem_base + (uint32_t)(stage * kSpStageBytes);
          for (int r = 0; r < 7; ++r) {
            const uint32_t a_lo = (((sa + (uint32_t)r * row_pitch) >> 4) & 0x3fffu) | (1u << 16);
            const uint32_t b_lo = (((w_base + (uint32_t)((m * 7 + r) * kSpWBytes)) >> 4) & 0x3fffu) | (1u << 16);
            sp_umma2(d_tmem, a_lo, a_hi, b_lo, b_hi, p.idesc, (uint32_t)(r != 0));
          }
          umma_commit(&aempty[stage]);
          if (q == 3) umma_commit(&tfull[buf]);
        }
        __syncwarp();
        if (++stage == p.n_stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue + pool: 16 warps =====================
    // Two warps per scheduler could not hide the LDS / TMEM latencies of this section (ncu: 34 % issue
    // slots, 7 cycles per issued instruction); 16 warps also give every (lane quarter, class) its own warp.
    const int et = threadIdx.x - 64;  // 0..511
    const int quarter = warp & 3;     // TMEM lane quarter this warp may access
    const int q = (warp - 2) >> 2;    // column class this warp moves out of TMEM
    const int i_row = quarter * 4 + (lane >> 3);  // conv row of this thread's accumulator lane
    const int jj = lane & 7;                      // column slot
    const int x_col = 4 * jj + ((q + 1) & 3);     // column inside the patch
    const uint32_t ninf = FMT == GHND_F16 ? 0xfc00fc00u : 0xff80ff80u;
    for (int i = et; i < p.n_models * 32; i += kSpEpiThreads) {
      if (FMT == GHND_F16) {
        sbias[i] = pack2_t<FMT>(__ldg(p.bias + 2 * i), __ldg(p.bias + 2 * i + 1));
      } else {
        sbias[2 * i] = __float_as_uint(__ldg(p.bias + 2 * i));
        sbias[2 * i + 1] = __float_as_uint(__ldg(p.bias + 2 * i + 1));
      }
    }
    named_bar_sync(1, kSpEpiThreads);
    int it = 0;
    for (int job = blockIdx.x; job < p.total_jobs; job += gridDim.x, ++it) {
      const int buf = it & 1;
      int t, m, img, rem, ty, tx;
      fd_divmod(p.fd_models, job, t, m);
      fd_divmod(p.fd_tiles_per_img, t, img, rem);
      fd_divmod(p.fd_tiles_x, rem, ty, tx);
      const int ph0 = kSpPoolRows * ty, pw0 = kSpPoolCols * tx;
      const int r_base = 2 * ph0 - 1, c_base = 2 * pw0 - 1;
      mbar_wait(&tfull[buf], ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      // ---- phase A: accumulators -> bias + ReLU -> swizzled 16-bit tile; -inf outside the image, so
      //      that the pool below needs no border tests ----
      {
        const uint32_t* bm = sbias + m * 32 * kBw;
        uint32_t r[64];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 256 + q * 64);
        tmem_ld32(taddr, r);
        tmem_ld32(taddr + 32, r + 32);
        tmem_ld_wait();
        const int key = (i_row + jj) & 7;
        uint8_t* px = tile + (size_t)(i_row * kSpCols + x_col) * 128;
        const bool inside = (unsigned)(r_base + i_row) < (unsigned)p.Hc && (unsigned)(c_base + x_col) < (unsigned)p.Wc;
        if (inside) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 o;
            o.x = sp_bias_relu<FMT>(__uint_as_float(r[8 * j]), __uint_as_float(r[8 * j + 1]), bm + (4 * j + 0) * kBw);
            o.y = sp_bias_relu<FMT>(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]), bm + (4 * j + 1) * kBw);
            o.z = sp_bias_relu<FMT>(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]), bm + (4 * j + 2) * kBw);
            o.w = sp_bias_relu<FMT>(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]), bm + (4 * j + 3) * kBw);
            *reinterpret_cast<uint4*>(px + ((j ^ key) << 4)) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(px + (j << 4)) = make_uint4(ninf, ninf, ninf, ninf);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[buf]);  // TMEM set free for the MMAs of the job after next
      named_bar_sync(1, kSpEpiThreads);
      // ---- phase B: 3x3 s2 max-pool of the tile -> 7 x 14 pooled pixels ----
      uint4* yo = reinterpret_cast<uint4*>(p.y[m]);
      uint2* ao = reinterpret_cast<uint2*>(p.argmax[m]);
      for (int item = et; item < kSpPoolRows * kSpPoolCols * 8; item += kSpEpiThreads) {
        const int cg = item & 7, pix = item >> 3;
        const int a = pix / kSpPoolCols, b = pix - a * kSpPoolCols;
        const int ph = ph0 + a, pw = pw0 + b;
        if (ph >= p.Ho || pw >= p.Wo) continue;
        uint4 v[9];
#pragma unroll
        for (int dr = 0; dr < 3; ++dr)
#pragma unroll
          for (int ds = 0; ds < 3; ++ds) {
            const int i = 2 * a + dr, x = 2 * b + ds;
            const int key = (i + (x >> 2)) & 7;
            v[dr * 3 + ds] = *reinterpret_cast<const uint4*>(tile + (size_t)(i * kSpCols + x) * 128 + ((cg ^ key) << 4));
          }
        const int64_t o = (((int64_t)img * p.Ho + ph) * p.Wo + pw) * 8 + cg;
        uint32_t best[4] = {ninf, ninf, ninf, ninf};
        if (ao == nullptr) {  // no argmax wanted (frozen model): plain maximum
#pragma unroll
          for (int tpi = 0; tpi < 9; ++tpi) {
            best[0] = sp_max2<FMT>(best[0], v[tpi].x);
            best[1] = sp_max2<FMT>(best[1], v[tpi].y);
            best[2] = sp_max2<FMT>(best[2], v[tpi].z);
            best[3] = sp_max2<FMT>(best[3], v[tpi].w);
          }
          yo[o] = make_uint4(best[0], best[1], best[2], best[3]);
          continue;
        }
        uint32_t idx[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int tpi = 0; tpi < 9; ++tpi) {
          const uint32_t code = (uint32_t)tpi * 0x00010001u;
          sp_pool_tap<FMT>(best[0], idx[0], v[tpi].x, code);
          sp_pool_tap<FMT>(best[1], idx[1], v[tpi].y, code);
          sp_pool_tap<FMT>(best[2], idx[2], v[tpi].z, code);
          sp_pool_tap<FMT>(best[3], idx[3], v[tpi].w, code);
        }
        yo[o] = make_uint4(best[0], best[1], best[2], best[3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {  // fold the ReLU mask into the code (see maxpool_kernel)
          uint32_t pos;
          if (FMT == GHND_F16) pos = __hgt2_mask(*reinterpret_cast<const __half2*>(&best[e]), __float2half2_rn(0.f));
          else pos = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&best[e]), __float2bfloat162_rn(0.f));
          idx[e] = (idx[e] & pos) | (0x00ff00ffu & ~pos);
        }
        uint2 am;
        am.x = __byte_perm(idx[0], idx[1], 0x6420);
        am.y = __byte_perm(idx[2], idx[3], 0x6420);
        ao[o] = am;
      }
      named_bar_sync(1, kSpEpiThreads);  // the tile may be overwritten by the next job's phase A
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace ghnd

struct ghnd_stem_pool_plan {
  ghnd::StemPoolParams p;
  int grid;
  size_t smem;
};

extern "C" {

int ghnd_stem_pool_plan_create(const void* x_packed, int x_fmt, const void* w_packed, int w_fmt,
                               const float* bias, int n_models, void* const* y, void* const* argmax,
                               int y_fmt, int N, int Hp, int Wp, ghnd_stem_pool_plan_t** out) {
  using namespace ghnd;
  GHND_CHECK_ARG(out && x_packed && w_packed && bias && y, "stem_pool_plan_create: null argument");
  *out = nullptr;
  GHND_CHECK_ARG(n_models == 1 || n_models == 2, "stem_pool: n_models %d unsupported", n_models);
  GHND_CHECK_ARG(N > 0 && Hp > 0 && Wp > 0 && Hp % 2 == 0 && Wp % 8 == 0,
                 "stem_pool: padded size must be a multiple of 2 x 8 (Hp=%d Wp=%d)", Hp, Wp);
  GHND_CHECK_ARG((x_fmt == GHND_F16 || x_fmt == GHND_BF16) && x_fmt == w_fmt && y_fmt == x_fmt,
                 "stem_pool: image, weights and output must share one 16-bit format");
  for (int m = 0; m < n_models; ++m) GHND_CHECK_ARG(y[m] != nullptr, "stem_pool: null output %d", m);
  ghnd_stem_pool_plan* plan = new ghnd_stem_pool_plan();
  StemPoolParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  const int Hc = Hp / 2, Wc = Wp / 2;
  p.n_models = n_models;
  p.n_img = N;
  p.Hc = Hc;
  p.Wc = Wc;
  p.Ho = (Hc + 1) / 2;
  p.Wo = (Wc + 1) / 2;
  p.fmt = y_fmt;
  p.bias = bias;
  for (int m = 0; m < n_models; ++m) {
    p.y[m] = y[m];
    p.argmax[m] = argmax ? static_cast<uint8_t*>(argmax[m]) : nullptr;
  }
  const int tiles_y = (p.Ho + kSpPoolRows - 1) / kSpPoolRows;
  p.tiles_x = (p.Wo + kSpPoolCols - 1) / kSpPoolCols;
  p.tiles_per_img = tiles_y * p.tiles_x;
  p.total_jobs = N * p.tiles_per_img * n_models;
  p.fd_models = make_fastdiv(n_models);
  p.fd_tiles_per_img = make_fastdiv(p.tiles_per_img);
  p.fd_tiles_x = make_fastdiv(p.tiles_x);
  p.idesc = make_idesc(x_fmt, w_fmt, 0, 0, 128, 64);
  const int rows = Hp + 6, RP = (Wp + 8) * 4;  // packed image rows / row pitch in elements
  int rc = GHND_OK;
  for (int q = 0; q < 4 && rc == GHND_OK; ++q) {
    const int J = (Wc - q + 3) / 4;
    const uint8_t* base = static_cast<const uint8_t*>(x_packed) + (size_t)q * 8 * 2;
    uint64_t dims[4] = {32, (uint64_t)(J > 0 ? J : 1), (uint64_t)rows, (uint64_t)N};
    uint64_t str[4] = {2, 64, (uint64_t)RP * 2, (uint64_t)rows * RP * 2};
    uint32_t box[4] = {32, 8, (uint32_t)kSpHaloRows, 1};
    rc = encode_tmap(&p.tmap_a[q], 2, 4, const_cast<uint8_t*>(base), dims, str, box, 64);
  }
  if (rc == GHND_OK) {
    uint64_t dims[2] = {7 * 32, (uint64_t)(64 * n_models)};
    uint64_t str[2] = {2, 7 * 32 * 2};
    uint32_t box[2] = {32, 64};
    rc = encode_tmap(&p.tmap_w, 2, 2, const_cast<void*>(w_packed), dims, str, box, 64);
  }
  if (rc != GHND_OK) {
    delete plan;
    return rc;
  }
  const int fixed = n_models * 7 * kSpWBytes + kSpTileBytes + 512 /*bias*/ + 512 /*barriers*/;
  int stages = (227 * 1024 - fixed) / kSpStageBytes;
  if (stages > kSpMaxStages) stages = kSpMaxStages;
  if (stages < 4) {
    delete plan;
    set_error("stem_pool: shared memory leaves %d stages", stages);
    return GHND_ERR_UNSUPPORTED;
  }
  p.n_stages = stages;
  plan->smem = (size_t)stages * kSpStageBytes + fixed;
  plan->grid = p.total_jobs < num_sms() ? p.total_jobs : num_sms();
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(stem_pool_kernel<GHND_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(stem_pool_kernel<GHND_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      delete plan;
      return cuda_fail(e, "cudaFuncSetAttribute(stem_pool_kernel)");
    }
    attr_set = true;
  }
  *out = plan;
  return GHND_OK;
}

int ghnd_stem_pool_plan_run(const ghnd_stem_pool_plan_t* plan, void* stream) {
  using namespace ghnd;
  GHND_CHECK_ARG(plan != nullptr, "stem_pool_plan_run: null plan");
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)plan->grid, 1, 1);
  cfg.blockDim = dim3(kSpThreads, 1, 1);
  cfg.dynamicSmemBytes = plan->smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool no_pdl = getenv("GHND_NO_PDL") != nullptr;
  cfg.attrs = attr;
  cfg.numAttrs = no_pdl ? 0 : 1;
  cudaError_t e = plan->p.fmt == GHND_F16 ? cudaLaunchKernelEx(&cfg, stem_pool_kernel<GHND_F16>, plan->p)
                                          : cudaLaunchKernelEx(&cfg, stem_pool_kernel<GHND_BF16>, plan->p);
  if (e != cudaSuccess) return cuda_fail(e, "stem_pool_kernel");
  return GHND_OK;
}

void ghnd_stem_pool_plan_destroy(ghnd_stem_pool_plan_t* plan) { delete plan; }
}
