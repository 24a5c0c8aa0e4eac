// Weight gradient of the student's wide k=2 convs on tcgen05 (sm_100a).
//
//   dw[k][tap][c] = sum_{pixels} dy[pixel][k] * x[pixel + tap_offset][c]
//
// Both operands are read straight out of the NHWC tensors by TMA as [32 pixels][64 channels]
// 128B-swizzled boxes, i.e. they sit in shared memory "MN-major" (channel contiguous, the GEMM
// reduction axis = pixel is the strided one); the instruction descriptor's a_major/b_major bits
// select that layout, so no transpose pass is needed.
//
// One CTA owns a 128 (rows operand) x NB (cols operand, <=128) output block for ALL taps (<=4):
// the accumulators fill TMEM (4 x 128 columns).  The pixel axis is split over CTAs; every CTA adds
// its partial block into the fp32 dw with red.global.add.f32 (dw is zeroed by the plan first).
// Zero padding of x and the ragged right edge of dy both come from TMA out-of-bounds zero fill.
#include <vector>

#include "common.cuh"

namespace ghnd {

static constexpr int kWgThreads = 224;  // warp0/6 TMA producers, warp1 MMA, warps 2..5 epilogue
static constexpr int kWgPix = 32;                  // pixels (GEMM-K) per pipeline stage
static constexpr int kWgBox = kWgPix * 128;        // bytes of one [32 px][64 ch] box = 4 KB
static constexpr int kWgMaxTaps = 4;

struct WgradParams {
  CUtensorMap tmap_row;   // rows operand tensor (GEMM M): box {64, 32, 1, 1}
  CUtensorMap tmap_col;   // cols operand tensor (GEMM N)
  int rows_is_dy;         // 1: rows = dy channels (k), cols = x channels (c); 0: swapped
  int n_taps, S;
  int pad;
  int m_tiles, n_chunks;  // output blocks: 128-row tiles x NB-col chunks
  int rows_valid;         // valid rows in a tile (64 or 128)
  int row_boxes;          // 64-channel boxes per row tile (1 or 2)
  int nb;                 // columns per chunk (64 or 128)
  int col_boxes;          // nb / 64
  int splits;             // CTAs per output block (pixel-axis split)
  int n_img, ho, wo;      // dy geometry
  int tiles_w;            // ceil(wo / 32)
  int total_pix_tiles;    // n_img * ho * tiles_w
  int stage_bytes, n_stages;
  int tap_dh[kWgMaxTaps], tap_dw[kWgMaxTaps];  // x offset of every tap (already minus pad)
  FastDiv fd_tiles_per_img, fd_tiles_w;
  uint32_t idesc;
  float* dw;              // [K][taps][C]
  int K, C;
};


// the two K=16 MMAs of one 32-pixel stage for one tap (second descriptor = +2048 B)
__device__ __forceinline__ void umma_pair_wg(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo,
                                             uint32_t desc_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 a, b;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.eq.b32 q, 0, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
      "add.u32 a, %1, 128;\n\t"
      "add.u32 b, %2, 128;\n\t"
      "mov.b64 da, {a, %3};\n\t"
      "mov.b64 db, {b, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, q;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1)
    wgrad_tc_kernel(const __grid_constant__ WgradParams p) {
  // aligned by declaration; plain pointer arithmetic keeps the shared state space (LDS/STS, not generic)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (smem_u32(smem_raw) & 1023u) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.n_stages * p.stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.n_stages;
  uint64_t* done_bar = bars + 2 * p.n_stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmap_row);
    prefetch_tmap(&p.tmap_col);
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&full_bar[i], 2);  // two producer warps, each posts its own byte count
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  // block decomposition
  const int block = blockIdx.x / p.splits;
  const int split = blockIdx.x - block * p.splits;
  const int m_tile = block / p.n_chunks;
  const int n_chunk = block - m_tile * p.n_chunks;
  const int t_begin = (int)(((int64_t)p.total_pix_tiles * split) / p.splits);
  const int t_end = (int)(((int64_t)p.total_pix_tiles * (split + 1)) / p.splits);
  const int a_bytes = p.row_boxes * kWgBox;
  const int b_tap_bytes = p.col_boxes * kWgBox;

  if (warp == 0 || warp == 6) {
    // Two TMA producer warps (whole warp converged, one elected lane issues).  A single thread
    // issuing all ~10 small boxes of a stage was the bottleneck of this kernel (ncu: MMA warp
    // starved while the producer never waited for a free slot), so the taps are split in halves.
    const int pw = warp == 0 ? 0 : 1;
    const int tap_lo = pw == 0 ? 0 : (p.n_taps + 1) / 2;
    const int tap_hi = pw == 0 ? (p.n_taps + 1) / 2 : p.n_taps;
    int stage = 0;
    uint32_t phase = 0;
    // (img, h, wt) of the first tile by multiply-high division, then carried incrementally
    int img, rem, h, wt;
    fd_divmod(p.fd_tiles_per_img, t_begin, img, rem);
    fd_divmod(p.fd_tiles_w, rem, h, wt);
    // bytes this warp posts per stage: its taps of the shifted operand (+ the unshifted one for pw 0)
    const uint32_t shifted_bytes = (uint32_t)((tap_hi - tap_lo) * (p.rows_is_dy ? b_tap_bytes : a_bytes));
    const uint32_t fixed_bytes = pw == 0 ? (uint32_t)(p.rows_is_dy ? a_bytes : b_tap_bytes) : 0u;
    for (int t = t_begin; t < t_end; ++t) {
      const int w0 = wt * kWgPix;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
        uint8_t* sb = sa + a_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], shifted_bytes + fixed_bytes);
        if (p.rows_is_dy) {
          // rows operand = dy (unshifted), cols operand = x shifted per tap
          if (pw == 0)
            for (int b = 0; b < p.row_boxes; ++b)
              tma_load_4d(sa + b * kWgBox, &p.tmap_row, &full_bar[stage], m_tile * 128 + b * 64, w0, h,
                          img);
          for (int tap = tap_lo; tap < tap_hi; ++tap) {
            const int dh = p.tap_dh[tap], dw = p.tap_dw[tap];
            for (int b = 0; b < p.col_boxes; ++b)
              tma_load_4d(sb + tap * b_tap_bytes + b * kWgBox, &p.tmap_col, &full_bar[stage],
                          n_chunk * p.nb + b * 64, w0 + dw, h + dh, img);
          }
        } else {
          // rows = x channels: A region holds n_taps shifted x tiles, B region the single dy tile
          for (int tap = tap_lo; tap < tap_hi; ++tap) {
            const int dh = p.tap_dh[tap], dw = p.tap_dw[tap];
            for (int b = 0; b < p.row_boxes; ++b)
              tma_load_4d(sa + tap * a_bytes + b * kWgBox, &p.tmap_row, &full_bar[stage],
                          m_tile * 128 + b * 64, w0 + dw, h + dh, img);
          }
          if (pw == 0) {
            uint8_t* sd = sa + p.n_taps * a_bytes;
            for (int b = 0; b < p.col_boxes; ++b)
              tma_load_4d(sd + b * kWgBox, &p.tmap_col, &full_bar[stage], n_chunk * p.nb + b * 64, w0,
                          h, img);
          }
        }
      }
      __syncwarp();
      if (++wt == p.tiles_w) {
        wt = 0;
        if (++h == p.ho) {
          h = 0;
          ++img;
        }
      }
      if (++stage == p.n_stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // MMA issuer: whole warp converged, one elected lane issues; descriptors built from uniform
    // 32-bit halves so the UTCHMMAs issue back to back from uniform registers
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t smem_base = smem_u32(smem);
    // rows operand with a single 64-channel box: rows 64..127 alias rows 0..63 (LBO = 0)
    const uint32_t a_lbo = p.row_boxes == 2 ? kWgBox : 0;
    const uint32_t b_lbo = kWgBox;
    const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | ((uint32_t)UMMA_SW128 << 29);
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_base + (uint32_t)(stage * p.stage_bytes);
        for (int tap = 0; tap < p.n_taps; ++tap) {
          uint32_t a_addr, b_addr;
          if (p.rows_is_dy) {
            a_addr = sa;
            b_addr = sa + a_bytes + tap * b_tap_bytes;
          } else {
            a_addr = sa + tap * a_bytes;
            b_addr = sa + p.n_taps * a_bytes;
          }
          const uint32_t a_lo = ((a_addr >> 4) & 0x3fffu) | ((a_lbo >> 4) << 16);
          const uint32_t b_lo = ((b_addr >> 4) & 0x3fffu) | ((b_lbo >> 4) << 16);
          // 16 pixels = two 8-row swizzle groups = 2048 B -> +128 in (addr >> 4)
          umma_pair_wg(tmem_base + (uint32_t)(tap * 128), a_lo, b_lo, desc_hi, p.idesc,
                       (uint32_t)(t != t_begin));
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == p.n_stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else if (warp < 6) {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    if (t_end > t_begin) {
      mbar_wait(done_bar, 0);
      tc_fence_after();
      const int taps = p.n_taps;
      for (int tap = 0; tap < taps; ++tap) {
        for (int c = 0; c < p.nb; c += 32) {
          uint32_t r[32];
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tap * 128 + c), r);
          tmem_ld_wait();
          if (row < p.rows_valid) {
            if (p.rows_is_dy) {
              const int k = m_tile * 128 + row;
              float* dst = p.dw + ((size_t)k * taps + tap) * p.C + n_chunk * p.nb + c;
#pragma unroll
              for (int j = 0; j < 32; j += 4)  // 16-byte vector reductions (dst is 128 B aligned)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j),
                             "f"(__uint_as_float(r[j])), "f"(__uint_as_float(r[j + 1])),
                             "f"(__uint_as_float(r[j + 2])), "f"(__uint_as_float(r[j + 3]))
                             : "memory");
            } else {
              const int cc = m_tile * 128 + row;
              const int k0 = n_chunk * p.nb + c;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                atomicAdd(p.dw + ((size_t)(k0 + j) * taps + tap) * p.C + cc, __uint_as_float(r[j]));
            }
          }
        }
      }
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

struct ghnd_wgrad_plan {
  ghnd::WgradParams p;
  int grid;
  size_t smem;
  size_t dw_bytes;
};

extern "C" {

int ghnd_wgrad_plan_create(const ghnd_wgrad_desc_t* d, ghnd_wgrad_plan_t** out) {
  using namespace ghnd;
  GHND_CHECK_ARG(d && out, "wgrad_plan_create: null argument");
  *out = nullptr;
  GHND_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0, "wgrad: bad geometry");
  GHND_CHECK_ARG(d->C % 64 == 0 && d->K % 64 == 0 && d->C > 0 && d->K > 0,
                 "wgrad: C and K must be multiples of 64 (C=%d K=%d)", d->C, d->K);
  GHND_CHECK_ARG(d->R >= 1 && d->S >= 1 && d->R * d->S <= kWgMaxTaps,
                 "wgrad: at most %d taps (got %dx%d)", kWgMaxTaps, d->R, d->S);
  GHND_CHECK_ARG(d->pad >= 0 && d->pad < 4, "wgrad: pad %d", d->pad);
  GHND_CHECK_ARG(d->x && d->dy && d->dw, "wgrad: null tensor");
  GHND_CHECK_ARG((d->x_fmt == GHND_F16 || d->x_fmt == GHND_BF16) &&
                     (d->dy_fmt == GHND_F16 || d->dy_fmt == GHND_BF16),
                 "wgrad: bad format");
  GHND_CHECK_ARG(d->x_fmt == d->dy_fmt,
                 "wgrad: x and dy must share one 16-bit format (tcgen05 kind::f16 rejects mixed "
                 "f16 x bf16 operands on sm_100a)");
  const int Ho = d->H + 2 * d->pad - d->R + 1;
  const int Wo = d->W + 2 * d->pad - d->S + 1;
  GHND_CHECK_ARG(Ho > 0 && Wo > 0, "wgrad: empty output");

  ghnd_wgrad_plan* plan = new ghnd_wgrad_plan();
  WgradParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  p.rows_is_dy = (d->K >= 128 || d->C < 128) ? 1 : 0;
  const int rows_ch = p.rows_is_dy ? d->K : d->C;
  const int cols_ch = p.rows_is_dy ? d->C : d->K;
  p.n_taps = d->R * d->S;
  p.S = d->S;
  p.pad = d->pad;
  p.rows_valid = rows_ch >= 128 ? 128 : 64;
  p.row_boxes = p.rows_valid / 64;
  p.m_tiles = (rows_ch + 127) / 128;
  p.nb = cols_ch % 128 == 0 ? 128 : 64;
  p.col_boxes = p.nb / 64;
  p.n_chunks = cols_ch / p.nb;
  p.n_img = d->N;
  p.ho = Ho;
  p.wo = Wo;
  p.tiles_w = (Wo + kWgPix - 1) / kWgPix;
  p.total_pix_tiles = d->N * Ho * p.tiles_w;
  p.fd_tiles_per_img = make_fastdiv(Ho * p.tiles_w);
  p.fd_tiles_w = make_fastdiv(p.tiles_w);
  for (int tap = 0; tap < p.n_taps; ++tap) {
    p.tap_dh[tap] = tap / d->S - d->pad;
    p.tap_dw[tap] = tap % d->S - d->pad;
  }
  const int blocks = p.m_tiles * p.n_chunks;
  int splits = num_sms() / blocks;
  if (splits < 1) splits = 1;
  if (splits > p.total_pix_tiles) splits = p.total_pix_tiles;
  p.splits = splits;
  // stage: rows operand boxes + cols operand boxes; the shifted (x) side is replicated per tap
  const int row_bytes = p.row_boxes * kWgBox, col_bytes = p.col_boxes * kWgBox;
  p.stage_bytes = p.rows_is_dy ? row_bytes + p.n_taps * col_bytes : p.n_taps * row_bytes + col_bytes;
  int stages = (200 * 1024) / p.stage_bytes;
  if (stages > 8) stages = 8;
  p.n_stages = stages;
  const int a_fmt = p.rows_is_dy ? d->dy_fmt : d->x_fmt;
  const int b_fmt = p.rows_is_dy ? d->x_fmt : d->dy_fmt;
  p.idesc = make_idesc(a_fmt, b_fmt, 1, 1, 128, p.nb);
  p.dw = d->dw;
  p.K = d->K;
  p.C = d->C;

  uint32_t box[4] = {64, (uint32_t)kWgPix, 1, 1};
  uint64_t dims_dy[4] = {(uint64_t)d->K, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)d->N};
  uint64_t str_dy[4] = {2, (uint64_t)d->K * 2, (uint64_t)Wo * d->K * 2, (uint64_t)Ho * Wo * d->K * 2};
  uint64_t dims_x[4] = {(uint64_t)d->C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
  uint64_t str_x[4] = {2, (uint64_t)d->C * 2, (uint64_t)d->W * d->C * 2,
                       (uint64_t)d->H * d->W * d->C * 2};
  CUtensorMap map_dy, map_x;
  int rc = encode_tmap(&map_dy, 2, 4, const_cast<void*>(d->dy), dims_dy, str_dy, box, 128);
  if (rc == GHND_OK) rc = encode_tmap(&map_x, 2, 4, const_cast<void*>(d->x), dims_x, str_x, box, 128);
  if (rc != GHND_OK) {
    delete plan;
    return rc;
  }
  p.tmap_row = p.rows_is_dy ? map_dy : map_x;
  p.tmap_col = p.rows_is_dy ? map_x : map_dy;
  plan->grid = blocks * splits;
  plan->smem = (size_t)p.n_stages * p.stage_bytes + 1024 + 256;
  plan->dw_bytes = (size_t)d->K * p.n_taps * d->C * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      delete plan;
      return cuda_fail(e, "cudaFuncSetAttribute(wgrad_tc_kernel)");
    }
    attr_set = true;
  }
  *out = plan;
  return GHND_OK;
}

int ghnd_wgrad_plan_run(const ghnd_wgrad_plan_t* plan, void* stream) {
  using namespace ghnd;
  GHND_CHECK_ARG(plan != nullptr, "wgrad_plan_run: null plan");
  cudaStream_t st = (cudaStream_t)stream;
  GHND_CUDA(cudaMemsetAsync(plan->p.dw, 0, plan->dw_bytes, st));
  wgrad_tc_kernel<<<plan->grid, kWgThreads, plan->smem, st>>>(plan->p);
  GHND_LAUNCH_CHECK("wgrad_tc_kernel");
  return GHND_OK;
}

void ghnd_wgrad_plan_destroy(ghnd_wgrad_plan_t* plan) { delete plan; }
}
