// Weight gradient of the student's wide k=2 convs on tcgen05 (sm_100a).
//
//   dw[k][tap][c] = sum_{pixels} dy[pixel][k] * x[pixel + tap_offset][c]
//
// Both operands are read straight out of the NHWC tensors by TMA as 128B-swizzled [pixels][64 channels]
// boxes, i.e. they sit in shared memory "MN-major" (channel contiguous, the GEMM reduction axis = pixel
// is the strided one); the instruction descriptor's a_major/b_major bits select that layout, so no
// transpose pass is needed.  A pipeline stage reduces over 2 image rows x 32 pixels of dy:
//   dy box  {64 ch, 32 px, 2 rows}                              =  8 KB per 64 channels
//   x  box  {64 ch, 32+S-1 px (rounded up to 8), 2+R-1 rows}    = 15 KB per 64 channels for 2x2 taps
// ONE x box serves every tap: tap (r, s) is the same box read through a descriptor whose start address is
// shifted by (r * box width + s) pixels = that many 128-byte rows (the 128B swizzle is a function of the
// absolute shared-memory address, so a shifted start stays consistent with what TMA wrote).  The first
// version loaded one shifted x box per tap (5-10 boxes of 4 KB per 32 pixels): 2.5x the L2 -> SM traffic
// and ~4x the TMA issues, 1.8 TB/s on the 64-channel dWs.
//
// One CTA owns a 128 (rows operand) x NB (cols operand, <=128) output block for ALL taps (<=4):
// the accumulators fill TMEM (4 x 128 columns).  The pixel axis is split over CTAs; every CTA adds
// its partial block into the fp32 dw with red.global.add.f32 (dw is zeroed by the plan first).
// Zero padding of x and the ragged right / bottom edge of dy both come from TMA out-of-bounds zero fill.
#include <stdlib.h>
#include <vector>

#include "common.cuh"

namespace ghnd {

static constexpr int kWgThreads = 224;  // warp0/6 TMA producers, warp1 MMA, warps 2..5 epilogue
static constexpr int kWgPix = 32;                  // dy pixels along W per box row
static constexpr int kWgRows = 2;                  // dy rows per box: a stage reduces over 2 x 32 pixels
static constexpr int kWgDyBox = kWgRows * kWgPix * 128;  // bytes of one dy box = 8 KB
static constexpr int kWgMaxTaps = 4;

struct WgradParams {
  CUtensorMap tmap_row;   // rows operand tensor (GEMM M): dy box {64, 32, 2, 1} or x box {64, xw, xr, 1}
  CUtensorMap tmap_col;   // cols operand tensor (GEMM N)
  int xw;                 // x box width in pixels (32 + S - 1, rounded up to 8)
  int x_box_bytes;        // xw * (2 + R - 1) * 128
  int debug;              // GHND_WGRAD_DEBUG: 1 = no MMAs, 2 = no TMA loads, 4 = no epilogue (timing experiments)
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
  int hp;                 // row pairs per image = ceil(ho / 2)
  int tiles_w;            // ceil(wo / 32)
  int total_pix_tiles;    // n_img * hp * tiles_w
  int stage_bytes, n_stages;
  FastDiv fd_tiles_per_img, fd_tiles_w;
  uint32_t idesc;
  float* dw;              // [K][taps][C]
  int K, C;
};


// two K=16 MMAs (32 pixels) of a stage for one tap (second descriptor = +2048 B); a stage issues two such pairs
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
  // stage = [rows operand boxes][cols operand boxes]; the dy side has 8 KB boxes, the x side x_box_bytes
  const int row_box = p.rows_is_dy ? kWgDyBox : p.x_box_bytes;
  const int col_box = p.rows_is_dy ? p.x_box_bytes : kWgDyBox;
  const int a_bytes = p.row_boxes * row_box;

  if (warp == 0 || warp == 6) {
    // Two TMA producer warps (whole warp converged, one elected lane issues): warp 0 loads the rows operand,
    // warp 6 the cols operand of every stage; each posts its own byte count on the stage's barrier.
    const bool is_rows = warp == 0;
    const bool is_dy = is_rows == (p.rows_is_dy != 0);
    const CUtensorMap* map = is_rows ? &p.tmap_row : &p.tmap_col;
    const int n_box = is_rows ? p.row_boxes : p.col_boxes;
    const int box_bytes = is_rows ? row_box : col_box;
    const int ch0 = is_rows ? m_tile * 128 : n_chunk * p.nb;
    const int off = is_dy ? 0 : -p.pad;  // x box origin = dy position - pad
    int stage = 0;
    uint32_t phase = 0;
    // (img, h, wt) of the first tile by multiply-high division, then carried incrementally
    int img, rem, h, wt;
    fd_divmod(p.fd_tiles_per_img, t_begin, img, rem);
    fd_divmod(p.fd_tiles_w, rem, h, wt);
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* dst = smem + (size_t)stage * p.stage_bytes + (is_rows ? 0 : a_bytes);
        mbar_arrive_expect_tx(&full_bar[stage], (p.debug & 2) ? 0u : (uint32_t)(n_box * box_bytes));
        for (int b = 0; b < n_box && !(p.debug & 2); ++b)
          tma_load_4d(dst + b * box_bytes, map, &full_bar[stage], ch0 + b * 64, wt * kWgPix + off,
                      kWgRows * h + off, img);
      }
      __syncwarp();
      if (++wt == p.tiles_w) {
        wt = 0;
        if (++h == p.hp) {
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
    const uint32_t a_lbo = p.row_boxes == 2 ? (uint32_t)row_box : 0u;
    const uint32_t b_lbo = (uint32_t)col_box;
    const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | ((uint32_t)UMMA_SW128 << 29);
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_base + (uint32_t)(stage * p.stage_bytes);
        const uint32_t sb = sa + (uint32_t)a_bytes;
        const uint32_t dy_base = p.rows_is_dy ? sa : sb, x_base = p.rows_is_dy ? sb : sa;
        for (int tap = 0; tap < p.n_taps && !(p.debug & 1); ++tap) {
          const int tr = tap / p.S, ts = tap - tr * p.S;
#pragma unroll
          for (int r = 0; r < kWgRows; ++r) {
            // dy row r of the stage against x row r + tr, shifted by ts pixels (128 B per pixel)
            const uint32_t dy_addr = dy_base + (uint32_t)(r * kWgPix * 128);
            const uint32_t x_addr = x_base + (uint32_t)(((r + tr) * p.xw + ts) * 128);
            const uint32_t a_addr = p.rows_is_dy ? dy_addr : x_addr;
            const uint32_t b_addr = p.rows_is_dy ? x_addr : dy_addr;
            const uint32_t a_lo = ((a_addr >> 4) & 0x3fffu) | ((a_lbo >> 4) << 16);
            const uint32_t b_lo = ((b_addr >> 4) & 0x3fffu) | ((b_lbo >> 4) << 16);
            // 16 pixels = two 8-row swizzle groups = 2048 B -> +128 in (addr >> 4)
            umma_pair_wg(tmem_base + (uint32_t)(tap * 128), a_lo, b_lo, desc_hi, p.idesc,
                         (uint32_t)(t != t_begin || r != 0));
          }
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
    if (t_end > t_begin && !(p.debug & 4)) {
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
  p.hp = (Ho + kWgRows - 1) / kWgRows;
  p.total_pix_tiles = d->N * p.hp * p.tiles_w;
  p.fd_tiles_per_img = make_fastdiv(p.hp * p.tiles_w);
  p.fd_tiles_w = make_fastdiv(p.tiles_w);
  const int blocks = p.m_tiles * p.n_chunks;
  int splits = num_sms() / blocks;
  if (splits < 1) splits = 1;
  if (splits > p.total_pix_tiles) splits = p.total_pix_tiles;
  p.splits = splits;
  // stage: rows operand boxes + cols operand boxes (dy: 8 KB, x: one halo box serving every tap)
  p.xw = (kWgPix + d->S - 1 + 7) / 8 * 8;
  p.x_box_bytes = p.xw * (kWgRows + d->R - 1) * 128;
  const int row_bytes = p.row_boxes * (p.rows_is_dy ? kWgDyBox : p.x_box_bytes);
  const int col_bytes = p.col_boxes * (p.rows_is_dy ? p.x_box_bytes : kWgDyBox);
  p.stage_bytes = row_bytes + col_bytes;
  int smem_kb = 200;
  if (const char* e = getenv("GHND_WGRAD_SMEM_KB")) smem_kb = atoi(e);
  int stages = (smem_kb * 1024) / p.stage_bytes;
  if (stages > 12) stages = 12;
  p.n_stages = stages;
  {
    const char* e = getenv("GHND_WGRAD_DEBUG");
    p.debug = e ? atoi(e) : 0;
  }
  const int a_fmt = p.rows_is_dy ? d->dy_fmt : d->x_fmt;
  const int b_fmt = p.rows_is_dy ? d->x_fmt : d->dy_fmt;
  p.idesc = make_idesc(a_fmt, b_fmt, 1, 1, 128, p.nb);
  p.dw = d->dw;
  p.K = d->K;
  p.C = d->C;

  uint32_t box_dy[4] = {64, (uint32_t)kWgPix, (uint32_t)kWgRows, 1};
  uint32_t box_x[4] = {64, (uint32_t)p.xw, (uint32_t)(kWgRows + d->R - 1), 1};
  uint64_t dims_dy[4] = {(uint64_t)d->K, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)d->N};
  uint64_t str_dy[4] = {2, (uint64_t)d->K * 2, (uint64_t)Wo * d->K * 2, (uint64_t)Ho * Wo * d->K * 2};
  uint64_t dims_x[4] = {(uint64_t)d->C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
  uint64_t str_x[4] = {2, (uint64_t)d->C * 2, (uint64_t)d->W * d->C * 2,
                       (uint64_t)d->H * d->W * d->C * 2};
  CUtensorMap map_dy, map_x;
  int rc = encode_tmap(&map_dy, 2, 4, const_cast<void*>(d->dy), dims_dy, str_dy, box_dy, 128);
  if (rc == GHND_OK) rc = encode_tmap(&map_x, 2, 4, const_cast<void*>(d->x), dims_x, str_x, box_x, 128);
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
    // keep the maximum shared-memory split whatever the ring needs, so that streaming kernels of the other
    // stream (which ask for the same split) can join the SMs this kernel runs on
    cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                         (int)cudaSharedmemCarveoutMaxShared);
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
