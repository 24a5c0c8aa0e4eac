// Weight gradient of the stem conv1 (7x7 s2 p3, 3->64) on tcgen05 (sm_100a).
//
//   dw[k][r][s*4+c] = sum_{n,ho,wo} g[n][ho][wo][k] * xp[n][2ho+r][2wo+s][c]       (xp = packed image)
//
// GEMM view: M = (filter row r, 32-element window = 8 px x 4 ch) -> 7 x 32 = 224 rows in two
// 128-row tiles, N = 64 output channels, K = output pixels (~1.1 M per 4-image batch).  Both
// operands come straight from global memory by TMA with the pixel axis as the strided (GEMM-K)
// axis, i.e. they sit in shared memory MN-major:
//   A        : ONE box {32 elem, 32 px, 13 image rows} per stage  -> 13 atoms [32 px][64 B], SWIZZLE_64B
//   B        : box {64 ch,  32 px, 4 output rows} of g            -> 4 x [32 px][128 B], SWIZZLE_128B
// A stage covers FOUR output rows ho0..ho0+3 of 32 pixels: output row ho0+i reads the image rows 2(ho0+i)+r, i.e.
// atoms 2i..2i+6 of the 13 resident ones -- the MMA descriptors of row i simply start 2i atoms further.  With one
// output row per stage every image row was fetched 3.5 times per column class (7 atoms per 32 pixels: 14 KB of A
// for 4 KB of B, 625 MB of L2 -> SM traffic per launch, the limit of this kernel); now it is 13 atoms per 4 rows.
// Output columns are split in four classes wo = 4j+q so that consecutive pixels of a class start
// 64 B apart in the packed row (no overlapping TMA rows), exactly like the forward stem kernel.
// Every CTA owns a contiguous range of pixel tiles, accumulates the whole 224x64 dw in TMEM
// (2 x 64 columns) and adds it once into the fp32 workspace [64][196] (index r*28 + s*4 + c).
#include "common.cuh"

namespace ghnd {

static constexpr int kSwtThreads = 192;      // warp0 TMA, warp1 MMA, warps 2..5 epilogue
static constexpr int kSwtPix = 32;           // pixels along W per stage
static constexpr int kSwtRows = 4;           // output rows per stage: GEMM-K = 4 x 32 pixels
static constexpr int kSwtAtom = kSwtPix * 64;   // one [32 px][64 B] A atom = 2 KB
static constexpr int kSwtImgRows = 2 * (kSwtRows - 1) + 7;  // 13 image rows feed 4 output rows
static constexpr int kSwtABytes = (kSwtImgRows + 1) * kSwtAtom;  // + 1 unused atom (M = 2 x 128 reads 8 per row)
static constexpr int kSwtBRow = kSwtPix * 128;  // g of one output row: 4 KB
static constexpr int kSwtBBytes = kSwtRows * kSwtBRow;
static constexpr int kSwtStage = kSwtABytes + kSwtBBytes;  // 44 KB
static constexpr int kSwtStages = 4;

struct StemWgradParams {
  CUtensorMap tmap_x[4];  // per column class q
  CUtensorMap tmap_g[4];
  int n_img, ho, tiles_j;  // ho = groups of kSwtRows output rows per image
  int total_tiles;        // n_img * ho * 4 * tiles_j
  FastDiv fd_tiles_j, fd_ho;
  uint32_t idesc;
  float* accum;           // [64][196]
};

// the two K=16 MMAs of one 32-pixel stage for one M tile: A advances 16 rows x 64 B, B 16 x 128 B
__device__ __forceinline__ void umma_pair_stem(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo,
                                               uint32_t a_hi, uint32_t b_hi, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 a, b;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.eq.b32 q, 0, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "add.u32 a, %1, 64;\n\t"
      "add.u32 b, %2, 128;\n\t"
      "mov.b64 da, {a, %3};\n\t"
      "mov.b64 db, {b, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, q;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(a_hi), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(kSwtThreads, 1)
    stem_wgrad_tc_kernel(const __grid_constant__ StemWgradParams p) {
  // aligned by declaration; plain pointer arithmetic keeps the shared state space (LDS/STS, not generic)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (smem_u32(smem_raw) & 1023u) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kSwtStages * kSwtStage);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kSwtStages;
  uint64_t* done_bar = bars + 2 * kSwtStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int q = 0; q < 4; ++q) {
      prefetch_tmap(&p.tmap_x[q]);
      prefetch_tmap(&p.tmap_g[q]);
    }
    for (int i = 0; i < kSwtStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  // the atom after the last image row is read by the MMAs of the stage's last output row (rows 96..127 of
  // M tile 1, discarded; for the other output rows those rows see a real image row): defined contents once
  for (int i = threadIdx.x; i < kSwtStages * (kSwtAtom / 16); i += kSwtThreads) {
    const int st = i / (kSwtAtom / 16), o = i - st * (kSwtAtom / 16);
    reinterpret_cast<uint4*>(smem + (size_t)st * kSwtStage + kSwtImgRows * kSwtAtom)[o] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  const int t_begin = (int)(((int64_t)p.total_tiles * blockIdx.x) / gridDim.x);
  const int t_end = (int)(((int64_t)p.total_tiles * (blockIdx.x + 1)) / gridDim.x);

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int t = t_begin; t < t_end; ++t) {
      int jt, rest, ho, img;
      fd_divmod(p.fd_tiles_j, t, rest, jt);
      const int q = rest & 3;
      fd_divmod(p.fd_ho, rest >> 2, img, ho);
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* sa = smem + (size_t)stage * kSwtStage;
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(kSwtImgRows * kSwtAtom + kSwtBBytes));
        // ho = group of four output rows: image rows 8*ho .. 8*ho + 12, g rows 4*ho .. 4*ho + 3 (rows past the
        // tensors are zero-filled by TMA)
        tma_load_4d(sa, &p.tmap_x[q], &full_bar[stage], 0, jt * kSwtPix, 2 * kSwtRows * ho, img);
        tma_load_4d(sa + kSwtABytes, &p.tmap_g[q], &full_bar[stage], 0, jt * kSwtPix, kSwtRows * ho, img);
      }
      __syncwarp();
      if (++stage == kSwtStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t smem_base = smem_u32(smem);
    // A: MN-major SWIZZLE_64B, LBO = atom stride (2 KB), SBO = 8 K-rows x 64 B
    const uint32_t a_hi = (512u >> 4) | (1u << 14) | ((uint32_t)UMMA_SW64 << 29);
    // B: MN-major SWIZZLE_128B, single 64-channel atom, SBO = 8 K-rows x 128 B
    const uint32_t b_hi = (1024u >> 4) | (1u << 14) | ((uint32_t)UMMA_SW128 << 29);
    for (int t = t_begin; t < t_end; ++t) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_base + (uint32_t)(stage * kSwtStage);
#pragma unroll
        for (int i = 0; i < kSwtRows; ++i) {  // output row i of the stage: atoms 2i .. 2i + 7, g row i
          const uint32_t b_lo =
              (((sa + kSwtABytes + i * kSwtBRow) >> 4) & 0x3fffu) | ((uint32_t)(kSwtBRow >> 4) << 16);
#pragma unroll
          for (int m = 0; m < 2; ++m) {
            const uint32_t a_lo =
                (((sa + (2 * i + m * 4) * kSwtAtom) >> 4) & 0x3fffu) | ((uint32_t)(kSwtAtom >> 4) << 16);
            umma_pair_stem(tmem_base + (uint32_t)(m * 64), a_lo, b_lo, a_hi, b_hi, p.idesc,
                           (uint32_t)(t != t_begin || i != 0));
          }
        }
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == kSwtStages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (elect_one()) umma_commit(done_bar);
    __syncwarp();
  } else if (t_end > t_begin) {
    const int quarter = warp & 3;
    mbar_wait(done_bar, 0);
    tc_fence_after();
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      uint32_t v[64];
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(m * 64);
      tmem_ld32(t_addr, v);
      tmem_ld32(t_addr + 32, v + 32);
      tmem_ld_wait();
      const int r = m * 4 + quarter;  // filter row of this warp's atom; lane = s*4 + c
      if (r < 7 && lane < 28) {
        float* dst = p.accum + r * 28 + lane;
#pragma unroll
        for (int k = 0; k < 64; ++k) atomicAdd(dst + (size_t)k * 196, __uint_as_float(v[k]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// conv_narrow.cu
void launch_stem_wgrad_finish(const float* accum, const float* scale, float* dw, cudaStream_t st);

}  // namespace ghnd

struct ghnd_stem_wgrad_plan {
  ghnd::StemWgradParams p;
  const float* scale;
  float* dw;
  int grid;
  size_t smem;
};

extern "C" {

int ghnd_stem_wgrad_plan_create(const void* x_packed, const void* g, int fmt, const float* scale_o,
                                float* dw_oihw, int N, int Hp, int Wp, void* workspace,
                                size_t workspace_bytes, ghnd_stem_wgrad_plan_t** out) {
  using namespace ghnd;
  GHND_CHECK_ARG(out && x_packed && g && dw_oihw && workspace, "stem_wgrad_plan_create: null pointer");
  *out = nullptr;
  GHND_CHECK_ARG(fmt == GHND_F16 || fmt == GHND_BF16, "stem_wgrad_plan: bad format");
  GHND_CHECK_ARG(N > 0 && Hp > 0 && Wp > 0 && Hp % 2 == 0 && Wp % 8 == 0,
                 "stem_wgrad_plan: padded size must be even x multiple of 8 (Hp=%d Wp=%d)", Hp, Wp);
  GHND_CHECK_ARG(workspace_bytes >= ghnd_stem_wgrad_workspace_bytes(),
                 "stem_wgrad_plan: workspace too small");
  const int Ho = Hp / 2, Wo = Wp / 2;
  const int rows = Hp + 6, RP = (Wp + 8) * 4;  // packed image rows / row pitch in elements
  ghnd_stem_wgrad_plan* plan = new ghnd_stem_wgrad_plan();
  StemWgradParams& p = plan->p;
  memset(&p, 0, sizeof(p));
  int rc = GHND_OK;
  const int J0 = (Wo + 3) / 4;
  for (int q = 0; q < 4 && rc == GHND_OK; ++q) {
    const int J = (Wo - q + 3) / 4;  // columns wo = 4j+q of this class
    {
      const uint8_t* base = static_cast<const uint8_t*>(x_packed) + (size_t)q * 8 * 2;
      uint64_t dims[4] = {32, (uint64_t)J, (uint64_t)rows, (uint64_t)N};
      uint64_t str[4] = {2, 64, (uint64_t)RP * 2, (uint64_t)rows * RP * 2};
      uint32_t box[4] = {32, (uint32_t)kSwtPix, (uint32_t)kSwtImgRows, 1};
      rc = encode_tmap(&p.tmap_x[q], 2, 4, const_cast<uint8_t*>(base), dims, str, box, 64);
    }
    if (rc == GHND_OK) {
      const uint8_t* base = static_cast<const uint8_t*>(g) + (size_t)q * 64 * 2;
      uint64_t dims[4] = {64, (uint64_t)J, (uint64_t)Ho, (uint64_t)N};
      uint64_t str[4] = {2, 4 * 128, (uint64_t)Wo * 128, (uint64_t)Ho * Wo * 128};
      uint32_t box[4] = {64, (uint32_t)kSwtPix, (uint32_t)kSwtRows, 1};
      rc = encode_tmap(&p.tmap_g[q], 2, 4, const_cast<uint8_t*>(base), dims, str, box, 128);
    }
  }
  if (rc != GHND_OK) {
    delete plan;
    return rc;
  }
  p.n_img = N;
  p.ho = (Ho + kSwtRows - 1) / kSwtRows;
  p.tiles_j = (J0 + kSwtPix - 1) / kSwtPix;
  p.total_tiles = N * p.ho * 4 * p.tiles_j;
  p.fd_tiles_j = make_fastdiv(p.tiles_j);
  p.fd_ho = make_fastdiv(p.ho);
  p.idesc = make_idesc(fmt, fmt, 1, 1, 128, 64);
  p.accum = static_cast<float*>(workspace);
  plan->scale = scale_o;
  plan->dw = dw_oihw;
  plan->grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  plan->smem = (size_t)kSwtStages * kSwtStage + 1024 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(stem_wgrad_tc_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      delete plan;
      return cuda_fail(e, "cudaFuncSetAttribute(stem_wgrad_tc_kernel)");
    }
    attr_set = true;
  }
  *out = plan;
  return GHND_OK;
}

int ghnd_stem_wgrad_plan_run(const ghnd_stem_wgrad_plan_t* plan, void* stream) {
  using namespace ghnd;
  GHND_CHECK_ARG(plan != nullptr, "stem_wgrad_plan_run: null plan");
  cudaStream_t st = (cudaStream_t)stream;
  GHND_CUDA(cudaMemsetAsync(plan->p.accum, 0, ghnd_stem_wgrad_workspace_bytes(), st));
  stem_wgrad_tc_kernel<<<plan->grid, kSwtThreads, plan->smem, st>>>(plan->p);
  GHND_LAUNCH_CHECK("stem_wgrad_tc_kernel");
  launch_stem_wgrad_finish(plan->p.accum, plan->scale, plan->dw, st);
  GHND_LAUNCH_CHECK("stem_wgrad_finish_kernel");
  return GHND_OK;
}

void ghnd_stem_wgrad_plan_destroy(ghnd_stem_wgrad_plan_t* plan) { delete plan; }
}
