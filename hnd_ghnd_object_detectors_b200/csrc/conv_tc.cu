// Wide convolutions as implicit GEMM on tcgen05 tensor cores (sm_100a).
//
//   D[pixel][out_ch] = sum_{tap, c} A_tap[pixel][c] * B[out_ch][tap*Cin + c]
//
// * M tile  = a TH x TW spatial patch of one image (TH*TW <= 128 rows of the UMMA M=128 tile);
//             for 1x1/stride-1 convs the whole N*H*W pixel axis is flattened (TW=128, TH=1).
// * A tiles = one TMA box {64 ch, TW, TH, 1} per (tap, 64-channel chunk) taken from the NHWC
//             tensor at the tap's shifted coordinates; TMA out-of-bounds zero fill implements the
//             convolution padding.  Stride-2 convs read through parity-lattice views of the tensor
//             (one descriptor per (h,w) parity), so only plain tiled TMA is needed.
// * B tiles = TMA box {64 k, BLOCK_N rows} of the K-major packed weights.
// * both land in 128B-swizzled K-major shared memory and feed tcgen05.mma.cta_group::1.kind::f16
//   (M=128, N=BLOCK_N, K=16) issued by one thread; fp32 accumulators live in TMEM, double
//   buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
// * persistent CTAs (one per SM), warp-specialised: warp0 TMA producer, warp1 MMA issuer (+TMEM
//   alloc), warps 2..5 epilogue (tcgen05.ld -> bias/residual/ReLU/mask -> 16-bit NHWC stores).
//
// The same kernel runs the data-gradient: dgrad of a stride-1 conv is a conv of dy with mirrored
// tap offsets over the transposed weights; dgrad of a stride-2 conv is four launches, one per
// output parity class, each using the taps that hit that class.
#include <vector>

#include "common.cuh"

namespace ghnd {

static constexpr int kConvThreads = 192;
static constexpr int kBlockM = 128;
static constexpr int kMaxTaps = 9;
// GEMM-K per pipeline stage: 64 x 16-bit = 128 B rows (SWIZZLE_128B) for the wide convs, or
// 32 x 16-bit = 64 B rows (SWIZZLE_64B) for the stem whose im2col row is 7 px x 4 ch (+4 zero).

struct ConvTap {
  int16_t dh, dw;   // offset added to the tile origin, in the coordinates of view `map`
  int16_t map;      // which A descriptor (parity view)
  int16_t wk;       // tap index inside the packed weights (K offset = wk * Cin_gemm)
};

struct ConvKernelParams {
  CUtensorMap tmap_a[4];
  CUtensorMap tmap_b;
  ConvTap taps[kMaxTaps];
  int n_taps;
  int cin;          // GEMM reduction channels per tap (multiple of 64)
  int cout;         // GEMM output channels (multiple of block_n)
  int block_n;      // 64 / 128 / 256
  int n_stages;
  int kblock;       // 64 or 32
  int a_bytes;      // 128 rows * kblock * 2
  int stage_bytes;  // a_bytes + block_n*kblock*2
  int a_box_bytes;  // TW*TH*kblock*2
  // tile grid over the GEMM-M space
  int n_img;        // images (1 when flattened)
  int tiles_h, tiles_w, th, tw;
  int ho, wo;       // extent of the M space per image (rows / cols of output positions)
  int n_tiles_n;
  int total_tiles;
  // dst addressing: GEMM position (n, i, j) -> dst pixel (n, i*osh + ooh, j*osw + oow) in [n_img][OH][OW]
  int oh_full, ow_full, osh, osw, ooh, oow;
  uint32_t idesc;
  // epilogue
  void* dst;
  const float* bias;
  const void* residual;
  const void* mask;
  int dst_fmt, res_fmt, mask_fmt;
  int relu, accumulate;
};

__global__ void __launch_bounds__(kConvThreads, 1)
    conv_tc_kernel(const __grid_constant__ ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.n_stages * p.stage_bytes);
  uint64_t* full_bar = bars;                    // [n_stages]
  uint64_t* empty_bar = bars + p.n_stages;      // [n_stages]
  uint64_t* tfull_bar = bars + 2 * p.n_stages;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) prefetch_tmap(&p.tmap_a[i]);
    prefetch_tmap(&p.tmap_b);
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int k_chunks = p.cin / p.kblock;
  const int row_bytes = p.kblock * 2;
  const int k_iters = p.n_taps * k_chunks;
  const int tiles_per_img = p.tiles_h * p.tiles_w;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles_n;
        const int m_tile = tile / p.n_tiles_n;
        const int img = m_tile / tiles_per_img;
        const int rem = m_tile - img * tiles_per_img;
        const int h0 = (rem / p.tiles_w) * p.th;
        const int w0 = (rem % p.tiles_w) * p.tw;
        for (int t = 0; t < p.n_taps; ++t) {
          const ConvTap tap = p.taps[t];
          for (int kc = 0; kc < k_chunks; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
            uint8_t* sb = sa + p.a_bytes;
            mbar_arrive_expect_tx(&full_bar[stage],
                                  (uint32_t)(p.a_box_bytes + p.block_n * row_bytes));
            tma_load_4d(sa, &p.tmap_a[tap.map], &full_bar[stage], kc * p.kblock, w0 + tap.dw,
                        h0 + tap.dh, img);
            tma_load_2d(sb, &p.tmap_b, &full_bar[stage], tap.wk * p.cin + kc * p.kblock,
                        n_tile * p.block_n);
            if (++stage == p.n_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint32_t sw_layout = p.kblock == 64 ? UMMA_SW128 : UMMA_SW64;
      const uint32_t sbo = 8u * (uint32_t)row_bytes;  // 8-row swizzle atom
      const int k_steps = p.kblock / 16;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(&tempty_bar[buf], (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.block_n);
        for (int ki = 0; ki < k_iters; ++ki) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * p.stage_bytes);
          const uint32_t sb = sa + (uint32_t)p.a_bytes;
          const uint64_t adesc = make_smem_desc(sa, 16, sbo, sw_layout);
          const uint64_t bdesc = make_smem_desc(sb, 16, sbo, sw_layout);
          for (int k = 0; k < k_steps; ++k) {
            // advance 16 elements (32 B) along K inside the swizzle row: +2 in (addr>>4)
            umma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc,
                     (uint32_t)((ki | k) != 0));
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (++stage == p.n_stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[buf]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // ===================== epilogue (4 warps, TMEM lane quarter = warp % 4) =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int r_h = row / p.tw;
    const int r_w = row - r_h * p.tw;
    const bool row_in_tile = r_h < p.th;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const int n_tile = tile % p.n_tiles_n;
      const int m_tile = tile / p.n_tiles_n;
      const int img = m_tile / tiles_per_img;
      const int rem = m_tile - img * tiles_per_img;
      const int gi = (rem / p.tiles_w) * p.th + r_h;
      const int gj = (rem % p.tiles_w) * p.tw + r_w;
      const bool valid = row_in_tile && gi < p.ho && gj < p.wo;
      const size_t pix =
          ((size_t)img * p.oh_full + (size_t)(gi * p.osh + p.ooh)) * p.ow_full + (gj * p.osw + p.oow);
      const size_t off = pix * (size_t)p.cout + (size_t)n_tile * p.block_n;

      mbar_wait(&tfull_bar[buf], use & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * p.block_n);
      for (int c = 0; c < p.block_n; c += 32) {
        uint32_t r[32];
        tmem_ld32(t_addr + (uint32_t)c, r);
        tmem_ld_wait();
        if (valid) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (p.bias != nullptr) {
            const float4* b4 = reinterpret_cast<const float4*>(p.bias + n_tile * p.block_n + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 b = __ldg(b4 + j);
              v[4 * j] += b.x;
              v[4 * j + 1] += b.y;
              v[4 * j + 2] += b.z;
              v[4 * j + 3] += b.w;
            }
          }
          if (p.residual != nullptr) {
            const uint4* r4 =
                reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.residual) + off + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 q = __ldg(r4 + j);
              uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 f = unpack2(w[e], p.res_fmt);
                v[8 * j + 2 * e] += f.x;
                v[8 * j + 2 * e + 1] += f.y;
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.mask != nullptr) {
            const uint4* m4 =
                reinterpret_cast<const uint4*>(static_cast<const uint16_t*>(p.mask) + off + c);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 q = __ldg(m4 + j);
              uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 f = unpack2(w[e], p.mask_fmt);
                if (!(f.x > 0.f)) v[8 * j + 2 * e] = 0.f;
                if (!(f.y > 0.f)) v[8 * j + 2 * e + 1] = 0.f;
              }
            }
          }
          uint4* d4 = reinterpret_cast<uint4*>(static_cast<uint16_t*>(p.dst) + off + c);
          if (p.accumulate) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 q = d4[j];
              uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float2 f = unpack2(w[e], p.dst_fmt);
                v[8 * j + 2 * e] += f.x;
                v[8 * j + 2 * e + 1] += f.y;
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack2(v[8 * j], v[8 * j + 1], p.dst_fmt);
            o.y = pack2(v[8 * j + 2], v[8 * j + 3], p.dst_fmt);
            o.z = pack2(v[8 * j + 4], v[8 * j + 5], p.dst_fmt);
            o.w = pack2(v[8 * j + 6], v[8 * j + 7], p.dst_fmt);
            d4[j] = o;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: plan = list of launches (1, or 4 parity classes for a stride-2 dgrad)
// ------------------------------------------------------------------------------------------------
struct ConvLaunch {
  ConvKernelParams p;
  int grid;
  size_t smem;
};

}  // namespace ghnd

struct ghnd_conv_plan {
  std::vector<ghnd::ConvLaunch> launches;
};

namespace ghnd {

static void choose_tile(int ho, int wo, int* th, int* tw) {
  double best = -1.0;
  int bth = 1, btw = 1;
  for (int w = 1; w <= 128 && w <= wo; ++w) {
    int h = 128 / w;
    if (h > ho) h = ho;
    if (h < 1) continue;
    const int tiles = ((ho + h - 1) / h) * ((wo + w - 1) / w);
    const double eff = (double)ho * wo / ((double)tiles * 128.0);
    if (eff > best + 1e-9) {
      best = eff;
      bth = h;
      btw = w;
    }
  }
  *th = bth;
  *tw = btw;
}

// A-operand view of an NHWC tensor restricted to the (ph,pw) parity lattice when sub==2.
static int make_a_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int sub, int ph,
                      int pw, int th, int tw) {
  const int hv = sub == 1 ? H : (H - ph + 1) / 2;
  const int wv = sub == 1 ? W : (W - pw + 1) / 2;
  if (hv <= 0 || wv <= 0) {
    // empty lattice (H==1 or W==1): encode a 1x1 view; taps using it are never generated
    return make_a_map(m, base, N, H, W, C, 1, 0, 0, th, tw);
  }
  const uint8_t* b = static_cast<const uint8_t*>(base) + ((size_t)ph * W + pw) * C * 2;
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)wv, (uint64_t)hv, (uint64_t)N};
  uint64_t str[4] = {2, (uint64_t)sub * C * 2, (uint64_t)sub * W * C * 2, (uint64_t)H * W * C * 2};
  uint32_t box[4] = {64, (uint32_t)tw, (uint32_t)th, 1};
  return encode_tmap(m, 2, 4, const_cast<uint8_t*>(b), dims, str, box, 128);
}

// Tile-N choice from a two-term cost model measured in round 1 (profiles/r1_launches.md): the
// kernel is bound either by the L2->SM operand traffic (every tile streams its A and B stages,
// ~8 TB/s aggregate) or by the tcgen05 issue rate (128xNx16 per N/2 cycles per SM, wave-quantised).
static int pick_block_n(int cout, int m_tiles, int k_iters, int row_bytes) {
  const int sms = num_sms();
  int best = 64;
  double best_t = 1e30;
  for (int bn = 256; bn >= 64; bn >>= 1) {
    if (cout % bn != 0) continue;
    const double tiles = (double)m_tiles * (cout / bn);
    const double traffic = tiles * k_iters * (128.0 * row_bytes + (double)bn * row_bytes);
    const double t_l2 = traffic / 8.0e12;
    const double waves = (double)((int64_t)(tiles + sms - 1) / sms);
    const double mma_cycles = k_iters * (row_bytes / 32) * (bn < 128 ? 48.0 : bn / 2.0);
    const double t_mma = waves * mma_cycles / 1.8e9;
    const double t = (t_l2 > t_mma ? t_l2 : t_mma) + 0.15 * (t_l2 < t_mma ? t_l2 : t_mma);
    if (t < best_t) {
      best_t = t;
      best = bn;
    }
  }
  return best;
}

static int finish_launch(ConvLaunch* L, const ghnd_conv_desc_t* d, int gemm_cin, int gemm_cout,
                         const void* weights, int total_taps, int kblock = 64) {
  ConvKernelParams& p = L->p;
  p.cin = gemm_cin;
  p.cout = gemm_cout;
  p.kblock = kblock;
  const int row_bytes = kblock * 2;
  p.a_bytes = kBlockM * row_bytes;
  const int m_tiles = p.n_img * p.tiles_h * p.tiles_w;
  p.block_n = pick_block_n(gemm_cout, m_tiles, p.n_taps * (gemm_cin / kblock), row_bytes);
  p.n_tiles_n = gemm_cout / p.block_n;
  p.total_tiles = m_tiles * p.n_tiles_n;
  p.stage_bytes = p.a_bytes + p.block_n * row_bytes;
  int stages = (200 * 1024) / p.stage_bytes;
  if (stages > 8) stages = 8;
  p.n_stages = stages;
  p.a_box_bytes = p.tw * p.th * row_bytes;
  p.idesc = make_idesc(d->src_fmt, d->w_fmt, 0, 0, kBlockM, p.block_n);
  // weights: 2D [gemm_cout rows][total_taps * gemm_cin]
  uint64_t dims[2] = {(uint64_t)total_taps * gemm_cin, (uint64_t)gemm_cout};
  uint64_t str[2] = {2, (uint64_t)total_taps * gemm_cin * 2};
  uint32_t box[2] = {(uint32_t)kblock, (uint32_t)p.block_n};
  int rc = encode_tmap(&p.tmap_b, 2, 2, const_cast<void*>(weights), dims, str, box, row_bytes);
  if (rc != GHND_OK) return rc;
  p.dst = d->dst;
  p.bias = d->bias;
  p.residual = d->residual;
  p.mask = d->mask;
  p.dst_fmt = d->dst_fmt;
  p.res_fmt = d->res_fmt;
  p.mask_fmt = d->mask_fmt;
  p.relu = d->relu;
  p.accumulate = d->accumulate;
  L->grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  L->smem = (size_t)p.n_stages * p.stage_bytes + 1024 /*align*/ + 256 /*barriers*/;
  return GHND_OK;
}

static bool fmt_ok(int f) { return f == GHND_F16 || f == GHND_BF16; }

}  // namespace ghnd

extern "C" {

int ghnd_conv_plan_create(const ghnd_conv_desc_t* d, ghnd_conv_plan_t** out) {
  using namespace ghnd;
  GHND_CHECK_ARG(d && out, "conv_plan_create: null argument");
  *out = nullptr;
  GHND_CHECK_ARG(d->kind == GHND_CONV_FWD || d->kind == GHND_CONV_DGRAD, "conv: bad kind %d",
                 d->kind);
  GHND_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0, "conv: bad geometry N=%d H=%d W=%d", d->N, d->H,
                 d->W);
  GHND_CHECK_ARG(d->C > 0 && d->C % 64 == 0 && d->K > 0 && d->K % 64 == 0,
                 "conv: tensor-core path needs C and K multiples of 64 (C=%d K=%d); narrow convs "
                 "use ghnd_conv_narrow_*",
                 d->C, d->K);
  GHND_CHECK_ARG(d->R >= 1 && d->S >= 1 && d->R * d->S <= kMaxTaps, "conv: kernel %dx%d unsupported",
                 d->R, d->S);
  GHND_CHECK_ARG(d->stride == 1 || d->stride == 2, "conv: stride %d unsupported", d->stride);
  GHND_CHECK_ARG(d->pad >= 0 && d->pad < 4, "conv: pad %d unsupported", d->pad);
  GHND_CHECK_ARG(d->src && d->weights && d->dst, "conv: null tensor pointer");
  GHND_CHECK_ARG(fmt_ok(d->src_fmt) && fmt_ok(d->w_fmt) && fmt_ok(d->dst_fmt), "conv: bad format");
  GHND_CHECK_ARG(d->src_fmt == d->w_fmt,
                 "conv: src and weights must share one 16-bit format (tcgen05 kind::f16 rejects "
                 "mixed f16 x bf16 operands on sm_100a)");
  GHND_CHECK_ARG(!d->residual || fmt_ok(d->res_fmt), "conv: bad residual format");
  GHND_CHECK_ARG(!d->mask || fmt_ok(d->mask_fmt), "conv: bad mask format");
  GHND_CHECK_ARG(d->stats == nullptr, "conv: fused statistics not available in this build");
  GHND_CHECK_ARG(((uintptr_t)d->src % 16) == 0 && ((uintptr_t)d->weights % 16) == 0 &&
                     ((uintptr_t)d->dst % 16) == 0 && ((uintptr_t)d->residual % 16) == 0 &&
                     ((uintptr_t)d->mask % 16) == 0 && ((uintptr_t)d->bias % 16) == 0,
                 "conv: tensors must be 16-byte aligned");
  const int Ho = (d->H + 2 * d->pad - d->R) / d->stride + 1;
  const int Wo = (d->W + 2 * d->pad - d->S) / d->stride + 1;
  GHND_CHECK_ARG(Ho > 0 && Wo > 0, "conv: empty output");

  ghnd_conv_plan* plan = new ghnd_conv_plan();
  int rc = GHND_OK;
  const int taps_total = d->R * d->S;

  if (d->kind == GHND_CONV_FWD) {
    ConvLaunch L;
    memset(&L, 0, sizeof(L));
    ConvKernelParams& p = L.p;
    const bool flat = (d->R == 1 && d->S == 1 && d->stride == 1 && d->pad == 0);
    if (flat) {
      const int64_t npix = (int64_t)d->N * d->H * d->W;
      p.n_img = 1;
      p.th = 1;
      p.tw = 128;
      p.ho = 1;
      p.wo = (int)npix;
      p.tiles_h = 1;
      p.tiles_w = (int)((npix + 127) / 128);
      p.oh_full = 1;
      p.ow_full = (int)npix;
      p.osh = p.osw = 1;
      rc = make_a_map(&p.tmap_a[0], d->src, 1, 1, (int)npix, d->C, 1, 0, 0, 1, 128);
      for (int i = 1; i < 4 && rc == GHND_OK; ++i) p.tmap_a[i] = p.tmap_a[0];
      p.n_taps = 1;
      p.taps[0] = ConvTap{0, 0, 0, 0};
    } else {
      p.n_img = d->N;
      p.ho = Ho;
      p.wo = Wo;
      choose_tile(Ho, Wo, &p.th, &p.tw);
      p.tiles_h = (Ho + p.th - 1) / p.th;
      p.tiles_w = (Wo + p.tw - 1) / p.tw;
      p.oh_full = Ho;
      p.ow_full = Wo;
      p.osh = p.osw = 1;
      if (d->stride == 1) {
        rc = make_a_map(&p.tmap_a[0], d->src, d->N, d->H, d->W, d->C, 1, 0, 0, p.th, p.tw);
        for (int i = 1; i < 4 && rc == GHND_OK; ++i) p.tmap_a[i] = p.tmap_a[0];
      } else {
        for (int ph = 0; ph < 2 && rc == GHND_OK; ++ph)
          for (int pw = 0; pw < 2 && rc == GHND_OK; ++pw)
            rc = make_a_map(&p.tmap_a[ph * 2 + pw], d->src, d->N, d->H, d->W, d->C, 2, ph, pw, p.th,
                            p.tw);
      }
      int nt = 0;
      for (int r = 0; r < d->R; ++r)
        for (int s = 0; s < d->S; ++s) {
          const int oh = r - d->pad, ow = s - d->pad;
          ConvTap t;
          if (d->stride == 1) {
            t = ConvTap{(int16_t)oh, (int16_t)ow, 0, (int16_t)(r * d->S + s)};
          } else {
            const int ph = oh & 1, pw = ow & 1;
            t = ConvTap{(int16_t)((oh - ph) / 2), (int16_t)((ow - pw) / 2), (int16_t)(ph * 2 + pw),
                        (int16_t)(r * d->S + s)};
          }
          p.taps[nt++] = t;
        }
      p.n_taps = nt;
    }
    if (rc == GHND_OK) rc = finish_launch(&L, d, d->C, d->K, d->weights, taps_total);
    if (rc == GHND_OK) plan->launches.push_back(L);
  } else {
    // DGRAD: src = dy [N,Ho,Wo,K], dst = dx [N,H,W,C], weights [C][R][S][K]
    const int sub = d->stride;
    for (int ph = 0; ph < sub && rc == GHND_OK; ++ph) {
      for (int pw = 0; pw < sub && rc == GHND_OK; ++pw) {
        ConvLaunch L;
        memset(&L, 0, sizeof(L));
        ConvKernelParams& p = L.p;
        // dx rows h = sub*i + ph, i in [0, hi)
        const int hi = (d->H - ph + sub - 1) / sub;
        const int wi = (d->W - pw + sub - 1) / sub;
        if (hi <= 0 || wi <= 0) continue;
        int nt = 0;
        for (int r = 0; r < d->R; ++r) {
          if (((ph + d->pad - r) % sub) != 0) continue;
          for (int s = 0; s < d->S; ++s) {
            if (((pw + d->pad - s) % sub) != 0) continue;
            // dy row = (h + pad - r)/sub = i + (ph + pad - r)/sub
            p.taps[nt++] = ConvTap{(int16_t)((ph + d->pad - r) / sub),
                                   (int16_t)((pw + d->pad - s) / sub), 0, (int16_t)(r * d->S + s)};
          }
        }
        if (nt == 0) {
          if (d->accumulate) continue;  // nothing to add for this parity class
          delete plan;
          set_error("conv dgrad: parity class (%d,%d) receives no taps; use accumulate=1", ph, pw);
          return GHND_ERR_UNSUPPORTED;
        }
        p.n_taps = nt;
        const bool flat = (sub == 1 && d->R == 1 && d->S == 1 && d->pad == 0);
        if (flat) {
          const int64_t npix = (int64_t)d->N * d->H * d->W;
          p.n_img = 1;
          p.th = 1;
          p.tw = 128;
          p.ho = 1;
          p.wo = (int)npix;
          p.tiles_h = 1;
          p.tiles_w = (int)((npix + 127) / 128);
          p.oh_full = 1;
          p.ow_full = (int)npix;
          p.osh = p.osw = 1;
          rc = make_a_map(&p.tmap_a[0], d->src, 1, 1, (int)npix, d->K, 1, 0, 0, 1, 128);
        } else {
          p.n_img = d->N;
          p.ho = hi;
          p.wo = wi;
          choose_tile(hi, wi, &p.th, &p.tw);
          p.tiles_h = (hi + p.th - 1) / p.th;
          p.tiles_w = (wi + p.tw - 1) / p.tw;
          p.oh_full = d->H;
          p.ow_full = d->W;
          p.osh = p.osw = sub;
          p.ooh = ph;
          p.oow = pw;
          rc = make_a_map(&p.tmap_a[0], d->src, d->N, Ho, Wo, d->K, 1, 0, 0, p.th, p.tw);
        }
        for (int i = 1; i < 4 && rc == GHND_OK; ++i) p.tmap_a[i] = p.tmap_a[0];
        if (rc == GHND_OK) rc = finish_launch(&L, d, d->K, d->C, d->weights, taps_total);
        if (rc == GHND_OK) plan->launches.push_back(L);
      }
    }
  }
  if (rc != GHND_OK) {
    delete plan;
    return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) {
      delete plan;
      return cuda_fail(e, "cudaFuncSetAttribute(conv_tc_kernel)");
    }
    attr_set = true;
  }
  *out = plan;
  return GHND_OK;
}

int ghnd_conv_plan_run(const ghnd_conv_plan_t* plan, void* stream) {
  using namespace ghnd;
  GHND_CHECK_ARG(plan != nullptr, "conv_plan_run: null plan");
  for (const ConvLaunch& L : plan->launches) {
    conv_tc_kernel<<<L.grid, kConvThreads, L.smem, (cudaStream_t)stream>>>(L.p);
    GHND_LAUNCH_CHECK("conv_tc_kernel");
  }
  return GHND_OK;
}

int ghnd_conv_plan_launches(const ghnd_conv_plan_t* plan) {
  return plan ? (int)plan->launches.size() : 0;
}

void ghnd_conv_plan_destroy(ghnd_conv_plan_t* plan) { delete plan; }
}

// ------------------------------------------------------------------------------------------------
// Stem conv1 (7x7 s2 p3, 3->64) on the same kernel: im2col row = 7 px x 4 ch (+4 zero) = 32
// elements (64 B, SWIZZLE_64B), one "tap" per filter row, output columns split in 4 classes
// (wo = 4j+q) so consecutive GEMM rows are 64 B apart in the packed image (plain tiled TMA).
// ------------------------------------------------------------------------------------------------
struct ghnd_stem_plan {
  std::vector<ghnd::ConvLaunch> launches;
};

extern "C" {

int ghnd_stem_conv_plan_create(const void* x_packed, int x_fmt, const void* w_packed, int w_fmt,
                               const float* bias, void* y, int y_fmt, int N, int Hp, int Wp,
                               ghnd_stem_plan_t** out) {
  using namespace ghnd;
  GHND_CHECK_ARG(out && x_packed && w_packed && y, "stem_conv_plan_create: null argument");
  *out = nullptr;
  GHND_CHECK_ARG(N > 0 && Hp > 0 && Wp > 0 && Hp % 2 == 0 && Wp % 8 == 0,
                 "stem conv: padded size must be even x multiple of 8 (Hp=%d Wp=%d)", Hp, Wp);
  GHND_CHECK_ARG(fmt_ok(x_fmt) && fmt_ok(w_fmt) && fmt_ok(y_fmt), "stem conv: bad format");
  GHND_CHECK_ARG(x_fmt == w_fmt, "stem conv: image and weights must share one 16-bit format");
  const int Ho = Hp / 2, Wo = Wp / 2;
  const int rows = Hp + 6, RP = (Wp + 8) * 4;  // packed image rows / row pitch in elements
  ghnd_stem_plan* plan = new ghnd_stem_plan();
  int rc = GHND_OK;
  ghnd_conv_desc_t d;
  memset(&d, 0, sizeof(d));
  d.src_fmt = x_fmt;
  d.w_fmt = w_fmt;
  d.dst = y;
  d.dst_fmt = y_fmt;
  d.bias = bias;
  d.relu = 1;
  for (int q = 0; q < 4 && rc == GHND_OK; ++q) {
    const int J = (Wo - q + 3) / 4;
    if (J <= 0) continue;
    ConvLaunch L;
    memset(&L, 0, sizeof(L));
    ConvKernelParams& p = L.p;
    p.n_img = N;
    p.ho = Ho;
    p.wo = J;
    choose_tile(Ho, J, &p.th, &p.tw);
    p.tiles_h = (Ho + p.th - 1) / p.th;
    p.tiles_w = (J + p.tw - 1) / p.tw;
    p.oh_full = Ho;
    p.ow_full = Wo;
    p.osh = 1;
    p.osw = 4;
    p.ooh = 0;
    p.oow = q;
    for (int ph = 0; ph < 2 && rc == GHND_OK; ++ph) {
      const uint8_t* base = static_cast<const uint8_t*>(x_packed) + ((size_t)ph * RP + q * 8) * 2;
      uint64_t dims[4] = {32, (uint64_t)J, (uint64_t)((rows - ph + 1) / 2), (uint64_t)N};
      uint64_t str[4] = {2, 64, (uint64_t)2 * RP * 2, (uint64_t)rows * RP * 2};
      uint32_t box[4] = {32, (uint32_t)p.tw, (uint32_t)p.th, 1};
      rc = encode_tmap(&p.tmap_a[ph], 2, 4, const_cast<uint8_t*>(base), dims, str, box, 64);
    }
    p.tmap_a[2] = p.tmap_a[0];
    p.tmap_a[3] = p.tmap_a[1];
    p.n_taps = 7;
    for (int r = 0; r < 7; ++r) {
      const int ph = r & 1;
      p.taps[r] = ConvTap{(int16_t)((r - ph) / 2), 0, (int16_t)ph, (int16_t)r};
    }
    if (rc == GHND_OK) rc = finish_launch(&L, &d, 32, 64, w_packed, 7, 32);
    if (rc == GHND_OK) plan->launches.push_back(L);
  }
  if (rc != GHND_OK) {
    delete plan;
    return rc;
  }
  cudaError_t e =
      cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) {
    delete plan;
    return cuda_fail(e, "cudaFuncSetAttribute(conv_tc_kernel)");
  }
  *out = plan;
  return GHND_OK;
}

int ghnd_stem_conv_plan_run(const ghnd_stem_plan_t* plan, void* stream) {
  using namespace ghnd;
  GHND_CHECK_ARG(plan != nullptr, "stem_conv_plan_run: null plan");
  for (const ConvLaunch& L : plan->launches) {
    conv_tc_kernel<<<L.grid, kConvThreads, L.smem, (cudaStream_t)stream>>>(L.p);
    GHND_LAUNCH_CHECK("conv_tc_kernel(stem)");
  }
  return GHND_OK;
}

void ghnd_stem_plan_destroy(ghnd_stem_plan_t* plan) { delete plan; }
}
