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
// * epilogue = 64-channel column chunks: tcgen05.ld -> (+bias, +residual, ReLU, mask, +dst) ->
//   16-bit rows written into a 128B-swizzled staging tile -> ONE TMA tensor store per chunk, so
//   global writes are whole 128-byte pixel rows and the ragged tile edge is clipped by the TMA unit.
//   Residual / mask / accumulate operands arrive the same way (TMA loads issued by a dedicated
//   warp, double buffered), never as per-thread strided global loads.
// * persistent CTAs (one per SM), warp-specialised: warp0 TMA producer, warp1 MMA issuer (+TMEM
//   alloc), warp2 epilogue-operand loader, warps 3..6 epilogue.
//
// The same kernel runs the data-gradient: dgrad of a stride-1 conv is a conv of dy with mirrored
// tap offsets over the transposed weights; dgrad of a stride-2 conv is four launches, one per
// output parity class, each using the taps that hit that class and storing through a lattice view.
#include <vector>

#include "common.cuh"

namespace ghnd {

static constexpr int kConvThreads = 224;
static constexpr int kEpiThreads = 128;
static constexpr int kEpiWarp0 = 3;
static constexpr int kBlockM = 128;
static constexpr int kMaxTaps = 9;
static constexpr int kChunkBytes = kBlockM * 128;  // one [128 rows][64 ch] 16-bit staging tile
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
  CUtensorMap tmap_out;   // dst viewed on the GEMM-M lattice, box {64, TW, TH, 1}
  CUtensorMap tmap_in0;   // residual (added before ReLU/mask) or dst itself (accumulate, added last)
  CUtensorMap tmap_in1;   // ReLU-mask source
  ConvTap taps[kMaxTaps];
  int n_taps;
  int cin;          // GEMM reduction channels per tap (multiple of kblock)
  int cout;         // GEMM output channels (multiple of block_n)
  int block_n;      // 64 / 128 / 256
  int n_stages;
  int kblock;       // 64 or 32
  int a_bytes;      // 128 rows * kblock * 2
  int stage_bytes;  // a_bytes + block_n*kblock*2
  int a_box_bytes;  // TW*TH*kblock*2
  int io_box_bytes; // TW*TH*128
  // tile grid over the GEMM-M space
  int n_img;        // images (1 when flattened)
  int tiles_h, tiles_w, th, tw;
  int n_tiles_n;
  int total_tiles;
  uint32_t idesc;
  // epilogue
  const float* bias;
  int has_in0, in0_post, has_in1;
  int out_fmt, in0_fmt, in1_fmt;
  int relu;
};

__device__ __forceinline__ uint32_t swz_off(int row, int chunk16) {
  // byte offset of 16-byte chunk `chunk16` of `row` inside a 128B-swizzled [rows][128 B] tile
  return (uint32_t)(row * 128 + ((chunk16 ^ (row & 7)) << 4));
}

__global__ void __launch_bounds__(kConvThreads, 1)
    conv_tc_kernel(const __grid_constant__ ConvKernelParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  const int n_in = p.has_in0 + p.has_in1;
  uint8_t* epi_out = smem + (size_t)p.n_stages * p.stage_bytes;  // [2][kChunkBytes]
  uint8_t* epi_in = epi_out + 2 * kChunkBytes;                   // [2][n_in][kChunkBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_in + (size_t)2 * n_in * kChunkBytes);
  uint64_t* full_bar = bars;                    // [n_stages]
  uint64_t* empty_bar = bars + p.n_stages;      // [n_stages]
  uint64_t* tfull_bar = bars + 2 * p.n_stages;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]
  uint64_t* ifull_bar = tempty_bar + 2;         // [2]
  uint64_t* iempty_bar = ifull_bar + 2;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(iempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) prefetch_tmap(&p.tmap_a[i]);
    prefetch_tmap(&p.tmap_b);
    prefetch_tmap(&p.tmap_out);
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kEpiThreads);
      mbar_init(&ifull_bar[i], 1);
      mbar_init(&iempty_bar[i], kEpiThreads);
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
  const int n_chunks = p.block_n >> 6;

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
  } else if (warp == 2) {
    // ===================== epilogue-operand loader =====================
    if (lane == 0 && n_in > 0) {
      uint32_t cnt = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles_n;
        const int m_tile = tile / p.n_tiles_n;
        const int img = m_tile / tiles_per_img;
        const int rem = m_tile - img * tiles_per_img;
        const int h0 = (rem / p.tiles_w) * p.th;
        const int w0 = (rem % p.tiles_w) * p.tw;
        for (int c = 0; c < n_chunks; ++c, ++cnt) {
          const int ib = cnt & 1;
          mbar_wait(&iempty_bar[ib], ((cnt >> 1) & 1) ^ 1);
          uint8_t* dst = epi_in + (size_t)ib * n_in * kChunkBytes;
          mbar_arrive_expect_tx(&ifull_bar[ib], (uint32_t)(n_in * p.io_box_bytes));
          const int ch = n_tile * p.block_n + c * 64;
          if (p.has_in0) tma_load_4d(dst, &p.tmap_in0, &ifull_bar[ib], ch, w0, h0, img);
          if (p.has_in1)
            tma_load_4d(dst + p.has_in0 * kChunkBytes, &p.tmap_in1, &ifull_bar[ib], ch, w0, h0, img);
        }
      }
    }
  } else {
    // ===================== epilogue (4 warps, TMEM lane quarter = warp % 4) =====================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int etid = threadIdx.x - kEpiWarp0 * 32;
    int it = 0;
    uint32_t cnt = 0;  // chunk counter (staging / operand double buffers)
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      const int n_tile = tile % p.n_tiles_n;
      const int m_tile = tile / p.n_tiles_n;
      const int img = m_tile / tiles_per_img;
      const int rem = m_tile - img * tiles_per_img;
      const int h0 = (rem / p.tiles_w) * p.th;
      const int w0 = (rem % p.tiles_w) * p.tw;

      mbar_wait(&tfull_bar[buf], use & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * p.block_n);
      for (int c = 0; c < n_chunks; ++c, ++cnt) {
        uint32_t r[64];
        tmem_ld32(t_addr + (uint32_t)(c * 64), r);
        tmem_ld32(t_addr + (uint32_t)(c * 64 + 32), r + 32);
        tmem_ld_wait();
        if (c == n_chunks - 1) {  // accumulator fully read: hand the TMEM buffer back to the MMA warp
          tc_fence_before();
          mbar_arrive(&tempty_bar[buf]);
        }
        const int ch = n_tile * p.block_n + c * 64;
        float v[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias != nullptr) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + ch);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 b = __ldg(b4 + j);
            v[4 * j] += b.x;
            v[4 * j + 1] += b.y;
            v[4 * j + 2] += b.z;
            v[4 * j + 3] += b.w;
          }
        }
        const int ib = cnt & 1;
        const uint8_t* in_base = epi_in + (size_t)ib * n_in * kChunkBytes;
        if (n_in > 0) mbar_wait(&ifull_bar[ib], (cnt >> 1) & 1);
        if (p.has_in0 && !p.in0_post) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 q = *reinterpret_cast<const uint4*>(in_base + swz_off(row, j));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack2(w[e], p.in0_fmt);
              v[8 * j + 2 * e] += f.x;
              v[8 * j + 2 * e + 1] += f.y;
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (p.has_in1) {
          const uint8_t* m_base = in_base + p.has_in0 * kChunkBytes;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 q = *reinterpret_cast<const uint4*>(m_base + swz_off(row, j));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack2(w[e], p.in1_fmt);
              if (!(f.x > 0.f)) v[8 * j + 2 * e] = 0.f;
              if (!(f.y > 0.f)) v[8 * j + 2 * e + 1] = 0.f;
            }
          }
        }
        if (p.has_in0 && p.in0_post) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 q = *reinterpret_cast<const uint4*>(in_base + swz_off(row, j));
            const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack2(w[e], p.in0_fmt);
              v[8 * j + 2 * e] += f.x;
              v[8 * j + 2 * e + 1] += f.y;
            }
          }
        }
        if (n_in > 0) mbar_arrive(&iempty_bar[ib]);  // operand buffer consumed
        // ---- stage the 64-channel rows and store them with one TMA tensor store ----
        const int ob = cnt & 1;
        uint8_t* o_base = epi_out + (size_t)ob * kChunkBytes;
        if (etid == 0) bulk_wait_read<1>();  // the store that used this staging buffer has drained
        named_bar_sync(1, kEpiThreads);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 o;
          o.x = pack2(v[8 * j], v[8 * j + 1], p.out_fmt);
          o.y = pack2(v[8 * j + 2], v[8 * j + 3], p.out_fmt);
          o.z = pack2(v[8 * j + 4], v[8 * j + 5], p.out_fmt);
          o.w = pack2(v[8 * j + 6], v[8 * j + 7], p.out_fmt);
          *reinterpret_cast<uint4*>(o_base + swz_off(row, j)) = o;
        }
        fence_proxy_async();
        named_bar_sync(1, kEpiThreads);
        if (etid == 0) {
          tma_store_4d(&p.tmap_out, o_base, ch, w0, h0, img);
          bulk_commit();
        }
      }
    }
    if (etid == 0) bulk_wait_all();
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

// View of an NHWC 16-bit tensor [N][H][W][C] restricted to the lattice h = sub_h*i + ph,
// w = sub_w*j + pw.  inner = elements of the innermost box dimension (64 -> SWIZZLE_128B).
static int make_lattice_map(CUtensorMap* m, const void* base, int N, int H, int W, int C, int sub_h,
                            int sub_w, int ph, int pw, int th, int tw, int inner = 64) {
  int hv = (H - ph + sub_h - 1) / sub_h;
  int wv = (W - pw + sub_w - 1) / sub_w;
  if (hv <= 0 || wv <= 0) {  // empty lattice: encode the full view, taps using it are never generated
    return make_lattice_map(m, base, N, H, W, C, 1, 1, 0, 0, th, tw, inner);
  }
  const uint8_t* b = static_cast<const uint8_t*>(base) + ((size_t)ph * W + pw) * C * 2;
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)wv, (uint64_t)hv, (uint64_t)N};
  uint64_t str[4] = {2, (uint64_t)sub_w * C * 2, (uint64_t)sub_h * W * C * 2, (uint64_t)H * W * C * 2};
  uint32_t box[4] = {(uint32_t)inner, (uint32_t)tw, (uint32_t)th, 1};
  return encode_tmap(m, 2, 4, const_cast<uint8_t*>(b), dims, str, box, inner * 2);
}

// Tile-N choice from a two-term cost model measured in round 1 (profiles/): the kernel is bound
// either by the L2->SM operand traffic (every tile streams its A and B stages, ~8 TB/s aggregate)
// or by the tcgen05 issue rate (128xNx16 per N/2 cycles per SM, wave-quantised).
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

// Geometry of the destination lattice of one launch (also used for the residual / mask views).
struct IoGeom {
  int N, H, W;  // full dst tensor [N][H][W][cout]
  int sub_h, sub_w, ph, pw;
};

static int finish_launch(ConvLaunch* L, const ghnd_conv_desc_t* d, int gemm_cin, int gemm_cout,
                         const void* weights, int total_taps, const IoGeom& g, int kblock = 64) {
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
  p.a_box_bytes = p.tw * p.th * row_bytes;
  p.io_box_bytes = p.tw * p.th * 128;
  p.idesc = make_idesc(d->src_fmt, d->w_fmt, 0, 0, kBlockM, p.block_n);
  // weights: 2D [gemm_cout rows][total_taps * gemm_cin]
  uint64_t dims[2] = {(uint64_t)total_taps * gemm_cin, (uint64_t)gemm_cout};
  uint64_t str[2] = {2, (uint64_t)total_taps * gemm_cin * 2};
  uint32_t box[2] = {(uint32_t)kblock, (uint32_t)p.block_n};
  int rc = encode_tmap(&p.tmap_b, 2, 2, const_cast<void*>(weights), dims, str, box, row_bytes);
  if (rc != GHND_OK) return rc;
  // epilogue operands, all on the destination lattice
  rc = make_lattice_map(&p.tmap_out, d->dst, g.N, g.H, g.W, gemm_cout, g.sub_h, g.sub_w, g.ph, g.pw,
                        p.th, p.tw);
  if (rc != GHND_OK) return rc;
  p.tmap_in0 = p.tmap_out;
  p.tmap_in1 = p.tmap_out;
  p.has_in0 = p.has_in1 = p.in0_post = 0;
  if (d->residual != nullptr && d->accumulate) {
    set_error("conv: residual and accumulate cannot be combined in one launch");
    return GHND_ERR_UNSUPPORTED;
  }
  if (d->residual != nullptr) {
    p.has_in0 = 1;
    p.in0_fmt = d->res_fmt;
    rc = make_lattice_map(&p.tmap_in0, d->residual, g.N, g.H, g.W, gemm_cout, g.sub_h, g.sub_w, g.ph,
                          g.pw, p.th, p.tw);
  } else if (d->accumulate) {
    p.has_in0 = 1;
    p.in0_post = 1;
    p.in0_fmt = d->dst_fmt;
  }
  if (rc != GHND_OK) return rc;
  if (d->mask != nullptr) {
    p.has_in1 = 1;
    p.in1_fmt = d->mask_fmt;
    rc = make_lattice_map(&p.tmap_in1, d->mask, g.N, g.H, g.W, gemm_cout, g.sub_h, g.sub_w, g.ph, g.pw,
                          p.th, p.tw);
    if (rc != GHND_OK) return rc;
  }
  p.bias = d->bias;
  p.out_fmt = d->dst_fmt;
  p.relu = d->relu;
  const int n_in = p.has_in0 + p.has_in1;
  const int epi_bytes = (2 + 2 * n_in) * kChunkBytes;
  int stages = (222 * 1024 - epi_bytes) / p.stage_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) {
    set_error("conv: shared memory budget leaves %d pipeline stages", stages);
    return GHND_ERR_UNSUPPORTED;
  }
  p.n_stages = stages;
  L->grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  L->smem = (size_t)p.n_stages * p.stage_bytes + epi_bytes + 1024 /*align*/ + 256 /*barriers*/;
  return GHND_OK;
}

static bool fmt_ok(int f) { return f == GHND_F16 || f == GHND_BF16; }

static int set_conv_attr() {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_tc_kernel)");
    attr_set = true;
  }
  return GHND_OK;
}

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
    IoGeom g;
    const bool flat = (d->R == 1 && d->S == 1 && d->stride == 1 && d->pad == 0);
    if (flat) {
      const int64_t npix = (int64_t)d->N * d->H * d->W;
      p.n_img = 1;
      p.th = 1;
      p.tw = 128;
      p.tiles_h = 1;
      p.tiles_w = (int)((npix + 127) / 128);
      g = IoGeom{1, 1, (int)npix, 1, 1, 0, 0};
      rc = make_lattice_map(&p.tmap_a[0], d->src, 1, 1, (int)npix, d->C, 1, 1, 0, 0, 1, 128);
      for (int i = 1; i < 4 && rc == GHND_OK; ++i) p.tmap_a[i] = p.tmap_a[0];
      p.n_taps = 1;
      p.taps[0] = ConvTap{0, 0, 0, 0};
    } else {
      p.n_img = d->N;
      choose_tile(Ho, Wo, &p.th, &p.tw);
      p.tiles_h = (Ho + p.th - 1) / p.th;
      p.tiles_w = (Wo + p.tw - 1) / p.tw;
      g = IoGeom{d->N, Ho, Wo, 1, 1, 0, 0};
      if (d->stride == 1) {
        rc = make_lattice_map(&p.tmap_a[0], d->src, d->N, d->H, d->W, d->C, 1, 1, 0, 0, p.th, p.tw);
        for (int i = 1; i < 4 && rc == GHND_OK; ++i) p.tmap_a[i] = p.tmap_a[0];
      } else {
        for (int ph = 0; ph < 2 && rc == GHND_OK; ++ph)
          for (int pw = 0; pw < 2 && rc == GHND_OK; ++pw)
            rc = make_lattice_map(&p.tmap_a[ph * 2 + pw], d->src, d->N, d->H, d->W, d->C, 2, 2, ph, pw,
                                  p.th, p.tw);
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
    if (rc == GHND_OK) rc = finish_launch(&L, d, d->C, d->K, d->weights, taps_total, g);
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
        IoGeom g;
        const bool flat = (sub == 1 && d->R == 1 && d->S == 1 && d->pad == 0);
        if (flat) {
          const int64_t npix = (int64_t)d->N * d->H * d->W;
          p.n_img = 1;
          p.th = 1;
          p.tw = 128;
          p.tiles_h = 1;
          p.tiles_w = (int)((npix + 127) / 128);
          g = IoGeom{1, 1, (int)npix, 1, 1, 0, 0};
          rc = make_lattice_map(&p.tmap_a[0], d->src, 1, 1, (int)npix, d->K, 1, 1, 0, 0, 1, 128);
        } else {
          p.n_img = d->N;
          choose_tile(hi, wi, &p.th, &p.tw);
          p.tiles_h = (hi + p.th - 1) / p.th;
          p.tiles_w = (wi + p.tw - 1) / p.tw;
          g = IoGeom{d->N, d->H, d->W, sub, sub, ph, pw};
          rc = make_lattice_map(&p.tmap_a[0], d->src, d->N, Ho, Wo, d->K, 1, 1, 0, 0, p.th, p.tw);
        }
        for (int i = 1; i < 4 && rc == GHND_OK; ++i) p.tmap_a[i] = p.tmap_a[0];
        if (rc == GHND_OK) rc = finish_launch(&L, d, d->K, d->C, d->weights, taps_total, g);
        if (rc == GHND_OK) plan->launches.push_back(L);
      }
    }
  }
  if (rc == GHND_OK) rc = set_conv_attr();
  if (rc != GHND_OK) {
    delete plan;
    return rc;
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
    choose_tile(Ho, J, &p.th, &p.tw);
    p.tiles_h = (Ho + p.th - 1) / p.th;
    p.tiles_w = (J + p.tw - 1) / p.tw;
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
    const IoGeom g{N, Ho, Wo, 1, 4, 0, q};
    if (rc == GHND_OK) rc = finish_launch(&L, &d, 32, 64, w_packed, 7, g, 32);
    if (rc == GHND_OK) plan->launches.push_back(L);
  }
  if (rc == GHND_OK) rc = set_conv_attr();
  if (rc != GHND_OK) {
    delete plan;
    return rc;
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
