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
#include <stdlib.h>
#include <vector>

#include "common.cuh"

namespace ghnd {

// warp roles: 0 TMA producer, 1 MMA issuer (+TMEM alloc), 2 epilogue-operand loader,
// 3..6 epilogue group 0, 7..10 epilogue group 1 (TMEM lane quarter = warp % 4 in both groups).
static constexpr int kConvThreads = 352;
static constexpr int kEpiGroupThreads = 128;
static constexpr int kEpiThreads = 2 * kEpiGroupThreads;
static constexpr int kEpiWarp0 = 3;
static constexpr int kBlockM = 128;
static constexpr int kMaxTaps = 9;
static constexpr int kMaxStages = 8;
static constexpr int kMaxRing = 8;
static constexpr int kMaxASlots = 4;
static constexpr int kChunkBytes = kBlockM * 128;  // one [128 rows][64 ch] 16-bit staging tile
// GEMM-K per "unit": 64 x 16-bit = 128 B rows (SWIZZLE_128B) for the wide convs, or
// 32 x 16-bit = 64 B rows (SWIZZLE_64B) for the stem whose im2col row is 7 px x 4 ch (+4 zero).
// A pipeline stage holds `units_per_stage` units so that one full-barrier wait + one commit are
// amortised over >= ~256 tensor-pipe cycles even for N=64 tiles.

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
  int b_bytes;      // block_n rows * kblock * 2
  int units_per_stage;
  int stage_bytes;  // units_per_stage * (a_bytes + b_bytes)
  int a_box_bytes;  // TW*TH*kblock*2
  int io_box_bytes; // TW*TH*128
  int ring;         // depth of the epilogue-operand ring (chunks)
  // tile grid over the GEMM-M space
  int n_img;        // images (1 when flattened)
  int tiles_h, tiles_w, th, tw;
  int lim_h, lim_w; // extent of the output lattice (row validity for the fused statistics)
  int n_tiles_n;
  int total_tiles;
  FastDiv fd_tiles_n, fd_tiles_per_img, fd_tiles_w, fd_tw, fd_ring;
  uint32_t idesc;
  // epilogue
  const float* bias;
  int has_in0, in0_post, has_in1;
  int out_fmt, in0_fmt, in1_fmt;
  int relu;
  double* stats;    // optional [2*cout]: sum / sum of squares of the stored (rounded) output
  int stats_mode;   // 1: sum of the output, sum of output * in1 operand (BN backward reductions)
  int in1_mask;     // 0: the in1 operand is only read by the statistics
  // stem "halo" mode: the im2col operand of ALL filter rows of a tile is one TMA box of
  // 2*th+5 image rows (read at row shifts by the MMA descriptors) and the weights stay resident
  int halo;
  int w_res_bytes;  // n_taps * b_bytes of resident weights behind the pipeline stages
  int epi_half;     // f16 outputs without mask / statistics: bias, residual and ReLU in packed half2
                    // after ONE fp32 -> f16 conversion of the accumulator (half the epilogue math)
  int epi_debug;    // what-if switches for profiling ONLY (results become wrong): 1 = no TMA store,
                    // 2 = no proxy fence, 4 = no tcgen05.ld, 8 = no bias/ReLU/operand math, 16 = no weight loads
  int epi_prefetch; // issue the next chunk's tcgen05.ld as soon as the current chunk is staged
  int epi_bufs;     // output staging tiles per epilogue group (2: the TMA store of chunk i drains
                    // while chunk i+1 is staged)
  // conv "halo" mode (MODE 2; stride-1 convs with more than one tap): per 64-channel chunk ONE dense TMA
  // box of (th+R-1) x (tw+S-1) pixels; every tap is an MMA descriptor at a row shift of that tile
  // (tw == 8: one 8-row swizzle group per output row, group stride = halo row pitch).  Cuts the
  // L2 -> shared-memory traffic of the A operand by n_taps x (th*tw)/(halo pixels): 3.3x for 2x2, 6.2x for
  // 3x3.  b_resident: the whole weight slab of this CTA's n-tile stays in shared memory.
  int a_halo_bytes;   // halo tile bytes rounded up to 1024
  int a_slots;        // depth of the halo ring
  int halo_w;         // halo row pitch in pixels (tw + S - 1)
  int org_dh, org_dw; // offset of the halo box origin from the tile origin
  int tap_off[kMaxTaps];  // byte offset of tap t's first row inside the halo tile
  int b_resident;
  int pdl_late_wait;  // independent of the previous launch (see launch_conv): wait for it at the END
  int mc;             // launched as clusters of 2 CTAs that work on the SAME n-tile of two neighbouring m-tiles
                      // (generic mode).  1: each CTA fetches half of every weight tile and TMA-multicasts it to
                      // both (measured neutral: the SM still ingests the whole tile); 2: CTA pair, ONE M=256
                      // cta_group::2 MMA over both CTAs, each holds only its half of the weight tile (b_bytes)
  int m_tiles;        // real m-tiles (mc: total_tiles counts m-tile PAIRS x n-tiles)
};

__device__ __forceinline__ int fd_ring_r(const ConvKernelParams& p, uint32_t cnt) {
  return (int)cnt - fd_div(p.fd_ring, (int)cnt) * (int)p.fd_ring.d;
}

__device__ __forceinline__ uint32_t swz_off(int row, int chunk16) {
  // byte offset of 16-byte chunk `chunk16` of `row` inside a 128B-swizzled [rows][128 B] tile
  return (uint32_t)(row * 128 + ((chunk16 ^ (row & 7)) << 4));
}

// One GEMM-K unit = KSTEPS tcgen05.mma (K=16 each) whose descriptors differ only in the 14-bit
// start-address field of the low word (+2 = 32 bytes along the swizzled row per step).  Written as
// one asm block on 32-bit descriptor halves so that ptxas keeps everything in uniform registers and
// issues the UTCHMMAs back to back.
template <int KSTEPS>
__device__ __forceinline__ void umma_unit(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo,
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
      "add.u32 a, %1, 2;\n\t"
      "add.u32 b, %2, 2;\n\t"
      "mov.b64 da, {a, %3};\n\t"
      "mov.b64 db, {b, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, q;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
  if (KSTEPS == 4) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        ".reg .b64 da, db;\n\t"
        ".reg .b32 a, b;\n\t"
        "setp.eq.b32 q, 0, 0;\n\t"
        "add.u32 a, %1, 4;\n\t"
        "add.u32 b, %2, 4;\n\t"
        "mov.b64 da, {a, %3};\n\t"
        "mov.b64 db, {b, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, q;\n\t"
        "add.u32 a, %1, 6;\n\t"
        "add.u32 b, %2, 6;\n\t"
        "mov.b64 da, {a, %3};\n\t"
        "mov.b64 db, {b, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, q;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc)
        : "memory");
  }
}

// Two K=16 steps with separate high words for A and B (different stride-byte-offsets).
__device__ __forceinline__ void umma_unit2_ab(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi,
                                              uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
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

// 2-D TMA load multicast to the CTAs of the cluster named by `mask`: data and the barrier's complete_tx land at
// the same shared-memory offsets in every destination CTA.
__device__ __forceinline__ void tma_load_2d_mc(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, "
      "{%3, %4}], [%2], %5;" ::"r"(smem_u32(smem)),
      "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// tcgen05.commit arriving on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// ---- CTA pair (cta_group::2): one M=256 MMA spans both CTAs of a cluster; each CTA supplies its own 128 rows of
// A and HALF of the weight tile, so the weight bytes every SM ingests are halved.  Only the leader (rank 0)
// issues MMAs; both CTAs' TMA loads complete on the LEADER's full barrier; commits are multicast. ----
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tma_load_2d_2sm(void* smem, const CUtensorMap* m, uint32_t bar_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(smem_u32(smem)),
      "l"(m), "r"(bar_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* smem, const CUtensorMap* m, uint32_t bar_addr, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5, %6}], [%2];" ::"r"(smem_u32(smem)),
      "l"(m), "r"(bar_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// umma_unit with cta_group::2 (see umma_unit)
template <int KSTEPS>
__device__ __forceinline__ void umma_unit_2sm(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 da, db;\n\t"
      ".reg .b32 a, b;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "setp.eq.b32 q, 0, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t"
      "add.u32 a, %1, 2;\n\t"
      "add.u32 b, %2, 2;\n\t"
      "mov.b64 da, {a, %3};\n\t"
      "mov.b64 db, {b, %3};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, q;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
  if (KSTEPS == 4) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        ".reg .b64 da, db;\n\t"
        ".reg .b32 a, b;\n\t"
        "setp.eq.b32 q, 0, 0;\n\t"
        "add.u32 a, %1, 4;\n\t"
        "add.u32 b, %2, 4;\n\t"
        "mov.b64 da, {a, %3};\n\t"
        "mov.b64 db, {b, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, q;\n\t"
        "add.u32 a, %1, 6;\n\t"
        "add.u32 b, %2, 6;\n\t"
        "mov.b64 da, {a, %3};\n\t"
        "mov.b64 db, {b, %3};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, q;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc)
        : "memory");
  }
}
// One barrier arrival per WARP: every lane has done its part (and its fences) before the __syncwarp, lane 0 arrives.
// (Per-thread arrivals were 128-256 shared-memory atomics on one word per chunk / tile.)
__device__ __forceinline__ void mbar_arrive_warp(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Persistent tile loop.  Plain: tile = blockIdx.x, += gridDim.x, (m_tile, n_tile) = divmod(tile, n_tiles_n).
// mc: the CTA pair {2j, 2j+1} walks pair-tiles j, j + gridDim.x/2, ...; CTA rank r takes m-tile 2*m_pair + r.  A
// pair whose second m-tile does not exist repeats the last one (dup: everything but the store / statistics),
// so that both CTAs run the same sequence of pipeline stages.
struct TileWalk {
  int first, step, rank;
};
template <typename P>
__device__ __forceinline__ TileWalk tile_walk(const P& p) {
  TileWalk w;
  w.first = p.mc ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  w.step = p.mc ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  w.rank = (int)(blockIdx.x & 1u);
  return w;
}
template <typename P>
__device__ __forceinline__ bool tile_decode(const P& p, const TileWalk& w, int tile, int& m_tile, int& n_tile) {
  fd_divmod(p.fd_tiles_n, tile, m_tile, n_tile);
  if (!p.mc) return false;
  m_tile = 2 * m_tile + w.rank;
  if (m_tile < p.m_tiles) return false;
  m_tile = p.m_tiles - 1;
  return true;
}

// Four K=16 steps (one 64-channel unit) with separate high words for A and B.
__device__ __forceinline__ void umma_unit4_ab(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi,
                                              uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
  umma_unit2_ab(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, accumulate);
  umma_unit2_ab(d_tmem, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1u);
}

// v[0..63] += the 64 channels of `row` of a swizzled 16-bit operand tile
template <int FMT>
__device__ __forceinline__ void epi_add_rows(float* v, const uint8_t* base, int row) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 q = *reinterpret_cast<const uint4*>(base + swz_off(row, j));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack2_t<FMT>(w[e]);
      v[8 * j + 2 * e] += f.x;
      v[8 * j + 2 * e + 1] += f.y;
    }
  }
}
// v[i] = mask[i] > 0 ? v[i] : 0
template <int FMT>
__device__ __forceinline__ void epi_mask_rows(float* v, const uint8_t* base, int row) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint4 q = *reinterpret_cast<const uint4*>(base + swz_off(row, j));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = unpack2_t<FMT>(w[e]);
      if (!(f.x > 0.f)) v[8 * j + 2 * e] = 0.f;
      if (!(f.y > 0.f)) v[8 * j + 2 * e + 1] = 0.f;
    }
  }
}
template <int FMT>
__device__ __forceinline__ void epi_stage_rows(const float* v, uint8_t* base, int row) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint4 o;
    o.x = pack2_t<FMT>(v[8 * j], v[8 * j + 1]);
    o.y = pack2_t<FMT>(v[8 * j + 2], v[8 * j + 3]);
    o.z = pack2_t<FMT>(v[8 * j + 4], v[8 * j + 5]);
    o.w = pack2_t<FMT>(v[8 * j + 6], v[8 * j + 7]);
    *reinterpret_cast<uint4*>(base + swz_off(row, j)) = o;
  }
}
// Per-channel sum / sum of squares of the staged (rounded) tile: this thread owns the channel
// pair (2*lane, 2*lane+1) over the 32 rows of its warp's quarter; `valid` = ballot of row validity.
template <int FMT>
__device__ __forceinline__ void epi_stats_rows(const uint8_t* base, int quarter, int lane,
                                               uint32_t valid, float* sstat, int ch, int cout) {
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  // this lane's 4 bytes of row r: chunk (lane >> 2) ^ (r & 7); r & 7 == i & 7 (quarter * 32 is a multiple of 8)
  const uint8_t* col = base + quarter * 32 * 128 + ((lane & 3) << 2);
  const int chunk = lane >> 2;
  if (valid == 0xffffffffu) {  // interior tile (uniform): no per-row predicate, two independent chains
    float t0 = 0.f, t1 = 0.f, u0 = 0.f, u1 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float2 f = unpack2_t<FMT>(*reinterpret_cast<const uint32_t*>(col + i * 128 + ((chunk ^ (i & 7)) << 4)));
      const float2 g =
          unpack2_t<FMT>(*reinterpret_cast<const uint32_t*>(col + (i + 1) * 128 + ((chunk ^ ((i + 1) & 7)) << 4)));
      s0 += f.x;
      s1 += f.y;
      q0 = fmaf(f.x, f.x, q0);
      q1 = fmaf(f.y, f.y, q1);
      t0 += g.x;
      t1 += g.y;
      u0 = fmaf(g.x, g.x, u0);
      u1 = fmaf(g.y, g.y, u1);
    }
    s0 += t0;
    s1 += t1;
    q0 += u0;
    q1 += u1;
  } else {
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(col + i * 128 + ((chunk ^ (i & 7)) << 4));
      if ((valid >> i) & 1u) {
        const float2 f = unpack2_t<FMT>(w);
        s0 += f.x;
        s1 += f.y;
        q0 = fmaf(f.x, f.x, q0);
        q1 = fmaf(f.y, f.y, q1);
      }
    }
  }
  // `sstat` is PRIVATE to this warp and each lane owns its channel pair: plain read-modify-write.
  // (Shared-memory float atomicAdd is a compare-and-swap spin loop in SASS -- ATOMS.CAST.SPIN --
  // and eight warps hitting the same 64 addresses after every chunk made these launches the
  // slowest tensor-core kernels of the step, 0.31-0.39 of their HBM bound.)
  float2* a = reinterpret_cast<float2*>(sstat + ch + 2 * lane);
  float2* b = reinterpret_cast<float2*>(sstat + cout + ch + 2 * lane);
  float2 av = *a, bv = *b;
  av.x += s0;
  av.y += s1;
  bv.x += q0;
  bv.y += q1;
  *a = av;
  *b = bv;
}

// BN-backward variant: sum of the staged output g and of g * a, a = the in1 operand tile (same
// swizzled layout, its own 16-bit format).
template <int FMT, int AFMT>
__device__ __forceinline__ void epi_stats2_rows(const uint8_t* base, const uint8_t* abase, int quarter,
                                                int lane, uint32_t valid, float* sstat, int ch,
                                                int cout) {
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
  for (int i = 0; i < 32; ++i) {
    const int r = quarter * 32 + i;
    const uint32_t off = r * 128 + ((((lane >> 2) ^ (r & 7))) << 4) + ((lane & 3) << 2);
    const uint32_t w = *reinterpret_cast<const uint32_t*>(base + off);
    const uint32_t a = *reinterpret_cast<const uint32_t*>(abase + off);
    if ((valid >> i) & 1u) {
      const float2 f = unpack2_t<FMT>(w);
      const float2 g = unpack2_t<AFMT>(a);
      s0 += f.x;
      s1 += f.y;
      q0 = fmaf(f.x, g.x, q0);
      q1 = fmaf(f.y, g.y, q1);
    }
  }
  // `sstat` is PRIVATE to this warp and each lane owns its channel pair: plain read-modify-write.
  // (Shared-memory float atomicAdd is a compare-and-swap spin loop in SASS -- ATOMS.CAST.SPIN --
  // and eight warps hitting the same 64 addresses after every chunk made these launches the
  // slowest tensor-core kernels of the step, 0.31-0.39 of their HBM bound.)
  float2* a = reinterpret_cast<float2*>(sstat + ch + 2 * lane);
  float2* b = reinterpret_cast<float2*>(sstat + cout + ch + 2 * lane);
  float2 av = *a, bv = *b;
  av.x += s0;
  av.y += s1;
  bv.x += q0;
  bv.y += q1;
  *a = av;
  *b = bv;
}

// Packed 16-bit-pair helpers on raw 32-bit words (FMT = GHND_F16 or GHND_BF16)
template <int FMT>
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) {
  if (FMT == GHND_F16) {
    const __half2 r = __hadd2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  const __nv_bfloat162 r =
      __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
template <int FMT>
__device__ __forceinline__ uint32_t h2_relu(uint32_t a) {
  if (FMT == GHND_F16) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), __float2half2_rn(0.f));
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), __float2bfloat162_rn(0.f));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// Packed epilogue of one accumulator row (64 channels), all arithmetic in the OUTPUT format after one
// fp32 -> 16-bit conversion of the accumulator:
//   h = cvt(acc); h += bias; h += in0 (residual, !in0_post); ReLU; h &= (mask > 0); h += in0 (in0_post)
// bias16: 64 values (128 B) of this chunk in the output format; in0: swizzled operand tile in the
// output format; mask: swizzled f16 tile (the forward activation).
template <int FMT>
__device__ __forceinline__ void epi_half_rows(const uint32_t* r, const uint8_t* bias16, const uint8_t* in0,
                                              int in0_post, const uint8_t* mask, int relu, uint32_t* out,
                                              int row) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e)
      h[e] = pack2_t<FMT>(__uint_as_float(r[8 * j + 2 * e]), __uint_as_float(r[8 * j + 2 * e + 1]));
    if (bias16 != nullptr) {
      const uint4 b = *reinterpret_cast<const uint4*>(bias16 + j * 16);
      h[0] = h2_add<FMT>(h[0], b.x);
      h[1] = h2_add<FMT>(h[1], b.y);
      h[2] = h2_add<FMT>(h[2], b.z);
      h[3] = h2_add<FMT>(h[3], b.w);
    }
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    if (in0 != nullptr) q = *reinterpret_cast<const uint4*>(in0 + swz_off(row, j));
    if (in0 != nullptr && !in0_post) {
      h[0] = h2_add<FMT>(h[0], q.x);
      h[1] = h2_add<FMT>(h[1], q.y);
      h[2] = h2_add<FMT>(h[2], q.z);
      h[3] = h2_add<FMT>(h[3], q.w);
    }
    if (relu) {
#pragma unroll
      for (int e = 0; e < 4; ++e) h[e] = h2_relu<FMT>(h[e]);
    }
    if (FMT == GHND_BF16 && mask != nullptr) {  // data gradients: ReLU mask from the f16 activation
      const uint4 m = *reinterpret_cast<const uint4*>(mask + swz_off(row, j));
      const __half2 z = __float2half2_rn(0.f);
      h[0] &= __hgt2_mask(*reinterpret_cast<const __half2*>(&m.x), z);
      h[1] &= __hgt2_mask(*reinterpret_cast<const __half2*>(&m.y), z);
      h[2] &= __hgt2_mask(*reinterpret_cast<const __half2*>(&m.z), z);
      h[3] &= __hgt2_mask(*reinterpret_cast<const __half2*>(&m.w), z);
    }
    if (in0 != nullptr && in0_post) {
      h[0] = h2_add<FMT>(h[0], q.x);
      h[1] = h2_add<FMT>(h[1], q.y);
      h[2] = h2_add<FMT>(h[2], q.z);
      h[3] = h2_add<FMT>(h[3], q.w);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) out[4 * j + e] = h[e];
  }
}
__device__ __forceinline__ void epi_stage_packed(const uint32_t* h, uint8_t* stage, int row) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<uint4*>(stage + swz_off(row, j)) =
        make_uint4(h[4 * j], h[4 * j + 1], h[4 * j + 2], h[4 * j + 3]);
}

// EPI (0 fp32, 1 packed f16, 2 packed bf16 + mask) / HALO select the epilogue and the operand pipeline (stem halo vs
// im2col units) at COMPILE time: each instantiation carries only its own paths.  The kernel is
// sensitive to code size (a build with all variants in one 7.2k-instruction kernel lost 10-20 % on
// the epilogue-bound layers against a 5.1k-instruction build, same algorithm).
// STATS: the launch accumulates per-channel statistics (kept out of the other instantiations)
template <int EPI, int MODE, bool STATS, bool PAIR = false>
__global__ void __launch_bounds__(kConvThreads, 1)
    conv_tc_kernel(const __grid_constant__ ConvKernelParams p) {
  constexpr bool HALO = MODE == 1;   // stem halo mode
  constexpr bool CHALO = MODE == 2;  // conv halo mode
  // 1024-byte alignment for the 128B swizzle atoms comes from the declaration; every pointer below is
  // derived from this array by plain arithmetic so that ptxas keeps the accesses in the shared state
  // space (LDS / STS).  The first version rounded the base up through uintptr_t: the casts hid the
  // address space and EVERY epilogue access became a generic LD.E / ST.E (ncu source page, round 2).
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (smem_u32(smem_raw) & 1023u) __trap();
  const int n_in = p.has_in0 + p.has_in1;
  // MODE 2: [a_slots halo tiles] precede the (B-only) pipeline stages / the resident weights
  uint8_t* stages0 = smem + (CHALO ? (size_t)p.a_slots * p.a_halo_bytes : 0);
  uint8_t* wres = stages0 + (size_t)p.n_stages * p.stage_bytes;            // halo modes: resident weights
  uint8_t* epi_out = wres + ((HALO || CHALO) ? p.w_res_bytes : 0);       // [2 groups][kChunkBytes]
  uint8_t* epi_in = epi_out + (size_t)2 * p.epi_bufs * kChunkBytes;        // [ring][n_in][kChunkBytes]
  float* sbias = reinterpret_cast<float*>(epi_in + (size_t)p.ring * n_in * kChunkBytes);  // [cout]
  float* sstat = sbias + (p.bias != nullptr ? p.cout : 0);                 // [8 epilogue warps][2*cout] when stats
  uint64_t* bars = reinterpret_cast<uint64_t*>(sstat + (STATS ? kEpiThreads / 32 * 2 * p.cout : 0));
  uint64_t* full_bar = bars;                        // [kMaxStages]
  uint64_t* empty_bar = full_bar + kMaxStages;      // [kMaxStages]
  uint64_t* tfull_bar = empty_bar + kMaxStages;     // [2]
  uint64_t* tempty_bar = tfull_bar + 2;             // [2]
  uint64_t* ifull_bar = tempty_bar + 2;             // [kMaxRing]
  uint64_t* iempty_bar = ifull_bar + kMaxRing;      // [kMaxRing]
  uint64_t* wfull_bar = iempty_bar + kMaxRing;      // [1] resident weights have landed (halo modes)
  uint64_t* afull_bar = wfull_bar + 1;              // [kMaxASlots] conv halo tiles
  uint64_t* aempty_bar = afull_bar + kMaxASlots;    // [kMaxASlots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aempty_bar + kMaxASlots);

  // warp index through a shuffle so that the compiler treats the role branches as warp-uniform
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) prefetch_tmap(&p.tmap_a[i]);
    prefetch_tmap(&p.tmap_b);
    prefetch_tmap(&p.tmap_out);
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], p.mc == 1 ? 2 : 1);  // mc 1: a stage is free when BOTH CTAs' MMAs have read it
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PAIR ? 2 * (kEpiThreads / 32) : kEpiThreads / 32);  // pair: both CTAs' epilogues
    }
    for (int i = 0; i < kMaxRing; ++i) {
      mbar_init(&ifull_bar[i], 1);
      mbar_init(&iempty_bar[i], kEpiGroupThreads / 32);
    }
    mbar_init(wfull_bar, 1);
    for (int i = 0; i < kMaxASlots; ++i) {
      mbar_init(&afull_bar[i], 1);
      mbar_init(&aempty_bar[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_2sm(tmem_slot, 512);
    else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (p.mc) cluster_sync_all();  // the peer's barriers are initialised before anything is multicast to them
  const TileWalk walk = tile_walk(p);
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  // Programmatic dependent launch: everything above overlapped the previous kernel's tail; from
  // here on global memory written by it is read, so wait for its completion + flush.
  // pdl_late_wait: this launch shares no data with the one before it in the stream (parity classes 2..4 of a
  // stride-2 dgrad: same inputs, disjoint outputs), and that one waited for everything earlier before it let
  // this one start -- so the launches of the plan run side by side on the SMs.  The wait moves to the end of
  // the kernel, which keeps completion transitive: whoever waits for this launch has waited for all of them.
  if (!p.pdl_late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  const int k_chunks = p.cin / p.kblock;
  const int n_units = p.n_taps * k_chunks;
  const int upst = p.units_per_stage;
  const int n_chunks = p.block_n >> 6;

  if (HALO && warp == 0) {
    // ===================== TMA producer, stem halo mode ==========================================
    // weights once; then ONE box per tile: 2*th+5 image rows x tw column slots x 64 B
    if (elect_one()) {
      mbar_arrive_expect_tx(wfull_bar, (uint32_t)p.w_res_bytes);
      for (int r = 0; r < p.n_taps; ++r)
        tma_load_2d(wres + (size_t)r * p.b_bytes, &p.tmap_b, wfull_bar, p.taps[r].wk * p.cin, 0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int img, rem, h0, w0;
      fd_divmod(p.fd_tiles_per_img, tile, img, rem);
      fd_divmod(p.fd_tiles_w, rem, h0, w0);
      h0 *= p.th;
      w0 *= p.tw;
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)p.a_box_bytes);
        tma_load_4d(smem + (size_t)stage * p.stage_bytes, &p.tmap_a[0], &full_bar[stage], 0, w0, 2 * h0,
                    img);
      }
      __syncwarp();
      if (++stage == p.n_stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (HALO && warp == 1) {
    // ===================== MMA issuer, stem halo mode ============================================
    // filter row r reads the halo tile at a shift of r image rows (r * tw * 64 B); output rows are
    // two image rows apart (SBO of A = 2 * tw * 64 B).  tw == 8: one 8-row swizzle atom per image row,
    // so every shift is a multiple of the 512-byte SWIZZLE_64B period.
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t w_base = smem_u32(wres);
    const uint32_t row_pitch = (uint32_t)p.tw * 64u;
    const uint32_t a_hi = ((2u * row_pitch) >> 4) | (1u << 14) | ((uint32_t)UMMA_SW64 << 29);
    const uint32_t b_hi = (512u >> 4) | (1u << 14) | ((uint32_t)UMMA_SW64 << 29);
    mbar_wait(wfull_bar, 0);
    tc_fence_after();
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&tempty_bar[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.block_n);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_base + (uint32_t)(stage * p.stage_bytes);
        for (int r = 0; r < p.n_taps; ++r) {
          const uint32_t a_lo = (((sa + (uint32_t)r * row_pitch) >> 4) & 0x3fffu) | (1u << 16);
          const uint32_t b_lo = (((w_base + (uint32_t)(r * p.b_bytes)) >> 4) & 0x3fffu) | (1u << 16);
          umma_unit2_ab(d_tmem, a_lo, a_hi, b_lo, b_hi, p.idesc, (uint32_t)(r != 0));
        }
        umma_commit(&empty_bar[stage]);  // the halo tile may be overwritten when these retire
        umma_commit(&tfull_bar[buf]);    // accumulator complete -> epilogue
      }
      __syncwarp();
      if (++stage == p.n_stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (CHALO && warp == 0) {
    // ===================== TMA producer, conv halo mode ==========================================
    // per tile and 64-channel chunk: one halo box; then (unless the weights are resident) this chunk's
    // B tiles, tap by tap, through the stage ring.  Loads are issued in the order the MMA warp consumes.
    const int n_tile0 = (int)blockIdx.x - fd_div(p.fd_tiles_n, (int)blockIdx.x) * (int)p.fd_tiles_n.d;
    if (p.b_resident) {
      if (elect_one()) {
        mbar_arrive_expect_tx(wfull_bar, (uint32_t)p.w_res_bytes);
        for (int t = 0; t < p.n_taps; ++t)
          for (int kc = 0; kc < k_chunks; ++kc)
            tma_load_2d(wres + (size_t)(t * k_chunks + kc) * p.b_bytes, &p.tmap_b, wfull_bar,
                        p.taps[t].wk * p.cin + kc * p.kblock, n_tile0 * p.block_n);
      }
      __syncwarp();
    }
    int stage = 0, aslot = 0;
    uint32_t phase = 0, aphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int n_tile, m_tile, img, rem, h0, w0;
      fd_divmod(p.fd_tiles_n, tile, m_tile, n_tile);
      fd_divmod(p.fd_tiles_per_img, m_tile, img, rem);
      fd_divmod(p.fd_tiles_w, rem, h0, w0);
      h0 *= p.th;
      w0 *= p.tw;
      for (int kc = 0; kc < k_chunks; ++kc) {
        mbar_wait(&aempty_bar[aslot], aphase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&afull_bar[aslot], (uint32_t)p.a_box_bytes);
          tma_load_4d(smem + (size_t)aslot * p.a_halo_bytes, &p.tmap_a[0], &afull_bar[aslot], kc * p.kblock,
                      w0 + p.org_dw, h0 + p.org_dh, img);
        }
        __syncwarp();
        if (++aslot == p.a_slots) {
          aslot = 0;
          aphase ^= 1;
        }
        if (!p.b_resident) {
          for (int t0 = 0; t0 < p.n_taps; t0 += upst) {
            const int nu = min(upst, p.n_taps - t0);
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (elect_one()) {
              uint8_t* sb = stages0 + (size_t)stage * p.stage_bytes;
              mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(nu * p.b_bytes));
              for (int j = 0; j < nu; ++j)
                tma_load_2d(sb + j * p.b_bytes, &p.tmap_b, &full_bar[stage],
                            p.taps[t0 + j].wk * p.cin + kc * p.kblock, n_tile * p.block_n);
            }
            __syncwarp();
            if (++stage == p.n_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (CHALO && warp == 1) {
    // ===================== MMA issuer, conv halo mode ============================================
    // tap (r, s) reads the chunk's halo tile at a shift of (r * halo_w + s) rows of 128 B; output row i of
    // the th x 8 tile is the 8-row group i, groups are halo_w rows apart (SBO of A).
    int stage = 0, aslot = 0, it = 0;
    uint32_t phase = 0, aphase = 0;
    const uint32_t a_base = smem_u32(smem), s_base = smem_u32(stages0), w_base = smem_u32(wres);
    const uint32_t a_hi = (((uint32_t)p.halo_w * 128u) >> 4) | (1u << 14) | ((uint32_t)UMMA_SW128 << 29);
    const uint32_t b_hi = (1024u >> 4) | (1u << 14) | ((uint32_t)UMMA_SW128 << 29);
    if (p.b_resident) {
      mbar_wait(wfull_bar, 0);
      tc_fence_after();
    }
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(&tempty_bar[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.block_n);
      for (int kc = 0; kc < k_chunks; ++kc) {
        mbar_wait(&afull_bar[aslot], aphase);
        tc_fence_after();
        const uint32_t a_tile = a_base + (uint32_t)(aslot * p.a_halo_bytes);
        if (p.b_resident) {
          if (elect_one()) {
            for (int t = 0; t < p.n_taps; ++t) {
              const uint32_t a_lo = (((a_tile + (uint32_t)p.tap_off[t]) >> 4) & 0x3fffu) | (1u << 16);
              const uint32_t b_lo =
                  (((w_base + (uint32_t)((t * k_chunks + kc) * p.b_bytes)) >> 4) & 0x3fffu) | (1u << 16);
              umma_unit4_ab(d_tmem, a_lo, a_hi, b_lo, b_hi, p.idesc, (uint32_t)((kc | t) != 0));
            }
            umma_commit(&aempty_bar[aslot]);
          }
          __syncwarp();
        } else {
          for (int t0 = 0; t0 < p.n_taps; t0 += upst) {
            const int nu = min(upst, p.n_taps - t0);
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t sb = s_base + (uint32_t)(stage * p.stage_bytes);
              for (int j = 0; j < nu; ++j) {
                const int t = t0 + j;
                const uint32_t a_lo = (((a_tile + (uint32_t)p.tap_off[t]) >> 4) & 0x3fffu) | (1u << 16);
                const uint32_t b_lo = (((sb + (uint32_t)(j * p.b_bytes)) >> 4) & 0x3fffu) | (1u << 16);
                umma_unit4_ab(d_tmem, a_lo, a_hi, b_lo, b_hi, p.idesc, (uint32_t)((kc | t) != 0));
              }
              umma_commit(&empty_bar[stage]);
              if (t0 + nu >= p.n_taps) umma_commit(&aempty_bar[aslot]);
            }
            __syncwarp();
            if (++stage == p.n_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        if (++aslot == p.a_slots) {
          aslot = 0;
          aphase ^= 1;
        }
      }
      if (elect_one()) umma_commit(&tfull_bar[buf]);  // accumulator complete -> epilogue
      __syncwarp();
    }
  } else if (!HALO && !CHALO && warp == 0) {
    // ===================== TMA producer (whole warp converged, one elected lane issues) ==========
    int stage = 0;
    uint32_t phase = 0;
    const int b_half = p.b_bytes >> 1;
    for (int tile = walk.first; tile < p.total_tiles; tile += walk.step) {
      int n_tile, m_tile, img, rem, h0, w0;
      tile_decode(p, walk, tile, m_tile, n_tile);
      fd_divmod(p.fd_tiles_per_img, m_tile, img, rem);
      fd_divmod(p.fd_tiles_w, rem, h0, w0);
      h0 *= p.th;
      w0 *= p.tw;
      int t = 0, kc = 0;
      for (int u = 0; u < n_units; u += upst) {
        const int nu = min(upst, n_units - u);
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + (size_t)stage * p.stage_bytes;
          uint8_t* sb = sa + upst * p.a_bytes;
          int tt = t, kk = kc;
          if (PAIR) {
            // pair: both CTAs' bytes complete on the LEADER's barrier; only the leader posts the expectation
            const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[stage]), 0u);
            if (walk.rank == 0)
              mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(2 * nu * (p.a_box_bytes + p.b_bytes)));
            for (int j = 0; j < nu; ++j) {
              const ConvTap tap = p.taps[tt];
              tma_load_4d_2sm(sa + j * p.a_bytes, &p.tmap_a[tap.map], lead_bar, kk * p.kblock, w0 + tap.dw,
                              h0 + tap.dh, img);
              tma_load_2d_2sm(sb + j * p.b_bytes, &p.tmap_b, lead_bar, tap.wk * p.cin + kk * p.kblock,
                              n_tile * p.block_n + walk.rank * (p.block_n >> 1));
              if (++kk == k_chunks) {
                kk = 0;
                ++tt;
              }
            }
          } else {
          const bool no_b = (p.epi_debug & 16) != 0;  // what-if: no weight loads (results are wrong)
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)(nu * (p.a_box_bytes + (no_b ? 0 : p.b_bytes))));
          for (int j = 0; j < nu; ++j) {
            const ConvTap tap = p.taps[tt];
            tma_load_4d(sa + j * p.a_bytes, &p.tmap_a[tap.map], &full_bar[stage], kk * p.kblock,
                        w0 + tap.dw, h0 + tap.dh, img);
            if (no_b) {
            } else if (p.mc)  // this CTA's half of the weight tile, to both CTAs (tmap_b's box is block_n / 2 rows)
              tma_load_2d_mc(sb + j * p.b_bytes + walk.rank * b_half, &p.tmap_b, &full_bar[stage],
                             tap.wk * p.cin + kk * p.kblock, n_tile * p.block_n + walk.rank * (p.block_n >> 1),
                             (uint16_t)3);
            else
              tma_load_2d(sb + j * p.b_bytes, &p.tmap_b, &full_bar[stage],
                          tap.wk * p.cin + kk * p.kblock, n_tile * p.block_n);
            if (++kk == k_chunks) {
              kk = 0;
              ++tt;
            }
          }
          }
        }
        __syncwarp();
        // every lane advances (t, kc) by the units of this stage (the elected lane may change)
        kc += nu;
        while (kc >= k_chunks) {
          kc -= k_chunks;
          ++t;
        }
        if (++stage == p.n_stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (!HALO && !CHALO && warp == 1) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) ============
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t sbo = 8u * (uint32_t)(p.kblock * 2);  // 8-row swizzle atom
    const uint32_t desc_hi = (sbo >> 4) | (1u << 14) | ((p.kblock == 64 ? (uint32_t)UMMA_SW128
                                                                         : (uint32_t)UMMA_SW64)
                                                        << 29);
    const uint32_t a_step = (uint32_t)p.a_bytes >> 4, b_step = (uint32_t)p.b_bytes >> 4;
    constexpr bool pair = PAIR;
    // pair: the leader issues the M=256 MMAs for both CTAs; the peer's MMA warp only owns its TMEM allocation
    for (int tile = (pair && walk.rank != 0) ? p.total_tiles : walk.first; tile < p.total_tiles;
         tile += walk.step, ++it) {
      const int buf = it & 1;
      mbar_wait(&tempty_bar[buf], (((uint32_t)it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.block_n);
      for (int u = 0; u < n_units; u += upst) {
        const int nu = min(upst, n_units - u);
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_base + (uint32_t)(stage * p.stage_bytes);
          // descriptor low word: start address >> 4 | LBO(16 B) << 16
          uint32_t a_lo = ((sa >> 4) & 0x3fffu) | (1u << 16);
          uint32_t b_lo = (((sa + (uint32_t)(upst * p.a_bytes)) >> 4) & 0x3fffu) | (1u << 16);
          if (pair) {  // kblock == 64 (host)
            for (int j = 0; j < nu; ++j) {
              umma_unit_2sm<4>(d_tmem, a_lo, b_lo, desc_hi, p.idesc, (uint32_t)((u + j) != 0));
              a_lo += a_step;
              b_lo += b_step;
            }
          } else if (p.kblock == 64) {
            for (int j = 0; j < nu; ++j) {
              umma_unit<4>(d_tmem, a_lo, b_lo, desc_hi, p.idesc, (uint32_t)((u + j) != 0));
              a_lo += a_step;
              b_lo += b_step;
            }
          } else {
            for (int j = 0; j < nu; ++j) {
              umma_unit<2>(d_tmem, a_lo, b_lo, desc_hi, p.idesc, (uint32_t)((u + j) != 0));
              a_lo += a_step;
              b_lo += b_step;
            }
          }
          // frees the smem slot when these MMAs retire (mc: in both CTAs -- the peer's multicast writes it too)
          if (pair) umma_commit_2sm(&empty_bar[stage], (uint16_t)3);
          else if (p.mc) umma_commit_mc(&empty_bar[stage], (uint16_t)3);
          else umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == p.n_stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (elect_one()) {  // accumulator complete -> epilogue (pair: of both CTAs)
        if (pair) umma_commit_2sm(&tfull_bar[buf], (uint16_t)3);
        else umma_commit(&tfull_bar[buf]);
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ===================== epilogue-operand loader =====================
    if (n_in > 0) {
      uint32_t cnt = 0;
      int slot = 0;
      uint32_t rphase = 0;
      for (int tile = walk.first; tile < p.total_tiles; tile += walk.step) {
        int n_tile, m_tile, img, rem, h0, w0;
        tile_decode(p, walk, tile, m_tile, n_tile);
        fd_divmod(p.fd_tiles_per_img, m_tile, img, rem);
        fd_divmod(p.fd_tiles_w, rem, h0, w0);
        h0 *= p.th;
        w0 *= p.tw;
        for (int c = 0; c < n_chunks; ++c, ++cnt) {
          mbar_wait(&iempty_bar[slot], rphase ^ 1);
          if (elect_one()) {
            uint8_t* dst = epi_in + (size_t)slot * n_in * kChunkBytes;
            mbar_arrive_expect_tx(&ifull_bar[slot], (uint32_t)(n_in * p.io_box_bytes));
            const int ch = n_tile * p.block_n + c * 64;
            if (p.has_in0) tma_load_4d(dst, &p.tmap_in0, &ifull_bar[slot], ch, w0, h0, img);
            if (p.has_in1)
              tma_load_4d(dst + p.has_in0 * kChunkBytes, &p.tmap_in1, &ifull_bar[slot], ch, w0, h0,
                          img);
          }
          __syncwarp();
          if (++slot == p.ring) {
            slot = 0;
            rphase ^= 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue: two groups of 4 warps, alternating 64-channel chunks ========
    const int group = (warp - kEpiWarp0) >> 2;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int etid = threadIdx.x - (kEpiWarp0 + 4 * group) * 32;
    const int bar_id = 1 + group;
    uint8_t* const o_base0 = epi_out + (size_t)group * p.epi_bufs * kChunkBytes;
    uint32_t n_staged = 0;  // chunks staged by this group so far (selects the staging tile)
    // bias / statistics scratch are only needed here: fill them off the producer/MMA critical path
    if (p.bias != nullptr) {
      if (EPI != 0) {
        uint16_t* hb = reinterpret_cast<uint16_t*>(sbias);
        for (int i = threadIdx.x - kEpiWarp0 * 32; i < p.cout; i += kEpiThreads)
          hb[i] = float_to_h16(__ldg(p.bias + i), p.out_fmt);
      } else {
        for (int i = threadIdx.x - kEpiWarp0 * 32; i < p.cout; i += kEpiThreads) sbias[i] = __ldg(p.bias + i);
      }
    }
    if (STATS)
      for (int i = threadIdx.x - kEpiWarp0 * 32; i < kEpiThreads / 32 * 2 * p.cout; i += kEpiThreads) sstat[i] = 0.f;
    float* const wstat = sstat + (size_t)(warp - kEpiWarp0) * 2 * p.cout;  // this warp's private partial sums
    named_bar_sync(3, kEpiThreads);
    int it = 0;
    uint32_t cnt0 = 0;  // global chunk counter at the start of the tile
    // Accumulator registers of the chunk being processed.  Once a chunk is staged in shared memory
    // they are dead, so the NEXT chunk's tcgen05.ld is issued right there (pre = true) and its
    // latency overlaps the proxy fence, the barrier, the TMA store and the statistics of the
    // current chunk instead of heading the next iteration (ncu: 36 % of the epilogue's stall
    // samples were the wait for tcgen05.ld).
    uint32_t r[64];
    bool pre = false;
    const bool prefetch_on = p.epi_prefetch != 0;
    for (int tile = walk.first; tile < p.total_tiles; tile += walk.step, ++it, cnt0 += n_chunks) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)it >> 1;
      int n_tile, m_tile, img, rem, h0, w0;
      const bool dup = tile_decode(p, walk, tile, m_tile, n_tile);  // mc: repeated tile, nothing is stored
      fd_divmod(p.fd_tiles_per_img, m_tile, img, rem);
      fd_divmod(p.fd_tiles_w, rem, h0, w0);
      h0 *= p.th;
      w0 *= p.tw;
      uint32_t valid = 0xffffffffu;
      if (STATS) {
        int rh, rw;
        fd_divmod(p.fd_tw, row, rh, rw);
        valid = __ballot_sync(0xffffffffu, rh < p.th && h0 + rh < p.lim_h && w0 + rw < p.lim_w);
      }
      // chunks of this tile handled by this group: c = c_first, c_first + 2, ...
      const int c_first = (int)((cnt0 ^ (uint32_t)group) & 1u);
      int c_last = -1;
      for (int c = c_first; c < n_chunks; c += 2) c_last = c;

      mbar_wait(&tfull_bar[buf], use & 1u);
      tc_fence_after();
      if (c_last < 0) {  // nothing to read from this accumulator: release our share right away
        tc_fence_before();
        if (PAIR) {  // one REMOTE arrival per warp on the leader's barrier (256 per tile and CTA cost more than
          __syncwarp();  // the epilogue of a short tile)
          if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[buf]), 0u));
        } else {
          mbar_arrive_warp(&tempty_bar[buf], lane);
        }
      }
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * p.block_n);
      for (int c = c_first; c < n_chunks; c += 2) {
        const uint32_t cnt = cnt0 + (uint32_t)c;
        if (!pre && !(p.epi_debug & 4)) {
          tmem_ld32(t_addr + (uint32_t)(c * 64), r);
          tmem_ld32(t_addr + (uint32_t)(c * 64 + 32), r + 32);
        }
        pre = false;
        tmem_ld_wait();
        if (c == c_last) {  // accumulator fully read by this thread: hand TMEM back to the MMA warp
          tc_fence_before();
          if (PAIR) {
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[buf]), 0u));
          } else {
            mbar_arrive_warp(&tempty_bar[buf], lane);
          }
        }
        const int ch = n_tile * p.block_n + c * 64;
        if (EPI != 0) {
          // ---- packed-half2 path: f16 output, optional bias / f16 residual / ReLU ----
          const int slot = fd_ring_r(p, cnt);
          const uint8_t* in_base = epi_in + (size_t)slot * n_in * kChunkBytes;
          if (n_in > 0) mbar_wait(&ifull_bar[slot], (uint32_t)fd_div(p.fd_ring, (int)cnt) & 1u);
          uint32_t hp[32];
          {
            const uint8_t* b16 = p.bias != nullptr ? reinterpret_cast<const uint8_t*>(sbias) + ch * 2 : nullptr;
            const uint8_t* i0 = p.has_in0 ? in_base : nullptr;
            if (EPI == 1) {
              epi_half_rows<GHND_F16>(r, b16, i0, 0, nullptr, p.relu, hp, row);
            } else {
              const uint8_t* mk = (p.has_in1 && p.in1_mask) ? in_base + p.has_in0 * kChunkBytes : nullptr;
              epi_half_rows<GHND_BF16>(r, b16, i0, p.in0_post, mk, p.relu, hp, row);
            }
          }
          if (n_in > 0) {
            // The operand tile was read through the generic proxy (ld.shared) and will be OVERWRITTEN by
            // the async proxy (the loader warp's next TMA load): a cross-proxy write-after-read.  The
            // mbarrier orders the two warps but not the two proxies -- without this fence a late TMA
            // write could land before these reads were performed (seen as a few corrupted 64-channel
            // rows in the "+res" 1x1 convs, only when another kernel shared the GPU: graph replay with
            // the side stream, scripts/debug/race_hunt.py).
            fence_proxy_async();
            mbar_arrive_warp(&iempty_bar[slot], lane);  // operand buffer consumed
          }
          uint8_t* o_base = o_base0;
          if (p.epi_bufs == 2) {
            o_base += (size_t)(n_staged & 1u) * kChunkBytes;
            if (etid == 0) bulk_wait_read<1>();
          } else {
            if (etid == 0) bulk_wait_read<0>();
          }
          ++n_staged;
          named_bar_sync(bar_id, kEpiGroupThreads);
          epi_stage_packed(hp, o_base, row);
          fence_proxy_async();
          named_bar_sync(bar_id, kEpiGroupThreads);
          if (etid == 0 && !dup) {
            tma_store_4d(&p.tmap_out, o_base, ch, w0, h0, img);
            bulk_commit();
          }
          if (STATS && EPI == 1)  // plain sum / sum of squares of the staged f16 tile (stats_mode 0)
            epi_stats_rows<GHND_F16>(o_base, quarter, lane, valid, wstat, ch, p.cout);
          continue;
        }
        float v[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias != nullptr && !(p.epi_debug & 8)) {
          const float4* b4 = reinterpret_cast<const float4*>(sbias + ch);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 b = b4[j];
            v[4 * j] += b.x;
            v[4 * j + 1] += b.y;
            v[4 * j + 2] += b.z;
            v[4 * j + 3] += b.w;
          }
        }
        const int slot = fd_ring_r(p, cnt);
        const uint8_t* in_base = epi_in + (size_t)slot * n_in * kChunkBytes;
        if (n_in > 0) mbar_wait(&ifull_bar[slot], (uint32_t)fd_div(p.fd_ring, (int)cnt) & 1u);
        if (p.has_in0 && !p.in0_post) {
          if (p.in0_fmt == GHND_F16) epi_add_rows<GHND_F16>(v, in_base, row);
          else epi_add_rows<GHND_BF16>(v, in_base, row);
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        const uint8_t* m_base = in_base + p.has_in0 * kChunkBytes;
        if (p.has_in1 && p.in1_mask) {
          if (p.in1_fmt == GHND_F16) epi_mask_rows<GHND_F16>(v, m_base, row);
          else epi_mask_rows<GHND_BF16>(v, m_base, row);
        }
        if (p.has_in0 && p.in0_post) {
          if (p.in0_fmt == GHND_F16) epi_add_rows<GHND_F16>(v, in_base, row);
          else epi_add_rows<GHND_BF16>(v, in_base, row);
        }
        // operand buffer consumed (the BN-backward statistics still read the in1 tile below)
        const bool late_release = STATS && p.stats_mode == 1;
        if (n_in > 0 && !late_release) {
          fence_proxy_async();  // generic-proxy reads before the async-proxy refill (see the packed path)
          mbar_arrive_warp(&iempty_bar[slot], lane);
        }
        // ---- stage the 64-channel rows and store them with one TMA tensor store ----
        // the store that last used this staging tile has finished reading it
        uint8_t* o_base = o_base0;
        if (p.epi_bufs == 2) {
          o_base += (size_t)(n_staged & 1u) * kChunkBytes;
          if (etid == 0) bulk_wait_read<1>();
        } else {
          if (etid == 0) bulk_wait_read<0>();
        }
        ++n_staged;
        named_bar_sync(bar_id, kEpiGroupThreads);
        if (p.out_fmt == GHND_F16) epi_stage_rows<GHND_F16>(v, o_base, row);
        else epi_stage_rows<GHND_BF16>(v, o_base, row);
        if (prefetch_on) {
          if (c + 2 < n_chunks) {  // this group's next chunk of the same accumulator
            tmem_ld32(t_addr + (uint32_t)((c + 2) * 64), r);
            tmem_ld32(t_addr + (uint32_t)((c + 2) * 64 + 32), r + 32);
            pre = true;
          } else if (tile + walk.step < p.total_tiles) {
            // first chunk of the next tile, if this group has one there and that accumulator is
            // already complete (never wait here: the MMA warp may still need our own release)
            const int c_next = (int)(((cnt0 + (uint32_t)n_chunks) ^ (uint32_t)group) & 1u);
            if (c_next < n_chunks) {
              const int nbuf = buf ^ 1;
              const uint32_t nuse = (uint32_t)(it + 1) >> 1;
              const bool ready = mbar_try_wait(&tfull_bar[nbuf], nuse & 1u);
              if (__all_sync(0xffffffffu, ready)) {
                tc_fence_after();
                const uint32_t n_addr =
                    tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(nbuf * p.block_n);
                tmem_ld32(n_addr + (uint32_t)(c_next * 64), r);
                tmem_ld32(n_addr + (uint32_t)(c_next * 64 + 32), r + 32);
                pre = true;
              }
            }
          }
        }
        if (!(p.epi_debug & 2)) fence_proxy_async();
        named_bar_sync(bar_id, kEpiGroupThreads);
        if (etid == 0 && !(p.epi_debug & 1) && !dup) {
          tma_store_4d(&p.tmap_out, o_base, ch, w0, h0, img);
          bulk_commit();
        }
        if (STATS) {
          if (p.stats_mode == 1) {
            // bf16 gradient out, f16 activation operand (the only combination the plans accept)
            epi_stats2_rows<GHND_BF16, GHND_F16>(o_base, m_base, quarter, lane, valid, wstat, ch, p.cout);
            fence_proxy_async();
            mbar_arrive_warp(&iempty_bar[slot], lane);  // now the operand buffer may be refilled
          } else if (p.out_fmt == GHND_F16) {
            epi_stats_rows<GHND_F16>(o_base, quarter, lane, valid, wstat, ch, p.cout);
          } else {
            epi_stats_rows<GHND_BF16>(o_base, quarter, lane, valid, wstat, ch, p.cout);
          }
        }
      }
    }
    if (etid == 0) bulk_wait_all();
    if (STATS) {
      named_bar_sync(3, kEpiThreads);  // both groups: all shared-memory partial sums are in
      for (int i = threadIdx.x - kEpiWarp0 * 32; i < 2 * p.cout; i += kEpiThreads) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kEpiThreads / 32; ++w) t += sstat[(size_t)w * 2 * p.cout + i];
        atomicAdd(p.stats + i, (double)t);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.mc) cluster_sync_all();  // no CTA leaves while its peer may still arrive on its barriers
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
  if (p.pdl_late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// host side: plan = list of launches (1, or 4 parity classes for a stride-2 dgrad)
// ------------------------------------------------------------------------------------------------
struct ConvLaunch {
  ConvKernelParams p;
  int grid;
  size_t smem;
  // A-operand source tensor [src_n][src_h][src_w][gemm_cin] (conv halo mode re-encodes its tensor map)
  const void* src;
  int src_n, src_h, src_w;
};

}  // namespace ghnd

struct ghnd_conv_plan {
  std::vector<ghnd::ConvLaunch> launches;
  bool stats_zeroed = false;  // GHND_SUMS_ZEROED: the caller zeroes the statistics buffer once per step
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

static const int kSmemBudget = 227 * 1024 - 512 /*barriers + tmem slot*/;  // aligned by declaration: no slack

// Tile-N choice from a cost model fitted to a forced-block_n sweep of every conv of the GHND step
// on B200 (scripts/gpu_bn_sweep.sh, profiles/r1_summary.md): a launch is bound by the slowest of
//   shared-memory ingest  (every tile streams its A and B stages + epilogue operands; ~40 B/clk/SM),
//   tcgen05 issue         (128xNx16 per max(N/2, 32) clk per SM, wave-quantised),
//   the epilogue          (~600 clk per 64-channel chunk, two groups in parallel),
//   HBM                   (unique bytes at ~6.2 TB/s),
// plus 10 % of the non-dominant terms; a shallow (depth-2) operand ring costs another 8 %.
// output staging tiles per epilogue group (GHND_EPI_BUFS=1|2 overrides)
static int epi_bufs_default() {
  static const int v = [] {
    const char* e = getenv("GHND_EPI_BUFS");
    const int n = e ? atoi(e) : 1;  // A/B on the B200: the second tile costs pipeline stages, no gain
    return n == 2 ? 2 : 1;
  }();
  return v;
}

static int pick_block_n(int cout, int m_tiles, int n_units, int n_taps, int row_bytes, int n_in,
                        bool has_bias) {
  const int sms = num_sms();
  if (const char* force = getenv("GHND_BLOCK_N")) {  // tuning switch: force a tile width where legal
    const int bn = atoi(force);
    if ((bn == 64 || bn == 128 || bn == 256) && cout % bn == 0) return bn;
  }
  const double clk = 1.9e9;
  const double px = (double)m_tiles * 128.0;
  const double cin = (double)n_units / n_taps * (row_bytes / 2);
  int best = 64;
  double best_t = 1e30;
  for (int bn = 256; bn >= 64; bn >>= 1) {
    if (cout % bn != 0) continue;
    const double tiles = (double)m_tiles * (cout / bn);
    const double waves = (double)((int64_t)(tiles + sms - 1) / sms);
    const double stage = 128.0 * row_bytes + (double)bn * row_bytes;
    const int fixed = 2 * epi_bufs_default() * kChunkBytes + (has_bias ? cout * 4 : 0);
    const bool ring2 = n_in > 0 && (kSmemBudget - fixed - 3 * (int)stage) / (n_in * kChunkBytes) < 4;
    const double per_tile = n_units * stage + (double)n_in * bn * 256.0;
    const double t_ing = waves * per_tile / (40.0 * clk);
    const double mma_clk = bn / 2.0 > 32.0 ? bn / 2.0 : 32.0;
    const double t_mma = waves * n_units * (row_bytes / 32) * mma_clk / clk;
    const double t_epi = waves * (bn / 64) * 600.0 * (1.0 + 0.5 * n_in) / clk / 2.0;
    const double t_hbm = (px * cout * 2.0 * (1 + n_in) + px * cin * 2.0 * (n_taps == 1 ? 1.0 : 1.2)) / 6.2e12;
    double mx = t_ing;
    if (t_mma > mx) mx = t_mma;
    if (t_epi > mx) mx = t_epi;
    if (t_hbm > mx) mx = t_hbm;
    double t = mx + 0.1 * (t_ing + t_mma + t_epi + t_hbm - mx);
    if (ring2) t *= 1.08;
    // two epilogue operands (residual + mask, or mask + accumulate): the forced-N sweep of the final
    // build has N=128 ahead of N=256 on every such launch (profiles/r1g_block_n_sweep.txt)
    if (n_in >= 2 && bn == 256) t *= 1.15;
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

// Shared-memory plan of one launch: [stages][2 staging tiles][operand ring][bias][stats][barriers].

// Conv halo mode planning (after the generic fields are set): returns true when the launch was switched
// to MODE 2.  Requires stride-1 taps on ONE view, th x 8 tiles, a packed epilogue.
static bool plan_conv_halo(ConvLaunch* L, const ghnd_conv_desc_t* d, int gemm_cin, int gemm_cout, int n_in,
                           int fixed_no_stages) {
  ConvKernelParams& p = L->p;
  static const bool off = [] {
    const char* e = getenv("GHND_CONV_HALO");
    return e != nullptr && atoi(e) == 0;
  }();
  if (off || p.n_taps < 2 || p.tw != 8 || p.kblock != 64 || p.epi_half == 0) return false;
  int dh0 = p.taps[0].dh, dh1 = dh0, dw0 = p.taps[0].dw, dw1 = dw0;
  for (int t = 0; t < p.n_taps; ++t) {
    if (p.taps[t].map != 0) return false;
    dh0 = p.taps[t].dh < dh0 ? p.taps[t].dh : dh0;
    dh1 = p.taps[t].dh > dh1 ? p.taps[t].dh : dh1;
    dw0 = p.taps[t].dw < dw0 ? p.taps[t].dw : dw0;
    dw1 = p.taps[t].dw > dw1 ? p.taps[t].dw : dw1;
  }
  const int hh = p.th + (dh1 - dh0), hw = p.tw + (dw1 - dw0);
  if (hw > 256 || hh > 256) return false;
  const int k_chunks = gemm_cin / 64;
  const int a_box = hh * hw * 128;
  const int a_halo = (a_box + 1023) / 1024 * 1024;
  // tile width: the largest N whose whole weight slab stays resident next to two halo tiles; else the
  // cost model's N with the weights streamed tap by tap
  const int ring_bytes = p.ring * n_in * kChunkBytes;
  const int avail = kSmemBudget - fixed_no_stages - (n_in > 0 ? ring_bytes : 0);
  int bn_res = 0;
  for (int bn = 256; bn >= 64; bn >>= 1) {
    if (gemm_cout % bn != 0) continue;
    if (p.n_taps * k_chunks * bn * 128 + 2 * a_halo <= avail) {
      bn_res = bn;
      break;
    }
  }
  if (const char* force = getenv("GHND_BLOCK_N")) {
    const int bn = atoi(force);
    if ((bn == 64 || bn == 128 || bn == 256) && gemm_cout % bn == 0)
      bn_res = (p.n_taps * k_chunks * bn * 128 + 2 * a_halo <= avail) ? bn : 0, p.block_n = bn;
  }
  static const bool no_res = getenv("GHND_CONV_HALO_NO_RESIDENT") != nullptr;  // A/B switch
  if (no_res) bn_res = 0;
  // Measured on the B200 (profiles/r2_summary.md): the mode pays when the weights are resident and the CTA
  // keeps (nearly) the whole output width -- the small-C convs whose shared-memory ingest was dominated by
  // re-fetched weights and 4x / 9x re-fetched pixels (3x3 C64->K64 -27 %, 2x2 C256->K64 -24 %, its dgrad
  // -22 %, 3x3 C128->K128 -12 %).  It LOSES on the deep convs: with the weights streamed the 16x8 tiling
  // (74 % full on a 50x84 map, one more wave) costs more than the saved A traffic (3x3 C256->K256 +44 %),
  // and a resident slab that forces four N=64 tiles per pixel block leaves only two halo slots in flight
  // (2x2 C256->K256 +44 %).  GHND_CONV_HALO_ALL=1 lifts the restriction (A/B switch).
  static const bool halo_all = getenv("GHND_CONV_HALO_ALL") != nullptr;
  if (!halo_all && (bn_res == 0 || gemm_cout / bn_res > 2)) return false;
  if (bn_res) p.block_n = bn_res;
  p.b_bytes = p.block_n * 128;
  p.n_tiles_n = gemm_cout / p.block_n;
  const int m_tiles = p.n_img * p.tiles_h * p.tiles_w;
  p.total_tiles = m_tiles * p.n_tiles_n;
  p.fd_tiles_n = make_fastdiv(p.n_tiles_n);
  p.idesc = make_idesc(d->src_fmt, d->w_fmt, 0, 0, kBlockM, p.block_n);
  // the weight map's box follows block_n
  {
    const int total_taps = d->R * d->S;
    uint64_t dims[2] = {(uint64_t)total_taps * gemm_cin, (uint64_t)gemm_cout};
    uint64_t str[2] = {2, (uint64_t)total_taps * gemm_cin * 2};
    uint32_t box[2] = {64u, (uint32_t)p.block_n};
    if (encode_tmap(&p.tmap_b, 2, 2, const_cast<void*>(d->weights), dims, str, box, 128) != GHND_OK) return false;
  }
  {  // A: one dense box of hh x hw pixels x 64 channels (out-of-bounds rows / columns zero-filled)
    uint64_t dims[4] = {(uint64_t)gemm_cin, (uint64_t)L->src_w, (uint64_t)L->src_h, (uint64_t)L->src_n};
    uint64_t str[4] = {2, (uint64_t)gemm_cin * 2, (uint64_t)L->src_w * gemm_cin * 2,
                       (uint64_t)L->src_h * L->src_w * gemm_cin * 2};
    uint32_t box[4] = {64u, (uint32_t)hw, (uint32_t)hh, 1u};
    if (encode_tmap(&p.tmap_a[0], 2, 4, const_cast<void*>(L->src), dims, str, box, 128) != GHND_OK) return false;
    for (int i = 1; i < 4; ++i) p.tmap_a[i] = p.tmap_a[0];
  }
  p.a_box_bytes = a_box;
  p.a_halo_bytes = a_halo;
  p.halo_w = hw;
  p.org_dh = dh0;
  p.org_dw = dw0;
  for (int t = 0; t < p.n_taps; ++t) p.tap_off[t] = ((p.taps[t].dh - dh0) * hw + (p.taps[t].dw - dw0)) * 128;
  int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  if (bn_res) {
    p.b_resident = 1;
    p.w_res_bytes = p.n_taps * k_chunks * p.b_bytes;
    p.n_stages = 0;
    p.stage_bytes = 0;
    p.units_per_stage = 1;
    int slots = (avail - p.w_res_bytes) / a_halo;
    p.a_slots = slots > kMaxASlots ? kMaxASlots : slots;
    grid -= grid % p.n_tiles_n;  // every CTA keeps ONE n-tile (its resident weights)
    if (grid < p.n_tiles_n) return false;
  } else {
    p.b_resident = 0;
    p.w_res_bytes = 0;
    // >= ~256 tensor-pipe cycles of B per stage, at most all taps of a chunk
    const int unit_cycles = 4 * (p.block_n / 2);
    int u = (256 + unit_cycles - 1) / unit_cycles;
    if (u > p.n_taps) u = p.n_taps;
    if (u < 1) u = 1;
    p.units_per_stage = u;
    p.stage_bytes = u * p.b_bytes;
    p.a_slots = k_chunks >= 2 ? 3 : 2;
    int stages = (avail - p.a_slots * a_halo) / p.stage_bytes;
    if (stages < 2) {
      p.a_slots = 2;
      stages = (avail - p.a_slots * a_halo) / p.stage_bytes;
    }
    if (stages < 2) return false;
    p.n_stages = stages > kMaxStages ? kMaxStages : stages;
  }
  L->grid = grid;
  L->smem = (size_t)p.a_slots * a_halo + (size_t)p.n_stages * p.stage_bytes + (size_t)p.w_res_bytes +
            (size_t)fixed_no_stages + (size_t)(n_in > 0 ? ring_bytes : 0) + 512;
  return true;
}

static int max_conv_pairs();  // CTA pairs the device can hold at once (0: cluster launches unavailable)

static int finish_launch(ConvLaunch* L, const ghnd_conv_desc_t* d, int gemm_cin, int gemm_cout,
                         const void* weights, int total_taps, const IoGeom& g, int kblock = 64,
                         int force_block_n = 0, bool try_halo = false) {
  ConvKernelParams& p = L->p;
  p.cin = gemm_cin;
  p.cout = gemm_cout;
  p.kblock = kblock;
  const int row_bytes = kblock * 2;
  p.a_bytes = kBlockM * row_bytes;
  const int m_tiles = p.n_img * p.tiles_h * p.tiles_w;
  const int n_units = p.n_taps * (gemm_cin / kblock);
  p.has_in0 = p.has_in1 = p.in0_post = 0;
  if (d->residual != nullptr && d->accumulate) {
    set_error("conv: residual and accumulate cannot be combined in one launch");
    return GHND_ERR_UNSUPPORTED;
  }
  if (d->residual != nullptr || d->accumulate) p.has_in0 = 1;
  if (d->mask != nullptr) p.has_in1 = 1;
  const int n_in = p.has_in0 + p.has_in1;
  p.block_n = force_block_n ? force_block_n
                            : pick_block_n(gemm_cout, m_tiles, n_units, p.n_taps, row_bytes, n_in,
                                           d->bias != nullptr);
  p.b_bytes = p.block_n * row_bytes;
  p.n_tiles_n = gemm_cout / p.block_n;
  p.total_tiles = m_tiles * p.n_tiles_n;
  p.fd_tiles_n = make_fastdiv(p.n_tiles_n);
  p.fd_tiles_per_img = make_fastdiv(p.tiles_h * p.tiles_w);
  p.fd_tiles_w = make_fastdiv(p.tiles_w);
  p.fd_tw = make_fastdiv(p.tw);
  // units per stage: >= ~256 tensor-pipe cycles (block_n/2 per K=16 step) behind every barrier wait
  {
    const int unit_cycles = (kblock / 16) * (p.block_n / 2);
    int u = (256 + unit_cycles - 1) / unit_cycles;
    if (u > 4) u = 4;
    if (u > n_units) u = n_units;
    if (u < 1) u = 1;
    p.units_per_stage = u;
  }
  p.stage_bytes = p.units_per_stage * (p.a_bytes + p.b_bytes);
  p.a_box_bytes = p.tw * p.th * row_bytes;
  p.io_box_bytes = p.tw * p.th * 128;
  p.lim_h = (g.H - g.ph + g.sub_h - 1) / g.sub_h;
  p.lim_w = (g.W - g.pw + g.sub_w - 1) / g.sub_w;
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
  if (d->residual != nullptr) {
    p.in0_fmt = d->res_fmt;
    rc = make_lattice_map(&p.tmap_in0, d->residual, g.N, g.H, g.W, gemm_cout, g.sub_h, g.sub_w, g.ph,
                          g.pw, p.th, p.tw);
  } else if (d->accumulate) {
    p.in0_post = 1;
    p.in0_fmt = d->dst_fmt;
  }
  if (rc != GHND_OK) return rc;
  if (d->mask != nullptr) {
    p.in1_fmt = d->mask_fmt;
    rc = make_lattice_map(&p.tmap_in1, d->mask, g.N, g.H, g.W, gemm_cout, g.sub_h, g.sub_w, g.ph, g.pw,
                          p.th, p.tw);
    if (rc != GHND_OK) return rc;
  }
  p.bias = d->bias;
  p.out_fmt = d->dst_fmt;
  p.relu = d->relu;
  p.halo = 0;
  p.w_res_bytes = 0;
  {
    static const bool pf = [] {
      const char* e = getenv("GHND_EPI_PREFETCH");  // A/B on the B200: no gain, off by default
      return e != nullptr && atoi(e) != 0;
    }();
    p.epi_prefetch = pf ? 1 : 0;
    const char* dbg = getenv("GHND_EPI_DEBUG");
    p.epi_debug = dbg ? atoi(dbg) : 0;
    // epi_half: 0 fp32 epilogue; 1 packed f16 (forward: bias / f16 residual / ReLU); 2 packed bf16 (data
    // gradients: bf16 residual or accumulate, ReLU mask from the f16 activation).  GHND_EPI_HALF=0|1|2
    // caps the level.
    static const int half_level = [] {
      const char* e = getenv("GHND_EPI_HALF");
      return e == nullptr ? 2 : atoi(e);
    }();
    p.epi_half = 0;
    const bool plain_stats = d->stats == nullptr || d->stats_mode == 0;
    if (plain_stats && p.epi_debug == 0 && !p.epi_prefetch) {
      if (half_level >= 1 && d->dst_fmt == GHND_F16 && d->mask == nullptr && !d->accumulate &&
          (d->residual == nullptr || d->res_fmt == GHND_F16))
        p.epi_half = 1;
      else if (d->stats == nullptr && half_level >= 2 && d->dst_fmt == GHND_BF16 && (d->mask == nullptr || d->mask_fmt == GHND_F16) &&
               (d->residual == nullptr || d->res_fmt == GHND_BF16))
        p.epi_half = 2;
    }
  }
  p.stats = d->stats;
  p.stats_mode = d->stats != nullptr ? d->stats_mode : 0;
  p.in1_mask = (d->mask != nullptr && !(d->stats != nullptr && d->stats_mode == 1 && d->mask_stats_only)) ? 1 : 0;
  // shared-memory split: >= 3 pipeline stages first, then the operand ring (deep enough to keep
  // ~64 KB of residual / mask loads in flight per SM), the rest goes to more stages
  // Two staging tiles per epilogue group when the pipeline keeps >= 3 stages (and the operand ring
  // its 4 slots) next to them; otherwise one.
  p.epi_bufs = epi_bufs_default();
  if (p.epi_bufs == 2) {
    const int fixed2 = 4 * kChunkBytes + (d->bias ? gemm_cout * 4 : 0) + (d->stats ? gemm_cout * 8 * (kEpiThreads / 32) : 0);
    const int ring2 = n_in > 0 ? ((kSmemBudget - fixed2 - 3 * p.stage_bytes) / (n_in * kChunkBytes) >= 4 ? 4 : 2) : 0;
    const int ring1 = n_in > 0 ? ((kSmemBudget - (fixed2 - 2 * kChunkBytes) - 3 * p.stage_bytes) /
                                              (n_in * kChunkBytes) >= 4 ? 4 : 2) : 0;
    const int stages2 = (kSmemBudget - fixed2 - ring2 * n_in * kChunkBytes) / p.stage_bytes;
    if (stages2 < 3 || ring2 < ring1) p.epi_bufs = 1;
  }
  const int fixed = 2 * p.epi_bufs * kChunkBytes + (d->bias ? gemm_cout * 4 : 0) +
                    (d->stats ? gemm_cout * 8 * (kEpiThreads / 32) : 0);
  const int stages_per_tile = (n_units + p.units_per_stage - 1) / p.units_per_stage;
  int ring = 0;
  if (n_in > 0) {
    // The ring depth must be EVEN: the two epilogue groups take alternate chunks, so with an even
    // depth every slot is always consumed by the same group.  With an odd depth a slot alternates
    // between the groups, and a group that runs ahead can wait for parity p of a slot barrier that
    // is still one whole phase behind -- a parity wait then passes spuriously (mbarrier parity
    // aliasing).  Seen on B200 as a rare cudaErrorLaunchFailure in the C64->K256 "+res" convs.
    ring = (kSmemBudget - fixed - 3 * p.stage_bytes) / (n_in * kChunkBytes);
    ring = ring >= 4 ? 4 : 2;
    // experiment: forced ring depth (even, <= kMaxRing), the operand stages get what is left (>= 2)
    static const int ring_env = [] {
      const char* e = getenv("GHND_CONV_RING");
      return e == nullptr ? 0 : atoi(e);
    }();
    if (ring_env >= 2 && ring_env <= kMaxRing && ring_env % 2 == 0 &&
        (kSmemBudget - fixed - ring_env * n_in * kChunkBytes) / p.stage_bytes >= 2)
      ring = ring_env;
  }
  p.ring = ring > 0 ? ring : 1;
  p.fd_ring = make_fastdiv(p.ring);
  int stages = (kSmemBudget - fixed - ring * n_in * kChunkBytes) / p.stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  (void)stages_per_tile;
  if (stages < 2) {
    set_error("conv: shared memory budget leaves %d pipeline stages", stages);
    return GHND_ERR_UNSUPPORTED;
  }
  p.n_stages = stages;
  L->grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  L->smem = (size_t)p.n_stages * p.stage_bytes + (size_t)fixed + (size_t)ring * n_in * kChunkBytes +
            512 /*barriers*/;
  p.a_halo_bytes = 0;
  p.b_resident = 0;
  p.mc = 0;
  p.m_tiles = m_tiles;
  if (try_halo && plan_conv_halo(L, d, gemm_cin, gemm_cout, n_in, fixed)) return GHND_OK;
  // CTA pairs (generic mode).  The trunk convs re-read their block_n x K weight slab for every m-tile, and that
  // ingest (L2 -> SM) bounds most of them.  Two CTAs of a cluster take the same n-tile of two neighbouring
  // m-tiles in lock step:
  //   GHND_CONV_PAIR=1 (mc 2): ONE cta_group::2 MMA (M = 256) spans the pair, each CTA loads and holds only HALF
  //     of every weight tile -- the weight bytes per SM are halved, and the smaller stages deepen the pipeline;
  //   GHND_CONV_MC=1   (mc 1): cta_group::1 MMAs, each CTA fetches half of the tile and TMA-multicasts it to both.
  //     Measured neutral on the B200 (712.6 vs 711.9 img/s): every SM still ingests the whole tile.
  // Not with fused statistics (a repeated tile would count twice) and not below one pair per SM pair.
  static const int mc_env = [] {
    const char* e = getenv("GHND_CONV_MC");
    return e == nullptr ? 0 : atoi(e);
  }();
  // Per-layer A/B on the B200 (profiles/r2_summary.md): the pair wins where the K loop dominates the tile
  // (K >= 1024: 3x3 convs, the 1x1 reductions of layers 3-4, the 2x2 C256 dgrad: -8 .. -24 %) and loses where
  // the tile is epilogue-bound (1x1 expansions with a residual, K <= 512: +20 .. +69 %, the MMAs of the next
  // tile wait for BOTH CTAs' epilogues) -> default: pairs for launches with K >= 1024.  GHND_CONV_PAIR=0 off,
  // 1 = every eligible launch with at least one pair per SM pair, 2 = every eligible launch.
  static const int pair_env = [] {
    const char* e = getenv("GHND_CONV_PAIR");
    return e == nullptr ? -1 : atoi(e);
  }();
  // (K = 512: the masked data-gradient launches gain 9 .. 17 %, the forward ones lose 4 .. 19 %)
  const int k_total = n_units * kblock;
  const bool pair_auto = k_total >= 1024 || (k_total >= 512 && d->mask != nullptr && d->dst_fmt == GHND_BF16);
  const int pair_want = pair_env >= 0 ? pair_env : (pair_auto ? 1 : 0);
  const int want = pair_want ? pair_want : mc_env;
  if (want != 0 && !p.halo && kblock == 64 && d->stats == nullptr && p.block_n >= 128 && m_tiles >= 2 &&
      num_sms() % 2 == 0 && (want == 2 || (int64_t)((m_tiles + 1) / 2) * p.n_tiles_n >= num_sms() / 2) &&
      max_conv_pairs() > 0) {
    uint32_t hbox[2] = {(uint32_t)kblock, (uint32_t)(p.block_n / 2)};
    if (encode_tmap(&p.tmap_b, 2, 2, const_cast<void*>(weights), dims, str, hbox, row_bytes) == GHND_OK) {
      p.mc = pair_want ? 2 : 1;
      p.total_tiles = ((m_tiles + 1) / 2) * p.n_tiles_n;
      L->grid = 2 * p.total_tiles < num_sms() ? 2 * p.total_tiles : num_sms();
      if (p.mc == 2) {
        p.b_bytes = (p.block_n / 2) * row_bytes;  // this CTA's half of a weight tile
        p.stage_bytes = p.units_per_stage * (p.a_bytes + p.b_bytes);
        int st2 = (kSmemBudget - fixed - ring * n_in * kChunkBytes) / p.stage_bytes;
        if (st2 > kMaxStages) st2 = kMaxStages;
        p.n_stages = st2;
        L->smem = (size_t)p.n_stages * p.stage_bytes + (size_t)fixed + (size_t)ring * n_in * kChunkBytes + 512;
        p.idesc = make_idesc(d->src_fmt, d->w_fmt, 0, 0, 2 * kBlockM, p.block_n);
        if (L->grid > 2 * max_conv_pairs()) L->grid = 2 * max_conv_pairs();
      }
    } else {
      rc = encode_tmap(&p.tmap_b, 2, 2, const_cast<void*>(weights), dims, str, box, row_bytes);
      if (rc != GHND_OK) return rc;
    }
  }
  return GHND_OK;
}

// Launch with the programmatic-stream-serialization attribute: the kernel's prologue (barrier init,
// TMEM allocation, tensor-map prefetch) may start while the previous kernel of the stream drains;
// the kernel itself waits (griddepcontrol.wait) before touching global memory.
static cudaError_t launch_conv(const ConvLaunch& L, cudaStream_t st) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)L.grid, 1, 1);
  cfg.blockDim = dim3(kConvThreads, 1, 1);
  cfg.dynamicSmemBytes = L.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  static const bool no_pdl = getenv("GHND_NO_PDL") != nullptr;  // debugging switch
  int na = 0;
  if (!no_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (L.p.mc) {  // CTA pairs sharing their weight tiles (see finish_launch)
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (L.p.mc == 2) {  // CTA pair: the instantiations with cta_group::2 instructions (cluster launches only)
    if (L.p.epi_half == 1) return cudaLaunchKernelEx(&cfg, conv_tc_kernel<1, 0, false, true>, L.p);
    if (L.p.epi_half == 2) return cudaLaunchKernelEx(&cfg, conv_tc_kernel<2, 0, false, true>, L.p);
    return cudaLaunchKernelEx(&cfg, conv_tc_kernel<0, 0, false, true>, L.p);
  }
  if (L.p.a_halo_bytes > 0) {  // conv halo mode (MODE 2): packed epilogues only
    if (L.p.epi_half == 2) return cudaLaunchKernelEx(&cfg, conv_tc_kernel<2, 2, false>, L.p);
    if (L.p.stats != nullptr) return cudaLaunchKernelEx(&cfg, conv_tc_kernel<1, 2, true>, L.p);
    return cudaLaunchKernelEx(&cfg, conv_tc_kernel<1, 2, false>, L.p);
  }
  if (L.p.epi_half == 1) {
    if (L.p.halo) return cudaLaunchKernelEx(&cfg, conv_tc_kernel<1, 1, false>, L.p);  // stem: no stats
    if (L.p.stats != nullptr) return cudaLaunchKernelEx(&cfg, conv_tc_kernel<1, 0, true>, L.p);
    return cudaLaunchKernelEx(&cfg, conv_tc_kernel<1, 0, false>, L.p);
  }
  if (L.p.epi_half == 2 && !L.p.halo) return cudaLaunchKernelEx(&cfg, conv_tc_kernel<2, 0, false>, L.p);
  if (L.p.halo) return cudaLaunchKernelEx(&cfg, conv_tc_kernel<0, 1, false>, L.p);
  if (L.p.stats != nullptr) return cudaLaunchKernelEx(&cfg, conv_tc_kernel<0, 0, true>, L.p);
  return cudaLaunchKernelEx(&cfg, conv_tc_kernel<0, 0, false>, L.p);
}

static bool fmt_ok(int f) { return f == GHND_F16 || f == GHND_BF16; }

static int set_conv_attr() {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaSuccess;
#define GHND_SET_SMEM(...)                                                                        \
  if (e == cudaSuccess)                                                                           \
    e = cudaFuncSetAttribute(conv_tc_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                             227 * 1024)
    GHND_SET_SMEM(0, 0, false);
    GHND_SET_SMEM(0, 0, true);
    GHND_SET_SMEM(0, 1, false);
    GHND_SET_SMEM(1, 0, false);
    GHND_SET_SMEM(1, 0, true);
    GHND_SET_SMEM(1, 1, false);
    GHND_SET_SMEM(2, 0, false);
    GHND_SET_SMEM(1, 2, false);
    GHND_SET_SMEM(1, 2, true);
    GHND_SET_SMEM(2, 2, false);
    GHND_SET_SMEM(0, 0, false, true);
    GHND_SET_SMEM(1, 0, false, true);
    GHND_SET_SMEM(2, 0, false, true);
#undef GHND_SET_SMEM
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_tc_kernel)");
    attr_set = true;
  }
  return GHND_OK;
}

// As many CTA pairs as the device can hold at once (the persistent loop takes any even grid).  A device that
// cannot co-schedule a pair of these CTAs (MIG slices, odd SM masks) keeps the single-CTA kernels.
static int max_conv_pairs() {
  static const int n_pairs = [] {
    if (set_conv_attr() != GHND_OK) return 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(num_sms() & ~1), 1, 1);
    cfg.blockDim = dim3(kConvThreads, 1, 1);
    cfg.dynamicSmemBytes = 227 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, conv_tc_kernel<1, 0, false, true>, &cfg) != cudaSuccess) {
      cudaGetLastError();
      return 0;
    }
    return n;
  }();
  return n_pairs;
}

}  // namespace ghnd

extern "C" {

int ghnd_conv_plan_create(const ghnd_conv_desc_t* d_in, ghnd_conv_plan_t** out) {
  using namespace ghnd;
  GHND_CHECK_ARG(d_in && out, "conv_plan_create: null argument");
  *out = nullptr;
  // GHND_SUMS_ZEROED in stats_mode: the caller zeroes `stats` once per step, the plan only accumulates
  ghnd_conv_desc_t d_copy = *d_in;
  const bool stats_zeroed = (d_copy.stats_mode & GHND_SUMS_ZEROED) != 0;
  d_copy.stats_mode &= ~GHND_SUMS_ZEROED;
  const ghnd_conv_desc_t* d = &d_copy;
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
  GHND_CHECK_ARG(d->stats == nullptr || d->stats_mode == 0 || d->stats_mode == 1, "conv: bad stats_mode %d",
                 d->stats_mode);
  GHND_CHECK_ARG(d->stats == nullptr || d->stats_mode != 0 ||
                     (d->kind == GHND_CONV_FWD && d->stride == 1 && d->K <= 256),
                 "conv: fused output statistics need a stride-1 forward conv with K <= 256");
  GHND_CHECK_ARG(d->stats == nullptr || d->stats_mode != 1 ||
                     (d->dst_fmt == GHND_BF16 && d->mask_fmt == GHND_F16),
                 "conv: BN-backward statistics need a bf16 output and an f16 mask operand");
  GHND_CHECK_ARG(d->stats == nullptr || d->stats_mode != 1 ||
                     (d->stride == 1 && d->mask != nullptr && !d->accumulate &&
                      (d->kind == GHND_CONV_FWD ? d->K : d->C) <= 256),
                 "conv: BN-backward statistics need a stride-1 launch with a mask operand, no "
                 "accumulate and <= 256 output channels");
  GHND_CHECK_ARG(((uintptr_t)d->src % 16) == 0 && ((uintptr_t)d->weights % 16) == 0 &&
                     ((uintptr_t)d->dst % 16) == 0 && ((uintptr_t)d->residual % 16) == 0 &&
                     ((uintptr_t)d->mask % 16) == 0 && ((uintptr_t)d->bias % 16) == 0,
                 "conv: tensors must be 16-byte aligned");
  const int Ho = (d->H + 2 * d->pad - d->R) / d->stride + 1;
  const int Wo = (d->W + 2 * d->pad - d->S) / d->stride + 1;
  GHND_CHECK_ARG(Ho > 0 && Wo > 0, "conv: empty output");

  ghnd_conv_plan* plan = new ghnd_conv_plan();
  plan->stats_zeroed = stats_zeroed;
  int rc = GHND_OK;
  const int taps_total = d->R * d->S;

  if (d->kind == GHND_CONV_FWD) {
   bool halo_failed = false;
   for (int pass = 0; pass < 2; ++pass) {  // pass 0 tries the conv halo geometry, pass 1 is the generic tiling
    ConvLaunch L;
    memset(&L, 0, sizeof(L));
    ConvKernelParams& p = L.p;
    IoGeom g;
    bool want_halo = false;
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
      // conv halo mode (stride 1, several taps): th x 8 tiles, see plan_conv_halo
      want_halo = d->stride == 1 && taps_total > 1 && Wo >= 8 && Ho >= 4;
      if (want_halo && !halo_failed) {
        p.th = Ho < 16 ? Ho : 16;
        p.tw = 8;
      }
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
    L.src = d->src;
    L.src_n = d->N;
    L.src_h = d->H;
    L.src_w = d->W;
    const bool try_halo = want_halo && !halo_failed;
    if (rc == GHND_OK) rc = finish_launch(&L, d, d->C, d->K, d->weights, taps_total, g, 64, 0, try_halo);
    if (rc == GHND_OK && try_halo && L.p.a_halo_bytes == 0) {
      halo_failed = true;  // not eligible (epilogue kind / shared memory): rebuild with the generic tiling
      continue;
    }
    if (rc == GHND_OK) plan->launches.push_back(L);
    break;
   }
  } else {
    // DGRAD: src = dy [N,Ho,Wo,K], dst = dx [N,H,W,C], weights [C][R][S][K]
    const int sub = d->stride;
    bool halo_failed = false;
    for (int ph = 0; ph < sub && rc == GHND_OK; ++ph) {
      for (int pw = 0; pw < sub && rc == GHND_OK; ++pw) {
        ConvLaunch L;
        memset(&L, 0, sizeof(L));
        ConvKernelParams& p = L.p;
        bool want_halo = false;
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
          // stride-1 dgrad with several taps: conv halo geometry first (th x 8 tiles)
          want_halo = sub == 1 && nt > 1 && wi >= 8 && hi >= 4;
          if (want_halo && !halo_failed) {
            p.th = hi < 16 ? hi : 16;
            p.tw = 8;
          }
          p.tiles_h = (hi + p.th - 1) / p.th;
          p.tiles_w = (wi + p.tw - 1) / p.tw;
          g = IoGeom{d->N, d->H, d->W, sub, sub, ph, pw};
          rc = make_lattice_map(&p.tmap_a[0], d->src, d->N, Ho, Wo, d->K, 1, 1, 0, 0, p.th, p.tw);
        }
        for (int i = 1; i < 4 && rc == GHND_OK; ++i) p.tmap_a[i] = p.tmap_a[0];
        L.src = d->src;
        L.src_n = d->N;
        L.src_h = Ho;
        L.src_w = Wo;
        const bool try_halo = want_halo && !halo_failed;
        if (rc == GHND_OK) rc = finish_launch(&L, d, d->K, d->C, d->weights, taps_total, g, 64, 0, try_halo);
        if (rc == GHND_OK && try_halo && L.p.a_halo_bytes == 0) {
          halo_failed = true;  // rebuild this parity class with the generic tiling
          --pw;
          continue;
        }
        halo_failed = false;
        if (rc == GHND_OK) {
          // parity classes after the first: concurrent with their predecessor (disjoint lattice outputs)
          static const bool serial = getenv("GHND_S2_SERIAL") != nullptr;  // A/B switch
          if (sub > 1 && !serial && !plan->launches.empty()) L.p.pdl_late_wait = 1;
          plan->launches.push_back(L);
        }
      }
    }
  }
  // Stride-2 dgrad: the parity-class launches run side by side (pdl_late_wait), but each one asked for every SM,
  // so the later ones only trickled onto SMs the earlier ones had left and every CTA saw about one tile
  // (prologue + pipeline fill + a bare epilogue per tile).  Give each launch a share of the SMs in proportion to its
  // work (tiles x taps): all of them are resident at once and every CTA streams several tiles through the
  // mainloop / epilogue overlap.  GHND_S2_SHARE=0: the old full-width grids.
  static const bool s2_share = [] {
    const char* e = getenv("GHND_S2_SHARE");
    return e == nullptr || atoi(e) != 0;
  }();
  if (rc == GHND_OK && s2_share && d->kind == GHND_CONV_DGRAD && d->stride > 1 && plan->launches.size() > 1) {
    bool late = true;
    for (size_t i = 1; i < plan->launches.size(); ++i) late = late && plan->launches[i].p.pdl_late_wait;
    double total = 0.0;
    std::vector<double> work;
    for (const ConvLaunch& L : plan->launches) {
      work.push_back((double)(L.p.mc ? 2 : 1) * L.p.total_tiles * L.p.n_taps);
      total += work.back();
    }
    const int sms = num_sms();
    // (a launch with several waves of tiles already keeps its CTAs busy: 200x336 C128 measured 48.6 -> 49.7 us)
    bool small = true;
    for (const ConvLaunch& L : plan->launches) small = small && L.p.total_tiles * (L.p.mc ? 2 : 1) < 2 * sms;
    if (late && small && total > 0.0) {
      int used = 0;
      std::vector<int> grid(work.size());
      for (size_t i = 0; i < work.size(); ++i) {
        const ConvLaunch& L = plan->launches[i];
        const int unit = L.p.mc ? 2 : 1;
        int g = (int)(sms * work[i] / total) / unit * unit;
        if (g < unit) g = unit;
        if (g > L.grid) g = L.grid;
        grid[i] = g;
        used += g;
      }
      // SMs left by the rounding go to the launches with the most work per CTA
      for (bool grew = true; grew && used < sms;) {
        grew = false;
        size_t best = work.size();
        double load = 0.0;
        for (size_t i = 0; i < work.size(); ++i) {
          const int unit = plan->launches[i].p.mc ? 2 : 1;
          if (grid[i] + unit > plan->launches[i].grid || used + unit > sms) continue;
          if (work[i] / grid[i] > load) {
            load = work[i] / grid[i];
            best = i;
          }
        }
        if (best < work.size()) {
          const int unit = plan->launches[best].p.mc ? 2 : 1;
          grid[best] += unit;
          used += unit;
          grew = true;
        }
      }
      if (used <= sms)
        for (size_t i = 0; i < work.size(); ++i) plan->launches[i].grid = grid[i];
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
    if (L.p.stats != nullptr && !plan->stats_zeroed)
      GHND_CUDA(cudaMemsetAsync(L.p.stats, 0, (size_t)2 * L.p.cout * sizeof(double),
                                (cudaStream_t)stream));
    GHND_CUDA(launch_conv(L, (cudaStream_t)stream));
  }
  return GHND_OK;
}

int ghnd_conv_plan_run_range(const ghnd_conv_plan_t* plan, int first, int count, void* stream) {
  using namespace ghnd;
  GHND_CHECK_ARG(plan != nullptr, "conv_plan_run_range: null plan");
  GHND_CHECK_ARG(first >= 0 && count >= 0 && (size_t)(first + count) <= plan->launches.size(),
                 "conv_plan_run_range: [%d, %d) outside the plan's %d launches", first, first + count,
                 (int)plan->launches.size());
  for (int i = first; i < first + count; ++i) {
    const ConvLaunch& L = plan->launches[i];
    if (L.p.stats != nullptr && i == 0 && !plan->stats_zeroed)
      GHND_CUDA(cudaMemsetAsync(L.p.stats, 0, (size_t)2 * L.p.cout * sizeof(double),
                                (cudaStream_t)stream));
    GHND_CUDA(launch_conv(L, (cudaStream_t)stream));
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
  return ghnd_stem_conv_plan_create_k(x_packed, x_fmt, w_packed, w_fmt, bias, y, y_fmt, N, Hp, Wp, 64,
                                      out);
}

int ghnd_stem_conv_plan_create_k(const void* x_packed, int x_fmt, const void* w_packed, int w_fmt,
                                 const float* bias, void* y, int y_fmt, int N, int Hp, int Wp, int K,
                                 ghnd_stem_plan_t** out) {
  using namespace ghnd;
  GHND_CHECK_ARG(out && x_packed && w_packed && y, "stem_conv_plan_create: null argument");
  *out = nullptr;
  GHND_CHECK_ARG(K == 64 || K == 128 || K == 256, "stem conv: K=%d output channels unsupported", K);
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
    static const bool no_halo = getenv("GHND_NO_STEM_HALO") != nullptr;  // A/B + debugging switch
    if (!no_halo && J >= 8 && K <= 256) {
      // halo mode: tile = 16 output rows x 8 column slots; one TMA box of 37 image rows per tile,
      // weights resident in shared memory (see the kernel's halo branches)
      p.th = 16;
      p.tw = 8;
      p.tiles_h = (Ho + p.th - 1) / p.th;
      p.tiles_w = (J + p.tw - 1) / p.tw;
      const int halo_rows = 2 * p.th + 5;
      const uint8_t* base = static_cast<const uint8_t*>(x_packed) + (size_t)q * 8 * 2;
      uint64_t dims[4] = {32, (uint64_t)J, (uint64_t)rows, (uint64_t)N};
      uint64_t str[4] = {2, 64, (uint64_t)RP * 2, (uint64_t)rows * RP * 2};
      uint32_t box[4] = {32, (uint32_t)p.tw, (uint32_t)halo_rows, 1};
      rc = encode_tmap(&p.tmap_a[0], 2, 4, const_cast<uint8_t*>(base), dims, str, box, 64);
      for (int i = 1; i < 4; ++i) p.tmap_a[i] = p.tmap_a[0];
      p.n_taps = 7;
      for (int r = 0; r < 7; ++r) p.taps[r] = ConvTap{(int16_t)r, 0, 0, (int16_t)r};
      const IoGeom g{N, Ho, Wo, 1, 4, 0, q};
      if (rc == GHND_OK) rc = finish_launch(&L, &d, 32, K, w_packed, 7, g, 32, K);
      if (rc == GHND_OK) {
        p.halo = 1;
        p.w_res_bytes = 7 * p.b_bytes;
        p.a_box_bytes = halo_rows * p.tw * 64;
        p.stage_bytes = (p.a_box_bytes + 1023) / 1024 * 1024;
        const int fixed = 2 * p.epi_bufs * kChunkBytes + (bias ? K * 4 : 0);
        int stages = (kSmemBudget - fixed - p.w_res_bytes) / p.stage_bytes;
        if (stages > kMaxStages) stages = kMaxStages;
        if (stages < 2) {
          set_error("stem conv: shared memory budget leaves %d halo stages", stages);
          rc = GHND_ERR_UNSUPPORTED;
        }
        p.n_stages = stages;
        L.smem = (size_t)p.n_stages * p.stage_bytes + (size_t)p.w_res_bytes + (size_t)fixed + 512;
      }
      if (rc == GHND_OK) plan->launches.push_back(L);
      continue;
    }
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
    if (rc == GHND_OK) rc = finish_launch(&L, &d, 32, K, w_packed, 7, g, 32);
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
    GHND_CUDA(launch_conv(L, (cudaStream_t)stream));
  }
  return GHND_OK;
}

void ghnd_stem_plan_destroy(ghnd_stem_plan_t* plan) { delete plan; }
}
