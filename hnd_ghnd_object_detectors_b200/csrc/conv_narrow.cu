// Narrow (bottleneck-side) convolutions and the stem weight gradient: HBM-bound SIMT kernels.
//   enc7  Conv2d(64->bch,k2,p1)  and  dec2  Conv2d(bch->64,k2,p0)   (resnet_layer.py:50,55)
// The wide side of each conv is an NHWC 16-bit tensor (one 128-byte pixel row per 64 channels,
// read/written as 8 lanes x 16 B so a warp touches 4 whole pixels = 512 contiguous bytes); the
// narrow side is the planar fp32 bottleneck tensor the quantizer works on.
#include <stdlib.h>

#include "common.cuh"

namespace ghnd {

static constexpr int kNarrowMaxC = 16;   // bottleneck channels supported (bch <= 15 in the reference)
static constexpr int kNarrowMaxTaps = 4;

struct NarrowTaps {
  int n_taps;
  int dh[kNarrowMaxTaps], dw[kNarrowMaxTaps];
};

// out[n][co][i][j] (planar fp32, co < CN) = sum_{tap,ci} wt[tap][ci][co] * in[n][i+dh][j+dw][ci]
// `in` NHWC 16-bit with CW wide channels; lanes = CW/8 threads cooperate on one pixel.
template <int CN_T>
__global__ void __launch_bounds__(256)
    narrow_out_kernel(const uint4* __restrict__ in, int in_fmt, const float* __restrict__ wt,
                      float* __restrict__ out, int N, int Hi, int Wi, int CW, int CN, int Ho, int Wo,
                      NarrowTaps taps) {
  extern __shared__ float s_w[];  // [tap][CW][CN]
  const int nw = taps.n_taps * CW * CN;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) s_w[i] = wt[i];
  __syncthreads();
  const int lanes = CW >> 3;
  const int cg = threadIdx.x % lanes;
  const int64_t npix = (int64_t)N * Ho * Wo;
  const int64_t pstep = (int64_t)gridDim.x * (blockDim.x / lanes);
  for (int64_t p0 = (int64_t)blockIdx.x * (blockDim.x / lanes); p0 < npix; p0 += pstep) {
    const int64_t p = p0 + threadIdx.x / lanes;
    const bool live = p < npix;
    int n = 0, i = 0, j = 0;
    if (live) {
      j = (int)(p % Wo);
      const int64_t t = p / Wo;
      i = (int)(t % Ho);
      n = (int)(t / Ho);
    }
    float acc[CN_T];
#pragma unroll
    for (int k = 0; k < CN_T; ++k) acc[k] = 0.f;
    if (live) {
      for (int t = 0; t < taps.n_taps; ++t) {
        const int h = i + taps.dh[t], w = j + taps.dw[t];
        if (h < 0 || h >= Hi || w < 0 || w >= Wi) continue;
        const uint4 v = __ldg(in + (((int64_t)n * Hi + h) * Wi + w) * lanes + cg);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
        const float* wrow = s_w + ((size_t)t * CW + cg * 8) * CN;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 f = unpack2(u[e], in_fmt);
#pragma unroll
          for (int k = 0; k < CN_T; ++k) {
            if (k < CN) {
              acc[k] = fmaf(f.x, wrow[(2 * e) * CN + k], acc[k]);
              acc[k] = fmaf(f.y, wrow[(2 * e + 1) * CN + k], acc[k]);
            }
          }
        }
      }
    }
    // reduce over the `lanes` threads of the pixel (lanes is a power of two <= 32)
#pragma unroll
    for (int k = 0; k < CN_T; ++k) {
      float a = acc[k];
      for (int o = lanes >> 1; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if (live && cg == 0 && k < CN) out[(((int64_t)n * CN + k) * Ho + i) * Wo + j] = a;
    }
  }
}

// out[n][i][j][co] (NHWC 16-bit, co < CW) = sum_{tap,ci<CN} wt[tap][ci][co] * f(in[n][ci][i+dh][j+dw])
// f(v) = pre ? (relu ? max(v*sc+sh,0) : v*sc+sh) : v ; padding is applied after f (zero taps).
__global__ void __launch_bounds__(256)
    narrow_in_kernel(const float* __restrict__ in, const float* __restrict__ pre, int pre_relu,
                     const float* __restrict__ wt, uint4* __restrict__ out, int out_fmt, int N, int Hi,
                     int Wi, int CN, int CW, int Ho, int Wo, NarrowTaps taps) {
  extern __shared__ float s_w[];  // [tap][CN][CW]
  const int nw = taps.n_taps * CN * CW;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) s_w[i] = wt[i];
  __syncthreads();
  const int lanes = CW >> 3;
  const int cg = threadIdx.x % lanes;
  const int64_t npix = (int64_t)N * Ho * Wo;
  const int64_t pstep = (int64_t)gridDim.x * (blockDim.x / lanes);
  for (int64_t p = (int64_t)blockIdx.x * (blockDim.x / lanes) + threadIdx.x / lanes; p < npix;
       p += pstep) {
    const int j = (int)(p % Wo);
    const int64_t t0 = p / Wo;
    const int i = (int)(t0 % Ho);
    const int n = (int)(t0 / Ho);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int t = 0; t < taps.n_taps; ++t) {
      const int h = i + taps.dh[t], w = j + taps.dw[t];
      if (h < 0 || h >= Hi || w < 0 || w >= Wi) continue;
      for (int c = 0; c < CN; ++c) {
        float v = __ldg(in + (((int64_t)n * CN + c) * Hi + h) * Wi + w);
        if (pre != nullptr) {
          v = fmaf(v, __ldg(pre + c), __ldg(pre + CN + c));
          if (pre_relu) v = fmaxf(v, 0.f);
        }
        const float* wrow = s_w + ((size_t)t * CN + c) * CW + cg * 8;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(v, wrow[e], acc[e]);
      }
    }
    uint4 o;
    o.x = pack2(acc[0], acc[1], out_fmt);
    o.y = pack2(acc[2], acc[3], out_fmt);
    o.z = pack2(acc[4], acc[5], out_fmt);
    o.w = pack2(acc[6], acc[7], out_fmt);
    out[p * lanes + cg] = o;
  }
}

// acc[ca][tap][cb] += f(a[n][ca][i][j]) * b[n][i+dh][j+dw][cb]   for ca in [ca0, ca0+3)
__global__ void __launch_bounds__(256)
    wgrad_narrow_kernel(const float* __restrict__ a, const float* __restrict__ pre, int pre_relu,
                        const uint4* __restrict__ b, int b_fmt, float* __restrict__ accum, int N,
                        int Ha, int Wa, int Ca, int ca0, int Hb, int Wb, int Cb, NarrowTaps taps) {
  const int lanes = Cb >> 3;
  const int cg = threadIdx.x % lanes;
  const int64_t npix = (int64_t)N * Ha * Wa;
  const int64_t pstep = (int64_t)gridDim.x * (blockDim.x / lanes);
  float acc[3][kNarrowMaxTaps][8];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int t = 0; t < kNarrowMaxTaps; ++t)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[c][t][e] = 0.f;
  for (int64_t p = (int64_t)blockIdx.x * (blockDim.x / lanes) + threadIdx.x / lanes; p < npix;
       p += pstep) {
    const int j = (int)(p % Wa);
    const int64_t t0 = p / Wa;
    const int i = (int)(t0 % Ha);
    const int n = (int)(t0 / Ha);
    float av[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int ca = ca0 + c;
      float v = 0.f;
      if (ca < Ca) {
        v = __ldg(a + (((int64_t)n * Ca + ca) * Ha + i) * Wa + j);
        if (pre != nullptr) {
          v = fmaf(v, __ldg(pre + ca), __ldg(pre + Ca + ca));
          if (pre_relu) v = fmaxf(v, 0.f);
        }
      }
      av[c] = v;
    }
#pragma unroll
    for (int t = 0; t < kNarrowMaxTaps; ++t) {
      if (t >= taps.n_taps) break;
      const int h = i + taps.dh[t], w = j + taps.dw[t];
      if (h < 0 || h >= Hb || w < 0 || w >= Wb) continue;
      const uint4 v = __ldg(b + (((int64_t)n * Hb + h) * Wb + w) * lanes + cg);
      const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack2(u[e], b_fmt);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          acc[c][t][2 * e] = fmaf(av[c], f.x, acc[c][t][2 * e]);
          acc[c][t][2 * e + 1] = fmaf(av[c], f.y, acc[c][t][2 * e + 1]);
        }
      }
    }
  }
  // fold the pixel lanes of the warp that share a channel group, then the warps of the block in
  // shared memory, then ONE global atomic per value and block
  __shared__ float s_acc[3 * kNarrowMaxTaps * 64];
  for (int i = threadIdx.x; i < 3 * kNarrowMaxTaps * 64; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int t = 0; t < kNarrowMaxTaps; ++t)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = acc[c][t][e];
        for (int o = 16; o >= lanes; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) < lanes && Cb <= 64)
          atomicAdd(&s_acc[(c * kNarrowMaxTaps + t) * 64 + cg * 8 + e], v);
        else if ((threadIdx.x & 31) < lanes && t < taps.n_taps && ca0 + c < Ca)
          atomicAdd(accum + ((size_t)(ca0 + c) * taps.n_taps + t) * Cb + cg * 8 + e, v);
      }
  __syncthreads();
  if (Cb <= 64) {
    for (int i = threadIdx.x; i < 3 * kNarrowMaxTaps * 64; i += blockDim.x) {
      const int cb = i & 63, t = (i >> 6) % kNarrowMaxTaps, c = i / (64 * kNarrowMaxTaps);
      if (cb < Cb && t < taps.n_taps && ca0 + c < Ca)
        atomicAdd(accum + ((size_t)(ca0 + c) * taps.n_taps + t) * Cb + cb, s_acc[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core variants for the 64-channel wide side (the only width the model uses).  The FLOPs
// are negligible; what the SIMT kernels above lose is issue slots (fp32 FMA + LDS + 64-bit index
// math per pixel), so these kernels map the small GEMMs onto mma.sync.m16n8k16 and are left with
// nothing but the streaming loads/stores:
//   narrow_out : [16 pixels x (tap,64 ch)] x [(tap,64 ch) x 8*NT narrow ch]   (A straight from NHWC)
//   narrow_in  : [16 pixels x (narrow ch,tap)] x [(narrow ch,tap) x 64 ch]    (D straight to NHWC)
// The GEMM-K / GEMM-N orderings are free, so they are chosen such that every lane's fragment
// registers are whole 16-byte pieces of a pixel row (no shuffles, no shared-memory staging).
// fp32 operands (weights; the narrow activations) enter as hi + lo 16-bit pairs, the lo part
// scaled by 2^11 (f16) / 2^8 (bf16) and accumulated separately, which keeps ~fp32 accuracy.
// Fragment layout of mma.m16n8k16 (g = lane/4, t = lane%4):
//   A regs: {row g, k 2t..2t+1} {row g+8, k 2t..} {row g, k 2t+8..} {row g+8, k 2t+8..}
//   B regs: {k 2t..2t+1, col g} {k 2t+8..2t+9, col g}      D: {row g, col 2t..2t+1} {row g+8, ..}
// ------------------------------------------------------------------------------------------------
template <int FMT>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                         uint32_t b1) {
  if (FMT == GHND_F16) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
        "{%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
template <int FMT>
struct LoScale {
  static constexpr float kUp = FMT == GHND_F16 ? 2048.f : 256.f;
  static constexpr float kDown = FMT == GHND_F16 ? (1.f / 2048.f) : (1.f / 256.f);
};
// v ~= hi + lo / kUp with hi, lo 16-bit
template <int FMT>
__device__ __forceinline__ void split16(float v, uint32_t& hi, uint32_t& lo) {
  const uint16_t h = float_to_h16(v, FMT);
  hi = h;
  lo = float_to_h16((v - h16_to_float(h, FMT)) * LoScale<FMT>::kUp, FMT);
}

// out[n][co][i][j] (planar fp32) = sum_{tap,c} wt[tap][c][co] * in[n][i+dh][j+dw][c], 64 wide channels.
// A CTA works on blocks of 8 output rows x 16 output columns: warp w owns row 8*ib + w, its 16
// pixels are one m16 tile; lane (g,t) loads, for pixels g and g+8 and every tap, the two 16-byte
// pieces t and t+4 of the pixel row.  Neighbouring rows / columns of a k2 conv share three of
// their four taps: with the 2-D blocking those repeats hit L1 (the flat pixel order sent every
// tap to L2: 4x the tensor in L2 traffic, ~25 % of HBM speed).  GEMM-K step (tap, s): lane t's
// logical k {2t,2t+1 | 2t+8,2t+9} are channels base..base+3, base = 32*(s>>1) + 8t + 4*(s&1).
// kMinMax: also publish the CTA's min / max of the stored values (NaN-propagating like torch.min)
// as partial[2*blockIdx.x .. +1] -- the quantizer's first pass, fused into the producer.
template <int FMT, int NT, bool kMinMax>
__global__ void __launch_bounds__(256, 2)
    narrow_out_mma_kernel(const uint4* __restrict__ in, const float* __restrict__ wt,
                          float* __restrict__ out, int N, int Hi, int Wi, int CN, int Ho, int Wo,
                          NarrowTaps taps, float* __restrict__ partial) {
  __shared__ uint4 s_b[16 * NT * 32];  // [k-step][n-tile][lane] = {b0 hi, b1 hi, b0 lo, b1 lo}
  __shared__ float s_mm[16];
  for (int idx = threadIdx.x; idx < 16 * NT * 32; idx += blockDim.x) {
    const int l = idx & 31, nt = (idx >> 5) % NT, ks = idx / (32 * NT);
    const int g = l >> 2, t = l & 3;
    const int tap = ks >> 2, s = ks & 3;
    const int base = (s >> 1) * 32 + 8 * t + (s & 1) * 4;
    const int co = nt * 8 + g;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float w = (co < CN && tap < taps.n_taps) ? wt[((size_t)tap * 64 + base + e) * CN + co] : 0.f;
      split16<FMT>(w, hi[e], lo[e]);
    }
    s_b[idx] = make_uint4(hi[0] | (hi[1] << 16), hi[2] | (hi[3] << 16), lo[0] | (lo[1] << 16),
                          lo[2] | (lo[3] << 16));
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const uint32_t JB = (uint32_t)(Wo + 15) >> 4, IB = (uint32_t)(Ho + 7) >> 3;
  const uint32_t ctiles = (uint32_t)N * IB * JB;
  float mn = INFINITY, mx = -INFINITY;
  bool has_nan = false;
  for (uint32_t ct = blockIdx.x; ct < ctiles; ct += gridDim.x) {
    const uint32_t q = ct / JB, jb = ct - q * JB;
    const uint32_t n = q / IB, ib = q - n * IB;
    const int i = (int)(ib * 8) + warp;
    const int j0 = (int)(jb * 16) + g;
    const bool row_ok = i < Ho;
    uint4 v[kNarrowMaxTaps][2][2];
#pragma unroll
    for (int tap = 0; tap < kNarrowMaxTaps; ++tap) {
      const int h = i + taps.dh[tap];
      const bool hok = row_ok && tap < taps.n_taps && h >= 0 && h < Hi;
      const uint32_t rowbase = ((uint32_t)((int)n * Hi + h) * (uint32_t)Wi) * 8u + (uint32_t)t;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int j = j0 + 8 * r, w = j + taps.dw[tap];
        const bool ok = hok && j < Wo && w >= 0 && w < Wi;
        const uint4* src = in + (rowbase + (uint32_t)w * 8u);
        v[tap][r][0] = ok ? __ldg(src) : make_uint4(0, 0, 0, 0);
        v[tap][r][1] = ok ? __ldg(src + 4) : make_uint4(0, 0, 0, 0);
      }
    }
    float dhi[NT][4], dlo[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dhi[nt][e] = dlo[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 16; ++ks) {
      const int tap = ks >> 2, s = ks & 3, hf = s >> 1;
      uint32_t a[4];
      if ((s & 1) == 0) {
        a[0] = v[tap][0][hf].x;
        a[1] = v[tap][1][hf].x;
        a[2] = v[tap][0][hf].y;
        a[3] = v[tap][1][hf].y;
      } else {
        a[0] = v[tap][0][hf].z;
        a[1] = v[tap][1][hf].z;
        a[2] = v[tap][0][hf].w;
        a[3] = v[tap][1][hf].w;
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const uint4 b = s_b[(ks * NT + nt) * 32 + lane];
        mma16816<FMT>(dhi[nt], a, b.x, b.y);
        mma16816<FMT>(dlo[nt], a, b.z, b.w);
      }
    }
    if (row_ok) {
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int co = nt * 8 + 2 * t + e;
          if (co < CN) {
            float* dst = out + ((size_t)((int)n * CN + co) * Ho + i) * Wo;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              const int j = j0 + 8 * r;
              if (j < Wo) {
                const float val = dhi[nt][2 * r + e] + dlo[nt][2 * r + e] * LoScale<FMT>::kDown;
                dst[j] = val;
                if (kMinMax) {
                  has_nan |= val != val;
                  mn = fminf(mn, val);
                  mx = fmaxf(mx, val);
                }
              }
            }
          }
        }
    }
  }
  if (kMinMax) {
    if (has_nan) mn = mx = __int_as_float(0x7fc00000);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float omn = __shfl_xor_sync(0xffffffffu, mn, o), omx = __shfl_xor_sync(0xffffffffu, mx, o);
      mn = (mn != mn) ? mn : ((omn != omn) ? omn : fminf(mn, omn));
      mx = (mx != mx) ? mx : ((omx != omx) ? omx : fmaxf(mx, omx));
    }
    if (lane == 0) {
      s_mm[2 * warp] = mn;
      s_mm[2 * warp + 1] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) {
        const float omn = s_mm[2 * w], omx = s_mm[2 * w + 1];
        mn = (mn != mn) ? mn : ((omn != omn) ? omn : fminf(mn, omn));
        mx = (mx != mx) ? mx : ((omx != omx) ? omx : fmaxf(mx, omx));
      }
      partial[2 * blockIdx.x] = mn;
      partial[2 * blockIdx.x + 1] = mx;
    }
  }
}

// The same conv with the wide tensor staged through shared memory by TMA (sm_100a):
// a CTA works on blocks of 8 output rows x 16 output columns; ONE tensor-map box of (8 + dh span) x (16 + dw
// span) pixels x 64 channels (128B-swizzled, out-of-bounds = the conv's zero padding) feeds all taps of all
// eight warps, double buffered so that the box of block k+1 is in flight while block k is multiplied.  A
// fragments come from `ldmatrix.x4` (conflict-free on the swizzled rows): 16 ldmatrix + 32 mma.sync per
// 16-pixel row instead of 64 predicated 16-byte global loads (ncu of the direct-load kernel: 41 % of the issue
// slots, 4.7 warps stalled on the long scoreboard per issue, 1.7 TB/s).  GEMM-K step (tap, s) = channels
// 16s .. 16s+15 of that tap in natural order.
static constexpr int kNoStages = 4;
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&a)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
               : "r"(addr));
}
// Weights keep fp32 accuracy as hi + lo/kUp 16-bit halves.  Both halves of a channel sit in NEIGHBOURING columns
// of one n8 tile (column 2c = hi, 2c+1 = lo of channel 4*nt + c), so ONE mma.sync per k-step yields both partial
// sums in the same thread (the mma.sync rate, not memory, paced the first TMA version: 32 HMMA per 16 pixels
// for 3 useful output channels).  NT = ceil(CN / 4) n-tiles.
template <int FMT, int NT, bool kMinMax>
__global__ void __launch_bounds__(256, 2)
    narrow_out_tma_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ w_oihw, int wmode,
                          float* __restrict__ out, int N, int CN, int Ho, int Wo, NarrowTaps taps, int bw,
                          int dh0, int dw0, int tile_bytes, int box_bytes, float* __restrict__ partial) {
  extern __shared__ __align__(1024) uint8_t dsm[];  // [kNoStages][tile_bytes]
  __shared__ uint2 s_b[16 * NT * 32];  // [k-step][n-tile][lane] = {b0, b1}
  __shared__ float s_mm[16];
  __shared__ uint64_t s_bar[kNoStages];
  if (smem_u32(dsm) & 1023u) __trap();
  for (int idx = threadIdx.x; idx < 16 * NT * 32; idx += blockDim.x) {
    const int l = idx & 31, nt = (idx >> 5) % NT, ks = idx / (32 * NT);
    const int g = l >> 2, t = l & 3;
    const int tap = ks >> 2, sx = ks & 3;
    const int co = nt * 4 + (g >> 1), part = g & 1;
    uint32_t h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {  // b0 = k {2t, 2t+1}, b1 = k {2t+8, 2t+9}
      const int c = 16 * sx + 2 * t + (e & 1) + (e >> 1) * 8;
      // straight from the OIHW tensor (no weight-table launch): wmode 0 = this conv's own weight [CN][64][R][S],
      // wmode 1 = data gradient of the bch -> 64 conv, weight [64][CN][R][S]
      const int wi = (wmode == 0 ? (co * 64 + c) : (c * CN + co)) * taps.n_taps + tap;
      const float w = (co < CN && tap < taps.n_taps) ? __ldg(w_oihw + wi) : 0.f;
      uint32_t hi, lo;
      split16<FMT>(w, hi, lo);
      h[e] = part ? lo : hi;
    }
    s_b[idx] = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
  }
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap);
    for (int b = 0; b < kNoStages; ++b) mbar_init(&s_bar[b], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const uint32_t JB = (uint32_t)(Wo + 15) >> 4, IB = (uint32_t)(Ho + 7) >> 3;
  const uint32_t ctiles = (uint32_t)N * IB * JB;
  auto issue = [&](uint32_t ct, int buf) {
    const uint32_t q = ct / JB, jb = ct - q * JB;
    const uint32_t n = q / IB, ib = q - n * IB;
    mbar_arrive_expect_tx(&s_bar[buf], (uint32_t)box_bytes);
    tma_load_4d(dsm + (size_t)buf * tile_bytes, &tmap, &s_bar[buf], 0, (int)(jb * 16) + dw0, (int)(ib * 8) + dh0,
                (int)n);
  };
  // kNoStages - 1 boxes in flight per CTA (with one box ahead the kernel ran at the TMA round trip: 1.7 us
  // per block and CTA, 2.9 TB/s)
  if (threadIdx.x == 0)
    for (int b = 0; b < kNoStages - 1; ++b)
      if (blockIdx.x + (uint32_t)b * gridDim.x < ctiles) issue(blockIdx.x + (uint32_t)b * gridDim.x, b);
  // per-lane ldmatrix row: matrix mi = lane >> 3 (pixels +8 for odd mi, channels +8 for mi >= 2), row lane & 7.
  // The byte offset of every k-step's row inside a buffer depends on (warp, lane, tap, s) only: computed once
  // (the first version rebuilt them per block: ~270 of its 320 instructions per 16-pixel row were index
  // arithmetic, 38 % of the issue slots).  With one n-tile the B fragments live in registers, too.
  const int lm_px = ((lane >> 3) & 1) * 8 + (lane & 7), lm_ch = lane >> 4;
  uint32_t aoff[16];
  uint2 breg[NT == 1 ? 16 : 1];
#pragma unroll
  for (int ks = 0; ks < 16; ++ks) {
    const int tap = ks >> 2, sx = ks & 3;
    const int tdh = tap < taps.n_taps ? taps.dh[tap] : taps.dh[0], tdw = tap < taps.n_taps ? taps.dw[tap] : taps.dw[0];
    const int px = (warp + tdh - dh0) * bw + (tdw - dw0) + lm_px;
    aoff[ks] = (uint32_t)(px * 128 + (((2 * sx + lm_ch) ^ (px & 7)) << 4));
    if (NT == 1) breg[ks] = s_b[ks * 32 + lane];
  }
  float mn = INFINITY, mx = -INFINITY;
  bool has_nan = false;
  const int n_taps = taps.n_taps;
  int k = 0;
  for (uint32_t ct = blockIdx.x; ct < ctiles; ct += gridDim.x, ++k) {
    const int buf = k % kNoStages;
    // the buffer of block k - 1 was released by the barrier that ended it: refill it with block k + kNoStages - 1
    if (threadIdx.x == 0 && ct + (uint32_t)(kNoStages - 1) * gridDim.x < ctiles)
      issue(ct + (uint32_t)(kNoStages - 1) * gridDim.x, (k + kNoStages - 1) % kNoStages);
    const uint32_t q = ct / JB, jb = ct - q * JB;
    const uint32_t n = q / IB, ib = q - n * IB;
    const int i = (int)(ib * 8) + warp;
    const int j0 = (int)(jb * 16) + g;
    const bool row_ok = i < Ho;
    mbar_wait(&s_bar[buf], ((uint32_t)(k / kNoStages)) & 1u);
    const uint32_t tile = smem_u32(dsm) + (uint32_t)(buf * tile_bytes);
    float d[NT][4], d2[NT][4];  // two accumulator sets: even / odd k-steps (halves the mma dependency chain)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) d[nt][e] = d2[nt][e] = 0.f;
    if (row_ok) {
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) {
        if ((ks >> 2) < n_taps) {  // uniform
          uint32_t a[4];
          ldmatrix_x4(a, tile + aoff[ks]);
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const uint2 b = NT == 1 ? breg[ks] : s_b[(ks * NT + nt) * 32 + lane];
            if (ks & 1) mma16816<FMT>(d2[nt], a, b.x, b.y);
            else mma16816<FMT>(d[nt], a, b.x, b.y);
          }
        }
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) d[nt][e] += d2[nt][e];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int co = nt * 4 + t;  // this thread's columns 2t / 2t+1 = hi / lo sums of channel co
        if (co < CN) {
          float* dst = out + ((size_t)((int)n * CN + co) * Ho + i) * Wo;
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const int j = j0 + 8 * r;
            if (j < Wo) {
              const float val = d[nt][2 * r] + d[nt][2 * r + 1] * LoScale<FMT>::kDown;
              dst[j] = val;
              if (kMinMax) {
                has_nan |= val != val;
                mn = fminf(mn, val);
                mx = fmaxf(mx, val);
              }
            }
          }
        }
      }
    }
    fence_proxy_async();  // generic-proxy reads of this buffer before the TMA refill two blocks later
    __syncthreads();
  }
  if (kMinMax) {
    if (has_nan) mn = mx = __int_as_float(0x7fc00000);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float omn = __shfl_xor_sync(0xffffffffu, mn, o), omx = __shfl_xor_sync(0xffffffffu, mx, o);
      mn = (mn != mn) ? mn : ((omn != omn) ? omn : fminf(mn, omn));
      mx = (mx != mx) ? mx : ((omx != omx) ? omx : fmaxf(mx, omx));
    }
    if (lane == 0) {
      s_mm[2 * warp] = mn;
      s_mm[2 * warp + 1] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < 8; ++w) {
        const float omn = s_mm[2 * w], omx = s_mm[2 * w + 1];
        mn = (mn != mn) ? mn : ((omn != omn) ? omn : fminf(mn, omn));
        mx = (mx != mx) ? mx : ((omx != omx) ? omx : fmaxf(mx, omx));
      }
      partial[2 * blockIdx.x] = mn;
      partial[2 * blockIdx.x + 1] = mx;
    }
  }
}

// out[n][i][j][co] (NHWC 16-bit, 64 ch) = sum_{tap,ci<CN} wt[tap][ci][co] * f(in[n][ci][i+dh][j+dw]).
// GEMM-K index k = ci*4 + tap (KS steps of 16); D column (n-tile nn, col 2t+e) is channel
// 8t + 2nn + e (nn < 4) or 32 + 8t + 2(nn-4) + e, so lane t stores pieces t and t+4 of a pixel row.
template <int FMT, int KS>
__global__ void __launch_bounds__(256, 2)
    narrow_in_mma_kernel(const float* __restrict__ in, const float* __restrict__ pre, int pre_relu,
                         const float* __restrict__ wt, uint4* __restrict__ out, int N, int Hi, int Wi,
                         int CN, int Ho, int Wo, NarrowTaps taps) {
  __shared__ uint4 s_b[KS * 8 * 32];  // [k-step][n-tile][lane] = {b0 hi, b1 hi, b0 lo, b1 lo}
  for (int idx = threadIdx.x; idx < KS * 8 * 32; idx += blockDim.x) {
    const int l = idx & 31, nn = (idx >> 5) & 7, ks = idx >> 8;
    const int g = l >> 2, t = l & 3;
    const int co = nn < 4 ? 8 * (g >> 1) + 2 * nn + (g & 1) : 32 + 8 * (g >> 1) + 2 * (nn - 4) + (g & 1);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 16 * ks + 2 * t + (e & 1) + (e >> 1) * 8;
      const int ci = k >> 2, tap = k & 3;
      const float w = (ci < CN && tap < taps.n_taps) ? wt[((size_t)tap * CN + ci) * 64 + co] : 0.f;
      split16<FMT>(w, hi[e], lo[e]);
    }
    s_b[idx] = make_uint4(hi[0] | (hi[1] << 16), hi[2] | (hi[3] << 16), lo[0] | (lo[1] << 16),
                          lo[2] | (lo[3] << 16));
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // this lane's two taps (e = 0, 1) and its per-k-step channels / pre-transform
  int tdh[2], tdw[2];
  bool tok[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int tap = 2 * (t & 1) + e;
    tok[e] = tap < taps.n_taps;
    tdh[e] = tap == 0 ? taps.dh[0] : tap == 1 ? taps.dh[1] : tap == 2 ? taps.dh[2] : taps.dh[3];
    tdw[e] = tap == 0 ? taps.dw[0] : tap == 1 ? taps.dw[1] : tap == 2 ? taps.dw[2] : taps.dw[3];
  }
  float psc[KS][2], psh[KS][2];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int cs = 0; cs < 2; ++cs) {
      const int ci = 4 * ks + 2 * cs + (t >> 1);
      psc[ks][cs] = (pre != nullptr && ci < CN) ? __ldg(pre + ci) : 1.f;
      psh[ks][cs] = (pre != nullptr && ci < CN) ? __ldg(pre + CN + ci) : 0.f;
    }
  const bool has_pre = pre != nullptr;
  const uint32_t npix = (uint32_t)N * Ho * Wo;
  const uint32_t tiles = (npix + 15) >> 4;
  for (uint32_t tile = blockIdx.x * 8 + (threadIdx.x >> 5); tile < tiles; tile += gridDim.x * 8) {
    bool pv[2];
    int pn[2], pi[2], pj[2];
    uint32_t pp[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const uint32_t p = tile * 16 + g + 8 * r;
      pv[r] = p < npix;
      pp[r] = pv[r] ? p : 0;
      const uint32_t q = pp[r] / (uint32_t)Wo;
      pj[r] = (int)(pp[r] - q * Wo);
      pn[r] = (int)(q / (uint32_t)Ho);
      pi[r] = (int)(q - (uint32_t)pn[r] * Ho);
    }
    float dhi[8][4], dlo[8][4];
#pragma unroll
    for (int nn = 0; nn < 8; ++nn)
#pragma unroll
      for (int e = 0; e < 4; ++e) dhi[nn][e] = dlo[nn][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t ahi[4], alo[4];
#pragma unroll
      for (int cs = 0; cs < 2; ++cs) {
        const int ci = 4 * ks + 2 * cs + (t >> 1);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          uint32_t h2[2], l2[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int h = pi[r] + tdh[e], w = pj[r] + tdw[e];
            const bool ok = pv[r] && tok[e] && ci < CN && h >= 0 && h < Hi && w >= 0 && w < Wi;
            float val = 0.f;
            if (ok) {
              val = __ldg(in + (((size_t)pn[r] * CN + ci) * Hi + h) * Wi + w);
              if (has_pre) {
                val = fmaf(val, psc[ks][cs], psh[ks][cs]);
                if (pre_relu) val = fmaxf(val, 0.f);
              }
            }
            split16<FMT>(val, h2[e], l2[e]);
          }
          ahi[r + 2 * cs] = h2[0] | (h2[1] << 16);
          alo[r + 2 * cs] = l2[0] | (l2[1] << 16);
        }
      }
#pragma unroll
      for (int nn = 0; nn < 8; ++nn) {
        const uint4 b = s_b[(ks * 8 + nn) * 32 + lane];
        mma16816<FMT>(dhi[nn], ahi, b.x, b.y);
        mma16816<FMT>(dlo[nn], alo, b.x, b.y);
        mma16816<FMT>(dlo[nn], ahi, b.z, b.w);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      uint32_t o[8];
#pragma unroll
      for (int nn = 0; nn < 8; ++nn)
        o[nn] = pack2_t<FMT>(dhi[nn][2 * r] + dlo[nn][2 * r] * LoScale<FMT>::kDown,
                             dhi[nn][2 * r + 1] + dlo[nn][2 * r + 1] * LoScale<FMT>::kDown);
      if (pv[r]) {
        uint4* dst = out + (size_t)pp[r] * 8 + t;
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[4] = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
  }
}

// ---- narrow-side halo of an 8 x 16 block through a cp.async ring (planar fp32 rows are not 16-byte aligned, so
// no TMA): every thread copies up to three 4-byte elements per block, kHaloDepth blocks ahead; out-of-image
// elements are zero-filled by the copy (src-size 0).  A lane then builds its MMA registers from the raw values:
// f = BN/ReLU prologue, padding stays zero AFTER f (hence the validity flags), fp32 -> hi + lo/kUp 16-bit halves.
static constexpr int kHaloDepth = 3;
__device__ __forceinline__ void cp_async4_zfill(void* smem, const void* gmem, bool valid) {
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// one packed (left, right) register pair: hi halves and lo halves of f(v0), f(v1)
template <int FMT>
__device__ __forceinline__ void narrow_pair(float v0, float v1, bool ok0, bool ok1, float sc, float sh, bool has_pre,
                                            int pre_relu, uint32_t& hi, uint32_t& lo) {
  if (has_pre) {
    v0 = fmaf(v0, sc, sh);
    v1 = fmaf(v1, sc, sh);
    if (pre_relu) {
      v0 = fmaxf(v0, 0.f);
      v1 = fmaxf(v1, 0.f);
    }
  }
  v0 = ok0 ? v0 : 0.f;
  v1 = ok1 ? v1 : 0.f;
  uint32_t h0, l0, h1, l1;
  split16<FMT>(v0, h0, l0);
  split16<FMT>(v1, h1, l1);
  hi = h0 | (h1 << 16);
  lo = l0 | (l1 << 16);
}

// The same conv for bch <= 4 and horizontally paired taps (k2 kernels), restructured like narrow_out_tma:
// a CTA works on blocks of 8 output rows x 16 columns (warp = row), the narrow input halo of a block is
// transformed (BN + ReLU prologue), split into hi / lo 16-bit halves and staged ONCE in shared memory -- every
// word holds the PAIR (column c, c+1), and GEMM-K is ordered (channel, tap pair, left/right tap) so that an A
// register is one LDS -- instead of 8 predicated scalar global loads + pre-transform + split per k-step and
// lane; B fragments live in registers; no per-pixel divisions.  ~170 instead of ~450 instructions per 16 pixels.
template <int FMT>
__global__ void __launch_bounds__(256, 2)
    narrow_in_pair_kernel(const float* __restrict__ in, const float* __restrict__ pre, int pre_relu,
                          const float* __restrict__ w_oihw, int wmode, uint4* __restrict__ out, int N, int Hi,
                          int Wi, int CN, int Ho, int Wo, NarrowTaps taps, int dhmin, int dwmin, int HR, int HC) {
  extern __shared__ float s_raw[];  // [kHaloDepth + 1][4 ci][HR][HC + 1] raw fp32 halo ring
  const int HCp = HC + 1, plane = 4 * HR * HCp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  // tap of (pair p, half h): the pair's taps share dh and are one column apart; half 0 = the left one
  auto tap_of = [&](int p, int h) { return taps.dw[2 * p] < taps.dw[2 * p + 1] ? 2 * p + h : 2 * p + 1 - h; };
  // B fragments (weights, hi/lo): k = 4 ci + 2 p + h; column g of n-tile nn = channel 8 (g>>1) + 2 nn + (g&1)
  // (+32 for nn >= 4), so that lane t ends up with channels 8t .. 8t+7 (and 32 + the same) of a pixel
  uint4 breg[8];
#pragma unroll
  for (int nn = 0; nn < 8; ++nn) {
    const int co = nn < 4 ? 8 * (g >> 1) + 2 * nn + (g & 1) : 32 + 8 * (g >> 1) + 2 * (nn - 4) + (g & 1);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int kk = 2 * t + (e & 1) + (e >> 1) * 8;
      const int ci = kk >> 2, p = (kk >> 1) & 1, h = kk & 1;
      float w = 0.f;
      if (ci < CN && 2 * p + 1 < taps.n_taps) {
        const int tp = tap_of(p, h);
        // wmode 0: forward weight [64][CN][R][S]; wmode 1: data gradient of the 64 -> bch conv, weight [CN][64][R][S]
        w = __ldg(w_oihw + (wmode == 0 ? (co * CN + ci) : (ci * 64 + co)) * taps.n_taps + tp);
      }
      split16<FMT>(w, hi[e], lo[e]);
    }
    breg[nn] = make_uint4(hi[0] | (hi[1] << 16), hi[2] | (hi[3] << 16), lo[0] | (lo[1] << 16), lo[2] | (lo[3] << 16));
  }
  const uint32_t JB = (uint32_t)(Wo + 15) >> 4, IB = (uint32_t)(Ho + 7) >> 3;
  const uint32_t ctiles = (uint32_t)N * IB * JB;
  auto decode = [&](uint32_t ct, int& n, int& i0, int& j0) {
    const uint32_t q = ct / JB, jb = ct - q * JB;
    const uint32_t nn = q / IB, ib = q - nn * IB;
    n = (int)nn;
    i0 = (int)(ib * 8);
    j0 = (int)(jb * 16);
  };
  const size_t plane_in = (size_t)Hi * Wi;
  // element w of a halo: (ci, hr, hc) -> in[n][ci][i0 + dhmin + hr][j0 + dwmin + hc]
  int e_ci[3], e_hr[3], e_hc[3];
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int w = threadIdx.x + 256 * u;
    e_hc[u] = w % HCp;
    e_hr[u] = (w / HCp) % HR;
    e_ci[u] = w / (HCp * HR);
  }
  auto issue_halo = [&](uint32_t ct, int slot) {
    if (ct < ctiles) {
      int n, i0, j0;
      decode(ct, n, i0, j0);
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int w = threadIdx.x + 256 * u;
        if (w < plane) {
          const int ih = i0 + dhmin + e_hr[u], iw = j0 + dwmin + e_hc[u];
          const bool ok = e_ci[u] < CN && ih >= 0 && ih < Hi && iw >= 0 && iw < Wi;
          const float* src = ok ? in + ((size_t)n * CN + e_ci[u]) * plane_in + (size_t)ih * Wi + iw : in;
          cp_async4_zfill(s_raw + (size_t)slot * plane + w, src, ok);
        }
      }
    }
    cp_async_commit_group();
  };
  // this lane's A values: k = 2t, 2t+1 -> channel t >> 1, pair t & 1 (k + 8: channel + 2); pixels g and g + 8
  const bool has_pre = pre != nullptr;
  const int pp = t & 1, cA = t >> 1, cB = cA + 2;
  const bool pair_ok = 2 * pp + 1 < taps.n_taps;
  const int pdh = pair_ok ? taps.dh[2 * pp] : 0;
  const int pdw = pair_ok ? min(taps.dw[2 * pp], taps.dw[2 * pp + 1]) : 0;
  const int ar = warp + pdh - dhmin, ac = g + pdw - dwmin;  // halo row / first column
  const int awA = (cA * HR + ar) * HCp + ac, awB = (cB * HR + ar) * HCp + ac;
  const bool okA = pair_ok && cA < CN, okB = pair_ok && cB < CN;
  const float scA = (has_pre && cA < CN) ? __ldg(pre + cA) : 1.f, shA = (has_pre && cA < CN) ? __ldg(pre + CN + cA) : 0.f;
  const float scB = (has_pre && cB < CN) ? __ldg(pre + cB) : 1.f, shB = (has_pre && cB < CN) ? __ldg(pre + CN + cB) : 0.f;
#pragma unroll
  for (int d = 0; d < kHaloDepth; ++d) issue_halo(blockIdx.x + (uint32_t)d * gridDim.x, d);
  int k = 0;
  for (uint32_t ct = blockIdx.x; ct < ctiles; ct += gridDim.x, ++k) {
    cp_async_wait_group<kHaloDepth - 1>();
    __syncthreads();  // block k's halo visible; everybody is done with block k - 1 (its slot is refilled next)
    issue_halo(ct + (uint32_t)kHaloDepth * gridDim.x, (k + kHaloDepth) % (kHaloDepth + 1));
    int n, i0, j0;
    decode(ct, n, i0, j0);
    const int i = i0 + warp;
    const float* raw = s_raw + (size_t)(k % (kHaloDepth + 1)) * plane;
    const int ih = i + pdh, iw = j0 + g + pdw;  // image position of the left value of pixel g's pair
    const bool rok = ih >= 0 && ih < Hi;
    bool cok[4];
    cok[0] = rok && iw >= 0 && iw < Wi;
    cok[1] = rok && iw + 1 >= 0 && iw + 1 < Wi;
    cok[2] = rok && iw + 8 >= 0 && iw + 8 < Wi;
    cok[3] = rok && iw + 9 >= 0 && iw + 9 < Wi;
    uint32_t ahi[4], alo[4];
    narrow_pair<FMT>(raw[awA], raw[awA + 1], okA && cok[0], okA && cok[1], scA, shA, has_pre, pre_relu, ahi[0], alo[0]);
    narrow_pair<FMT>(raw[awA + 8], raw[awA + 9], okA && cok[2], okA && cok[3], scA, shA, has_pre, pre_relu, ahi[1], alo[1]);
    narrow_pair<FMT>(raw[awB], raw[awB + 1], okB && cok[0], okB && cok[1], scB, shB, has_pre, pre_relu, ahi[2], alo[2]);
    narrow_pair<FMT>(raw[awB + 8], raw[awB + 9], okB && cok[2], okB && cok[3], scB, shB, has_pre, pre_relu, ahi[3], alo[3]);
    if (i < Ho) {
      float dhi[8][4], dlo[8][4];
#pragma unroll
      for (int nn = 0; nn < 8; ++nn) {
#pragma unroll
        for (int e = 0; e < 4; ++e) dhi[nn][e] = dlo[nn][e] = 0.f;
        mma16816<FMT>(dhi[nn], ahi, breg[nn].x, breg[nn].y);
        mma16816<FMT>(dlo[nn], alo, breg[nn].x, breg[nn].y);
        mma16816<FMT>(dlo[nn], ahi, breg[nn].z, breg[nn].w);
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int j = j0 + g + 8 * r;
        uint32_t o[8];
#pragma unroll
        for (int nn = 0; nn < 8; ++nn)
          o[nn] = pack2_t<FMT>(dhi[nn][2 * r] + dlo[nn][2 * r] * LoScale<FMT>::kDown,
                               dhi[nn][2 * r + 1] + dlo[nn][2 * r + 1] * LoScale<FMT>::kDown);
        if (j < Wo) {
          uint4* dst = out + ((size_t)((size_t)n * Ho + i) * Wo + j) * 8 + t;
          dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
          dst[4] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
    }
  }
  cp_async_wait_group<0>();
}

// Weight gradient, wide-tensor stationary: the CTA walks 32-pixel chunks of rows of b (NHWC, 64
// ch); each thread owns one (pixel slot, 8-channel piece) whose 16 bytes arrive through a private
// cp.async ring (kWgStages chunks in flight, no barrier needed: a thread reads only what it
// copied itself).  The narrow values f(a[ca][q - tap]) of the chunk are staged once per chunk in
// shared memory (double buffered, one __syncthreads per chunk).
//   accum[ca][tap][cb] += f(a[n][ca][ib-dh][jb-dw]) * b[n][ib][jb][cb]   for ca in [ca0, ca0+3)
static constexpr int kWgStages = 4;
static constexpr int kWgChunk = 64;  // pixels per chunk: two pixel slots per thread
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int FMT>
__global__ void __launch_bounds__(256, 2)
    wgrad_narrow_bs_kernel(const float* __restrict__ a, const float* __restrict__ pre, int pre_relu,
                           const uint4* __restrict__ b, float* __restrict__ accum, int N, int Ha,
                           int Wa, int Ca, int ca0, int Hb, int Wb, NarrowTaps taps) {
  __shared__ uint4 s_ring[kWgStages][2][256];
  __shared__ float s_a[2][3 * kNarrowMaxTaps * kWgChunk];  // [buf][ca][tap][pixel]
  __shared__ float s_acc[3 * kNarrowMaxTaps * 64];
  const int cg = threadIdx.x & 7, slot = threadIdx.x >> 3;
  const int cpr = (Wb + kWgChunk - 1) / kWgChunk;  // chunks per row of b
  const int chunks = N * Hb * cpr;
  // Each CTA owns a contiguous range of chunks and walks it with incremental (n, row, chunk-in-row)
  // cursors (the strided assignment cost two integer divisions per chunk and pipeline position).
  const int per = chunks / (int)gridDim.x, extra = chunks - per * (int)gridDim.x;
  const int c_begin = (int)blockIdx.x * per + min((int)blockIdx.x, extra);
  const int c_end = c_begin + per + ((int)blockIdx.x < extra ? 1 : 0);
  struct Cursor {
    int c, n, ib, jc;
  };
  auto make_cursor = [&](int c) {
    Cursor k;
    k.c = c;
    const int row = c / cpr;
    k.jc = c - row * cpr;
    k.n = row / Hb;
    k.ib = row - k.n * Hb;
    return k;
  };
  auto advance = [&](Cursor& k) {
    ++k.c;
    if (++k.jc == cpr) {
      k.jc = 0;
      if (++k.ib == Hb) {
        k.ib = 0;
        ++k.n;
      }
    }
  };
  auto issue = [&](const Cursor& k, int stage) {
    if (k.c < c_end) {
      const int jb = k.jc * kWgChunk + slot;
      const uint4* src = b + (((size_t)k.n * Hb + k.ib) * Wb + jb) * 8 + cg;
      if (jb < Wb) cp_async16(&s_ring[stage][0][threadIdx.x], src);
      if (jb + 32 < Wb) cp_async16(&s_ring[stage][1][threadIdx.x], src + 32 * 8);
    }
    cp_async_commit();
  };
  // narrow values of a chunk, 3 channels x taps x 64 pixels: this thread fetches (channel k, its
  // tap, its pixel) for k = 0..2.  The loads are raw; the BN/ReLU pre-transform is applied when the
  // values are stored to shared memory one chunk later, so the load latency overlaps the FMAs.
  const int f_sl = threadIdx.x & 63, f_tap = (threadIdx.x >> 6) & 3;
  const bool f_tap_ok = f_tap < taps.n_taps;
  const int f_dh = taps.dh[f_tap], f_dw = taps.dw[f_tap];
  float f_sc[3], f_sh[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const bool okc = ca0 + k < Ca && pre != nullptr;
    f_sc[k] = okc ? __ldg(pre + ca0 + k) : 1.f;
    f_sh[k] = okc ? __ldg(pre + Ca + ca0 + k) : 0.f;
  }
  const bool has_pre = pre != nullptr;
  const size_t a_plane = (size_t)Ha * Wa;
  auto fetch_a = [&](const Cursor& cur, float (&val)[3], bool& ok) {
    const int ia = cur.ib - f_dh, ja = cur.jc * kWgChunk + f_sl - f_dw;
    ok = cur.c < c_end && f_tap_ok && ia >= 0 && ia < Ha && ja >= 0 && ja < Wa;
    const float* src = a + ((size_t)cur.n * Ca + ca0) * a_plane + (size_t)ia * Wa + ja;
#pragma unroll
    for (int k = 0; k < 3; ++k) val[k] = (ok && ca0 + k < Ca) ? __ldg(src + k * a_plane) : 0.f;
  };
  auto store_a = [&](int buf, const float (&val)[3], bool ok) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float v = val[k];
      if (has_pre) {
        v = fmaf(v, f_sc[k], f_sh[k]);
        if (pre_relu) v = fmaxf(v, 0.f);
      }
      s_a[buf][k * (kNarrowMaxTaps * kWgChunk) + threadIdx.x] = (ok && ca0 + k < Ca) ? v : 0.f;
    }
  };
  float acc[3][kNarrowMaxTaps][8];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int t = 0; t < kNarrowMaxTaps; ++t)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[c][t][e] = 0.f;
  for (int i = threadIdx.x; i < 3 * kNarrowMaxTaps * 64; i += blockDim.x) s_acc[i] = 0.f;
  Cursor cur = make_cursor(c_begin), ahead = cur, nxt_cur = cur;
#pragma unroll
  for (int s = 0; s < kWgStages; ++s) {
    issue(ahead, s);
    advance(ahead);
  }
  {
    float v0[3];
    bool ok0;
    fetch_a(cur, v0, ok0);
    store_a(0, v0, ok0);
  }
  advance(nxt_cur);
  __syncthreads();
  int stage = 0, buf = 0;
  for (; cur.c < c_end; advance(cur), advance(nxt_cur)) {
    float nxt[3];
    bool nxt_ok;
    fetch_a(nxt_cur, nxt, nxt_ok);  // consumed after this chunk's FMAs
    cp_async_wait<kWgStages - 1>();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int ps = slot + 32 * half;
      if (cur.jc * kWgChunk + ps < Wb) {
        const uint4 v = s_ring[stage][half][threadIdx.x];
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
        float f[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 x2 = unpack2_t<FMT>(u[e]);
          f[2 * e] = x2.x;
          f[2 * e + 1] = x2.y;
        }
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3)
#pragma unroll
          for (int t = 0; t < kNarrowMaxTaps; ++t) {
            const float av = s_a[buf][(c3 * kNarrowMaxTaps + t) * kWgChunk + ps];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[c3][t][e] = fmaf(av, f[e], acc[c3][t][e]);
          }
      }
    }
    issue(ahead, stage);
    advance(ahead);
    store_a(buf ^ 1, nxt, nxt_ok);
    __syncthreads();
    stage = stage + 1 == kWgStages ? 0 : stage + 1;
    buf ^= 1;
  }
  cp_async_wait<0>();
  // fold the 4 pixel slots of a warp, then the warps of the block, then one atomic per value
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int t = 0; t < kNarrowMaxTaps; ++t)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = acc[c][t][e];
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if ((threadIdx.x & 31) < 8) atomicAdd(&s_acc[(c * kNarrowMaxTaps + t) * 64 + cg * 8 + e], v);
      }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * kNarrowMaxTaps * 64; i += blockDim.x) {
    const int cb = i & 63, t = (i >> 6) % kNarrowMaxTaps, c = i / (64 * kNarrowMaxTaps);
    if (t < taps.n_taps && ca0 + c < Ca)
      atomicAdd(accum + ((size_t)(ca0 + c) * taps.n_taps + t) * 64 + cb, s_acc[i]);
  }
}

// The same weight gradient on mma.sync (sm_100a): GEMM M = (narrow channel, tap) = 4 x 4 rows, N = 64 wide
// channels, K = pixels.  The wide tensor arrives as TMA boxes of 8 rows x 16 pixels x 64 channels (128B-swizzled,
// 4-deep ring); `ldmatrix.x4.trans` turns its pixel-major rows into the k-major B fragments (pixels are the
// reduction axis), four per 16 pixels.  The narrow operand f(a) is staged once per block in shared memory as
// hi / lo 16-bit halves (fp32 accuracy), each word holding the PAIR (pixel p, p+1) so that an A register is one
// LDS whatever the tap's column shift.  16 mma.sync + 4 ldmatrix + 8 LDS per 16 pixels and warp replace the
// 384 fp32 FMAs per warp of the SIMT kernel (ncu: 20 M instructions, 57 % of the issue slots, 35 us).
static constexpr int kWnStages = 4;
static constexpr int kWnTile = 8 * 16 * 128;
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&b)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3])
               : "r"(addr));
}
template <int FMT>
__global__ void __launch_bounds__(256, 2)
    wgrad_narrow_tma_kernel(const __grid_constant__ CUtensorMap tmap_b, const float* __restrict__ a,
                            const float* __restrict__ pre, int pre_relu, float* __restrict__ accum, int N,
                            int Ha, int Wa, int Ca, int ca0, int Hb, int Wb, NarrowTaps taps, int dhmax,
                            int dwmax, int HR, int HC) {
  extern __shared__ __align__(1024) uint8_t dsm[];  // [kWnStages][16 KB] wide tiles, then the narrow halo ring
  if (smem_u32(dsm) & 1023u) __trap();
  float* s_raw = reinterpret_cast<float*>(dsm + kWnStages * kWnTile);  // [kHaloDepth + 1][4 ca][HR][HC + 1]
  __shared__ float s_red[16 * 64];
  __shared__ uint64_t s_bar[kWnStages];
  const int HCp = HC + 1, plane = 4 * HR * HCp;
  for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) s_red[i] = 0.f;
  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_b);
    for (int b = 0; b < kWnStages; ++b) mbar_init(&s_bar[b], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const uint32_t JB = (uint32_t)(Wb + 15) >> 4, IB = (uint32_t)(Hb + 7) >> 3;
  const uint32_t ctiles = (uint32_t)N * IB * JB;
  auto decode = [&](uint32_t ct, int& n, int& i0, int& j0) {
    const uint32_t q = ct / JB, jb = ct - q * JB;
    const uint32_t nn = q / IB, ib = q - nn * IB;
    n = (int)nn;
    i0 = (int)(ib * 8);
    j0 = (int)(jb * 16);
  };
  auto issue = [&](uint32_t ct, int buf) {
    int n, i0, j0;
    decode(ct, n, i0, j0);
    mbar_arrive_expect_tx(&s_bar[buf], (uint32_t)kWnTile);
    tma_load_4d(dsm + (size_t)buf * kWnTile, &tmap_b, &s_bar[buf], 0, j0, i0, n);
  };
  if (threadIdx.x == 0)
    for (int b = 0; b < kWnStages - 1; ++b)
      if (blockIdx.x + (uint32_t)b * gridDim.x < ctiles) issue(blockIdx.x + (uint32_t)b * gridDim.x, b);
  // narrow halo element (ca, hr, hc) = a[n][ca0 + ca][i0 + hr - dhmax][j0 + hc - dwmax] (see issue_halo in
  // narrow_in_pair_kernel)
  const size_t a_plane = (size_t)Ha * Wa;
  int e_ca[3], e_hr[3], e_hc[3];
#pragma unroll
  for (int u = 0; u < 3; ++u) {
    const int w = threadIdx.x + 256 * u;
    e_hc[u] = w % HCp;
    e_hr[u] = (w / HCp) % HR;
    e_ca[u] = w / (HCp * HR);
  }
  auto issue_halo = [&](uint32_t ct, int slot) {
    if (ct < ctiles) {
      int n, i0, j0;
      decode(ct, n, i0, j0);
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        const int w = threadIdx.x + 256 * u;
        if (w < plane) {
          const int ia = i0 + e_hr[u] - dhmax, ja = j0 + e_hc[u] - dwmax;
          const bool ok = ca0 + e_ca[u] < Ca && ia >= 0 && ia < Ha && ja >= 0 && ja < Wa;
          const float* src = ok ? a + ((size_t)n * Ca + ca0 + e_ca[u]) * a_plane + (size_t)ia * Wa + ja : a;
          cp_async4_zfill(s_raw + (size_t)slot * plane + w, src, ok);
        }
      }
    }
    cp_async_commit_group();
  };
  // per-lane constants: A rows m = g and g + 8 -> channel m >> 2, tap m & 3; B row offsets
  const bool has_pre = pre != nullptr;
  const int tap = g & 3, cA = g >> 2, cB = cA + 2;
  const bool tap_ok = tap < taps.n_taps;
  const int tdh = tap_ok ? taps.dh[tap] : 0, tdw = tap_ok ? taps.dw[tap] : 0;
  const int ar = warp - tdh + dhmax, ac = 2 * t - tdw + dwmax;
  const int awA = (cA * HR + ar) * HCp + ac, awB = (cB * HR + ar) * HCp + ac;
  const bool okA = tap_ok && ca0 + cA < Ca, okB = tap_ok && ca0 + cB < Ca;
  const float scA = (has_pre && ca0 + cA < Ca) ? __ldg(pre + ca0 + cA) : 1.f;
  const float shA = (has_pre && ca0 + cA < Ca) ? __ldg(pre + Ca + ca0 + cA) : 0.f;
  const float scB = (has_pre && ca0 + cB < Ca) ? __ldg(pre + ca0 + cB) : 1.f;
  const float shB = (has_pre && ca0 + cB < Ca) ? __ldg(pre + Ca + ca0 + cB) : 0.f;
  uint32_t boff[4];
  {
    const int px = (lane & 7) + 8 * ((lane >> 3) & 1);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      boff[q] = (uint32_t)((warp * 16 + px) * 128 + (((2 * q + (lane >> 4)) ^ (px & 7)) << 4));
  }
  float ahi[8][4], alo[8][4];
#pragma unroll
  for (int nn = 0; nn < 8; ++nn)
#pragma unroll
    for (int e = 0; e < 4; ++e) ahi[nn][e] = alo[nn][e] = 0.f;
#pragma unroll
  for (int d = 0; d < kHaloDepth; ++d) issue_halo(blockIdx.x + (uint32_t)d * gridDim.x, d);
  int k = 0;
  for (uint32_t ct = blockIdx.x; ct < ctiles; ct += gridDim.x, ++k) {
    const int buf = k % kWnStages;
    cp_async_wait_group<kHaloDepth - 1>();
    fence_proxy_async();  // generic-proxy reads of block k - 1's wide tile before its TMA refill below
    __syncthreads();      // block k's halo visible; everybody is done with block k - 1
    if (threadIdx.x == 0 && ct + (uint32_t)(kWnStages - 1) * gridDim.x < ctiles)
      issue(ct + (uint32_t)(kWnStages - 1) * gridDim.x, (k + kWnStages - 1) % kWnStages);
    issue_halo(ct + (uint32_t)kHaloDepth * gridDim.x, (k + kHaloDepth) % (kHaloDepth + 1));
    int n, i0, j0;
    decode(ct, n, i0, j0);
    const float* raw = s_raw + (size_t)(k % (kHaloDepth + 1)) * plane;
    const int ia = i0 + warp - tdh, ja = j0 + 2 * t - tdw;  // image position of A[.][k = 2t]
    const bool rok = ia >= 0 && ia < Ha;
    bool cok[4];
    cok[0] = rok && ja >= 0 && ja < Wa;
    cok[1] = rok && ja + 1 >= 0 && ja + 1 < Wa;
    cok[2] = rok && ja + 8 >= 0 && ja + 8 < Wa;
    cok[3] = rok && ja + 9 >= 0 && ja + 9 < Wa;
    uint32_t fh[4], fl[4];  // a0 = (m g, k 2t..), a1 = (m g+8, k 2t..), a2 = (m g, k 2t+8..), a3 = (m g+8, k 2t+8..)
    narrow_pair<FMT>(raw[awA], raw[awA + 1], okA && cok[0], okA && cok[1], scA, shA, has_pre, pre_relu, fh[0], fl[0]);
    narrow_pair<FMT>(raw[awB], raw[awB + 1], okB && cok[0], okB && cok[1], scB, shB, has_pre, pre_relu, fh[1], fl[1]);
    narrow_pair<FMT>(raw[awA + 8], raw[awA + 9], okA && cok[2], okA && cok[3], scA, shA, has_pre, pre_relu, fh[2], fl[2]);
    narrow_pair<FMT>(raw[awB + 8], raw[awB + 9], okB && cok[2], okB && cok[3], scB, shB, has_pre, pre_relu, fh[3], fl[3]);
    mbar_wait(&s_bar[buf], ((uint32_t)(k / kWnStages)) & 1u);
    const uint32_t tile = smem_u32(dsm) + (uint32_t)(buf * kWnTile);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t b[4];
      ldmatrix_x4_trans(b, tile + boff[q]);
      mma16816<FMT>(ahi[2 * q], fh, b[0], b[1]);
      mma16816<FMT>(alo[2 * q], fl, b[0], b[1]);
      mma16816<FMT>(ahi[2 * q + 1], fh, b[2], b[3]);
      mma16816<FMT>(alo[2 * q + 1], fl, b[2], b[3]);
    }
  }
  cp_async_wait_group<0>();
  // D[m][n]: rows g / g + 8, columns 8 nn + 2t + {0, 1}; fold the warps in shared memory, one atomic per value
#pragma unroll
  for (int nn = 0; nn < 8; ++nn)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int m = g + 8 * (e >> 1), n = 8 * nn + 2 * t + (e & 1);
      atomicAdd(&s_red[m * 64 + n], ahi[nn][e] + alo[nn][e] * LoScale<FMT>::kDown);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) {
    const int n = i & 63, m = i >> 6, ca = m >> 2, tp = m & 3;
    if (tp < taps.n_taps && ca0 + ca < Ca)
      atomicAdd(accum + ((size_t)(ca0 + ca) * taps.n_taps + tp) * 64 + n, s_red[i]);
  }
}

// accum[ca][tap][cb] -> OIHW dw
__global__ void wgrad_narrow_finish_kernel(const float* __restrict__ accum, float* __restrict__ dw,
                                           int a_is_output, int Ca, int Cb, int R, int S) {
  const int total = Ca * Cb * R * S;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int t = i;
    const int s = t % S;
    t /= S;
    const int r = t % R;
    t /= R;
    int o, c;
    if (a_is_output) {  // dw[k=ca][c=cb][r][s]
      c = t % Cb;
      o = t / Cb;
      dw[i] = accum[((size_t)o * R * S + r * S + s) * Cb + c];
    } else {  // dw[k=cb][c=ca][r][s]
      c = t % Ca;
      o = t / Ca;
      dw[i] = accum[((size_t)c * R * S + r * S + s) * Cb + o];
    }
  }
}

// build the smem weight tables on device from OIHW fp32
// mode 0: wt[tap][ci=c][co=k] = w[k][c][r][s]      (narrow_out fwd: wide in c, narrow out k)
// mode 1: wt[tap][ci=k][co=c] = w[k][c][r][s]      (narrow_out_dgrad / narrow_in flip: reduce over k)
__global__ void narrow_wtable_kernel(const float* __restrict__ w, float* __restrict__ wt, int K, int C,
                                     int R, int S, int mode) {
  const int total = K * C * R * S;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int t = i;
    const int s = t % S;
    t /= S;
    const int r = t % R;
    t /= R;
    const int c = t % C;
    const int k = t / C;
    const int tap = r * S + s;
    if (mode == 0)
      wt[((size_t)tap * C + c) * K + k] = w[i];
    else
      wt[((size_t)tap * K + k) * C + c] = w[i];
  }
}

// ------------------------------------------------------------------------------------------------
// stem conv1 weight gradient (SIMT): out[k][q] = sum_pix g[pix][k] * patch[pix][q],
// q = r*28 + s*4 + c over the packed 4-channel image; 8 k x 7 q register tile per thread.
// ------------------------------------------------------------------------------------------------
static constexpr int kSwPix = 64;  // output pixels (one row segment) per smem chunk
__global__ void __launch_bounds__(256)
    stem_wgrad_kernel(const uint16_t* __restrict__ xp, int x_fmt, const uint16_t* __restrict__ g,
                      int g_fmt, float* __restrict__ accum, int N, int Ho, int Wo, int rows, int cols) {
  __shared__ __align__(16) uint16_t s_g[kSwPix * 64];
  __shared__ __align__(16) uint16_t s_x[7 * (2 * kSwPix + 6) * 4];
  constexpr int RW = (2 * kSwPix + 6) * 4;  // smem patch row width in elements
  const int tk = threadIdx.x >> 5;          // 8 groups of 8 output channels
  const int tq = threadIdx.x & 31;
  int qoff[7];
  bool qok[7];
#pragma unroll
  for (int m = 0; m < 7; ++m) {
    const int q = tq + 32 * m;
    qok[m] = q < 196;
    const int r = qok[m] ? q / 28 : 0, s4 = qok[m] ? q % 28 : 0;
    qoff[m] = r * RW + s4;
  }
  float acc[8][7];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int m = 0; m < 7; ++m) acc[a][m] = 0.f;
  const int segs = (Wo + kSwPix - 1) / kSwPix;
  const int64_t chunks = (int64_t)N * Ho * segs;
  for (int64_t ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
    const int seg = (int)(ch % segs);
    const int64_t t = ch / segs;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const int wo0 = seg * kSwPix;
    const int np = min(kSwPix, Wo - wo0);
    __syncthreads();
    // g tile: np pixels x 64 channels
    for (int i = threadIdx.x; i < kSwPix * 8; i += blockDim.x) {
      const int px = i >> 3, v = i & 7;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (px < np)
        val = __ldg(reinterpret_cast<const uint4*>(g + (((int64_t)n * Ho + ho) * Wo + wo0 + px) * 64) + v);
      reinterpret_cast<uint4*>(s_g)[i] = val;
    }
    // image patch rows 2ho..2ho+6, cols 2wo0 .. 2wo0 + 2*kSwPix+5 (packed coords, frame included)
    for (int i = threadIdx.x; i < 7 * (2 * kSwPix + 6); i += blockDim.x) {
      const int r = i / (2 * kSwPix + 6), cpx = i % (2 * kSwPix + 6);
      const int row = 2 * ho + r, col = 2 * wo0 + cpx;
      uint2 val = make_uint2(0, 0);
      if (row < rows && col < cols)
        val = __ldg(reinterpret_cast<const uint2*>(xp + (((int64_t)n * rows + row) * cols + col) * 4));
      reinterpret_cast<uint2*>(s_x)[i] = val;
    }
    __syncthreads();
    for (int px = 0; px < np; ++px) {
      const uint4 gv = *reinterpret_cast<const uint4*>(s_g + px * 64 + tk * 8);
      const uint32_t gu[4] = {gv.x, gv.y, gv.z, gv.w};
      float gf[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack2(gu[e], g_fmt);
        gf[2 * e] = f.x;
        gf[2 * e + 1] = f.y;
      }
      float xf[7];
#pragma unroll
      for (int m = 0; m < 7; ++m) xf[m] = qok[m] ? h16_to_float(s_x[qoff[m] + 8 * px], x_fmt) : 0.f;
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int m = 0; m < 7; ++m) acc[a][m] = fmaf(gf[a], xf[m], acc[a][m]);
    }
  }
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int m = 0; m < 7; ++m)
      if (qok[m]) atomicAdd(accum + (size_t)(tk * 8 + a) * 196 + tq + 32 * m, acc[a][m]);
}

// accum[k][r*28 + s*4 + c] -> dw[k][c][r][s] * scale[k]
__global__ void stem_wgrad_finish_kernel(const float* __restrict__ accum,
                                         const float* __restrict__ scale, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 3 * 49) return;
  int t = i;
  const int s = t % 7;
  t /= 7;
  const int r = t % 7;
  t /= 7;
  const int c = t % 3;
  const int k = t / 3;
  dw[i] = accum[(size_t)k * 196 + r * 28 + s * 4 + c] * (scale ? scale[k] : 1.f);
}

// stem weights: OIHW [64][3][7][7] fp32 (x scale[k]) -> [64][7][32] 16-bit, q = s*4 + c, zero padded
__global__ void stem_pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ scale,
                                        uint16_t* __restrict__ dst, int fmt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 7 * 32) return;
  const int q = i & 31, r = (i >> 5) % 7, k = i / (7 * 32);
  const int s = q >> 2, c = q & 3;
  float v = 0.f;
  if (s < 7 && c < 3) v = w[(((size_t)k * 3 + c) * 7 + r) * 7 + s] * (scale ? scale[k] : 1.f);
  dst[i] = float_to_h16(v, fmt);
}

static int blocks_for_pixels(int64_t npix, int pix_per_block, int per_sm = 4) {
  int64_t b = (npix + pix_per_block - 1) / pix_per_block;
  const int64_t cap = (int64_t)num_sms() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

static bool pow2_lanes(int cw) {
  const int l = cw / 8;
  return cw % 8 == 0 && l >= 1 && l <= 32 && (l & (l - 1)) == 0;
}

void launch_stem_wgrad_finish(const float* accum, const float* scale, float* dw, cudaStream_t st) {
  stem_wgrad_finish_kernel<<<(64 * 3 * 49 + 255) / 256, 256, 0, st>>>(accum, scale, dw);
}

// mma.sync fast paths (64 wide channels, <= 16 narrow channels, pixel count in 32 bits)
static bool narrow_fast_ok(int c_wide, int c_narrow, int64_t npix) {
  return c_wide == 64 && c_narrow >= 1 && c_narrow <= 16 && npix > 0 && npix < (int64_t)1 << 30;
}
static int narrow_tile_blocks(int64_t npix) {
  const int64_t tiles = (npix + 15) / 16;
  int64_t b = (tiles + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 2;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}
static int narrow_out_blocks(int N, int Ho, int Wo) {
  const int64_t ctiles = (int64_t)N * ((Ho + 7) / 8) * ((Wo + 15) / 16);
  const int64_t cap = (int64_t)num_sms() * 2;
  return (int)(ctiles < cap ? ctiles : cap);
}
// returns the number of CTAs (= min/max partial pairs written when `partial` is given)
static int launch_narrow_out_mma(const void* in, int fmt, const float* w_oihw, int wmode, int R, int S, float* wt,
                                 float* out, int N, int Hi, int Wi, int CN, int Ho, int Wo, const NarrowTaps& taps,
                                 float* partial, cudaStream_t st) {
  const int blocks = narrow_out_blocks(N, Ho, Wo);
  // TMA-staged kernel (GHND_NARROW_TMA=0: the direct-load kernel)
  static const bool tma_off = [] {
    const char* e = getenv("GHND_NARROW_TMA");
    return e != nullptr && atoi(e) == 0;
  }();
  if (!tma_off && taps.n_taps >= 1) {
    int dh0 = taps.dh[0], dh1 = dh0, dw0 = taps.dw[0], dw1 = dw0;
    for (int t = 1; t < taps.n_taps; ++t) {
      dh0 = taps.dh[t] < dh0 ? taps.dh[t] : dh0;
      dh1 = taps.dh[t] > dh1 ? taps.dh[t] : dh1;
      dw0 = taps.dw[t] < dw0 ? taps.dw[t] : dw0;
      dw1 = taps.dw[t] > dw1 ? taps.dw[t] : dw1;
    }
    const int bw = 16 + dw1 - dw0, bh = 8 + dh1 - dh0;
    const int box_bytes = bw * bh * 128, tile_bytes = (box_bytes + 1023) / 1024 * 1024;
    CUtensorMap tmap;
    uint64_t dims[4] = {64, (uint64_t)Wi, (uint64_t)Hi, (uint64_t)N};
    uint64_t str[4] = {2, 128, (uint64_t)Wi * 128, (uint64_t)Hi * Wi * 128};
    uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, 1};
    if (bw <= 256 && bh <= 256 && kNoStages * tile_bytes <= 100 * 1024 &&
        encode_tmap(&tmap, 2, 4, const_cast<void*>(in), dims, str, box, 128) == GHND_OK) {
      const size_t smem = (size_t)kNoStages * tile_bytes;
#define GHND_NOT(FMT, NT, MM)                                                                                  \
  do {                                                                                                         \
    static bool attr = false;                                                                                  \
    if (!attr) {                                                                                               \
      cudaFuncSetAttribute(narrow_out_tma_kernel<FMT, NT, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                           100 * 1024);                                                                        \
      attr = true;                                                                                             \
    }                                                                                                          \
    narrow_out_tma_kernel<FMT, NT, MM><<<blocks, 256, smem, st>>>(tmap, w_oihw, wmode, out, N, CN, Ho, Wo, taps, \
                                                                  bw, dh0, dw0, tile_bytes, box_bytes,          \
                                                                  partial);                                     \
  } while (0)
#define GHND_NOT_MM(FMT, NT) \
  if (partial) GHND_NOT(FMT, NT, true); else GHND_NOT(FMT, NT, false)
      const int nt = (CN + 3) / 4;  // four channels (hi, lo column pairs) per n8 tile
      if (fmt == GHND_F16) {
        if (nt == 1) { GHND_NOT_MM(GHND_F16, 1); } else if (nt == 2) { GHND_NOT_MM(GHND_F16, 2); }
        else if (nt == 3) { GHND_NOT_MM(GHND_F16, 3); } else { GHND_NOT_MM(GHND_F16, 4); }
      } else {
        if (nt == 1) { GHND_NOT_MM(GHND_BF16, 1); } else if (nt == 2) { GHND_NOT_MM(GHND_BF16, 2); }
        else if (nt == 3) { GHND_NOT_MM(GHND_BF16, 3); } else { GHND_NOT_MM(GHND_BF16, 4); }
      }
#undef GHND_NOT_MM
#undef GHND_NOT
      return blocks;
    }
  }
  // direct-load kernel: needs the [tap][c][co] weight table
  if (wmode == 0) narrow_wtable_kernel<<<4, 256, 0, st>>>(w_oihw, wt, CN, 64, R, S, 0);
  else narrow_wtable_kernel<<<4, 256, 0, st>>>(w_oihw, wt, 64, CN, R, S, 1);
#define GHND_NO(FMT, NT, MM)                                                                        \
  narrow_out_mma_kernel<FMT, NT, MM><<<blocks, 256, 0, st>>>((const uint4*)in, wt, out, N, Hi, Wi, \
                                                             CN, Ho, Wo, taps, partial)
#define GHND_NO_MM(FMT, NT) \
  if (partial) GHND_NO(FMT, NT, true); else GHND_NO(FMT, NT, false)
  if (fmt == GHND_F16) {
    if (CN <= 8) { GHND_NO_MM(GHND_F16, 1); } else { GHND_NO_MM(GHND_F16, 2); }
  } else {
    if (CN <= 8) { GHND_NO_MM(GHND_BF16, 1); } else { GHND_NO_MM(GHND_BF16, 2); }
  }
#undef GHND_NO_MM
#undef GHND_NO
  return blocks;
}
static void launch_narrow_in_mma(const float* in, const float* pre, int pre_relu, const float* wt,
                                 void* out, int fmt, int N, int Hi, int Wi, int CN, int Ho, int Wo,
                                 const NarrowTaps& taps, cudaStream_t st) {
  const int blocks = narrow_tile_blocks((int64_t)N * Ho * Wo);
  const int ks = (CN + 3) / 4;
#define GHND_NI(FMT, KS)                                                                          \
  narrow_in_mma_kernel<FMT, KS><<<blocks, 256, 0, st>>>(in, pre, pre_relu, wt, (uint4*)out, N, Hi, \
                                                        Wi, CN, Ho, Wo, taps)
#define GHND_NI_FMT(FMT)            \
  switch (ks) {                     \
    case 1: GHND_NI(FMT, 1); break; \
    case 2: GHND_NI(FMT, 2); break; \
    case 3: GHND_NI(FMT, 3); break; \
    default: GHND_NI(FMT, 4); break; \
  }
  if (fmt == GHND_F16) {
    GHND_NI_FMT(GHND_F16)
  } else {
    GHND_NI_FMT(GHND_BF16)
  }
#undef GHND_NI_FMT
#undef GHND_NI
}

}  // namespace ghnd

extern "C" {
using namespace ghnd;

// workspace layout for the narrow convs: a small fp32 weight table
size_t ghnd_conv_narrow_workspace_bytes(int C, int K, int R, int S) {
  return (size_t)C * K * R * S * sizeof(float);
}

static int conv_narrow_out_impl(const void* x, int x_fmt, const float* w, float* y, int N, int H,
                                int W, int C, int K, int R, int S, int pad, void* workspace,
                                size_t workspace_bytes, float* minmax_partial, int partial_capacity,
                                int* n_partial, void* stream) {
  GHND_CHECK_ARG(x && w && y && workspace, "conv_narrow_out: null pointer");
  GHND_CHECK_ARG(pow2_lanes(C) && K >= 1 && K <= kNarrowMaxC && R * S <= kNarrowMaxTaps && R >= 1 &&
                     S >= 1,
                 "conv_narrow_out: unsupported shape C=%d K=%d %dx%d", C, K, R, S);
  GHND_CHECK_ARG(workspace_bytes >= ghnd_conv_narrow_workspace_bytes(C, K, R, S),
                 "conv_narrow_out: workspace too small");
  const int Ho = H + 2 * pad - R + 1, Wo = W + 2 * pad - S + 1;
  GHND_CHECK_ARG(Ho > 0 && Wo > 0, "conv_narrow_out: empty output");
  cudaStream_t st = (cudaStream_t)stream;
  NarrowTaps taps;
  taps.n_taps = R * S;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      taps.dh[r * S + s] = r - pad;
      taps.dw[r * S + s] = s - pad;
    }
  const int lanes = C / 8;
  const int64_t npix = (int64_t)N * Ho * Wo;
  const bool fast = narrow_fast_ok(C, K, npix) && (x_fmt == GHND_F16 || x_fmt == GHND_BF16);
  if (!fast) {  // the fast path builds its weight fragments itself (or its own table)
    narrow_wtable_kernel<<<4, 256, 0, st>>>(w, (float*)workspace, K, C, R, S, 0);
    GHND_LAUNCH_CHECK("narrow_wtable_kernel");
  }
  if (minmax_partial != nullptr) {
    GHND_CHECK_ARG(fast && n_partial != nullptr && partial_capacity >= narrow_out_blocks(N, Ho, Wo),
                   "conv_narrow_out_minmax: needs the 64-channel 16-bit path and %d partial pairs",
                   narrow_out_blocks(N, Ho, Wo));
  }
  if (fast) {
    const int blocks = launch_narrow_out_mma(x, x_fmt, w, 0, R, S, (float*)workspace, y, N, H, W, K, Ho, Wo,
                                             taps, minmax_partial, st);
    if (n_partial) *n_partial = blocks;
  } else if (K <= 3)
    narrow_out_kernel<3><<<blocks_for_pixels(npix, 256 / lanes), 256,
                           (size_t)R * S * C * K * sizeof(float), st>>>(
        (const uint4*)x, x_fmt, (const float*)workspace, y, N, H, W, C, K, Ho, Wo, taps);
  else
    narrow_out_kernel<kNarrowMaxC><<<blocks_for_pixels(npix, 256 / lanes), 256,
                                     (size_t)R * S * C * K * sizeof(float), st>>>(
        (const uint4*)x, x_fmt, (const float*)workspace, y, N, H, W, C, K, Ho, Wo, taps);
  GHND_LAUNCH_CHECK("narrow_out_kernel");
  return GHND_OK;
}

int ghnd_conv_narrow_out(const void* x, int x_fmt, const float* w, float* y, int N, int H, int W,
                         int C, int K, int R, int S, int pad, void* workspace,
                         size_t workspace_bytes, void* stream) {
  return conv_narrow_out_impl(x, x_fmt, w, y, N, H, W, C, K, R, S, pad, workspace, workspace_bytes,
                              nullptr, 0, nullptr, stream);
}

int ghnd_conv_narrow_out_minmax(const void* x, int x_fmt, const float* w, float* y, int N, int H,
                                int W, int C, int K, int R, int S, int pad, void* workspace,
                                size_t workspace_bytes, float* minmax_partial, int partial_capacity,
                                int* n_partial, void* stream) {
  GHND_CHECK_ARG(minmax_partial && n_partial, "conv_narrow_out_minmax: null min/max buffer");
  return conv_narrow_out_impl(x, x_fmt, w, y, N, H, W, C, K, R, S, pad, workspace, workspace_bytes,
                              minmax_partial, partial_capacity, n_partial, stream);
}

int ghnd_conv_narrow_out_dgrad(const void* dy, int dy_fmt, const float* w, float* dx, int N, int H,
                               int W, int C, int K, int R, int S, int pad, void* workspace,
                               size_t workspace_bytes, void* stream) {
  // forward conv: x[N][C(narrow)][H][W] -> y[N][Ho][Wo][K(wide)], w[K][C][R][S]
  GHND_CHECK_ARG(dy && w && dx && workspace, "conv_narrow_out_dgrad: null pointer");
  GHND_CHECK_ARG(pow2_lanes(K) && C >= 1 && C <= kNarrowMaxC && R * S <= kNarrowMaxTaps && R >= 1 &&
                     S >= 1,
                 "conv_narrow_out_dgrad: unsupported shape C=%d K=%d %dx%d", C, K, R, S);
  GHND_CHECK_ARG(workspace_bytes >= ghnd_conv_narrow_workspace_bytes(C, K, R, S),
                 "conv_narrow_out_dgrad: workspace too small");
  const int Ho = H + 2 * pad - R + 1, Wo = W + 2 * pad - S + 1;
  GHND_CHECK_ARG(Ho > 0 && Wo > 0, "conv_narrow_out_dgrad: empty output");
  cudaStream_t st = (cudaStream_t)stream;
  NarrowTaps taps;
  taps.n_taps = R * S;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      taps.dh[r * S + s] = pad - r;
      taps.dw[r * S + s] = pad - s;
    }
  const int lanes = K / 8;
  const int64_t npix = (int64_t)N * H * W;
  const bool dg_fast = narrow_fast_ok(K, C, npix) && (dy_fmt == GHND_F16 || dy_fmt == GHND_BF16);
  if (!dg_fast) {  // table wt[tap][ci=k][co=c]
    narrow_wtable_kernel<<<4, 256, 0, st>>>(w, (float*)workspace, K, C, R, S, 1);
    GHND_LAUNCH_CHECK("narrow_wtable_kernel");
  }
  if (dg_fast)
    launch_narrow_out_mma(dy, dy_fmt, w, 1, R, S, (float*)workspace, dx, N, Ho, Wo, C, H, W, taps, nullptr, st);
  else if (C <= 3)
    narrow_out_kernel<3><<<blocks_for_pixels(npix, 256 / lanes), 256,
                           (size_t)R * S * C * K * sizeof(float), st>>>(
        (const uint4*)dy, dy_fmt, (const float*)workspace, dx, N, Ho, Wo, K, C, H, W, taps);
  else
    narrow_out_kernel<kNarrowMaxC><<<blocks_for_pixels(npix, 256 / lanes), 256,
                                     (size_t)R * S * C * K * sizeof(float), st>>>(
        (const uint4*)dy, dy_fmt, (const float*)workspace, dx, N, Ho, Wo, K, C, H, W, taps);
  GHND_LAUNCH_CHECK("narrow_out_kernel(dgrad)");
  return GHND_OK;
}

int ghnd_conv_narrow_in(const float* x, const float* pre_scale_shift, int pre_relu, const float* w,
                        int flip, void* y, int y_fmt, int N, int H, int W, int C, int K, int R,
                        int S, int pad, void* workspace, size_t workspace_bytes, void* stream) {
  // flip=0: forward conv x[N][C][H][W] (narrow) -> y[N][Ho][Wo][K] (wide), w[K][C][R][S], pad
  // flip=1: dgrad of a narrow-OUT conv (wide C_w=K here -> narrow C here): x = dy_narrow
  //         [N][C][H][W], y = dx_wide [N][Hx][Wx][K]; w is that conv's OIHW weight [C][K][R][S].
  GHND_CHECK_ARG(x && w && y && workspace, "conv_narrow_in: null pointer");
  GHND_CHECK_ARG(pow2_lanes(K) && C >= 1 && C <= kNarrowMaxC && R * S <= kNarrowMaxTaps && R >= 1 &&
                     S >= 1,
                 "conv_narrow_in: unsupported shape C=%d K=%d %dx%d", C, K, R, S);
  GHND_CHECK_ARG(workspace_bytes >= ghnd_conv_narrow_workspace_bytes(C, K, R, S),
                 "conv_narrow_in: workspace too small");
  GHND_CHECK_ARG(y_fmt == GHND_F16 || y_fmt == GHND_BF16, "conv_narrow_in: bad format");
  cudaStream_t st = (cudaStream_t)stream;
  NarrowTaps taps;
  taps.n_taps = R * S;
  int Ho, Wo;
  {
    // paired-tap kernel: bch <= 4, 64 wide channels, taps in horizontally adjacent pairs (S == 2)
    static const bool pair_off = [] {
      const char* e = getenv("GHND_NARROW_TMA");
      return e != nullptr && atoi(e) == 0;
    }();
    const int ho = flip ? H - 2 * pad + R - 1 : H + 2 * pad - R + 1;
    const int wo = flip ? W - 2 * pad + S - 1 : W + 2 * pad - S + 1;
    if (!pair_off && K == 64 && C <= 4 && S == 2 && R <= 2 && ho > 0 && wo > 0 &&
        (int64_t)N * ho * wo < (int64_t)1 << 30) {
      int dhmin = 1 << 20, dhmax = -(1 << 20), dwmin = 1 << 20, dwmax = -(1 << 20);
      for (int r = 0; r < R; ++r)
        for (int s2 = 0; s2 < S; ++s2) {
          const int dh = flip ? pad - r : r - pad, dw = flip ? pad - s2 : s2 - pad;
          taps.dh[r * S + s2] = dh;
          taps.dw[r * S + s2] = dw;
          dhmin = dh < dhmin ? dh : dhmin;
          dhmax = dh > dhmax ? dh : dhmax;
          dwmin = dw < dwmin ? dw : dwmin;
          dwmax = dw > dwmax ? dw : dwmax;
        }
      const int HR = 8 + dhmax - dhmin, HC = 16 + dwmax - dwmin;
      const int64_t ctiles = (int64_t)N * ((ho + 7) / 8) * ((wo + 15) / 16);
      int blocks = num_sms() * 2;
      if (blocks > ctiles) blocks = (int)ctiles;
      const size_t smem = (size_t)(kHaloDepth + 1) * 4 * HR * (HC + 1) * sizeof(float);
      if (4 * HR * (HC + 1) <= 768) {
        if (y_fmt == GHND_F16)
          narrow_in_pair_kernel<GHND_F16><<<blocks, 256, smem, st>>>(x, pre_scale_shift, pre_relu, w, flip ? 1 : 0,
                                                                    (uint4*)y, N, H, W, C, ho, wo, taps, dhmin, dwmin,
                                                                    HR, HC);
        else
          narrow_in_pair_kernel<GHND_BF16><<<blocks, 256, smem, st>>>(x, pre_scale_shift, pre_relu, w, flip ? 1 : 0,
                                                                     (uint4*)y, N, H, W, C, ho, wo, taps, dhmin, dwmin,
                                                                     HR, HC);
        GHND_LAUNCH_CHECK("narrow_in_pair_kernel");
        return GHND_OK;
      }
    }
  }
  if (!flip) {
    Ho = H + 2 * pad - R + 1;
    Wo = W + 2 * pad - S + 1;
    // wt[tap][ci=c][co=k] = w[k][c][r][s] : mode 1 of the table builder with (K,C) as stored
    narrow_wtable_kernel<<<4, 256, 0, st>>>(w, (float*)workspace, K, C, R, S, 0);
    // mode 0 gives wt[tap][c][k] which is exactly [tap][ci][co]
    for (int r = 0; r < R; ++r)
      for (int s = 0; s < S; ++s) {
        taps.dh[r * S + s] = r - pad;
        taps.dw[r * S + s] = s - pad;
      }
  } else {
    // the narrow-out conv mapped wide[Hx][Wx] -> narrow[H][W] with H = Hx + 2*pad - R + 1
    Ho = H - 2 * pad + R - 1;
    Wo = W - 2 * pad + S - 1;
    // w is [Knarrow=C][Cwide=K][R][S]; need wt[tap][ci=narrow][co=wide] -> mode 1 with (K=C, C=K)
    narrow_wtable_kernel<<<4, 256, 0, st>>>(w, (float*)workspace, C, K, R, S, 1);
    for (int r = 0; r < R; ++r)
      for (int s = 0; s < S; ++s) {
        taps.dh[r * S + s] = pad - r;
        taps.dw[r * S + s] = pad - s;
      }
  }
  GHND_LAUNCH_CHECK("narrow_wtable_kernel");
  GHND_CHECK_ARG(Ho > 0 && Wo > 0, "conv_narrow_in: empty output");
  const int lanes = K / 8;
  const int64_t npix = (int64_t)N * Ho * Wo;
  if (narrow_fast_ok(K, C, npix))
    launch_narrow_in_mma(x, pre_scale_shift, pre_relu, (const float*)workspace, y, y_fmt, N, H, W, C,
                         Ho, Wo, taps, st);
  else
    narrow_in_kernel<<<blocks_for_pixels(npix, 256 / lanes), 256,
                       (size_t)R * S * C * K * sizeof(float), st>>>(
        x, pre_scale_shift, pre_relu, (const float*)workspace, (uint4*)y, y_fmt, N, H, W, C, K, Ho,
        Wo, taps);
  GHND_LAUNCH_CHECK("narrow_in_kernel");
  return GHND_OK;
}

size_t ghnd_wgrad_narrow_workspace_bytes(int Ca, int Cb, int R, int S) {
  return (size_t)Ca * Cb * R * S * sizeof(float);
}

int ghnd_wgrad_narrow(const float* a, const float* pre_scale_shift, int pre_relu, const void* b,
                      int b_fmt, float* dw_oihw, int a_is_output, int N, int Ha, int Wa, int Ca,
                      int Hb, int Wb, int Cb, int R, int S, int pad, void* workspace,
                      size_t workspace_bytes, void* stream) {
  GHND_CHECK_ARG(a && b && dw_oihw && workspace, "wgrad_narrow: null pointer");
  GHND_CHECK_ARG(pow2_lanes(Cb) && Ca >= 1 && Ca <= kNarrowMaxC && R * S <= kNarrowMaxTaps && R >= 1 &&
                     S >= 1,
                 "wgrad_narrow: unsupported shape Ca=%d Cb=%d %dx%d", Ca, Cb, R, S);
  GHND_CHECK_ARG(workspace_bytes >= ghnd_wgrad_narrow_workspace_bytes(Ca, Cb, R, S),
                 "wgrad_narrow: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  GHND_CUDA(cudaMemsetAsync(workspace, 0, ghnd_wgrad_narrow_workspace_bytes(Ca, Cb, R, S), st));
  NarrowTaps taps;
  taps.n_taps = R * S;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < S; ++s) {
      taps.dh[r * S + s] = a_is_output ? r - pad : pad - r;
      taps.dw[r * S + s] = a_is_output ? s - pad : pad - s;
    }
  const int lanes = Cb / 8;
  const int64_t npix = (int64_t)N * Ha * Wa;
  const int64_t chunks_b = (int64_t)N * Hb * ((Wb + kWgChunk - 1) / kWgChunk);
  const bool fast = Cb == 64 && chunks_b < (int64_t)1 << 30 && (b_fmt == GHND_F16 || b_fmt == GHND_BF16);
  static const bool tma_off = [] {
    const char* e = getenv("GHND_NARROW_TMA");
    return e != nullptr && atoi(e) == 0;
  }();
  if (fast && !tma_off) {
    int dh0 = taps.dh[0], dh1 = dh0, dw0 = taps.dw[0], dw1 = dw0;
    for (int t = 1; t < taps.n_taps; ++t) {
      dh0 = taps.dh[t] < dh0 ? taps.dh[t] : dh0;
      dh1 = taps.dh[t] > dh1 ? taps.dh[t] : dh1;
      dw0 = taps.dw[t] < dw0 ? taps.dw[t] : dw0;
      dw1 = taps.dw[t] > dw1 ? taps.dw[t] : dw1;
    }
    const int HR = 8 + dh1 - dh0, HC = 16 + dw1 - dw0;
    CUtensorMap tmap;
    uint64_t dims[4] = {64, (uint64_t)Wb, (uint64_t)Hb, (uint64_t)N};
    uint64_t str[4] = {2, 128, (uint64_t)Wb * 128, (uint64_t)Hb * Wb * 128};
    uint32_t box[4] = {64, 16, 8, 1};
    if (4 * HR * (HC + 1) <= 768 && encode_tmap(&tmap, 2, 4, const_cast<void*>(b), dims, str, box, 128) == GHND_OK) {
      const size_t smem = (size_t)kWnStages * kWnTile + (size_t)(kHaloDepth + 1) * 4 * HR * (HC + 1) * sizeof(float);
      static bool attr = false;
      if (!attr) {
        cudaFuncSetAttribute(wgrad_narrow_tma_kernel<GHND_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(wgrad_narrow_tma_kernel<GHND_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        attr = true;
      }
      const int64_t ctiles = (int64_t)N * ((Hb + 7) / 8) * ((Wb + 15) / 16);
      int blocks = num_sms() * 2;
      if (blocks > ctiles) blocks = (int)ctiles;
      for (int ca0 = 0; ca0 < Ca; ca0 += 4) {
        if (b_fmt == GHND_F16)
          wgrad_narrow_tma_kernel<GHND_F16><<<blocks, 256, smem, st>>>(tmap, a, pre_scale_shift, pre_relu,
                                                                      (float*)workspace, N, Ha, Wa, Ca, ca0, Hb, Wb,
                                                                      taps, dh1, dw1, HR, HC);
        else
          wgrad_narrow_tma_kernel<GHND_BF16><<<blocks, 256, smem, st>>>(tmap, a, pre_scale_shift, pre_relu,
                                                                       (float*)workspace, N, Ha, Wa, Ca, ca0, Hb, Wb,
                                                                       taps, dh1, dw1, HR, HC);
        GHND_LAUNCH_CHECK("wgrad_narrow_tma_kernel");
      }
      wgrad_narrow_finish_kernel<<<4, 256, 0, st>>>((const float*)workspace, dw_oihw, a_is_output, Ca, Cb, R, S);
      GHND_LAUNCH_CHECK("wgrad_narrow_finish_kernel");
      return GHND_OK;
    }
  }
  for (int ca0 = 0; ca0 < Ca; ca0 += 3) {
    if (fast) {
      int blocks = num_sms() * 2;
      if (blocks > chunks_b) blocks = (int)chunks_b;
      if (b_fmt == GHND_F16)
        wgrad_narrow_bs_kernel<GHND_F16><<<blocks, 256, 0, st>>>(
            a, pre_scale_shift, pre_relu, (const uint4*)b, (float*)workspace, N, Ha, Wa, Ca, ca0, Hb,
            Wb, taps);
      else
        wgrad_narrow_bs_kernel<GHND_BF16><<<blocks, 256, 0, st>>>(
            a, pre_scale_shift, pre_relu, (const uint4*)b, (float*)workspace, N, Ha, Wa, Ca, ca0, Hb,
            Wb, taps);
    } else {
      wgrad_narrow_kernel<<<blocks_for_pixels(npix, (256 / lanes) * 8, 2), 256, 0, st>>>(
          a, pre_scale_shift, pre_relu, (const uint4*)b, b_fmt, (float*)workspace, N, Ha, Wa, Ca, ca0,
          Hb, Wb, Cb, taps);
    }
    GHND_LAUNCH_CHECK("wgrad_narrow_kernel");
  }
  wgrad_narrow_finish_kernel<<<4, 256, 0, st>>>((const float*)workspace, dw_oihw, a_is_output, Ca, Cb,
                                                R, S);
  GHND_LAUNCH_CHECK("wgrad_narrow_finish_kernel");
  return GHND_OK;
}

int ghnd_stem_pack_weight(const float* w_oihw, const float* scale_o, void* dst, int dst_fmt,
                          void* stream) {
  GHND_CHECK_ARG(w_oihw && dst && (dst_fmt == GHND_F16 || dst_fmt == GHND_BF16),
                 "stem_pack_weight: bad argument");
  stem_pack_weight_kernel<<<(64 * 7 * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      w_oihw, scale_o, (uint16_t*)dst, dst_fmt);
  GHND_LAUNCH_CHECK("stem_pack_weight_kernel");
  return GHND_OK;
}

size_t ghnd_stem_wgrad_workspace_bytes(void) { return (size_t)64 * 196 * sizeof(float); }

int ghnd_stem_wgrad(const void* x_packed, int x_fmt, const void* g, int g_fmt, const float* scale_o,
                    float* dw_oihw, int N, int Hp, int Wp, void* workspace, size_t workspace_bytes,
                    void* stream) {
  GHND_CHECK_ARG(x_packed && g && dw_oihw && workspace, "stem_wgrad: null pointer");
  GHND_CHECK_ARG(N > 0 && Hp > 0 && Wp > 0 && Hp % 2 == 0 && Wp % 8 == 0, "stem_wgrad: bad geometry");
  GHND_CHECK_ARG(workspace_bytes >= ghnd_stem_wgrad_workspace_bytes(), "stem_wgrad: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  GHND_CUDA(cudaMemsetAsync(workspace, 0, ghnd_stem_wgrad_workspace_bytes(), st));
  const int Ho = Hp / 2, Wo = Wp / 2;
  const int segs = (Wo + kSwPix - 1) / kSwPix;
  const int64_t chunks = (int64_t)N * Ho * segs;
  int blocks = num_sms() * 2;
  if (blocks > chunks) blocks = (int)chunks;
  stem_wgrad_kernel<<<blocks, 256, 0, st>>>((const uint16_t*)x_packed, x_fmt, (const uint16_t*)g, g_fmt,
                                            (float*)workspace, N, Ho, Wo, Hp + 6, Wp + 8);
  GHND_LAUNCH_CHECK("stem_wgrad_kernel");
  stem_wgrad_finish_kernel<<<(64 * 3 * 49 + 255) / 256, 256, 0, st>>>((const float*)workspace, scale_o,
                                                                      dw_oihw);
  GHND_LAUNCH_CHECK("stem_wgrad_finish_kernel");
  return GHND_OK;
}
}
