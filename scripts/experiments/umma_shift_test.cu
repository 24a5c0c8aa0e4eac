// Micro-experiment (round 1): which shared-memory descriptor tricks does tcgen05.mma accept on sm_100a?
//  T1  K-major SWIZZLE_128B A tile loaded by ONE dense TMA box {64ch, HW=10, HR=18}; the MMA reads the
//      window shifted by (dh,dw) rows with SBO = HW*128 (not a multiple of 1024), base_offset 0 / dw.
//  T2  same but every halo row in its own 2 KB slot (SBO = 2048), base_offset 0 / dw.
//  T3  K-major SWIZZLE_NONE with LBO = 16 B: overlapping im2col rows (row r = bytes [16r, 16r+32)).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I.. umma_shift_test.cu -o umma_shift_test
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "../../hnd_ghnd_object_detectors_b200/csrc/common.cuh"
using namespace ghnd;

namespace ghnd {
void set_error(const char* fmt, ...) {}
int cuda_fail(cudaError_t e, const char* what) { printf("cuda fail %s: %s\n", what, cudaGetErrorString(e)); return 2; }
int num_sms() { return 148; }
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int encode_tmap(CUtensorMap* map, int elem_bytes, int rank, void* base, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  encode_tiled_fn fn = (encode_tiled_fn)p;
  cuuint64_t gd[5], gs[5]; cuuint32_t gb[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; gb[i] = box[i]; es[i] = 1; if (i > 0) gs[i - 1] = strides_bytes[i]; }
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, rank, base, gd, gs, gb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 2; }
  return 0;
}
}

struct P {
  CUtensorMap ta, tb;
  int mode;        // 1 dense halo, 2 slotted halo, 3 no-swizzle overlapping rows
  int dh, dw, base_off_mode;
  float* out;      // [128][64]
};

__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  uint64_t d = make_smem_desc(addr, lbo, sbo, layout);
  d |= (uint64_t)(base_off & 7) << 49;
  return d;
}

__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ P p) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sa = smem;              // up to 40 KB
  uint8_t* sb = smem + 40 * 1024;  // B tile [64 n][64 k] SW128 = 8 KB (mode 3: [64 n][32 k] SW64 4KB)
  uint64_t* bar = (uint64_t*)(smem + 52 * 1024);
  uint64_t* mbar = bar + 1;
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(mbar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = *slot;
  if (threadIdx.x == 0) {
    if (p.mode == 1) {
      mbar_arrive_expect_tx(bar, 18 * 10 * 128 + 64 * 128);
      tma_load_4d(sa, &p.ta, bar, 0, 0, 0, 0);
    } else if (p.mode == 2) {
      mbar_arrive_expect_tx(bar, 18 * 10 * 128 + 64 * 128);
      for (int r = 0; r < 18; ++r) tma_load_4d(sa + r * 2048, &p.ta, bar, 0, 0, r, 0);
    } else {
      mbar_arrive_expect_tx(bar, 2304 + 64 * 64);
      tma_load_2d(sa, &p.ta, bar, 0, 0);   // one raw image row segment: 1152 elements = 2304 B
    }
    tma_load_2d(sb, &p.tb, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    if (p.mode == 1 || p.mode == 2) {
      const uint32_t pitch = p.mode == 1 ? 10 * 128 : 2048;
      const uint32_t start = smem_u32(sa) + p.dh * pitch + p.dw * 128;
      const uint32_t bo = p.base_off_mode == 0 ? 0 : (p.base_off_mode == 1 ? (uint32_t)p.dw : ((start >> 7) & 7));
      const uint32_t idesc = make_idesc(GHND_F16, GHND_F16, 0, 0, 128, 64);
      for (int kk = 0; kk < 4; ++kk) {
        uint64_t ad = mkdesc(start + kk * 32, 16, pitch, UMMA_SW128, bo);
        uint64_t bd = mkdesc(smem_u32(sb) + kk * 32, 16, 1024, UMMA_SW128, 0);
        umma_f16(tb, ad, bd, idesc, kk != 0);
      }
    } else {
      // A: SWIZZLE_NONE K-major, rows 16 B apart (LBO = 16: K chunk c of row r at start + 16(r+c)), SBO = 128
      const uint32_t idesc = make_idesc(GHND_F16, GHND_F16, 0, 0, 128, 64);
      for (int kk = 0; kk < 2; ++kk) {
        uint64_t ad = mkdesc(smem_u32(sa) + kk * 32, 16, 128, UMMA_SW_NONE, 0);
        uint64_t bd = mkdesc(smem_u32(sb) + kk * 32, 16, 512, UMMA_SW64, 0);
        umma_f16(tb, ad, bd, idesc, kk != 0);
      }
    }
    umma_commit(mbar);
  }
  mbar_wait(mbar, 0);
  tc_fence_after();
  uint32_t v[64];
  tmem_ld32(tb + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld32(tb + ((uint32_t)(warp * 32) << 16) + 32, v + 32);
  tmem_ld_wait();
  for (int j = 0; j < 64; ++j) p.out[(warp * 32 + lane) * 64 + j] = __uint_as_float(v[j]);
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 64); }
}

static float h2f(uint16_t h) { __half x; memcpy(&x, &h, 2); return __half2float(x); }
static uint16_t f2h(float f) { __half x = __float2half_rn(f); uint16_t h; memcpy(&h, &x, 2); return h; }

int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  srand(1);
  // ---------------- T1/T2: image [18][10][64] fp16, weights [64 n][64 k] ----------------
  std::vector<uint16_t> img(18 * 10 * 64), w(64 * 64), row(1152), w3(64 * 32);
  for (auto& x : img) x = f2h((rand() % 17 - 8) / 8.f);
  for (auto& x : w) x = f2h((rand() % 9 - 4) / 4.f);
  for (auto& x : row) x = f2h((rand() % 17 - 8) / 8.f);
  for (auto& x : w3) x = f2h((rand() % 9 - 4) / 4.f);
  uint16_t *dimg, *dw, *drow, *dw3; float* dout;
  cudaMalloc(&dimg, img.size() * 2); cudaMalloc(&dw, w.size() * 2); cudaMalloc(&drow, row.size() * 2); cudaMalloc(&dw3, w3.size() * 2);
  cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(dimg, img.data(), img.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, w.data(), w.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(drow, row.data(), row.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw3, w3.data(), w3.size() * 2, cudaMemcpyHostToDevice);
  std::vector<float> out(128 * 64);
  for (int mode = 1; mode <= 3; ++mode) {
    for (int bom = 0; bom < (mode == 3 ? 1 : 3); ++bom) {
      for (int sh = 0; sh < (mode == 3 ? 1 : 4); ++sh) {
        const int dh = sh >> 1 ? 2 : 0, dw_ = (sh & 1) ? 1 : (sh == 2 ? 2 : 0);
        P p; memset(&p, 0, sizeof(p));
        p.mode = mode; p.dh = dh; p.dw = dw_; p.base_off_mode = bom; p.out = dout;
        if (mode == 1) {
          uint64_t d[4] = {64, 10, 18, 1}, s[4] = {2, 128, 1280, 18 * 1280}; uint32_t b[4] = {64, 10, 18, 1};
          encode_tmap(&p.ta, 2, 4, dimg, d, s, b, 128);
        } else if (mode == 2) {
          uint64_t d[4] = {64, 10, 18, 1}, s[4] = {2, 128, 1280, 18 * 1280}; uint32_t b[4] = {64, 10, 1, 1};
          encode_tmap(&p.ta, 2, 4, dimg, d, s, b, 128);
        } else {
          uint64_t d[2] = {1152, 1}, s[2] = {2, 2304}; uint32_t b[2] = {1152 / 1, 1};
          // box inner dim limit is 256 elements: use {256, ...}? a raw 2304 B row needs 5 boxes; instead view as [9][128]
          uint64_t d2[2] = {128, 9}, s2[2] = {2, 256}; uint32_t b2[2] = {128, 9};
          (void)d; (void)s; (void)b;
          encode_tmap(&p.ta, 2, 2, drow, d2, s2, b2, 0);
        }
        if (mode != 3) { uint64_t d[2] = {64, 64}, s[2] = {2, 128}; uint32_t b[2] = {64, 64}; encode_tmap(&p.tb, 2, 2, dw, d, s, b, 128); }
        else { uint64_t d[2] = {32, 64}, s[2] = {2, 64}; uint32_t b[2] = {32, 64}; encode_tmap(&p.tb, 2, 2, dw3, d, s, b, 64); }
        cudaMemset(dout, 0, 128 * 64 * 4);
        k<<<1, 128, 64 * 1024>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d bom %d shift(%d,%d): CUDA error %s\n", mode, bom, dh, dw_, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0;
        for (int r = 0; r < 128; ++r)
          for (int n = 0; n < 64; ++n) {
            double ref = 0;
            if (mode != 3) {
              const int i = r / 8, j = r % 8;
              for (int c = 0; c < 64; ++c) ref += h2f(img[((i + dh) * 10 + j + dw_) * 64 + c]) * h2f(w[n * 64 + c]);
            } else {
              for (int c = 0; c < 32; ++c) ref += h2f(row[8 * r + c]) * h2f(w3[n * 32 + c]);
            }
            const double err = fabs(ref - out[r * 64 + n]);
            if (err > maxerr) maxerr = err;
            if (err > 1e-2) ++bad;
          }
        printf("mode %d base_off_mode %d shift(dh=%d,dw=%d): max err %.4g, bad %d / 8192 -> %s\n", mode, bom, dh, dw_, maxerr, bad, bad ? "MISMATCH" : "OK");
      }
    }
  }
  return 0;
}
