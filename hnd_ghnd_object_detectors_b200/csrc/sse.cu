// Fused HND/GHND sum-of-squared-errors loss: one pass reads teacher + student activations,
// writes dL/ds and per-block partial sums; a one-block finalize folds them in double.
// Reference: src/distillation/loss.py:25-34 with nn.MSELoss(reduction='sum') per term.
#include "common.cuh"

namespace ghnd {

static constexpr int kSseThreads = 256;
static constexpr int kSseBlocksPerLevel = 592;  // 148 SMs x 4

struct SseArgs {
  const uint4* t[GHND_SSE_MAX_LEVELS];
  const uint4* s[GHND_SSE_MAX_LEVELS];
  uint4* g[GHND_SSE_MAX_LEVELS];
  int64_t n8[GHND_SSE_MAX_LEVELS];  // number of 8-element vectors
  float factor[GHND_SSE_MAX_LEVELS];
  int relu_mask[GHND_SSE_MAX_LEVELS];
  int n_levels;
  int in_fmt, grad_fmt;
};

__device__ __forceinline__ float sse_vec(uint4 tv, uint4 sv, float gscale, int in_fmt, int grad_fmt,
                                         uint4* gout, bool write, bool mask) {
  uint32_t tw[4] = {tv.x, tv.y, tv.z, tv.w};
  uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
  uint32_t gw[4];
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 a = unpack2(tw[j], in_fmt), b = unpack2(sw[j], in_fmt);
    float d0 = b.x - a.x, d1 = b.y - a.y;
    acc = fmaf(d0, d0, acc);
    acc = fmaf(d1, d1, acc);
    float g0 = gscale * d0, g1 = gscale * d1;
    if (mask) {
      if (!(b.x > 0.f)) g0 = 0.f;
      if (!(b.y > 0.f)) g1 = 0.f;
    }
    gw[j] = pack2(g0, g1, grad_fmt);
  }
  if (write) *gout = make_uint4(gw[0], gw[1], gw[2], gw[3]);
  return acc;
}

__global__ void __launch_bounds__(kSseThreads)
    sse_kernel(const __grid_constant__ SseArgs a, float* __restrict__ partial) {
  const int lvl = blockIdx.y;
  const uint4* __restrict__ t = a.t[lvl];
  const uint4* __restrict__ s = a.s[lvl];
  uint4* __restrict__ g = a.g[lvl];
  const int64_t n8 = a.n8[lvl];
  const float gscale = 2.0f * a.factor[lvl];
  const bool wr = g != nullptr;
  const bool mk = a.relu_mask[lvl] != 0;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  // two independent 16-byte load pairs in flight per iteration
  for (; i + nthreads < n8; i += 2 * nthreads) {
    uint4 t0 = ld_stream(t + i), s0 = ld_stream(s + i);
    uint4 t1 = ld_stream(t + i + nthreads), s1 = ld_stream(s + i + nthreads);
    uint4 g0, g1;
    acc += sse_vec(t0, s0, gscale, a.in_fmt, a.grad_fmt, &g0, wr, mk);
    acc += sse_vec(t1, s1, gscale, a.in_fmt, a.grad_fmt, &g1, wr, mk);
    if (wr) {
      g[i] = g0;
      g[i + nthreads] = g1;
    }
  }
  if (i < n8) {
    uint4 t0 = ld_stream(t + i), s0 = ld_stream(s + i);
    uint4 g0;
    acc += sse_vec(t0, s0, gscale, a.in_fmt, a.grad_fmt, &g0, wr, mk);
    if (wr) g[i] = g0;
  }
  __shared__ float red[kSseThreads / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < kSseThreads / 32 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) partial[(size_t)lvl * gridDim.x + blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256)
    sse_finalize_kernel(const float* __restrict__ partial, int blocks_per_level, SseArgs a,
                        float* __restrict__ loss_out) {
  // one warp per level (levels run side by side; the block used to walk them one after the other with two
  // block barriers each: 9.9 us on the critical chain between the loss kernel and the first data gradient)
  __shared__ double lvl_sum[GHND_SSE_MAX_LEVELS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int lvl = warp; lvl < a.n_levels; lvl += 8) {
    double acc = 0.0;
    for (int i = lane; i < blocks_per_level; i += 32)
      acc += (double)partial[(size_t)lvl * blocks_per_level + i];
    acc = warp_sum(acc);
    if (lane == 0) lvl_sum[lvl] = acc * (double)a.factor[lvl];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int lvl = 0; lvl < a.n_levels; ++lvl) {
      tot += lvl_sum[lvl];
      loss_out[1 + lvl] = (float)lvl_sum[lvl];
    }
    loss_out[0] = (float)tot;
  }
}

}  // namespace ghnd

extern "C" {

size_t ghnd_sse_workspace_bytes(void) {
  return (size_t)GHND_SSE_MAX_LEVELS * ghnd::kSseBlocksPerLevel * sizeof(float);
}

int ghnd_sse_fwd_bwd(const ghnd_sse_level_t* levels, int n_levels, int in_fmt, int grad_fmt,
                     float* loss_out, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace ghnd;
  GHND_CHECK_ARG(levels && loss_out && workspace, "sse_fwd_bwd: null pointer");
  GHND_CHECK_ARG(n_levels >= 1 && n_levels <= GHND_SSE_MAX_LEVELS, "sse_fwd_bwd: n_levels %d",
                 n_levels);
  GHND_CHECK_ARG(workspace_bytes >= ghnd_sse_workspace_bytes(), "sse_fwd_bwd: workspace too small");
  GHND_CHECK_ARG((in_fmt == GHND_F16 || in_fmt == GHND_BF16) &&
                     (grad_fmt == GHND_F16 || grad_fmt == GHND_BF16),
                 "sse_fwd_bwd: bad format");
  SseArgs a;
  memset(&a, 0, sizeof(a));
  a.n_levels = n_levels;
  a.in_fmt = in_fmt;
  a.grad_fmt = grad_fmt;
  for (int i = 0; i < n_levels; ++i) {
    GHND_CHECK_ARG(levels[i].teacher && levels[i].student, "sse_fwd_bwd: level %d null tensor", i);
    GHND_CHECK_ARG(levels[i].n > 0 && levels[i].n % 8 == 0,
                   "sse_fwd_bwd: level %d size %lld not a positive multiple of 8", i,
                   (long long)levels[i].n);
    GHND_CHECK_ARG(((uintptr_t)levels[i].teacher % 16) == 0 &&
                       ((uintptr_t)levels[i].student % 16) == 0 &&
                       ((uintptr_t)levels[i].grad % 16) == 0,
                   "sse_fwd_bwd: level %d tensors must be 16-byte aligned", i);
    a.t[i] = (const uint4*)levels[i].teacher;
    a.s[i] = (const uint4*)levels[i].student;
    a.g[i] = (uint4*)levels[i].grad;
    a.n8[i] = levels[i].n / 8;
    a.factor[i] = levels[i].factor;
    a.relu_mask[i] = levels[i].relu_mask;
  }
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(kSseBlocksPerLevel, n_levels);
  sse_kernel<<<grid, kSseThreads, 0, st>>>(a, (float*)workspace);
  GHND_LAUNCH_CHECK("sse_kernel");
  sse_finalize_kernel<<<1, 256, 0, st>>>((const float*)workspace, kSseBlocksPerLevel, a, loss_out);
  GHND_LAUNCH_CHECK("sse_finalize_kernel");
  return GHND_OK;
}
}
