// optim.cu — gradient clipping statistics, DGL gradient diagnostics and SGD-momentum over
// the flat fp32 parameter arena (reference main_dgl.py:129 clip_grad_norm_(40, L2),
// :132-143 sum_p mean|grad_p| per encoder, :154/:249 torch.optim.SGD(momentum, weight_decay)).
// One pass over the gradients produces the global L2 norm and both diagnostics; the clip
// coefficient stays on the device and is folded into the SGD kernel, which also writes the
// clipped gradient back (clip_grad_norm_ is in-place in the reference).
#include "common.cuh"
#include "api_version.h"

namespace gdl {

constexpr int kChunk = 8192;  // elements per block: fixed => deterministic partials
constexpr int kOptThreads = 256;

__device__ __forceinline__ int find_segment(const int64_t* __restrict__ seg_end, int nseg, int64_t i) {
  int lo = 0, hi = nseg - 1;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (i < seg_end[mid]) hi = mid; else lo = mid + 1;
  }
  return lo;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x < 32) {
    t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
  }
  return t;  // valid in warp 0
}

__global__ void __launch_bounds__(kOptThreads) grad_stats_kernel(
    const float* __restrict__ grad, int64_t numel, const int64_t* __restrict__ seg_end,
    const int32_t* __restrict__ seg_group, const float* __restrict__ seg_inv_numel, int nseg,
    float* __restrict__ partial) {
  pdl_enter();
  __shared__ float red[32];
  __shared__ int s_lo, s_hi;
  const int64_t start = (int64_t)blockIdx.x * kChunk;
  const int64_t end = start + kChunk < numel ? start + kChunk : numel;
  if (threadIdx.x == 0) {
    s_lo = find_segment(seg_end, nseg, start);
    s_hi = find_segment(seg_end, nseg, end - 1);
  }
  __syncthreads();
  const int lo = s_lo, hi = s_hi;
  float sq = 0.f, wa = 0.f, wv = 0.f;
  for (int64_t i = start + threadIdx.x; i < end; i += kOptThreads) {
    float g = grad[i];
    sq = fmaf(g, g, sq);
    int sgm = lo;
    if (lo != hi) {
      while (sgm < hi && i >= seg_end[sgm]) ++sgm;
    }
    int grp = seg_group[sgm];
    float w = fabsf(g) * seg_inv_numel[sgm];
    if (grp == 0) wa += w;
    else if (grp == 1) wv += w;
  }
  float t0 = block_sum(sq, red);
  float t1 = block_sum(wa, red);
  float t2 = block_sum(wv, red);
  if (threadIdx.x == 0) {
    partial[(int64_t)blockIdx.x * 3 + 0] = t0;
    partial[(int64_t)blockIdx.x * 3 + 1] = t1;
    partial[(int64_t)blockIdx.x * 3 + 2] = t2;
  }
}

__global__ void grad_stats_finalize_kernel(const float* __restrict__ partial, int nblk, float max_norm,
                                           float* __restrict__ stats) {
  pdl_enter();
  // single block; fixed-order strided accumulation in double, then a fixed tree
  __shared__ double red[3][256];
  double a0 = 0, a1 = 0, a2 = 0;
  for (int b = threadIdx.x; b < nblk; b += 256) {
    a0 += partial[(int64_t)b * 3 + 0];
    a1 += partial[(int64_t)b * 3 + 1];
    a2 += partial[(int64_t)b * 3 + 2];
  }
  red[0][threadIdx.x] = a0; red[1][threadIdx.x] = a1; red[2][threadIdx.x] = a2;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      red[0][threadIdx.x] += red[0][threadIdx.x + s];
      red[1][threadIdx.x] += red[1][threadIdx.x + s];
      red[2][threadIdx.x] += red[2][threadIdx.x + s];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    float norm = (float)sqrt(red[0][0]);
    float coef = max_norm / (norm + 1e-6f);  // torch.nn.utils.clip_grad_norm_ formula
    if (coef > 1.f) coef = 1.f;
    stats[0] = norm;
    stats[1] = coef;
    stats[2] = (float)red[1][0] * coef;
    stats[3] = (float)red[2][0] * coef;
  }
}

__global__ void __launch_bounds__(256) sgd_momentum_kernel(float* __restrict__ param,
                                                           float* __restrict__ grad,
                                                           float* __restrict__ buf, int64_t n4,
                                                           int64_t numel, float lr, float mu,
                                                           float wd, int first,
                                                           const float* __restrict__ stats) {
  pdl_enter();
  const float coef = stats ? stats[1] : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = reinterpret_cast<float4*>(param)[i];
    float4 g = reinterpret_cast<float4*>(grad)[i];
    float4 m = first ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(buf)[i];
    float* pp = &p.x; float* gg = &g.x; float* mm = &m.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gc = gg[k] * coef;
      gg[k] = gc;
      float d = fmaf(wd, pp[k], gc);
      mm[k] = first ? d : fmaf(mu, mm[k], d);
      pp[k] = fmaf(-lr, mm[k], pp[k]);
    }
    reinterpret_cast<float4*>(param)[i] = p;
    reinterpret_cast<float4*>(grad)[i] = g;
    reinterpret_cast<float4*>(buf)[i] = m;
  }
  // tail (numel not a multiple of 4)
  if (blockIdx.x == 0 && threadIdx.x < (numel & 3)) {
    int64_t i = n4 * 4 + threadIdx.x;
    float gc = grad[i] * coef;
    grad[i] = gc;
    float d = fmaf(wd, param[i], gc);
    float m = first ? d : fmaf(mu, buf[i], d);
    buf[i] = m;
    param[i] = fmaf(-lr, m, param[i]);
  }
}

static char g_last_error[512] = "";

void set_last_error(const char* msg) { snprintf(g_last_error, sizeof(g_last_error), "%s", msg); }
int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", where, cudaGetErrorString(e));
  return GDL_ECUDA;
}

}  // namespace gdl

using namespace gdl;

extern "C" int gdl_version(void) { return GDL_B200_VERSION; }
extern "C" const char* gdl_last_error_string(void) { return g_last_error; }

extern "C" int gdl_init(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
  if (prop.major != 10) {
    snprintf(g_last_error, sizeof(g_last_error),
             "gdl_b200 needs an sm_100 (B200) device; device %d is sm_%d%d", device, prop.major, prop.minor);
    return GDL_EARCH;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
  return GDL_OK;
}

extern "C" int64_t gdl_optim_scratch_floats(int64_t numel, int nseg) {
  (void)nseg;
  return ceil_div64(numel, kChunk) * 3;
}

extern "C" int gdl_grad_stats(const float* grad, int64_t numel, const int64_t* seg_end,
                              const int32_t* seg_group, const float* seg_inv_numel, int nseg,
                              float max_norm, float* scratch, float* stats_out, gdl_stream_t s) {
  GDL_REQUIRE(grad && seg_end && seg_group && seg_inv_numel && scratch && stats_out, "gdl_grad_stats: null pointer");
  GDL_REQUIRE(numel > 0 && nseg > 0, "gdl_grad_stats: bad shape");
  int nblk = (int)ceil_div64(numel, kChunk);
  launch_pdl(grad_stats_kernel, nblk, kOptThreads, 0, (cudaStream_t)s, grad, numel, seg_end, seg_group,
                                                              seg_inv_numel, nseg, scratch);
  GDL_CHECK_LAUNCH("grad_stats_kernel");
  launch_pdl(grad_stats_finalize_kernel, 1, 256, 0, (cudaStream_t)s, scratch, nblk, max_norm, stats_out);
  GDL_CHECK_LAUNCH("grad_stats_finalize_kernel");
  return GDL_OK;
}

extern "C" int gdl_sgd_momentum(float* param, float* grad, float* momentum_buf, int64_t numel,
                                float lr, float mu, float wd, int first_step, const float* stats,
                                gdl_stream_t s) {
  GDL_REQUIRE(param && grad && momentum_buf && numel > 0, "gdl_sgd_momentum: bad arguments");
  GDL_REQUIRE((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)momentum_buf) & 15) == 0,
              "gdl_sgd_momentum: arenas must be 16-byte aligned");
  int64_t n4 = numel / 4;
  int64_t blocks = ceil_div64(n4 > 0 ? n4 : 1, 256);
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  launch_pdl(sgd_momentum_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)s, param, grad, momentum_buf, n4, numel,
                                                                    lr, mu, wd, first_step, stats);
  GDL_CHECK_LAUNCH("sgd_momentum_kernel");
  return GDL_OK;
}
