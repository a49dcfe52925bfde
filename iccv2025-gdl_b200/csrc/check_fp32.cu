// check_fp32.cu — the FP32 CHECK MODE of the encoders (north_star: "losses within 1e-4 with an FP32-accumulate
// check mode").  NOT the product path and not a fallback: DGLStep(check_fp32=True) swaps the bf16 tcgen05 encoder
// engine for this one so that the WHOLE step (orchestration, head, truncation, clip, SGD — shared with the product
// path) can be compared free-running against the fp32 reference without bf16 storage noise.
// Layout is the reference's own: fp32 NCHW activations, fp32 OIHW weights (the nn.Parameters themselves, no packing).
// Kernels are plain CUDA-core loops with fp64 accumulation and fixed summation order (deterministic).
// Reference call sites: nn.Conv2d / BatchNorm2d / ReLU / MaxPool2d models/backbone.py:20-28,44-66,97-106,
// adaptive_avg_pool2d/3d models/basic_model.py:73-82, the frame fold backbone.py:162-164.
#include "common.cuh"

namespace gdl {

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; ++i) t += sh[i];  // fixed order
  return t;
}

struct CkConv {
  int N, Ci, Hi, Wi, Co, Ho, Wo, R, S, stride, pad;
};

__global__ void ck_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y, CkConv c) {
  const int64_t total = (int64_t)c.N * c.Co * c.Ho * c.Wo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int wo = int(i % c.Wo), ho = int((i / c.Wo) % c.Ho), co = int((i / ((int64_t)c.Wo * c.Ho)) % c.Co);
    const int n = int(i / ((int64_t)c.Wo * c.Ho * c.Co));
    double acc = 0.0;
    for (int ci = 0; ci < c.Ci; ++ci) {
      const float* xp = x + ((int64_t)n * c.Ci + ci) * c.Hi * c.Wi;
      const float* wp = w + ((int64_t)co * c.Ci + ci) * c.R * c.S;
      for (int r = 0; r < c.R; ++r) {
        const int hi = ho * c.stride - c.pad + r;
        if (hi < 0 || hi >= c.Hi) continue;
        for (int s = 0; s < c.S; ++s) {
          const int wi = wo * c.stride - c.pad + s;
          if (wi < 0 || wi >= c.Wi) continue;
          acc += (double)xp[hi * c.Wi + wi] * (double)wp[r * c.S + s];
        }
      }
    }
    y[i] = (float)acc;
  }
}

__global__ void ck_conv_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ w, const float* __restrict__ add,
                                     float* __restrict__ dx, CkConv c) {
  const int64_t total = (int64_t)c.N * c.Ci * c.Hi * c.Wi;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int wi = int(i % c.Wi), hi = int((i / c.Wi) % c.Hi), ci = int((i / ((int64_t)c.Wi * c.Hi)) % c.Ci);
    const int n = int(i / ((int64_t)c.Wi * c.Hi * c.Ci));
    double acc = add ? (double)add[i] : 0.0;
    for (int r = 0; r < c.R; ++r) {
      const int hn = hi + c.pad - r;
      if (hn < 0 || hn % c.stride) continue;
      const int ho = hn / c.stride;
      if (ho >= c.Ho) continue;
      for (int s = 0; s < c.S; ++s) {
        const int wn = wi + c.pad - s;
        if (wn < 0 || wn % c.stride) continue;
        const int wo = wn / c.stride;
        if (wo >= c.Wo) continue;
        for (int co = 0; co < c.Co; ++co)
          acc += (double)dy[(((int64_t)n * c.Co + co) * c.Ho + ho) * c.Wo + wo] *
                 (double)w[(((int64_t)co * c.Ci + ci) * c.R + r) * c.S + s];
      }
    }
    dx[i] = (float)acc;
  }
}

// one block per weight element (co, ci, r, s): threads stride over the N*Ho*Wo output pixels
__global__ void ck_conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, CkConv c) {
  __shared__ double sh[32];
  const int64_t e = blockIdx.x;
  const int s = int(e % c.S), r = int((e / c.S) % c.R), ci = int((e / ((int64_t)c.S * c.R)) % c.Ci);
  const int co = int(e / ((int64_t)c.S * c.R * c.Ci));
  const int64_t P = (int64_t)c.N * c.Ho * c.Wo;
  double acc = 0.0;
  for (int64_t p = threadIdx.x; p < P; p += blockDim.x) {
    const int wo = int(p % c.Wo), ho = int((p / c.Wo) % c.Ho), n = int(p / ((int64_t)c.Wo * c.Ho));
    const int hi = ho * c.stride - c.pad + r, wi = wo * c.stride - c.pad + s;
    if (hi < 0 || hi >= c.Hi || wi < 0 || wi >= c.Wi) continue;
    acc += (double)dy[(((int64_t)n * c.Co + co) * c.Ho + ho) * c.Wo + wo] *
           (double)x[(((int64_t)n * c.Ci + ci) * c.Hi + hi) * c.Wi + wi];
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) dw[e] = (float)acc;
}

// BatchNorm2d training forward, one block per channel (x NCHW, M = N*HW values per channel)
__global__ void ck_bn_fwd_kernel(const float* __restrict__ x, const float* __restrict__ res, float* __restrict__ y, int N, int C,
                                 int HW, const float* gamma, const float* beta, float eps, float momentum,
                                 float* running_mean, float* running_var, float* mean_out, float* invstd_out, int relu,
                                 int training) {
  __shared__ double sh[32];
  const int ch = blockIdx.x;
  const int64_t M = (int64_t)N * HW;
  double mean, var;
  if (training) {
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < M; i += blockDim.x) s += (double)x[((i / HW) * C + ch) * HW + i % HW];
    mean = block_sum(s, sh) / (double)M;
    double q = 0.0;
    for (int64_t i = threadIdx.x; i < M; i += blockDim.x) {
      const double d = (double)x[((i / HW) * C + ch) * HW + i % HW] - mean;
      q += d * d;
    }
    var = block_sum(q, sh) / (double)M;
    if (threadIdx.x == 0) {
      running_mean[ch] = (float)((1.0 - momentum) * running_mean[ch] + momentum * mean);
      running_var[ch] = (float)((1.0 - momentum) * running_var[ch] + momentum * var * ((double)M / (double)(M > 1 ? M - 1 : 1)));
    }
  } else {
    mean = running_mean[ch];
    var = running_var[ch];
  }
  const double invstd = 1.0 / sqrt(var + (double)eps);
  if (threadIdx.x == 0) {
    mean_out[ch] = (float)mean;
    invstd_out[ch] = (float)invstd;
  }
  const float mf = (float)mean, isf = (float)invstd, g = gamma[ch], b = beta[ch];
  for (int64_t i = threadIdx.x; i < M; i += blockDim.x) {
    const int64_t o = ((i / HW) * C + ch) * HW + i % HW;
    float v = (x[o] - mf) * isf * g + b;
    if (res) v += res[o];
    if (relu) v = fmaxf(v, 0.f);
    y[o] = v;
  }
}

// backward of y = [relu](bn(x) [+ res]): dz = dy * (y > 0) (written to dz_out when non-null: the gradient of the
// residual branch), dgamma = sum dz*xhat, dbeta = sum dz, dx = gamma*invstd*(dz - dbeta/M - xhat*dgamma/M)
__global__ void ck_bn_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
                                 float* __restrict__ dz_out, float* __restrict__ dx, int N, int C, int HW, const float* gamma,
                                 const float* mean, const float* invstd, float* dgamma, float* dbeta, int relu) {
  __shared__ double sh[32];
  const int ch = blockIdx.x;
  const int64_t M = (int64_t)N * HW;
  const float mf = mean[ch], isf = invstd[ch];
  double s1 = 0.0, s2 = 0.0;
  for (int64_t i = threadIdx.x; i < M; i += blockDim.x) {
    const int64_t o = ((i / HW) * C + ch) * HW + i % HW;
    const float dz = (relu && !(y[o] > 0.f)) ? 0.f : dy[o];
    s1 += (double)dz;
    s2 += (double)dz * (double)((x[o] - mf) * isf);
  }
  s1 = block_sum(s1, sh);
  s2 = block_sum(s2, sh);
  if (threadIdx.x == 0) {
    dbeta[ch] = (float)s1;
    dgamma[ch] = (float)s2;
  }
  const double gi = (double)gamma[ch] * (double)isf, m1 = s1 / (double)M, m2 = s2 / (double)M;
  for (int64_t i = threadIdx.x; i < M; i += blockDim.x) {
    const int64_t o = ((i / HW) * C + ch) * HW + i % HW;
    const float dz = (relu && !(y[o] > 0.f)) ? 0.f : dy[o];
    const double xh = (double)((x[o] - mf) * isf);
    if (dz_out) dz_out[o] = dz;
    dx[o] = (float)(gi * ((double)dz - m1 - xh * m2));
  }
}

// MaxPool2d(3, 2, 1): first maximum in scan order (torch CPU/CUDA semantics), arg-max as the flat input index h*W+w
__global__ void ck_maxpool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int32_t* __restrict__ idx, int64_t planes,
                                      int H, int W, int Ho, int Wo) {
  const int64_t total = planes * Ho * Wo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int wo = int(i % Wo), ho = int((i / Wo) % Ho);
    const int64_t pl = i / ((int64_t)Wo * Ho);
    const float* xp = x + pl * H * W;
    float best = -INFINITY;
    int bi = -1;
    for (int r = 0; r < 3; ++r) {
      const int h = ho * 2 - 1 + r;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int w = wo * 2 - 1 + s;
        if (w < 0 || w >= W) continue;
        const float v = xp[h * W + w];
        if (v > best || bi < 0) {
          best = v;
          bi = h * W + w;
        }
      }
    }
    y[i] = best;
    idx[i] = bi;
  }
}
__global__ void ck_maxpool_bwd_kernel(const float* __restrict__ dy, const int32_t* __restrict__ idx, float* __restrict__ dx,
                                      int64_t planes, int H, int W, int Ho, int Wo) {
  const int64_t total = planes * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = int(i % W), h = int((i / W) % H);
    const int64_t pl = i / ((int64_t)W * H);
    const int me = h * W + w;
    float acc = 0.f;
    for (int ho = (h + 1) / 2 - 1; ho <= (h + 1) / 2; ++ho) {  // windows with 2*ho-1 <= h <= 2*ho+1
      if (ho < 0 || ho >= Ho || h < 2 * ho - 1 || h > 2 * ho + 1) continue;
      for (int wo = (w + 1) / 2 - 1; wo <= (w + 1) / 2; ++wo) {
        if (wo < 0 || wo >= Wo || w < 2 * wo - 1 || w > 2 * wo + 1) continue;
        const int64_t o = (pl * Ho + ho) * Wo + wo;
        if (idx[o] == me) acc += dy[o];
      }
    }
    dx[i] = acc;
  }
}

// x [B*T][C][HW] -> out [B][C] = mean over (t, hw)   (adaptive_avg_pool2d / 3d after the (B,T) un-fold)
__global__ void ck_gap_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int B, int T, int C, int HW) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int ch = i % C, b = i / C;
  double s = 0.0;
  for (int t = 0; t < T; ++t) {
    const float* xp = x + ((int64_t)(b * T + t) * C + ch) * HW;
    for (int p = 0; p < HW; ++p) s += (double)xp[p];
  }
  out[i] = (float)(s / (double)(T * HW));
}
__global__ void ck_gap_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dx, int B, int T, int C, int HW) {
  const int64_t total = (int64_t)B * T * C * HW;
  const float inv = 1.f / (float)(T * HW);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = int((i / HW) % C);
    const int b = int(i / ((int64_t)HW * C * T));
    dx[i] = dout[b * C + ch] * inv;
  }
}
// src [B][C][T][HW] -> dst [B*T][C][HW]   (backbone.py:162-164 permute + contiguous + view)
__global__ void ck_fold_frames_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int T, int64_t HW) {
  const int64_t total = (int64_t)B * C * T * HW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i % HW;
    const int ch = int((i / HW) % C), t = int((i / (HW * C)) % T), b = int(i / (HW * C * T));
    dst[i] = src[(((int64_t)b * C + ch) * T + t) * HW + p];
  }
}

static inline int ck_grid(int64_t total, int threads) {
  int64_t g = (total + threads - 1) / threads;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
static inline CkConv ck_desc(const gdl_conv_desc* d, int ci_real) {
  CkConv c = {d->N, ci_real, d->Hi, d->Wi, d->Co, d->Ho, d->Wo, d->R, d->S, d->stride, d->pad};
  return c;
}

}  // namespace gdl

using namespace gdl;

extern "C" int gdl_check_conv_fwd(const gdl_conv_desc* d, int ci_real, const float* x, const float* w, float* y,
                                  gdl_stream_t s) {
  GDL_REQUIRE(d && x && w && y && ci_real > 0, "gdl_check_conv_fwd: null pointer");
  const CkConv c = ck_desc(d, ci_real);
  ck_conv_fwd_kernel<<<ck_grid((int64_t)c.N * c.Co * c.Ho * c.Wo, 256), 256, 0, (cudaStream_t)s>>>(x, w, y, c);
  GDL_CHECK_LAUNCH("ck_conv_fwd_kernel");
  return GDL_OK;
}
extern "C" int gdl_check_conv_dgrad(const gdl_conv_desc* d, int ci_real, const float* dy, const float* w, const float* add,
                                    float* dx, gdl_stream_t s) {
  GDL_REQUIRE(d && dy && w && dx && ci_real > 0, "gdl_check_conv_dgrad: null pointer");
  const CkConv c = ck_desc(d, ci_real);
  ck_conv_dgrad_kernel<<<ck_grid((int64_t)c.N * c.Ci * c.Hi * c.Wi, 256), 256, 0, (cudaStream_t)s>>>(dy, w, add, dx, c);
  GDL_CHECK_LAUNCH("ck_conv_dgrad_kernel");
  return GDL_OK;
}
extern "C" int gdl_check_conv_wgrad(const gdl_conv_desc* d, int ci_real, const float* x, const float* dy, float* dw,
                                    gdl_stream_t s) {
  GDL_REQUIRE(d && x && dy && dw && ci_real > 0, "gdl_check_conv_wgrad: null pointer");
  const CkConv c = ck_desc(d, ci_real);
  const int64_t elems = (int64_t)c.Co * c.Ci * c.R * c.S;
  GDL_REQUIRE(elems < ((int64_t)1 << 31), "gdl_check_conv_wgrad: too many weight elements");
  ck_conv_wgrad_kernel<<<(unsigned)elems, 128, 0, (cudaStream_t)s>>>(x, dy, dw, c);
  GDL_CHECK_LAUNCH("ck_conv_wgrad_kernel");
  return GDL_OK;
}
extern "C" int gdl_check_bn_fwd(const float* x, const float* res, float* y, int N, int C, int HW, const float* gamma,
                                const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                                float* mean, float* invstd, int relu, int training, gdl_stream_t s) {
  GDL_REQUIRE(x && y && gamma && beta && running_mean && running_var && mean && invstd, "gdl_check_bn_fwd: null pointer");
  ck_bn_fwd_kernel<<<C, 512, 0, (cudaStream_t)s>>>(x, res, y, N, C, HW, gamma, beta, eps, momentum, running_mean, running_var,
                                                   mean, invstd, relu, training);
  GDL_CHECK_LAUNCH("ck_bn_fwd_kernel");
  return GDL_OK;
}
extern "C" int gdl_check_bn_bwd(const float* dy, const float* y, const float* x, float* dz, float* dx, int N, int C, int HW,
                                const float* gamma, const float* mean, const float* invstd, float* dgamma, float* dbeta,
                                int relu, gdl_stream_t s) {
  GDL_REQUIRE(dy && x && dx && gamma && mean && invstd && dgamma && dbeta && (y || !relu), "gdl_check_bn_bwd: null pointer");
  ck_bn_bwd_kernel<<<C, 512, 0, (cudaStream_t)s>>>(dy, y, x, dz, dx, N, C, HW, gamma, mean, invstd, dgamma, dbeta, relu);
  GDL_CHECK_LAUNCH("ck_bn_bwd_kernel");
  return GDL_OK;
}
extern "C" int gdl_check_maxpool_fwd(const float* x, float* y, int32_t* argmax, int64_t planes, int H, int W, int Ho, int Wo,
                                     gdl_stream_t s) {
  GDL_REQUIRE(x && y && argmax, "gdl_check_maxpool_fwd: null pointer");
  ck_maxpool_fwd_kernel<<<ck_grid(planes * Ho * Wo, 256), 256, 0, (cudaStream_t)s>>>(x, y, argmax, planes, H, W, Ho, Wo);
  GDL_CHECK_LAUNCH("ck_maxpool_fwd_kernel");
  return GDL_OK;
}
extern "C" int gdl_check_maxpool_bwd(const float* dy, const int32_t* argmax, float* dx, int64_t planes, int H, int W, int Ho,
                                     int Wo, gdl_stream_t s) {
  GDL_REQUIRE(dy && dx && argmax, "gdl_check_maxpool_bwd: null pointer");
  ck_maxpool_bwd_kernel<<<ck_grid(planes * H * W, 256), 256, 0, (cudaStream_t)s>>>(dy, argmax, dx, planes, H, W, Ho, Wo);
  GDL_CHECK_LAUNCH("ck_maxpool_bwd_kernel");
  return GDL_OK;
}
extern "C" int gdl_check_gap_fwd(const float* x, float* out, int B, int T, int C, int HW, gdl_stream_t s) {
  GDL_REQUIRE(x && out, "gdl_check_gap_fwd: null pointer");
  ck_gap_fwd_kernel<<<(B * C + 127) / 128, 128, 0, (cudaStream_t)s>>>(x, out, B, T, C, HW);
  GDL_CHECK_LAUNCH("ck_gap_fwd_kernel");
  return GDL_OK;
}
extern "C" int gdl_check_gap_bwd(const float* dout, float* dx, int B, int T, int C, int HW, gdl_stream_t s) {
  GDL_REQUIRE(dout && dx, "gdl_check_gap_bwd: null pointer");
  ck_gap_bwd_kernel<<<ck_grid((int64_t)B * T * C * HW, 256), 256, 0, (cudaStream_t)s>>>(dout, dx, B, T, C, HW);
  GDL_CHECK_LAUNCH("ck_gap_bwd_kernel");
  return GDL_OK;
}
extern "C" int gdl_check_fold_frames(const float* src, float* dst, int B, int C, int T, int H, int W, gdl_stream_t s) {
  GDL_REQUIRE(src && dst, "gdl_check_fold_frames: null pointer");
  ck_fold_frames_kernel<<<ck_grid((int64_t)B * C * T * H * W, 256), 256, 0, (cudaStream_t)s>>>(src, dst, B, C, T, (int64_t)H * W);
  GDL_CHECK_LAUNCH("ck_fold_frames_kernel");
  return GDL_OK;
}
