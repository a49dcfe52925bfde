// film.cu — FiLM_DGL head (reference models/fusion_modules.py:126-178; despite the name an
// OUTER-PRODUCT head: fc: Linear(512*512 -> 512) applied to a (x) v, a (x) a and v (x) v, then
// fc_out: Linear(512 -> n)).  The 134 M-parameter fc is three dense contractions per step:
//
//   H  [3B, 512]     = Z  [3B, 262144] * W1^T          (forward, all three branches at once)
//   dW1[512, 262144] = dH_f^T [512, B] * Z_f [B, 262144]   (Lf only — the unimodal head gradient is wiped)
//   G  [2B, 262144]  = dH_{a,v} [2B, 512] * W1          (unimodal branches only — Lf saw detached features)
//
// They run on the tcgen05 flat-window kernels through two GEMM entry points:
//   gdl_gemm_nt_bf16 : C[M,N] bf16 = A[M,K] * B[N,K]^T   (conv_flat.cu single-tap mode; M = "pixels")
//   gdl_gemm_tn_f32  : C[M,N] f32  = At[K,M]^T * Bt[K,N] (conv_wgrad_flat.cu 1x1 mode; K = "pixels", split-K)
// with everything stored FEATURE-MAJOR so that both are natural: Zt [262144][ZB] (batch contiguous) and the
// bf16 shadow W1t [262144][512].  The rest are small HBM-bound kernels: the outer products, the
// contraction of G with a / v, transposes between the fp32 [512][262144] parameter layout and W1t.
#include "common.cuh"

namespace gdl {

int try_conv_flat(int kind, int N, int Hs, int Ws, int Cs, int64_t sW, int64_t sH, int64_t sN, const void* src,
                  const void* wt, int64_t wt_rows, int64_t wt_k, void* dst, int Hd, int Wd, int Cd,
                  const void* add_src, int add_mode, cudaStream_t s, float* stats = nullptr,
                  int* stats_rows = nullptr, const float* bias = nullptr, int relu = 0);
int64_t wgrad_flat_workspace_bytes(int N, int Ho, int Wo, int Ci, int Co, int R, int stride);
int try_wgrad_flat(int N, int Hi, int Wi, int Ho, int Wo, int Ci, int Co, int R, int stride, const void* x,
                   const void* dy, float* partial, int64_t workspace_bytes, cudaStream_t s, int transposed,
                   int* tap_splits);

constexpr int kGemmW = 128;  // the GEMM row index is folded into an (H, 128) "image"

// C[m][n] = sum over splits of partial[sp][m][n] (+ bias[n]); fixed order => deterministic
__global__ void gemm_tn_reduce_kernel(const float* __restrict__ partial, int splits, int64_t MN, int N,
                                      const float* __restrict__ bias, float* __restrict__ C, int accumulate) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= MN) return;
  // accumulate: C already holds the sum over the earlier K chunks (launch order = summation order: deterministic)
  float4 acc = accumulate ? *reinterpret_cast<const float4*>(C + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int sp = 0; sp < splits; ++sp) {
    const float4 v = *reinterpret_cast<const float4*>(partial + (size_t)sp * MN + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  if (bias != nullptr) {
    const int n = int(i % N);
    acc.x += bias[n]; acc.y += bias[n + 1]; acc.z += bias[n + 2]; acc.w += bias[n + 3];
  }
  *reinterpret_cast<float4*>(C + i) = acc;
}

// Zt[f = i*D + j][col]: col < B: a_i v_j (multimodal, detached) | B..2B: a_i a_j | 2B..3B: v_i v_j | rest 0.
// Step 1 transposes the features to [D][ZB-padded batch] so that step 2 reads them along the batch index.
__global__ void film_transpose_feat_kernel(const float* __restrict__ a, const float* __restrict__ v,
                                           float* __restrict__ at, float* __restrict__ vt, int B, int D, int Bp) {
  const int64_t total = (int64_t)D * Bp;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = int(idx / Bp), b = int(idx - (int64_t)i * Bp);
    at[idx] = b < B ? a[(int64_t)b * D + i] : 0.f;
    vt[idx] = b < B ? v[(int64_t)b * D + i] : 0.f;
  }
}
// Rows [f0, f0 + nf) of the feature-major matrix are written to Zt[0 .. nf) (a K chunk: see gdl_film_outer_chunk).
__global__ void __launch_bounds__(256) film_outer_kernel(const float* __restrict__ at, const float* __restrict__ vt,
                                                         bf16* __restrict__ Zt, int B, int D, int ZB, int Bp,
                                                         int variants, int64_t f0, int64_t nf) {
  const int groups = ZB / 8;
  const int64_t total = nf * groups;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int g = int(idx % groups);
    const int64_t fl = idx / groups, f = f0 + fl;
    const int i = int(f / D), j = int(f - (int64_t)i * D);
    const float* ai = at + (int64_t)i * Bp;
    const float* aj = at + (int64_t)j * Bp;
    const float* vi = vt + (int64_t)i * Bp;
    const float* vj = vt + (int64_t)j * Bp;
    float o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int col = g * 8 + c;
      const int var = col / B, b = col - var * B;
      float val = 0.f;
      if (var < variants) val = var == 0 ? ai[b] * vj[b] : (var == 1 ? ai[b] * aj[b] : vi[b] * vj[b]);
      o[c] = val;
    }
    *reinterpret_cast<uint4*>(Zt + fl * ZB + g * 8) = pack8(o);
  }
}

// dst bf16 [drows][dcols] (zero padded) from fp32 sources: rows [0,r0) from src0, [r0, r0+r1) from src1
// (each [r][cols] with row stride ld); transpose: dst[c][r] instead of dst[r][c].
__global__ void cast_pad_kernel(const float* __restrict__ src0, int r0, const float* __restrict__ src1, int r1, int cols,
                                int ld, int transpose, bf16* __restrict__ dst, int drows, int dcols) {
  const int64_t total = (int64_t)drows * dcols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int dr = int(idx / dcols), dc = int(idx - (int64_t)dr * dcols);
    const int r = transpose ? dc : dr, c = transpose ? dr : dc;
    float val = 0.f;
    if (c < cols) {
      if (r < r0) val = src0[(int64_t)r * ld + c];
      else if (r < r0 + r1) val = src1[(int64_t)(r - r0) * ld + c];
    }
    dst[idx] = __float2bfloat16_rn(val);
  }
}

// G [D*D][ldg] bf16 (column = batch row): dx[b][i] = sum_j G[i*D+j][c0+b] * y[b][j],
//                                        dy[b][j] = sum_i G[i*D+j][c0+b] * x[b][i]
// sum_mode 1: dx <- dx + dy (x and y are the same tensor: a (x) a), dy not written.
// One block per output index t (i for dx, j for dy).  Vector path (B, c0, ldg multiples of 8): a thread owns 8
// batch columns (one 16-byte load per G row) and one of KS interleaved k slices; slices are combined in a
// fixed order through shared memory.  xt / yt are the features transposed to [D][Bp].
constexpr int kContractThreads = 256;
// G holds the rows of i in [i0, i0 + ni) only (row (i - i0)*D + j): block t adds its part of both contractions to the
// outputs — dx[b][t] (+)= sum_j G[(t - i0)*D + j] y[b][j] when t lies in the chunk (complete: all j are in the chunk),
// dy[b][t] (+)= sum_{i in chunk} G[(i - i0)*D + t] x[b][i] — the first chunk writes (acc = 0), the later ones add
// (launch order = summation order: deterministic).  Whole matrix: i0 = 0, ni = D, acc = 0.
__global__ void __launch_bounds__(kContractThreads) film_contract_kernel(
    const bf16* __restrict__ G, int ldg, int c0, const float* __restrict__ xt, const float* __restrict__ yt,
    float* __restrict__ dx, float* __restrict__ dy, int B, int D, int Bp, int sum_mode, int i0, int ni, int acc) {
  __shared__ float red[2][kContractThreads][8];
  const int t = blockIdx.x;
  const int bgroups = B / 8;                       // vector path only
  const int KS = kContractThreads / bgroups;       // k slices
  const int bq = threadIdx.x % bgroups, ks = threadIdx.x / bgroups;
  const bool own = t >= i0 && t < i0 + ni;
  float sx[8], sy[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) sx[c] = sy[c] = 0.f;
  if (ks < KS) {
    // four G rows in flight per thread (the loop is a chain of L2 / DRAM latencies otherwise)
    auto fma8 = [&](const uint4& gu, const float* f, float (&acc)[8]) {
      float g[8];
      unpack8(gu, g);
      const float4 f0 = *reinterpret_cast<const float4*>(f), f1 = *reinterpret_cast<const float4*>(f + 4);
      const float fv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = fmaf(g[c], fv[c], acc[c]);
    };
    if (own) {
      const bf16* g0 = G + (int64_t)(t - i0) * D * ldg + c0 + bq * 8;
      int k = ks;
      for (; k + 3 * KS < D; k += 4 * KS) {
        uint4 u[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) u[q] = ld_stream16(g0 + (int64_t)(k + q * KS) * ldg);
#pragma unroll
        for (int q = 0; q < 4; ++q) fma8(u[q], yt + (int64_t)(k + q * KS) * Bp + bq * 8, sx);
      }
      for (; k < D; k += KS) fma8(ld_stream16(g0 + (int64_t)k * ldg), yt + (int64_t)k * Bp + bq * 8, sx);
    }
    {
      const bf16* g0 = G + (int64_t)t * ldg + c0 + bq * 8;
      int k = ks;
      for (; k + 3 * KS < ni; k += 4 * KS) {
        uint4 u[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) u[q] = ld_stream16(g0 + (int64_t)(k + q * KS) * D * ldg);
#pragma unroll
        for (int q = 0; q < 4; ++q) fma8(u[q], xt + (int64_t)(i0 + k + q * KS) * Bp + bq * 8, sy);
      }
      for (; k < ni; k += KS) fma8(ld_stream16(g0 + (int64_t)k * D * ldg), xt + (int64_t)(i0 + k) * Bp + bq * 8, sy);
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    red[0][threadIdx.x][c] = sx[c];
    red[1][threadIdx.x][c] = sy[c];
  }
  __syncthreads();
  if (ks == 0) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float ax = 0.f, ay = 0.f;
      for (int s2 = 0; s2 < KS; ++s2) {
        ax += red[0][s2 * bgroups + bq][c];
        ay += red[1][s2 * bgroups + bq][c];
      }
      const int b = bq * 8 + c;
      float* px = dx + (int64_t)b * D + t;
      if (sum_mode) {
        *px = acc ? *px + (ax + ay) : ax + ay;
      } else {
        float* py = dy + (int64_t)b * D + t;
        if (!acc || own) *px = acc ? *px + ax : ax;  // ax == 0 outside the chunk
        *py = acc ? *py + ay : ay;
      }
    }
  }
}
// scalar path for batches that are not a multiple of 8 (same chunk semantics)
__global__ void __launch_bounds__(256) film_contract_scalar_kernel(const bf16* __restrict__ G, int ldg, int c0,
                                                                   const float* __restrict__ xt,
                                                                   const float* __restrict__ yt, float* __restrict__ dx,
                                                                   float* __restrict__ dy, int B, int D, int Bp,
                                                                   int sum_mode, int i0, int ni, int acc) {
  const int t = blockIdx.x;
  const bool own = t >= i0 && t < i0 + ni;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float sx = 0.f, sy = 0.f;
    const bf16* gcol = G + c0 + b;
    if (own)
      for (int k = 0; k < D; ++k)
        sx = fmaf(__bfloat162float(gcol[((int64_t)(t - i0) * D + k) * ldg]), yt[(int64_t)k * Bp + b], sx);
    for (int k = 0; k < ni; ++k)
      sy = fmaf(__bfloat162float(gcol[((int64_t)k * D + t) * ldg]), xt[(int64_t)(i0 + k) * Bp + b], sy);
    float* px = dx + (int64_t)b * D + t;
    if (sum_mode) {
      *px = acc ? *px + (sx + sy) : sx + sy;
    } else {
      float* py = dy + (int64_t)b * D + t;
      if (!acc || own) *px = acc ? *px + sx : sx;
      *py = acc ? *py + sy : sy;
    }
  }
}

// 32x32 tiled transposes between the fp32 parameter layout [R][Cn] and the bf16 feature-major shadow [Cn][R]
// src is a column window of a row-major fp32 matrix with row stride lds: src[r][c] = src[r * lds + c], c in [0, Cn)
__global__ void transpose_f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int R, int64_t Cn,
                                             int64_t lds) {
  __shared__ float tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32;
  const int r0 = blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += 8) tile[k][threadIdx.x] = src[(int64_t)(r0 + k) * lds + c0 + threadIdx.x];
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += 8)
    dst[(c0 + k) * R + r0 + threadIdx.x] = __float2bfloat16_rn(tile[threadIdx.x][k]);
}
__global__ void transpose_bf16_to_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int R, int64_t Cn,
                                             int64_t ldd) {
  // src [Cn][R] -> dst [R][Cn] inside a row-major matrix with row stride ldd
  __shared__ float tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32;
  const int r0 = blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += 8)
    tile[k][threadIdx.x] = __bfloat162float(src[(c0 + k) * R + r0 + threadIdx.x]);
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += 8) dst[(int64_t)(r0 + k) * ldd + c0 + threadIdx.x] = tile[threadIdx.x][k];
}

}  // namespace gdl

using namespace gdl;

extern "C" int gdl_gemm_nt_bf16(const void* A, int64_t lda, const void* B, void* C, int64_t M, int N, int K,
                                gdl_stream_t s) {
  GDL_REQUIRE(A && B && C, "gdl_gemm_nt_bf16: null pointer");
  GDL_REQUIRE(M > 0 && M % kGemmW == 0 && N % 64 == 0 && K % 64 == 0 && lda >= K && lda % 8 == 0,
              "gdl_gemm_nt_bf16: need M % 128 == 0, N % 64 == 0, K % 64 == 0");
  const int H = int(M / kGemmW);
  int rc = try_conv_flat(3, 1, H, kGemmW, K, lda, (int64_t)kGemmW * lda, M * lda, A, B, N, K, C, H, kGemmW, N, nullptr,
                         0, (cudaStream_t)s);
  if (rc < 0) return rc;
  if (rc == 0) {
    set_last_error("gdl_gemm_nt_bf16: shape not supported by the flat kernel");
    return GDL_EINVAL;
  }
  return GDL_OK;
}

extern "C" int64_t gdl_gemm_tn_workspace_bytes(int M, int N, int64_t K) {
  if (K <= 0 || K % kGemmW != 0) return GDL_EINVAL;
  int64_t b = wgrad_flat_workspace_bytes(1, int(K / kGemmW), kGemmW, M, N, 1, 1);
  return b > 0 ? b : GDL_EINVAL;
}

static int gemm_tn_impl(const void* At, const void* Bt, const float* bias, float* C, int M, int N, int64_t K,
                        void* workspace, int64_t workspace_bytes, int accumulate, gdl_stream_t s) {
  GDL_REQUIRE(At && Bt && C && workspace, "gdl_gemm_tn_f32: null pointer");
  GDL_REQUIRE(K > 0 && K % kGemmW == 0 && N % 4 == 0, "gdl_gemm_tn_f32: need K % 128 == 0");
  const int H = int(K / kGemmW);
  int ns = try_wgrad_flat(1, H, kGemmW, H, kGemmW, M, N, 1, 1, At, Bt, (float*)workspace, workspace_bytes,
                          (cudaStream_t)s, 0, nullptr);
  if (ns < 0) return ns;
  if (ns == 0) {
    set_last_error("gdl_gemm_tn_f32: shape not supported (M, N multiples of 128 or M == 64) or workspace too small");
    return GDL_EINVAL;
  }
  const int64_t MN = (int64_t)M * N;
  gemm_tn_reduce_kernel<<<(unsigned)ceil_div64(MN / 4, 256), 256, 0, (cudaStream_t)s>>>((const float*)workspace, ns, MN,
                                                                                      N, bias, C, accumulate);
  GDL_CHECK_LAUNCH("gemm_tn_reduce_kernel");
  return GDL_OK;
}

extern "C" int gdl_gemm_tn_f32(const void* At, const void* Bt, const float* bias, float* C, int M, int N, int64_t K,
                               void* workspace, int64_t workspace_bytes, gdl_stream_t s) {
  return gemm_tn_impl(At, Bt, bias, C, M, N, K, workspace, workspace_bytes, 0, s);
}

extern "C" int gdl_gemm_tn_f32_acc(const void* At, const void* Bt, float* C, int M, int N, int64_t K, void* workspace,
                                   int64_t workspace_bytes, gdl_stream_t s) {
  return gemm_tn_impl(At, Bt, nullptr, C, M, N, K, workspace, workspace_bytes, 1, s);
}

static int film_bp(int B) { return (B + 7) / 8 * 8; }

extern "C" int64_t gdl_film_scratch_floats(int B, int D) { return 2 * (int64_t)D * film_bp(B); }

static int film_transpose(const float* a, const float* v, float* scratch, int B, int D, cudaStream_t s) {
  const int Bp = film_bp(B);
  const int64_t total = (int64_t)D * Bp;
  film_transpose_feat_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(a, v, scratch, scratch + total, B, D, Bp);
  GDL_CHECK_LAUNCH("film_transpose_feat_kernel");
  return GDL_OK;
}

static int film_outer_impl(const float* a, const float* v, void* Zt, int B, int D, int ZB, int variants, float* scratch,
                           int64_t f0, int64_t nf, bool transpose, gdl_stream_t s) {
  GDL_REQUIRE(a && v && Zt && scratch && B > 0 && D > 0 && ZB % 8 == 0 && variants >= 1 && variants <= 3 &&
                  variants * B <= ZB && f0 >= 0 && nf > 0 && f0 + nf <= (int64_t)D * D,
              "gdl_film_outer: bad arguments");
  if (transpose) {
    int rc = film_transpose(a, v, scratch, B, D, (cudaStream_t)s);
    if (rc != GDL_OK) return rc;
  }
  const int Bp = film_bp(B);
  const int64_t total = nf * (ZB / 8);
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  film_outer_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(scratch, scratch + (int64_t)D * Bp, (bf16*)Zt, B, D, ZB,
                                                                  Bp, variants, f0, nf);
  GDL_CHECK_LAUNCH("film_outer_kernel");
  return GDL_OK;
}

extern "C" int gdl_film_outer(const float* a, const float* v, void* Zt, int B, int D, int ZB, int variants,
                              float* scratch, gdl_stream_t s) {
  return film_outer_impl(a, v, Zt, B, D, ZB, variants, scratch, 0, (int64_t)D * D, true, s);
}

extern "C" int gdl_film_outer_chunk(const float* a, const float* v, void* Zt_chunk, int B, int D, int ZB, int variants,
                                    float* scratch, int64_t f0, int64_t nf, gdl_stream_t s) {
  // the transposed features in scratch are (re)built by the chunk that starts at feature 0
  return film_outer_impl(a, v, Zt_chunk, B, D, ZB, variants, scratch, f0, nf, f0 == 0, s);
}

extern "C" int gdl_cast_pad_bf16(const float* src0, int r0, const float* src1, int r1, int cols, int ld, int transpose,
                                 void* dst, int drows, int dcols, gdl_stream_t s) {
  GDL_REQUIRE(src0 && dst && r0 >= 0 && r1 >= 0 && (r1 == 0 || src1) && drows > 0 && dcols > 0,
              "gdl_cast_pad_bf16: bad arguments");
  const int64_t total = (int64_t)drows * dcols;
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  cast_pad_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(src0, r0, src1, r1, cols, ld, transpose, (bf16*)dst, drows,
                                                                dcols);
  GDL_CHECK_LAUNCH("cast_pad_kernel");
  return GDL_OK;
}

static int film_contract_impl(const void* G, int ldg, int c0, const float* x, const float* y, float* dx, float* dy, int B,
                              int D, int sum_mode, float* scratch, int i0, int ni, int acc, bool transpose,
                              gdl_stream_t s) {
  GDL_REQUIRE(G && x && y && dx && scratch && (sum_mode || dy) && B > 0 && D > 0 && i0 >= 0 && ni > 0 && i0 + ni <= D,
              "gdl_film_contract: bad arguments");
  if (transpose) {
    int rc = film_transpose(x, y, scratch, B, D, (cudaStream_t)s);
    if (rc != GDL_OK) return rc;
  }
  const int Bp = film_bp(B);
  const float* xt = scratch;
  const float* yt = scratch + (int64_t)D * Bp;
  if (B % 8 == 0 && c0 % 8 == 0 && ldg % 8 == 0 && B / 8 <= kContractThreads)
    film_contract_kernel<<<D, kContractThreads, 0, (cudaStream_t)s>>>((const bf16*)G, ldg, c0, xt, yt, dx, dy, B, D, Bp,
                                                                     sum_mode, i0, ni, acc);
  else
    film_contract_scalar_kernel<<<D, 256, 0, (cudaStream_t)s>>>((const bf16*)G, ldg, c0, xt, yt, dx, dy, B, D, Bp,
                                                                sum_mode, i0, ni, acc);
  GDL_CHECK_LAUNCH("film_contract_kernel");
  return GDL_OK;
}

extern "C" int gdl_film_contract(const void* G, int ldg, int c0, const float* x, const float* y, float* dx, float* dy,
                                 int B, int D, int sum_mode, float* scratch, gdl_stream_t s) {
  return film_contract_impl(G, ldg, c0, x, y, dx, dy, B, D, sum_mode, scratch, 0, D, 0, true, s);
}

extern "C" int gdl_film_contract_chunk(const void* G_chunk, int ldg, int c0, const float* x, const float* y, float* dx,
                                       float* dy, int B, int D, int sum_mode, float* scratch, int i0, int ni,
                                       int accumulate, int rebuild_scratch, gdl_stream_t s) {
  return film_contract_impl(G_chunk, ldg, c0, x, y, dx, dy, B, D, sum_mode, scratch, i0, ni, accumulate != 0,
                            rebuild_scratch != 0, s);
}

extern "C" int gdl_transpose_f32_to_bf16(const float* src, void* dst, int R, int64_t Cn, gdl_stream_t s) {
  GDL_REQUIRE(src && dst && R % 32 == 0 && Cn % 32 == 0, "gdl_transpose_f32_to_bf16: dims must be multiples of 32");
  dim3 grid((unsigned)(Cn / 32), R / 32), block(32, 8);
  transpose_f32_to_bf16_kernel<<<grid, block, 0, (cudaStream_t)s>>>(src, (bf16*)dst, R, Cn, Cn);
  GDL_CHECK_LAUNCH("transpose_f32_to_bf16_kernel");
  return GDL_OK;
}

extern "C" int gdl_transpose_bf16_to_f32(const void* src, float* dst, int R, int64_t Cn, gdl_stream_t s) {
  GDL_REQUIRE(src && dst && R % 32 == 0 && Cn % 32 == 0, "gdl_transpose_bf16_to_f32: dims must be multiples of 32");
  dim3 grid((unsigned)(Cn / 32), R / 32), block(32, 8);
  transpose_bf16_to_f32_kernel<<<grid, block, 0, (cudaStream_t)s>>>((const bf16*)src, dst, R, Cn, Cn);
  GDL_CHECK_LAUNCH("transpose_bf16_to_f32_kernel");
  return GDL_OK;
}

// Column window [c0, c0 + Cn) of the fp32 matrix dst [R][ldd] from the bf16 chunk src [Cn][R] (K-chunked FiLM weight gradient)
extern "C" int gdl_transpose_bf16_to_f32_window(const void* src, float* dst, int R, int64_t Cn, int64_t ldd, int64_t c0,
                                                gdl_stream_t s) {
  GDL_REQUIRE(src && dst && R % 32 == 0 && Cn % 32 == 0 && c0 >= 0 && c0 + Cn <= ldd,
              "gdl_transpose_bf16_to_f32_window: bad arguments");
  dim3 grid((unsigned)(Cn / 32), R / 32), block(32, 8);
  transpose_bf16_to_f32_kernel<<<grid, block, 0, (cudaStream_t)s>>>((const bf16*)src, dst + c0, R, Cn, ldd);
  GDL_CHECK_LAUNCH("transpose_bf16_to_f32_kernel(window)");
  return GDL_OK;
}
