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
                  const void* add_src, int add_mode, cudaStream_t s);
int64_t wgrad_flat_workspace_bytes(int N, int Ho, int Wo, int Ci, int Co, int R, int stride);
int try_wgrad_flat(int N, int Hi, int Wi, int Ho, int Wo, int Ci, int Co, int R, int stride, const void* x,
                   const void* dy, float* partial, int64_t workspace_bytes, cudaStream_t s);

constexpr int kGemmW = 128;  // the GEMM row index is folded into an (H, 128) "image"

// C[m][n] = sum over splits of partial[sp][m][n] (+ bias[n]); fixed order => deterministic
__global__ void gemm_tn_reduce_kernel(const float* __restrict__ partial, int splits, int64_t MN, int N,
                                      const float* __restrict__ bias, float* __restrict__ C) {
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= MN) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
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
// at / vt are the features transposed to [D][B] so that the batch index is contiguous.
__global__ void __launch_bounds__(256) film_outer_kernel(const float* __restrict__ a, const float* __restrict__ v,
                                                         bf16* __restrict__ Zt, int B, int D, int ZB, int variants) {
  const int groups = ZB / 8;
  const int64_t total = (int64_t)D * D * groups;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int g = int(idx % groups);
    const int64_t f = idx / groups;
    const int i = int(f / D), j = int(f - (int64_t)i * D);
    float o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const int col = g * 8 + c;
      const int var = col / B, b = col - var * B;
      float val = 0.f;
      if (var < variants) {
        const float ai = a[(int64_t)b * D + i], aj = a[(int64_t)b * D + j];
        const float vi = v[(int64_t)b * D + i], vj = v[(int64_t)b * D + j];
        val = var == 0 ? ai * vj : (var == 1 ? ai * aj : vi * vj);
      }
      o[c] = val;
    }
    *reinterpret_cast<uint4*>(Zt + f * ZB + g * 8) = pack8(o);
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
// One block per output index t (i for dx, j for dy), threads over the batch (coalesced rows of G).
__global__ void __launch_bounds__(256) film_contract_kernel(const bf16* __restrict__ G, int ldg, int c0,
                                                            const float* __restrict__ x, const float* __restrict__ y,
                                                            float* __restrict__ dx, float* __restrict__ dy, int B, int D,
                                                            int sum_mode) {
  const int t = blockIdx.x;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float sx = 0.f, sy = 0.f;
    const bf16* gcol = G + c0 + b;
    for (int k = 0; k < D; ++k) {
      sx = fmaf(__bfloat162float(gcol[((int64_t)t * D + k) * ldg]), y[(int64_t)b * D + k], sx);
      sy = fmaf(__bfloat162float(gcol[((int64_t)k * D + t) * ldg]), x[(int64_t)b * D + k], sy);
    }
    if (sum_mode) {
      dx[(int64_t)b * D + t] = sx + sy;
    } else {
      dx[(int64_t)b * D + t] = sx;
      dy[(int64_t)b * D + t] = sy;
    }
  }
}

// 32x32 tiled transposes between the fp32 parameter layout [R][Cn] and the bf16 feature-major shadow [Cn][R]
__global__ void transpose_f32_to_bf16_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int R, int64_t Cn) {
  __shared__ float tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32;
  const int r0 = blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += 8) tile[k][threadIdx.x] = src[(int64_t)(r0 + k) * Cn + c0 + threadIdx.x];
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += 8)
    dst[(c0 + k) * R + r0 + threadIdx.x] = __float2bfloat16_rn(tile[threadIdx.x][k]);
}
__global__ void transpose_bf16_to_f32_kernel(const bf16* __restrict__ src, float* __restrict__ dst, int R, int64_t Cn) {
  // src [Cn][R] -> dst [R][Cn]
  __shared__ float tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32;
  const int r0 = blockIdx.y * 32;
  for (int k = threadIdx.y; k < 32; k += 8)
    tile[k][threadIdx.x] = __bfloat162float(src[(c0 + k) * R + r0 + threadIdx.x]);
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += 8) dst[(int64_t)(r0 + k) * Cn + c0 + threadIdx.x] = tile[threadIdx.x][k];
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

extern "C" int gdl_gemm_tn_f32(const void* At, const void* Bt, const float* bias, float* C, int M, int N, int64_t K,
                               void* workspace, int64_t workspace_bytes, gdl_stream_t s) {
  GDL_REQUIRE(At && Bt && C && workspace, "gdl_gemm_tn_f32: null pointer");
  GDL_REQUIRE(K > 0 && K % kGemmW == 0 && N % 4 == 0, "gdl_gemm_tn_f32: need K % 128 == 0");
  const int H = int(K / kGemmW);
  int ns = try_wgrad_flat(1, H, kGemmW, H, kGemmW, M, N, 1, 1, At, Bt, (float*)workspace, workspace_bytes,
                          (cudaStream_t)s);
  if (ns < 0) return ns;
  if (ns == 0) {
    set_last_error("gdl_gemm_tn_f32: shape not supported (M, N multiples of 128 or M == 64) or workspace too small");
    return GDL_EINVAL;
  }
  const int64_t MN = (int64_t)M * N;
  gemm_tn_reduce_kernel<<<(unsigned)ceil_div64(MN / 4, 256), 256, 0, (cudaStream_t)s>>>((const float*)workspace, ns, MN,
                                                                                      N, bias, C);
  GDL_CHECK_LAUNCH("gemm_tn_reduce_kernel");
  return GDL_OK;
}

extern "C" int gdl_film_outer(const float* a, const float* v, void* Zt, int B, int D, int ZB, int variants,
                              gdl_stream_t s) {
  GDL_REQUIRE(a && v && Zt && B > 0 && D > 0 && ZB % 8 == 0 && variants >= 1 && variants <= 3 && variants * B <= ZB,
              "gdl_film_outer: bad arguments");
  const int64_t total = (int64_t)D * D * (ZB / 8);
  int64_t blocks = ceil_div64(total, 256);
  if (blocks > kNumSMs * 32) blocks = kNumSMs * 32;
  film_outer_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(a, v, (bf16*)Zt, B, D, ZB, variants);
  GDL_CHECK_LAUNCH("film_outer_kernel");
  return GDL_OK;
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

extern "C" int gdl_film_contract(const void* G, int ldg, int c0, const float* x, const float* y, float* dx, float* dy,
                                 int B, int D, int sum_mode, gdl_stream_t s) {
  GDL_REQUIRE(G && x && y && dx && (sum_mode || dy) && B > 0 && D > 0, "gdl_film_contract: bad arguments");
  film_contract_kernel<<<D, 256, 0, (cudaStream_t)s>>>((const bf16*)G, ldg, c0, x, y, dx, dy, B, D, sum_mode);
  GDL_CHECK_LAUNCH("film_contract_kernel");
  return GDL_OK;
}

extern "C" int gdl_transpose_f32_to_bf16(const float* src, void* dst, int R, int64_t Cn, gdl_stream_t s) {
  GDL_REQUIRE(src && dst && R % 32 == 0 && Cn % 32 == 0, "gdl_transpose_f32_to_bf16: dims must be multiples of 32");
  dim3 grid((unsigned)(Cn / 32), R / 32), block(32, 8);
  transpose_f32_to_bf16_kernel<<<grid, block, 0, (cudaStream_t)s>>>(src, (bf16*)dst, R, Cn);
  GDL_CHECK_LAUNCH("transpose_f32_to_bf16_kernel");
  return GDL_OK;
}

extern "C" int gdl_transpose_bf16_to_f32(const void* src, float* dst, int R, int64_t Cn, gdl_stream_t s) {
  GDL_REQUIRE(src && dst && R % 32 == 0 && Cn % 32 == 0, "gdl_transpose_bf16_to_f32: dims must be multiples of 32");
  dim3 grid((unsigned)(Cn / 32), R / 32), block(32, 8);
  transpose_bf16_to_f32_kernel<<<grid, block, 0, (cudaStream_t)s>>>((const bf16*)src, dst, R, Cn);
  GDL_CHECK_LAUNCH("transpose_bf16_to_f32_kernel");
  return GDL_OK;
}
