// conv_wgrad_flat.cu — weight gradient on the flat pixel grid of conv_flat.cu (reference
// models/backbone.py:44,47,142-145 conv3x3 / conv1x1 backward, reached from main_dgl.py:110):
//
//   dW[co][tap][ci] = sum over flat pixels q   dY[q][co] * X_plane(tap)[q + shift(tap)][ci]
//
// q enumerates the OUTPUT grid with one zero pad column per row and one zero pad row per image
// (q = n*IS + h*P + w, P = Wo+1, IS = (Ho+1)*P), so a filter tap is a constant row shift and the
// reduction runs over consecutive q regardless of the map size: 7x7 and 9x6 maps fill 77 % of every
// 128-pixel step where 16x8 pixel tiles fill 38-42 %.  Pad pixels contribute nothing because both
// windows are written by TMA with out-of-bounds zero fill.  Stride-2 convolutions read X through its
// four parity planes (strided tensor-map views on the output grid): tap (r,s) lives in plane
// ((r+1)&1, (s+1)&1) at shift floor((r-1)/2)*P + floor((s-1)/2).
//   Both operands are MN-major views of the row windows (rows = pixels = the reduction dimension);
//   D[(tap,ci) 128 rows][BN co] accumulates in TMEM over the CTA's pixel range (split-K), fp32
//   partials are written once per CTA and reduced in a fixed order (deterministic).
//   MODE 0 (Ci == 64): an M tile is TWO taps x 64 ci (second block = same window, LBO rows further).
//   MODE 1 (Ci >= 128): an M tile is one tap x 128 ci (two slab windows, LBO = slab stride).
#include <string.h>
#include "common.cuh"
#include "tma.cuh"

namespace gdl {
using namespace tc05;

constexpr int kWfThreads = 192;
constexpr int kWfTM = 128;  // pixels per pipeline stage
constexpr int kWfMaxUnits = 5;

struct WfUnit {
  int shift0;  // row shift of block 0
  int lbo;     // MODE 0: byte distance to the second tap's rows; MODE 1: unused
  int tap0, tap1;  // linear tap indices (r*S+s) of the two 64-row blocks; tap1 < 0: unused block
};
struct WfJobType {
  int plane, nunits, smin, smax;
  WfUnit u[kWfMaxUnits];
};
struct WfParams {
  CUtensorMap tm_x[4];
  CUtensorMap tm_dy;
  float* partial;  // [splits][Kp][Co]
  int N, Hs, Ws, Ci, Co, P, IS;
  int ntypes;
  WfJobType jt[4];
  int ci_tiles, co_tiles;
  int tiles_total, tiles_per_split;
  int xwin_bytes, dywin_bytes, stage_bytes, stages;
  int Kp;
};

__device__ __forceinline__ int wf_floor_div(int a, int b) {
  int q = a / b;
  return (a - q * b < 0) ? q - 1 : q;
}

template <int BN, int MODE>
__global__ void __launch_bounds__(kWfThreads, 1) conv_wgrad_flat_kernel(const __grid_constant__ WfParams p) {
  constexpr int NS = MODE == 0 ? 1 : 2;
  constexpr int NB = BN / 64;
  constexpr int kMaxSt = 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* empty = full + kMaxSt;
  uint64_t* tmem_full = empty + kMaxSt;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int ST = p.stages;
  const int t0 = blockIdx.x * p.tiles_per_split;
  const int t1 = min(t0 + p.tiles_per_split, p.tiles_total);
  int y = blockIdx.y;
  const int type = y % p.ntypes;
  y /= p.ntypes;
  const int co0 = (y % p.co_tiles) * BN;
  const int ci0 = (y / p.co_tiles) * (NS * 64);
  const WfJobType& jt = p.jt[type];

  if (tid == 0) {
    for (int i = 0; i < kMaxSt; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (tid == 5 * 32) {
    tma_prefetch_desc(&p.tm_x[jt.plane]);
    tma_prefetch_desc(&p.tm_dy);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);
  const int dy_off = NS * p.xwin_bytes;

  if (tid == 5 * 32) {
    // ------------------------------ TMA producer ------------------------------
    const int rows_img = p.Hs + 1;
    const uint32_t row_bytes = (uint32_t)p.P * 128u;
    const CUtensorMap* tmx = &p.tm_x[jt.plane];
    int it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const int q0 = t * kWfTM;
      const int st = it % ST;
      if (it >= ST) mbar_wait(&empty[st], ((it / ST) - 1) & 1);
      const int xa = wf_floor_div(q0 + jt.smin, p.P), xb = wf_floor_div(q0 + kWfTM + jt.smax - 1, p.P);
      const int da = q0 / p.P, db = (q0 + kWfTM - 1) / p.P;
      const int nx = xb - xa + 1, nd = db - da + 1;
      mbar_arrive_expect_tx(&full[st], (uint32_t)(NS * nx + NB * nd) * row_bytes);
      const uint32_t sbase = smem_base + st * p.stage_bytes;
      for (int i = 0; i < nx; ++i) {
        const int rho = xa + i;
        const int n = wf_floor_div(rho, rows_img);
        const int h = rho - n * rows_img;
#pragma unroll
        for (int sl = 0; sl < NS; ++sl)
          tma_load_4d(sbase + sl * p.xwin_bytes + i * row_bytes, tmx, &full[st], ci0 + sl * 64, 0, h, n);
      }
      for (int i = 0; i < nd; ++i) {
        const int rho = da + i;
        const int n = rho / rows_img;
        const int h = rho - n * rows_img;
#pragma unroll
        for (int b = 0; b < NB; ++b)
          tma_load_4d(sbase + dy_off + b * p.dywin_bytes + i * row_bytes, &p.tm_dy, &full[st], co0 + b * 64, 0, h, n);
      }
    }
  } else if (tid == 4 * 32) {
    // ------------------------------ MMA issuer ------------------------------
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
    const uint32_t hi = desc_hi_sw128(1024);  // 8-pixel K groups are dense rows in both windows
    const int nunits = jt.nunits;
    int it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const int q0 = t * kWfTM;
      const int st = it % ST;
      const int xa = wf_floor_div(q0 + jt.smin, p.P);
      const int ox = q0 - xa * p.P;            // stage row of pixel q0 in the X window
      const int od = q0 - (q0 / p.P) * p.P;    // stage row of pixel q0 in the dY window
      mbar_wait(&full[st], (it / ST) & 1);
      tc_fence_after();
      const uint32_t sbase = smem_base + st * p.stage_bytes;
      const uint32_t b_lo0 = desc_lo_sw128(sbase + dy_off + od * 128, p.dywin_bytes);
#pragma unroll 1
      for (int j = 0; j < kWfTM / 16; ++j) {
        for (int u = 0; u < nunits; ++u) {
          const uint32_t a_lo = desc_lo_sw128(sbase + (ox + jt.u[u].shift0 + 16 * j) * 128,
                                              MODE == 0 ? jt.u[u].lbo : p.xwin_bytes);
          mma_bf16_ss(tmem_base + u * BN, desc_join(a_lo, hi), desc_join(b_lo0 + j * (2048 >> 4), hi), idesc,
                      (it | j) != 0 ? 1u : 0u);
        }
      }
      mma_commit(&empty[st]);
    }
    mma_commit(tmem_full);
  } else if (warp < 4) {
    // ------------------------------ epilogue: fp32 partials ------------------------------
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int row = tid;
    const uint32_t trow = tmem_base + (uint32_t(warp * 32) << 16);
#pragma unroll 1
    for (int u = 0; u < jt.nunits; ++u) {
      int tap, ci;
      if (MODE == 0) {
        tap = row < 64 ? jt.u[u].tap0 : jt.u[u].tap1;
        ci = row & 63;
      } else {
        tap = jt.u[u].tap0;
        ci = ci0 + row;
      }
      const bool valid = tap >= 0 && t1 > t0;
      float* out = p.partial + ((size_t)blockIdx.x * p.Kp + (size_t)(tap < 0 ? 0 : tap) * p.Ci + ci) * p.Co + co0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(trow + u * BN + c0, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<uint4*>(out + c0 + g * 4) = make_uint4(r[g * 4], r[g * 4 + 1], r[g * 4 + 2], r[g * 4 + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct WfPlan {
  int mode, BN, ntypes, ci_tiles, co_tiles, gy, splits, tiles_total, tiles_per_split;
  int xwin_bytes, dywin_bytes, stage_bytes, stages, P, IS, taps;
  WfJobType jt[4];
};

static void add_unit(WfJobType& t, int shift0, int lbo, int tap0, int tap1) {
  WfUnit& u = t.u[t.nunits++];
  u.shift0 = shift0;
  u.lbo = lbo;
  u.tap0 = tap0;
  u.tap1 = tap1;
}

// R in {1,3}, stride in {1,2}; (Ho, Wo) is the output grid the flat index runs over.
static bool plan_wgrad_flat(int N, int Ho, int Wo, int Ci, int Co, int R, int stride, WfPlan& w) {
  if (Ci % 64 != 0 || Co % 64 != 0) return false;
  if (!(R == 3 || R == 1) || !(stride == 1 || stride == 2)) return false;
  memset(&w, 0, sizeof(w));
  const int P = Wo + 1;
  if (P > 256) return false;
  w.P = P;
  w.IS = (Ho + 1) * P;
  const int64_t Q = (int64_t)N * w.IS;
  if (Q + 4 * P + 1024 >= ((int64_t)1 << 31)) return false;
  w.taps = R * R;
  if (Ci == 64) {
    w.mode = 0;
    w.BN = Co == 64 ? 64 : 128;
    w.ci_tiles = 1;
  } else if (Ci % 128 == 0 && Co % 128 == 0) {
    w.mode = 1;
    w.BN = 128;
    w.ci_tiles = Ci / 128;
  } else {
    return false;
  }
  w.co_tiles = Co / w.BN;
  // --- job types: which taps share one X window ---
  struct T { int tap, plane, shift; };
  T taps[9];
  int nt = 0;
  for (int r = 0; r < R; ++r)
    for (int s = 0; s < R; ++s) {
      T t;
      t.tap = r * R + s;
      if (R == 1) {
        t.plane = 0;
        t.shift = 0;
      } else if (stride == 1) {
        t.plane = 0;
        t.shift = (r - 1) * P + (s - 1);
      } else {
        t.plane = ((r + 1) & 1) * 2 + ((s + 1) & 1);
        t.shift = (r == 0 ? -1 : 0) * P + (s == 0 ? -1 : 0);
      }
      taps[nt++] = t;
    }
  if (stride == 1 && w.mode == 1 && R == 3) {
    w.ntypes = 3;  // one filter row per CTA: 3 accumulators, narrow window
    for (int r = 0; r < 3; ++r) {
      w.jt[r].plane = 0;
      for (int s = 0; s < 3; ++s) add_unit(w.jt[r], taps[r * 3 + s].shift, 0, r * 3 + s, -1);
    }
  } else {
    // group by plane (stride 1: a single plane); MODE 0 pairs consecutive taps of the group
    const int nplanes = (stride == 2 && R == 3) ? 4 : 1;
    w.ntypes = 0;
    for (int pl = 0; pl < nplanes; ++pl) {
      T g[9];
      int ng = 0;
      for (int i = 0; i < nt; ++i)
        if (taps[i].plane == pl) g[ng++] = taps[i];
      if (ng == 0) continue;
      WfJobType& jt = w.jt[w.ntypes++];
      jt.plane = pl;
      if (w.mode == 0) {
        for (int i = 0; i < ng; i += 2) {
          if (i + 1 < ng)
            add_unit(jt, g[i].shift, (g[i + 1].shift - g[i].shift) * 128, g[i].tap, g[i + 1].tap);
          else
            add_unit(jt, g[i].shift, 128, g[i].tap, -1);
        }
      } else {
        for (int i = 0; i < ng; ++i) add_unit(jt, g[i].shift, 0, g[i].tap, -1);
      }
    }
  }
  int span = 0;
  for (int k = 0; k < w.ntypes; ++k) {
    WfJobType& jt = w.jt[k];
    if (jt.nunits * w.BN > 512) return false;
    jt.smin = 1 << 30;
    jt.smax = -(1 << 30);
    for (int i = 0; i < jt.nunits; ++i) {
      int lo = jt.u[i].shift0, hi = jt.u[i].shift0;
      if (w.mode == 0) hi += jt.u[i].lbo / 128;  // second block (also covers the ignored block of a single)
      if (lo < jt.smin) jt.smin = lo;
      if (hi > jt.smax) jt.smax = hi;
      if (w.mode == 0 && jt.u[i].lbo <= 0) return false;
    }
    if (jt.smax - jt.smin > span) span = jt.smax - jt.smin;
  }
  const int nx_max = (kWfTM + span - 1 + P - 1) / P + 1, nd_max = (kWfTM - 1 + P - 1) / P + 1;
  w.xwin_bytes = (nx_max * P * 128 + 1023) / 1024 * 1024;
  w.dywin_bytes = (nd_max * P * 128 + 1023) / 1024 * 1024;
  const int NS = w.mode == 0 ? 1 : 2;
  w.stage_bytes = NS * w.xwin_bytes + (w.BN / 64) * w.dywin_bytes;
  if (w.xwin_bytes >= (1 << 18) || w.dywin_bytes >= (1 << 18)) return false;  // LBO field
  w.stages = (227 * 1024 - 2048) / w.stage_bytes;
  if (w.stages > 4) w.stages = 4;
  if (w.stages < 2) return false;
  w.tiles_total = int((Q + kWfTM - 1) / kWfTM);
  w.gy = w.ci_tiles * w.co_tiles * w.ntypes;
  int splits = (2 * kNumSMs) / w.gy;
  if (splits < 1) splits = 1;
  if (splits > w.tiles_total) splits = w.tiles_total;
  w.tiles_per_split = (w.tiles_total + splits - 1) / splits;
  w.splits = (w.tiles_total + w.tiles_per_split - 1) / w.tiles_per_split;
  return true;
}

template <int BN, int MODE>
static int launch_wf(const WfParams& p, const WfPlan& w, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_flat_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv_wgrad_flat)");
    attr_set = true;
  }
  dim3 grid(w.splits, w.gy);
  conv_wgrad_flat_kernel<BN, MODE><<<grid, kWfThreads, w.stages * w.stage_bytes + 512 + 1024, s>>>(p);
  GDL_CHECK_LAUNCH("conv_wgrad_flat_kernel");
  return GDL_OK;
}

// Workspace (bytes) the flat path needs, or 0 when the shape is not eligible.
int64_t wgrad_flat_workspace_bytes(int N, int Ho, int Wo, int Ci, int Co, int R, int stride) {
  WfPlan w;
  if (!plan_wgrad_flat(N, Ho, Wo, Ci, Co, R, stride, w)) return 0;
  return (int64_t)w.splits * w.taps * Ci * Co * (int64_t)sizeof(float);
}

// x is the conv INPUT [N,Hi,Wi,Ci]; dy the output gradient [N,Ho,Wo,Co].  Returns the number of splits
// written (>0) when handled, 0 when not eligible, <0 on error.
int try_wgrad_flat(int N, int Hi, int Wi, int Ho, int Wo, int Ci, int Co, int R, int stride, const void* x,
                   const void* dy, float* partial, int64_t workspace_bytes, cudaStream_t s) {
  WfPlan w;
  if (!plan_wgrad_flat(N, Ho, Wo, Ci, Co, R, stride, w)) return 0;
  if (workspace_bytes < (int64_t)w.splits * w.taps * Ci * Co * (int64_t)sizeof(float)) return 0;
  WfParams p;
  memset(&p, 0, sizeof(p));
  const bf16* xb = (const bf16*)x;
  const int64_t sW = (int64_t)stride * Ci, sH = (int64_t)stride * Wi * Ci, sN = (int64_t)Hi * Wi * Ci;
  if (stride == 2 && R == 3) {
    // plane (a,b)[i,j] = x[2i+a, 2j+b]; its extent on the output grid is what exists of the input
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        const int ph = (Hi - a + 1) / 2, pw = (Wi - b + 1) / 2;
        if (ph <= 0 || pw <= 0) return 0;
        const CUtensorMap* t = tmap_view4(xb + ((int64_t)a * Wi + b) * Ci, Ci, pw, ph, N, sW, sH, sN, w.P);
        if (!t) return GDL_ECUDA;
        p.tm_x[a * 2 + b] = *t;
      }
  } else {
    const CUtensorMap* t = tmap_view4(xb, Ci, Wo, Ho, N, sW, sH, sN, w.P);
    if (!t) return GDL_ECUDA;
    p.tm_x[0] = *t;
  }
  const CUtensorMap* td = tmap_view4(dy, Co, Wo, Ho, N, Co, (int64_t)Wo * Co, (int64_t)Ho * Wo * Co, w.P);
  if (!td) return GDL_ECUDA;
  p.tm_dy = *td;
  p.partial = partial;
  p.N = N; p.Hs = Ho; p.Ws = Wo; p.Ci = Ci; p.Co = Co; p.P = w.P; p.IS = w.IS;
  p.ntypes = w.ntypes;
  for (int i = 0; i < 4; ++i) p.jt[i] = w.jt[i];
  p.ci_tiles = w.ci_tiles; p.co_tiles = w.co_tiles;
  p.tiles_total = w.tiles_total; p.tiles_per_split = w.tiles_per_split;
  p.xwin_bytes = w.xwin_bytes; p.dywin_bytes = w.dywin_bytes; p.stage_bytes = w.stage_bytes; p.stages = w.stages;
  p.Kp = w.taps * Ci;
  int rc;
  if (w.mode == 0)
    rc = w.BN == 64 ? launch_wf<64, 0>(p, w, s) : launch_wf<128, 0>(p, w, s);
  else
    rc = launch_wf<128, 1>(p, w, s);
  return rc == GDL_OK ? w.splits : rc;
}

}  // namespace gdl
