// conv_wgrad_halo.cu — TMA-fed weight gradient of the 3x3 / stride-1 / pad-1 convolutions
// (the backward of reference models/backbone.py:44,47 conv3x3, reached from main_dgl.py:110).
//
//   dW[co][tap][ci] = sum over pixels  dY[pix][co] * X[pix + tap][ci]
//
// Per 16x8 pixel tile ONE TMA halo box of X and ONE box of dY land in shared memory; both are
// MN-major operands (rows = pixels = the reduction dimension).  The filter taps are shifted
// views of the halo (start address + whole 128-byte rows), so X is read once per tile instead
// of nine times.  D[(tap,ci) 128 rows][co BN cols] accumulates in TMEM over the CTA's pixel
// range (split-K); fp32 partials are written once per CTA and reduced in a fixed order.
//   MODE 0 (Ci == 64): an M tile is TWO taps x 64 ci — the second 64-row block of the A
//       descriptor is the same halo shifted by one pixel (LBO = 128 B) or, across a filter row,
//       by pitch-2 pixels (LBO = 1024 B).  9 taps = 4 pairs + 1 single: 5 MMAs per 16 pixels.
//   MODE 1 (Ci >= 128): an M tile is one tap x 128 ci (two 64-channel slabs, LBO = slab stride);
//       one CTA owns one filter row (3 taps), grid.y enumerates (ci tile, co tile, filter row).
#include "common.cuh"
#include "tma.cuh"

namespace gdl {
using namespace tc05;

constexpr int kWgThreads = 192;
constexpr int kWgPitch = 10;  // halo pitch (tile width 8 + 2)

struct WgHaloParams {
  CUtensorMap tm_x;   // X {Ci, W, H, N}, box {64, 10, HR, 1}
  CUtensorMap tm_dy;  // dY {Co, W, H, N}, box {64, 8, 16, 1}
  float* partial;     // [splits][Kp = 9*Ci][Co]
  int N, H, W, Ci, Co;
  int tiles_h, tiles_w, tiles_total, tiles_per_split;
  int co_tiles;
};

template <int BN, int MODE>
struct WgHaloCfg {
  static constexpr int NS = MODE == 0 ? 1 : 2;             // 64-channel slabs per stage
  static constexpr int HR = MODE == 0 ? 18 : 16;           // halo rows loaded
  static constexpr int UNITS = MODE == 0 ? 5 : 3;          // MMAs per 16-pixel step
  static constexpr int SLAB_BOX = kWgPitch * HR * 128;
  static constexpr int SLAB_BYTES = (SLAB_BOX + 1023) / 1024 * 1024;
  static constexpr int DY_BYTES = (BN / 64) * 128 * 128;
  static constexpr int STAGE_BYTES = NS * SLAB_BYTES + DY_BYTES;
  static constexpr int ST = MODE == 0 ? 4 : 3;
  static constexpr int BAR_OFF = ST * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 512 + 1024;
  static constexpr int TMEM_COLS = 512;
  static_assert(UNITS * BN <= 512, "accumulators must fit TMEM");
  static_assert(TOTAL <= 227 * 1024, "smem budget");
};

template <int BN, int MODE>
__global__ void __launch_bounds__(kWgThreads, 1) conv3x3_wgrad_halo_kernel(const __grid_constant__ WgHaloParams p) {
  using L = WgHaloCfg<BN, MODE>;
  constexpr int ST = L::ST;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty = full + ST;
  uint64_t* tmem_full = empty + ST;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tiles_img = p.tiles_h * p.tiles_w;
  const int t0 = blockIdx.x * p.tiles_per_split;
  const int t1 = min(t0 + p.tiles_per_split, p.tiles_total);
  // work item: MODE 0: y = co tile.  MODE 1: y = ((ci_tile * co_tiles) + co_tile) * 3 + filter row
  int y = blockIdx.y;
  int frow = 0, ci0 = 0;
  if (MODE == 1) {
    frow = y % 3;
    y /= 3;
  }
  const int co0 = (y % p.co_tiles) * BN;
  if (MODE == 1) ci0 = (y / p.co_tiles) * 128;

  if (tid == 0) {
    for (int i = 0; i < ST; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, L::TMEM_COLS);
    tmem_relinquish();
  }
  if (tid == 5 * 32) {
    tma_prefetch_desc(&p.tm_x);
    tma_prefetch_desc(&p.tm_dy);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);

  if (tid == 5 * 32) {
    // ------------------------------ TMA producer ------------------------------
    int it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const int n = t / tiles_img;
      const int rem = t - n * tiles_img;
      const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
      const int h0 = th * 16, w0 = tw * 8;
      const int st = it % ST;
      if (it >= ST) mbar_wait(&empty[st], ((it / ST) - 1) & 1);
      const uint32_t sbase = smem_base + st * L::STAGE_BYTES;
      mbar_arrive_expect_tx(&full[st], L::NS * L::SLAB_BOX + L::DY_BYTES);
#pragma unroll
      for (int sl = 0; sl < L::NS; ++sl)
        tma_load_4d(sbase + sl * L::SLAB_BYTES, &p.tm_x, &full[st], ci0 + sl * 64, w0 - 1, h0 - 1 + frow, n);
#pragma unroll
      for (int b = 0; b < BN / 64; ++b)
        tma_load_4d(sbase + L::NS * L::SLAB_BYTES + b * 16384, &p.tm_dy, &full[st], co0 + b * 64, w0, h0, n);
    }
  } else if (tid == 4 * 32) {
    // ------------------------------ MMA issuer ------------------------------
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
    const uint32_t a_hi = desc_hi_sw128(kWgPitch * 128);  // next 8-pixel group = next image row of the halo
    const uint32_t b_hi = desc_hi_sw128(1024);            // dY tile rows are dense
    int it = 0;
    for (int t = t0; t < t1; ++t, ++it) {
      const int st = it % ST;
      mbar_wait(&full[st], (it / ST) & 1);
      tc_fence_after();
      const uint32_t sbase = smem_base + st * L::STAGE_BYTES;
      const uint32_t b_lo0 = desc_lo_sw128(sbase + L::NS * L::SLAB_BYTES, 16384);
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // 16 pixels = tile rows 2j, 2j+1
        const uint32_t row_off = 2 * j * kWgPitch * 128;
#pragma unroll
        for (int u = 0; u < L::UNITS; ++u) {
          uint32_t a_off, a_lbo;
          if (MODE == 0) {
            // taps 2u and 2u+1 (linear tap index r*3+s); the second block of the single tap 8 is ignored
            const int tap = 2 * u;
            const int r = tap / 3, s = tap - r * 3;
            a_off = (r * kWgPitch + s) * 128;
            a_lbo = (s == 2) ? (kWgPitch - 2) * 128 : 128;
          } else {
            a_off = u * 128;  // taps (frow, u): the halo was loaded starting at image row h0-1+frow
            a_lbo = L::SLAB_BYTES;
          }
          const uint32_t a_lo = desc_lo_sw128(sbase + a_off + row_off, a_lbo);
          mma_bf16_ss(tmem_base + u * BN, desc_join(a_lo, a_hi), desc_join(b_lo0 + j * (2048 >> 4), b_hi), idesc,
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
    const int Kp = 9 * p.Ci;
#pragma unroll 1
    for (int u = 0; u < L::UNITS; ++u) {
      int tap, ci;
      if (MODE == 0) {
        tap = 2 * u + (row >> 6);
        ci = row & 63;
      } else {
        tap = frow * 3 + u;
        ci = ci0 + row;
      }
      const bool valid = tap < 9 && t1 > t0;
      float* out = p.partial + ((size_t)blockIdx.x * Kp + (size_t)tap * p.Ci + ci) * p.Co + co0;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(trow + u * BN + c0, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(out + c0 + q * 4) = make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, L::TMEM_COLS);
}

struct WgHaloPlan {
  int mode, BN, tiles_h, tiles_w, tiles_total, gy, splits, tiles_per_split, co_tiles;
};

static bool plan_wgrad_halo(int N, int H, int W, int Ci, int Co, WgHaloPlan& w) {
  if (Ci % 64 != 0 || Co % 64 != 0) return false;
  w.tiles_h = (H + 15) / 16;
  w.tiles_w = (W + 7) / 8;
  w.tiles_total = N * w.tiles_h * w.tiles_w;
  if (Ci == 64) {
    w.mode = 0;
    w.BN = 64;
    w.co_tiles = Co / 64;
    w.gy = w.co_tiles;
  } else if (Ci % 128 == 0 && Co % 128 == 0) {
    w.mode = 1;
    w.BN = 128;
    w.co_tiles = Co / 128;
    w.gy = (Ci / 128) * w.co_tiles * 3;
  } else {
    return false;
  }
  int target = 2 * kNumSMs;
  int splits = target / w.gy;
  if (splits < 1) splits = 1;
  if (splits > w.tiles_total) splits = w.tiles_total;
  w.tiles_per_split = (w.tiles_total + splits - 1) / splits;
  w.splits = (w.tiles_total + w.tiles_per_split - 1) / w.tiles_per_split;
  return true;
}

template <int BN, int MODE>
static int launch_wg_halo(const WgHaloParams& p, const WgHaloPlan& w, cudaStream_t s) {
  using L = WgHaloCfg<BN, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_halo_kernel<BN, MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv3x3_wgrad_halo)");
    attr_set = true;
  }
  dim3 grid(w.splits, w.gy);
  conv3x3_wgrad_halo_kernel<BN, MODE><<<grid, kWgThreads, L::TOTAL, s>>>(p);
  GDL_CHECK_LAUNCH("conv3x3_wgrad_halo_kernel");
  return GDL_OK;
}

static int env_int2(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// Workspace (bytes) the halo path needs, or 0 when the shape is not eligible.
int64_t wgrad_halo_workspace_bytes(int N, int H, int W, int Ci, int Co) {
  WgHaloPlan w;
  if (!plan_wgrad_halo(N, H, W, Ci, Co, w)) return 0;
  return (int64_t)w.splits * 9 * Ci * Co * (int64_t)sizeof(float);
}

// Returns the number of splits written (>0) when handled, 0 when not eligible, <0 on error.
int try_wgrad3x3_halo(int N, int H, int W, int Ci, int Co, const void* x, const void* dy, float* partial,
                      int64_t workspace_bytes, cudaStream_t s) {
  static const int impl = env_int2("GDL_WGRAD_IMPL", 1);
  static const int min_util = env_int2("GDL_HALO_MIN_UTIL_PCT", 60);
  if (!impl) return 0;
  WgHaloPlan w;
  if (!plan_wgrad_halo(N, H, W, Ci, Co, w)) return 0;
  const int util = 100 * H * W / (w.tiles_h * 16 * w.tiles_w * 8);
  if (util < min_util) return 0;
  if (workspace_bytes < (int64_t)w.splits * 9 * Ci * Co * (int64_t)sizeof(float)) return 0;
  const CUtensorMap* tx = tmap_nhwc(x, N, H, W, Ci, kWgPitch, w.mode == 0 ? 18 : 16);
  const CUtensorMap* td = tmap_nhwc(dy, N, H, W, Co, 8, 16);
  if (!tx || !td) return GDL_ECUDA;
  WgHaloParams p;
  p.tm_x = *tx;
  p.tm_dy = *td;
  p.partial = partial;
  p.N = N; p.H = H; p.W = W; p.Ci = Ci; p.Co = Co;
  p.tiles_h = w.tiles_h; p.tiles_w = w.tiles_w; p.tiles_total = w.tiles_total;
  p.tiles_per_split = w.tiles_per_split;
  p.co_tiles = w.co_tiles;
  int rc = w.mode == 0 ? launch_wg_halo<64, 0>(p, w, s) : launch_wg_halo<128, 1>(p, w, s);
  return rc == GDL_OK ? w.splits : rc;
}

}  // namespace gdl
