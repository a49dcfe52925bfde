// conv_halo.cu — TMA-fed implicit-GEMM for the 3x3 / stride-1 / pad-1 convolutions (forward,
// and data-gradient with flipped taps), which carry ~3/4 of the encoder FLOPs
// (reference models/backbone.py:44,47 conv3x3 inside BasicBlock).
//
// One CTA computes a 16(h) x 8(w) pixel tile x BN output channels.  Per 64-channel input slab
// ONE TMA box {64 ch, 16 w, 18 h} (the tile plus its halo, zero-filled outside the image =
// padding) lands in shared memory as 288 rows of 128 swizzled bytes.  The nine filter taps are
// nine *shifted views* of that same halo: for tap (r,s) the A-operand descriptor starts at row
// r*16+s and strides 16 rows (2048 B) between 8-pixel groups, so the im2col matrix is never
// materialised and every input byte is fetched from L2 once per slab instead of nine times.
// Weights stream through a second TMA ring ([BN rows][64 k] tiles of the packed matrix).
//   warp 5 lane 0 : TMA producer        warp 4 lane 0 : tcgen05.mma issuer (+ TMEM alloc)
//   warps 0-3     : epilogue (TMEM -> registers -> bf16 NHWC, optional residual-gradient add)
#include <mutex>
#include <unordered_map>
#include <string.h>
#include "common.cuh"
#include "tma.cuh"

namespace gdl {
using namespace tc05;

// ------------------------------------------------------------------------------------------
// tensor-map cache
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct TmapKey {
  uint64_t v[6];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) h = (h ^ x) * 1099511628211ull;
    return (size_t)h;
  }
};
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap*, TmapHash> g_tmaps;

static const CUtensorMap* cached_tmap(const TmapKey& key, cuuint32_t rank, void* ptr, const cuuint64_t* dims,
                                      const cuuint64_t* strides, const cuuint32_t* box,
                                      CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) return it->second;
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_last_error("cuTensorMapEncodeTiled entry point not available");
    return nullptr;
  }
  CUtensorMap* m = nullptr;
  if (posix_memalign(reinterpret_cast<void**>(&m), 64, sizeof(CUtensorMap)) != 0) return nullptr;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, ptr, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[128];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    set_last_error(buf);
    free(m);
    return nullptr;
  }
  g_tmaps.emplace(key, m);
  return m;
}

const CUtensorMap* tmap_nhwc(const void* ptr, int N, int H, int W, int C, int box_w, int box_h) {
  TmapKey key{{(uint64_t)ptr, ((uint64_t)N << 32) | (uint32_t)H, ((uint64_t)W << 32) | (uint32_t)C,
               ((uint64_t)box_w << 32) | (uint32_t)box_h, 4, 0}};
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  return cached_tmap(key, 4, const_cast<void*>(ptr), dims, strides, box);
}

const CUtensorMap* tmap_rows(const void* ptr, int64_t rows, int64_t K, int box_rows) {
  TmapKey key{{(uint64_t)ptr, (uint64_t)rows, (uint64_t)K, (uint64_t)box_rows, 2, 0}};
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  return cached_tmap(key, 2, const_cast<void*>(ptr), dims, strides, box);
}

// Strided 4-D view {C, W, H, N} of a bf16 tensor (element strides sW, sH, sN; channels contiguous),
// box {64, box_w, box_h, box_n}: whole padded pixel rows / row bands / images per TMA operation
// (conv_flat.cu, conv_wgrad_flat.cu).  box_w may exceed W: the surplus pixels are out-of-bounds zero fill.
const CUtensorMap* tmap_view4(const void* ptr, int C, int W, int H, int N, int64_t sW, int64_t sH, int64_t sN,
                              int box_w, int box_h, int box_n) {
  TmapKey key{{(uint64_t)ptr, ((uint64_t)N << 32) | (uint32_t)H, ((uint64_t)W << 32) | (uint32_t)C,
               ((uint64_t)box_w << 32) | ((uint64_t)box_h << 16) | (uint64_t)box_n | 0x80000000ull,
               (uint64_t)sW ^ ((uint64_t)sH << 24), (uint64_t)sN}};
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sN * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n};
  return cached_tmap(key, 4, const_cast<void*>(ptr), dims, strides, box);
}

// 16-channel (32-byte rows, 32-byte swizzle) variants used by the space-to-depth stems.
const CUtensorMap* tmap_nhwc16(const void* ptr, int N, int H, int W, int box_w, int box_h) {
  TmapKey key{{(uint64_t)ptr, ((uint64_t)N << 32) | (uint32_t)H, ((uint64_t)W << 32) | 16u,
               ((uint64_t)box_w << 32) | (uint32_t)box_h, 4, 32}};
  cuuint64_t dims[4] = {16, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {32, (cuuint64_t)W * 32, (cuuint64_t)H * W * 32};
  cuuint32_t box[4] = {16, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  return cached_tmap(key, 4, const_cast<void*>(ptr), dims, strides, box, CU_TENSOR_MAP_SWIZZLE_32B);
}
const CUtensorMap* tmap_rows16(const void* ptr, int64_t rows, int64_t K, int box_rows) {
  TmapKey key{{(uint64_t)ptr, (uint64_t)rows, (uint64_t)K, (uint64_t)box_rows, 2, 32}};
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {16, (cuuint32_t)box_rows};
  return cached_tmap(key, 2, const_cast<void*>(ptr), dims, strides, box, CU_TENSOR_MAP_SWIZZLE_32B);
}

// ------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------
constexpr int kTileH = 16, kTileW = 8;
// Halo box = tile + 1-pixel border, stored densely: pitch 10 rows per image row.  8-pixel row
// groups of a shifted view then start at arbitrary 128-byte rows and straddle 1024-byte swizzle
// atoms — legal because the tensor core (like TMA) swizzles on absolute smem address bits.
constexpr int kHaloW = 10, kHaloH = 18;
constexpr int kHaloBoxBytes = kHaloW * kHaloH * 128;                       // 23040 delivered by TMA
constexpr int kHaloBytes = (kHaloBoxBytes + 1023) / 1024 * 1024;           // 23552 per stage
constexpr int kHaloThreads = 192;
constexpr int kOutBlkBytes = 128 * 128;            // one 64-channel block of a 128-pixel tile

struct HaloParams {
  CUtensorMap tm_x;    // source activations {C, W, H, N}, box {64,16,18,1}
  CUtensorMap tm_w;    // packed weights {K, rows}, box {64, BN}
  CUtensorMap tm_out;  // destination {Cd, W, H, N}, box {64,8,16,1}
  CUtensorMap tm_res;  // residual gradient (same geometry as tm_out); valid when add_mode == 1
  int add_mode;
  int N, H, W, Cs, Cd;
  int tiles_h, tiles_w, tiles_total;  // tiles_total includes the N-tile dimension (slowest)
  int flip;                           // 1: use tap (2-r, 2-s) of the weights (data gradient)
};

template <int BN, int WST>
struct HaloSmem {
  static constexpr int HST = BN == 64 ? 4 : 3;  // halo ring depth (covers the TMA round trip)
  static constexpr int W_BYTES = BN * 128;
  static constexpr int OUT_BYTES = (BN / 64) * kOutBlkBytes;
  static constexpr int W_OFF = HST * kHaloBytes;
  static constexpr int OUT_OFF = W_OFF + WST * W_BYTES;
  static constexpr int RES_OFF = OUT_OFF + OUT_BYTES;
  static constexpr int BAR_OFF = RES_OFF + OUT_BYTES;
  static constexpr int TOTAL = BAR_OFF + 512 + 1024;
  static_assert(TOTAL <= 227 * 1024, "halo kernel shared memory budget");
};

// Persistent: one CTA per SM loops over output tiles.  Three pipelines run concurrently:
// TMA producer (halo + weight rings, optional residual tile) -> MMA issuer (two TMEM accumulator
// buffers) -> epilogue warps (TMEM -> registers -> swizzled smem -> TMA store).
template <int BN, int WST>
__global__ void __launch_bounds__(kHaloThreads, 1) conv3x3_halo_kernel(const __grid_constant__ HaloParams p) {
  using L = HaloSmem<BN, WST>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* halo_full = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  constexpr int HST = L::HST;
  uint64_t* halo_empty = halo_full + HST;
  uint64_t* w_full = halo_empty + HST;
  uint64_t* w_empty = w_full + WST;
  uint64_t* tmem_full = w_empty + WST;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_full = tmem_empty + 2;
  uint64_t* res_empty = res_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_empty + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int slabs = p.Cs >> 6;
  const int kbs = slabs * 9;                  // weight tiles per output tile
  const bool resident = kbs <= WST && p.Cd == BN;  // all weight tiles stay in smem for the CTA lifetime
  const int tiles_img = p.tiles_h * p.tiles_w;
  const int tiles_n = p.N * tiles_img;        // tiles per N-tile

  if (tid == 0) {
    for (int i = 0; i < HST; ++i) {
      mbar_init(&halo_full[i], 1);
      mbar_init(&halo_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    for (int i = 0; i < WST; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    mbar_init(res_full, 1);
    mbar_init(res_empty, 128);
    fence_mbar_init();
  }
  if (warp == 4) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  if (tid == 5 * 32) {
    tma_prefetch_desc(&p.tm_x);
    tma_prefetch_desc(&p.tm_w);
    tma_prefetch_desc(&p.tm_out);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t smem_base = smem_u32(smem);

  if (tid == 5 * 32) {
    // ------------------------------ TMA producer ------------------------------
    int hcount = 0, wcount = 0, it = 0;
    for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x, ++it) {
      const int nt = t / tiles_n;
      int rem = t - nt * tiles_n;
      const int n = rem / tiles_img;
      rem -= n * tiles_img;
      const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
      const int h0 = th * kTileH, w0 = tw * kTileW, n0 = nt * BN;
      for (int slab = 0; slab < slabs; ++slab, ++hcount) {
        const int hs = hcount % HST;
        if (hcount >= HST) mbar_wait(&halo_empty[hs], ((hcount / HST) - 1) & 1);
        mbar_arrive_expect_tx(&halo_full[hs], kHaloBoxBytes);
        tma_load_4d(smem_base + hs * kHaloBytes, &p.tm_x, &halo_full[hs], slab * 64, w0 - 1, h0 - 1, n);
        if (resident && it > 0) continue;
        for (int tap = 0; tap < 9; ++tap, ++wcount) {
          const int st = wcount % WST;
          if (wcount >= WST) mbar_wait(&w_empty[st], ((wcount / WST) - 1) & 1);
          const int wtap = p.flip ? 8 - tap : tap;
          mbar_arrive_expect_tx(&w_full[st], L::W_BYTES);
          tma_load_2d(smem_base + L::W_OFF + st * L::W_BYTES, &p.tm_w, &w_full[st], wtap * p.Cs + slab * 64, n0);
        }
      }
      if (p.add_mode == 1) {  // residual-gradient tile, consumed by the epilogue of this tile
        if (it >= 1) mbar_wait(res_empty, (it - 1) & 1);
        mbar_arrive_expect_tx(res_full, L::OUT_BYTES);
#pragma unroll
        for (int b = 0; b < BN / 64; ++b)
          tma_load_4d(smem_base + L::RES_OFF + b * kOutBlkBytes, &p.tm_res, res_full, n0 + b * 64, w0, h0, n);
      }
    }
  } else if (tid == 4 * 32) {
    // ------------------------------ MMA issuer ------------------------------
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
    const uint32_t a_hi = desc_hi_sw128(kHaloW * 128), b_hi = desc_hi_sw128(1024);
    const uint32_t a_lo0 = desc_lo_sw128(smem_base, 16), b_lo0 = desc_lo_sw128(smem_base + L::W_OFF, 16);
    int hcount = 0, wcount = 0, it = 0;
    for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x, ++it) {
      const int acc = it & 1;
      if (it >= 2) {
        mbar_wait(&tmem_empty[acc], ((it >> 1) - 1) & 1);
        tc_fence_after();
      }
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int slab = 0; slab < slabs; ++slab, ++hcount) {
        const int hs = hcount % HST;
        mbar_wait(&halo_full[hs], (hcount / HST) & 1);
        tc_fence_after();
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          int st;
          if (resident) {
            st = slab * 9 + tap;
            if (it == 0) {
              mbar_wait(&w_full[st], 0);
              tc_fence_after();
            }
          } else {
            st = wcount % WST;
            mbar_wait(&w_full[st], (wcount / WST) & 1);
            tc_fence_after();
          }
          // shifted view of the halo: the tensor core swizzles on absolute smem address bits,
          // so a start address offset by whole 128-byte rows needs no base-offset field
          const int r = tap / 3, s = tap - r * 3;
          const uint32_t a_lo = a_lo0 + hs * (kHaloBytes >> 4) + (((r * kHaloW + s) * 128) >> 4);
          const uint32_t b_lo = b_lo0 + st * (L::W_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_bf16_ss(d_tmem, desc_join(a_lo + 2 * k, a_hi), desc_join(b_lo + 2 * k, b_hi), idesc,
                        (slab | tap | k) != 0 ? 1u : 0u);
          if (!resident) {
            mma_commit(&w_empty[st]);
            ++wcount;
          }
        }
        mma_commit(&halo_empty[hs]);
      }
      mma_commit(&tmem_full[acc]);
    }
  } else if (warp < 4) {
    // ------------------------------ epilogue ------------------------------
    const int row = tid;  // tile pixel: h = row / 8, w = row % 8  (== TMA box order)
    const int sw = row & 7;
    const uint32_t out_row = smem_base + L::OUT_OFF + row * 128;
    const uint32_t res_row = smem_base + L::RES_OFF + row * 128;
    int it = 0;
    for (int t = blockIdx.x; t < p.tiles_total; t += gridDim.x, ++it) {
      const int nt = t / tiles_n;
      int rem = t - nt * tiles_n;
      const int n = rem / tiles_img;
      rem -= n * tiles_img;
      const int th = rem / p.tiles_w, tw = rem - th * p.tiles_w;
      const int acc = it & 1;
      mbar_wait(&tmem_full[acc], (it >> 1) & 1);
      tc_fence_after();
      if (p.add_mode == 1) mbar_wait(res_full, it & 1);
      if (tid == 0) tma_store_wait_read();  // previous tile's store has drained the staging buffer
      named_bar_sync(1, 128);
      const uint32_t trow = tmem_base + (uint32_t(warp * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(trow + c0, r);
        tmem_ld_wait();
        const int blk = c0 >> 6;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[q * 8 + i]);
          const int chunk = ((c0 & 63) >> 3) + q;
          const uint32_t off = blk * kOutBlkBytes + ((chunk ^ sw) << 4);
          if (p.add_mode == 1) {
            uint4 u;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(u.x), "=r"(u.y), "=r"(u.z), "=r"(u.w)
                         : "r"(res_row + off));
            float a[8];
            unpack8(u, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] += a[i];
          }
          uint4 o = pack8(f);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(out_row + off), "r"(o.x), "r"(o.y),
                       "r"(o.z), "r"(o.w)
                       : "memory");
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      if (p.add_mode == 1) mbar_arrive(res_empty);
      fence_proxy_async();
      named_bar_sync(1, 128);
      if (tid == 0) {
#pragma unroll
        for (int b = 0; b < BN / 64; ++b)
          tma_store_4d(&p.tm_out, smem_base + L::OUT_OFF + b * kOutBlkBytes, nt * BN + b * 64, tw * kTileW,
                       th * kTileH, n);
        tma_store_commit();
      }
    }
    if (tid == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 2 * BN);
}

template <int BN, int WST>
static int launch_halo(const HaloParams& p, cudaStream_t s) {
  using L = HaloSmem<BN, WST>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_kernel<BN, WST>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(conv3x3_halo)");
    attr_set = true;
  }
  int grid = p.tiles_total < kNumSMs ? p.tiles_total : kNumSMs;
  conv3x3_halo_kernel<BN, WST><<<grid, kHaloThreads, L::TOTAL, s>>>(p);
  GDL_CHECK_LAUNCH("conv3x3_halo_kernel");
  return GDL_OK;
}

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// Returns 1 when the convolution was handled here, 0 when it is not eligible (caller falls
// through to the generic gather kernel), < 0 on error.
int try_conv3x3_halo(int N, int H, int W, int Cs, int Cd, const void* src, const void* wt, int64_t wt_rows,
                     int64_t wt_k, void* dst, const void* add_src, int add_mode, int flip, cudaStream_t s) {
  static const int impl = env_int("GDL_CONV_IMPL", 1);  // 0 = generic gather kernel only
  static const int min_util = env_int("GDL_HALO_MIN_UTIL_PCT", 60);
  if (!impl) return 0;
  if (Cs % 64 != 0 || Cd % 64 != 0 || add_mode == 2) return 0;
  const int tiles_h = (H + kTileH - 1) / kTileH, tiles_w = (W + kTileW - 1) / kTileW;
  const int util = 100 * H * W / (tiles_h * kTileH * tiles_w * kTileW);
  if (util < min_util) return 0;
  const int BN = (Cd % 128 == 0) ? 128 : 64;
  const CUtensorMap* tx = tmap_nhwc(src, N, H, W, Cs, kHaloW, kHaloH);
  const CUtensorMap* tw = tmap_rows(wt, wt_rows, wt_k, BN);
  const CUtensorMap* to = tmap_nhwc(dst, N, H, W, Cd, kTileW, kTileH);
  const CUtensorMap* tr = add_mode == 1 ? tmap_nhwc(add_src, N, H, W, Cd, kTileW, kTileH) : to;
  if (!tx || !tw || !to || !tr) return GDL_ECUDA;
  HaloParams p;
  p.tm_x = *tx;
  p.tm_w = *tw;
  p.tm_out = *to;
  p.tm_res = *tr;
  p.add_mode = add_mode;
  p.N = N; p.H = H; p.W = W; p.Cs = Cs; p.Cd = Cd;
  p.tiles_h = tiles_h; p.tiles_w = tiles_w;
  p.tiles_total = (Cd / BN) * N * tiles_h * tiles_w;
  p.flip = flip;
  int rc = (BN == 128) ? launch_halo<128, 5>(p, s) : launch_halo<64, 9>(p, s);
  return rc == GDL_OK ? 1 : rc;
}

}  // namespace gdl
