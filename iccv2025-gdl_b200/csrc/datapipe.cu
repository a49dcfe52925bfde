// datapipe.cu — visual half of the reference's per-item transform on the GPU, driven by the crop boxes and
// flip decisions that torchvision drew on the host (reference dataset/CramedDataset.py:76-89,96-101 and
// dataset/KSDataset.py:160-173,183-190):
//
//   RandomResizedCrop(224) | Resize((224,224))  ->  RandomHorizontalFlip  ->  ToTensor  ->  Normalize
//
// The resize is Pillow's 8-bit two-pass bilinear resample (torchvision's PIL backend: img.crop + img.resize,
// Pillow src/libImaging/Resample.c), restated bit for bit:
//   crop_coeff_kernel      precompute_coeffs + normalize_coeffs_8bpc per frame and axis, in fp64 with explicit
//                          round-to-nearest intrinsics (no FMA contraction: Pillow's x86-64 build has none)
//   crop_resample_kernel   one CTA per (band of output rows, frame): horizontal pass of the input rows the band
//                          needs into an 8-bit shared-memory tile (rounded and clipped like Pillow's temporary
//                          image), vertical pass out of shared memory, flip, uint8/255, (t-mean)/std with IEEE
//                          division, written as fp32 [B,3,T,S,S] — the tensor the reference's DataLoader delivers
// Source frames are uint8 HWC in a device-resident store (a 180 GB B200 holds the decoded CREMA-D / Kinetics-
// Sounds frame sets), so a training step uploads 24 bytes per frame instead of 602 KB.
// HBM-bound byte work: one read of the crop region, one write of the fp32 batch.
#include <stdlib.h>
#include "common.cuh"

namespace gdl {

constexpr int kCropKMax = 16;                 // coefficients per output sample: ceil(scale)*2+1 <= 16, scale <= 7.5
constexpr int kCropRowInts = 2 + kCropKMax;   // xmin, count, coefficients
constexpr int kCropBandRows = 16;
constexpr int kCropThreads = 256;
constexpr int kPrecisionBits = 32 - 8 - 2;    // Resample.c PRECISION_BITS for 8-bit channels

struct CropParams {  // one frame: which stored frame, crop box (top, left, height, width), horizontal flip
  int src, i, j, h, w, flip;
};

// table[frame][axis (0 = horizontal / width, 1 = vertical / height)][kCropRowInts][S]: entry q of output sample xx
// at [q][xx] (q = 0 first input index, 1 count, 2.. coefficients), so that a warp working on consecutive xx
// reads consecutive ints (with [S][q] every lane touched its own sector: 32 wavefronts per coefficient load)
__global__ void crop_coeff_kernel(const CropParams* __restrict__ params, int* __restrict__ table, int S) {
  const int f = blockIdx.x, axis = blockIdx.y;
  const CropParams p = params[f];
  const int in_size = axis == 0 ? p.w : p.h;
  int* tab = table + ((size_t)(f * 2 + axis) * S) * kCropRowInts;
  const double scale = __ddiv_rn((double)in_size, (double)S);
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = filterscale;  // bilinear support 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, filterscale);
  for (int xx = threadIdx.x; xx < S; xx += blockDim.x) {
    const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
    int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    if (xmax > kCropKMax) xmax = kCropKMax;  // unreachable: the launcher bounds the scale
    double k[kCropKMax];
    double ww = 0.0;
#pragma unroll
    for (int x = 0; x < kCropKMax; ++x) {
      double w = 0.0;
      if (x < xmax) {
        double a = __dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss);
        if (a < 0.0) a = -a;
        w = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
        ww = __dadd_rn(ww, w);
      }
      k[x] = w;
    }
    int* row = tab + xx;  // stride S between entries
    row[0] = xmin;
    row[S] = xmax;
#pragma unroll
    for (int x = 0; x < kCropKMax; ++x) {
      int q = 0;
      if (x < xmax) {
        const double v = ww != 0.0 ? __ddiv_rn(k[x], ww) : k[x];
        q = v < 0.0 ? (int)__dadd_rn(-0.5, __dmul_rn(v, (double)(1 << kPrecisionBits)))
                    : (int)__dadd_rn(0.5, __dmul_rn(v, (double)(1 << kPrecisionBits)));
      }
      row[(size_t)(2 + x) * S] = q;
    }
  }
}

__device__ __forceinline__ int clip8(int acc) {
  const int v = acc >> kPrecisionBits;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// out[b][c][t][y][x], frame f = b*T + t.  Dynamic shared memory: max_rows * 3 * S bytes (planar 8-bit tile).
__global__ void __launch_bounds__(kCropThreads) crop_resample_kernel(
    const uint8_t* __restrict__ store, const CropParams* __restrict__ params, const int* __restrict__ table,
    float* __restrict__ out, int Hs, int Ws, int T, int S, int max_rows, float m0, float m1, float m2, float s0,
    float s1, float s2) {
  extern __shared__ uint8_t tile[];  // [rows][3][S]
  const int f = blockIdx.y;
  const int y0 = blockIdx.x * kCropBandRows;
  const int y1 = min(y0 + kCropBandRows, S);
  const CropParams p = params[f];
  const int* tab_h = table + ((size_t)(f * 2 + 0) * S) * kCropRowInts;
  const int* tab_v = table + ((size_t)(f * 2 + 1) * S) * kCropRowInts;
  // rows of the (cropped) input this band's vertical windows read
  const int r0 = __ldg(tab_v + y0);
  const int r1 = __ldg(tab_v + (y1 - 1)) + __ldg(tab_v + S + (y1 - 1));
  const int nrows = min(r1 - r0, max_rows);
  const uint8_t* src = store + ((size_t)p.src * Hs + p.i) * Ws * 3 + (size_t)p.j * 3;
  const bool same_w = p.w == S, same_h = p.h == S;  // Pillow skips a pass whose size does not change
  // One warp per (row, channel) line, lanes over xx: no per-item integer division, coefficient and pixel reads
  // of a warp are contiguous / near-contiguous.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = kCropThreads / 32;
  // ---- horizontal pass: tile[r][c][xx] = clip8(2^21 + sum_k kk[xx][k] * in[r0+r][xmin+k][c])
  for (int rc = warp; rc < nrows * 3; rc += kWarps) {
    const int r = rc / 3, c = rc - r * 3;
    const uint8_t* row = src + (size_t)(r0 + r) * Ws * 3 + c;
    uint8_t* trow = tile + (size_t)rc * S;
    for (int xx = lane; xx < S; xx += 32) {
      int v;
      if (same_w) {
        v = row[xx * 3];
      } else {
        const int* k = tab_h + xx;
        const int xmin = __ldg(k), cnt = __ldg(k + S);
        const uint8_t* px = row + xmin * 3;
        int acc = 1 << (kPrecisionBits - 1);
        for (int x = 0; x < cnt; ++x) acc += (int)px[x * 3] * __ldg(k + (2 + x) * S);
        v = clip8(acc);
      }
      trow[xx] = (uint8_t)v;
    }
  }
  __syncthreads();
  // ---- vertical pass + flip + ToTensor + Normalize
  const int b = f / T, t = f - b * T;
  const int band = y1 - y0;
  for (int yc = warp; yc < band * 3; yc += kWarps) {
    const int c = yc / band, y = y0 + (yc - c * band);
    const int* k = tab_v + y;  // warp-uniform: broadcast reads
    const int ymin = __ldg(k) - r0, cnt = __ldg(k + S);
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    const uint8_t* tcol = tile + ((size_t)ymin * 3 + c) * S;
    float* orow = out + (((size_t)(b * 3 + c) * T + t) * S + y) * S;
    for (int xx = lane; xx < S; xx += 32) {
      int v;
      if (same_h) {
        v = tcol[xx];  // identity weights
      } else {
        int acc = 1 << (kPrecisionBits - 1);
        for (int x = 0; x < cnt; ++x) acc += (int)tcol[(size_t)x * 3 * S + xx] * __ldg(k + (2 + x) * S);
        v = clip8(acc);
      }
      const float val = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.f), mean), sd);
      orow[p.flip ? S - 1 - xx : xx] = val;
    }
  }
}

}  // namespace gdl

using namespace gdl;

extern "C" int64_t gdl_crop_table_ints(int frames, int S) {
  if (frames <= 0 || S <= 0) return GDL_EINVAL;
  return (int64_t)frames * 2 * S * kCropRowInts;
}

// Rows of shared memory a band may need for frames of Hs x Ws: the band's windows span at most
// kCropBandRows*scale + 2*support + 2 input rows.
static int crop_max_rows(int Hs, int S) {
  const double scale = (double)Hs / S;
  const double fs = scale < 1.0 ? 1.0 : scale;
  return (int)(kCropBandRows * scale + 2.0 * fs + 3.0);
}

extern "C" int gdl_crop_resize_normalize(const uint8_t* store, int64_t store_frames, int Hs, int Ws,
                                         const int32_t* params, int frames, int T, int S, const float* mean3,
                                         const float* std3, float* out, int32_t* table, gdl_stream_t s) {
  GDL_REQUIRE(store && params && out && table && mean3 && std3, "gdl_crop_resize_normalize: null pointer");
  GDL_REQUIRE(store_frames > 0 && Hs > 0 && Ws > 0 && frames > 0 && T > 0 && frames % T == 0 && S > 0 && S <= 1024,
              "gdl_crop_resize_normalize: bad shape");
  const int hmax = Hs > Ws ? Hs : Ws;
  const int ks = ((hmax + S - 1) / S) * 2 + 1;  // ceil(max scale) * 2 + 1 (>= 3)
  GDL_REQUIRE(ks <= kCropKMax, "gdl_crop_resize_normalize: frames larger than 7.5x the output size are not supported");
  const int max_rows = crop_max_rows(Hs, S);
  const size_t smem = (size_t)max_rows * 3 * S;
  GDL_REQUIRE(smem <= 200 * 1024, "gdl_crop_resize_normalize: band tile does not fit in shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(crop_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(crop_resample)");
    attr_set = true;
  }
  const CropParams* cp = reinterpret_cast<const CropParams*>(params);
  crop_coeff_kernel<<<dim3(frames, 2), 256, 0, (cudaStream_t)s>>>(cp, table, S);
  GDL_CHECK_LAUNCH("crop_coeff_kernel");
  dim3 grid((S + kCropBandRows - 1) / kCropBandRows, frames);
  crop_resample_kernel<<<grid, kCropThreads, smem, (cudaStream_t)s>>>(store, cp, table, out, Hs, Ws, T, S, max_rows,
                                                                     mean3[0], mean3[1], mean3[2], std3[0], std3[1],
                                                                     std3[2]);
  GDL_CHECK_LAUNCH("crop_resample_kernel");
  return GDL_OK;
}

// ------------------------------------------------------------------------------------------
// audio half of the data pipeline: spectrogram = log(|librosa.stft(clip(wave), n_fft, hop)| + 1e-7)
// (reference dataset/CramedDataset.py:60-66: 22 050 Hz * 3 s tiled, n_fft 512, hop 353;
//  KSDataset.py:138-150 / VGGSoundDataset.py:112-122: 16 kHz, random 5-s window of the clip tiled to >= 10 s,
//  n_fft 256, hop 128).  librosa.stft: center=True (n_fft/2 samples of padding on both sides), periodic Hann
//  window, frames multiplied by the float64 window and transformed by a float64 FFT, the result stored as complex64;
//  np.abs and np.log then run in float32.  The kernel follows those precisions: fp64 window product and radix-2 FFT
//  in shared memory (low-energy bins of real audio sit 100+ dB under the frame's peak, below an fp32 FFT's noise
//  floor), components rounded to fp32, hypotf, logf.  One CTA transforms kStftFrames consecutive frames of one item
//  and writes out[b][f][t0 .. t0+kStftFrames) so every 32-byte sector of the [F][frames] plane is written whole.
// ------------------------------------------------------------------------------------------
namespace gdl {

constexpr int kStftFrames = 8;
constexpr int kStftMaxFft = 1024;

struct StftItem {
  int32_t clip, start;
};

__global__ void __launch_bounds__(kStftMaxFft / 2) log_stft_kernel(const float* __restrict__ waves, int64_t clip_stride,
                                                                   const int32_t* __restrict__ clip_len,
                                                                   const StftItem* __restrict__ items, int L, int n_fft,
                                                                   int log2n, int hop, int frames, int pad_zero,
                                                                   float* __restrict__ out) {
  extern __shared__ double sm[];
  double* re = sm;                  // [n_fft]
  double* im = re + n_fft;          // [n_fft]
  double* twr = im + n_fft;         // [n_fft/2]  cos(2 pi k / n)
  double* twi = twr + n_fft / 2;    // [n_fft/2] -sin(2 pi k / n)
  float* obuf = reinterpret_cast<float*>(twi + n_fft / 2);  // [n_fft/2 + 1][kStftFrames]
  const int tid = threadIdx.x, half_n = n_fft >> 1, F = half_n + 1;
  const int b = blockIdx.y, t0 = blockIdx.x * kStftFrames;
  const StftItem it = items[b];
  const float* w = waves + (int64_t)it.clip * clip_stride;
  const int len = clip_len[it.clip];
  {
    double s, c;
    sincospi(2.0 * (double)tid / (double)n_fft, &s, &c);
    twr[tid] = c;
    twi[tid] = -s;
  }
  for (int ft = 0; ft < kStftFrames; ++ft) {
    const int t = t0 + ft;
    __syncthreads();  // twiddles ready / previous frame's spectrum consumed
    if (t < frames) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = tid + h * half_n;
        int j = t * hop - half_n + k;  // index into the un-padded item
        double v = 0.0;
        bool inside = true;
        if (j < 0) {
          inside = !pad_zero;
          j = -j;                       // np.pad(mode="reflect"): x[-j] = x[j]
        } else if (j >= L) {
          inside = !pad_zero;
          j = 2 * (L - 1) - j;
        }
        if (inside) {
          float x = w[(int)(((int64_t)it.start + j) % len)];  // np.tile(...)[start:start+L]
          x = fminf(fmaxf(x, -1.f), 1.f);                     // samples[> 1] = 1, samples[< -1] = -1
          double s, c;
          sincospi(2.0 * (double)k / (double)n_fft, &s, &c);
          v = (double)x * (0.5 - 0.5 * c);                    // scipy.signal.get_window("hann", n_fft, fftbins=True)
        }
        const int r = (int)(__brev((unsigned)k) >> (32 - log2n));
        re[r] = v;
        im[r] = 0.0;
      }
      for (int st = 0; st < log2n; ++st) {
        __syncthreads();
        const int hb = 1 << st;
        const int pos = tid & (hb - 1);
        const int i0 = ((tid >> st) << (st + 1)) + pos, i1 = i0 + hb;
        const int tw = pos << (log2n - 1 - st);
        const double wr = twr[tw], wi = twi[tw];
        const double ar = re[i1] * wr - im[i1] * wi, ai = re[i1] * wi + im[i1] * wr;
        const double br = re[i0], bi = im[i0];
        re[i0] = br + ar;
        im[i0] = bi + ai;
        re[i1] = br - ar;
        im[i1] = bi - ai;
      }
      __syncthreads();
      for (int f = tid; f < F; f += half_n) {
        const float mag = hypotf((float)re[f], (float)im[f]);  // np.abs of the complex64 bin
        obuf[f * kStftFrames + ft] = logf(mag + 1e-7f);
      }
    }
  }
  __syncthreads();
  const int nt = min(kStftFrames, frames - t0);
  float* ob = out + (int64_t)b * F * frames + t0;
  for (int i = tid; i < F * kStftFrames; i += half_n) {
    const int f = i / kStftFrames, ft = i - f * kStftFrames;
    if (ft < nt) ob[(int64_t)f * frames + ft] = obuf[i];
  }
}

}  // namespace gdl

extern "C" int gdl_log_stft(const float* waves, int64_t clip_stride, const int32_t* clip_len, const int32_t* params, int B,
                            int L, int n_fft, int hop, int pad_mode, float* out, gdl_stream_t s) {
  using namespace gdl;
  GDL_REQUIRE(waves && clip_len && params && out, "gdl_log_stft: null pointer");
  GDL_REQUIRE(B > 0 && hop > 0 && n_fft >= 64 && n_fft <= kStftMaxFft && (n_fft & (n_fft - 1)) == 0 && L > n_fft / 2,
              "gdl_log_stft: bad shape (n_fft a power of two in [64, 1024], L > n_fft/2)");
  GDL_REQUIRE(pad_mode == 0 || pad_mode == 1, "gdl_log_stft: pad_mode is 0 (reflect) or 1 (zeros)");
  int log2n = 0;
  while ((1 << log2n) < n_fft) ++log2n;
  const int frames = 1 + L / hop;
  const size_t smem = (size_t)(2 * n_fft + n_fft) * sizeof(double) + (size_t)(n_fft / 2 + 1) * kStftFrames * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(log_stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(log_stft)");
    attr_set = true;
  }
  dim3 grid((frames + kStftFrames - 1) / kStftFrames, B);
  log_stft_kernel<<<grid, n_fft / 2, smem, (cudaStream_t)s>>>(waves, clip_stride, clip_len,
                                                             reinterpret_cast<const StftItem*>(params), L, n_fft, log2n, hop,
                                                             frames, pad_mode, out);
  GDL_CHECK_LAUNCH("log_stft_kernel");
  return GDL_OK;
}
