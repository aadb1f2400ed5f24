// Prediction / evaluation kernels: fused argmax + confusion-matrix histogram (integer, bit-exact).
//
// Reference semantics (paths relative to the reference tree):
//   argmax mask        src/models/predict.py:113-130, src/models/train.py:227  (first maximal index,
//                      NaN counts as maximal — torch.argmax)
//   confusion matrix   src/analysis/metrics.py:17-27  (mask 0<=true<C [& != ignore]; bincount(C*true+pred))
//
// HBM-bound: every logit is read once (NCHW planes, 16/8-byte vectors), the histogram is privatised
// per warp in shared memory (uniform-warp aggregation for blocky label maps), then flushed with one
// 64-bit global atomic per non-empty bin per CTA.
#include "common.cuh"

namespace uda {
namespace {

constexpr int kEvalThreads = 256;

__device__ __forceinline__ void hist_add(unsigned int* sh, int bin, bool ok) {
  // warp-uniform fast path: all 32 lanes active & same bin -> one atomic of 32
  const unsigned full = 0xffffffffu;
  int b0 = __shfl_sync(full, bin, 0);
  bool uni = __all_sync(full, ok && bin == b0);
  if (uni) {
    if ((threadIdx.x & 31) == 0) atomicAdd(sh + b0, 32u);
  } else if (ok) {
    atomicAdd(sh + bin, 1u);
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kEvalThreads)
argmax_confmat_kernel(const T* __restrict__ logits, const long long* __restrict__ target,
                      long long* __restrict__ mask64, unsigned char* __restrict__ mask8,
                      unsigned long long* __restrict__ hist, unsigned long long* __restrict__ bad,
                      int B, int C, long long HW, long long ignore_index, int has_ignore, int ncopies) {
  extern __shared__ unsigned int sh_hist[];
  const int CC = C * C;
  const bool do_hist = (target != nullptr) && (hist != nullptr);
  if (do_hist && ncopies > 0) {
    for (int i = threadIdx.x; i < CC * ncopies; i += blockDim.x) sh_hist[i] = 0u;
    __syncthreads();
  }
  unsigned int* my = sh_hist + ((threadIdx.x >> 5) % (ncopies > 0 ? ncopies : 1)) * CC;
  const long long nvec_img = HW / VEC;
  const long long nvec = nvec_img * B;
  unsigned long long nbad = 0;
  // warp-uniform trip count so the full-mask shuffles in hist_add stay legal
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long base0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
  for (long long wbase = base0; wbase < nvec; wbase += stride) {
    const long long iv = wbase + (threadIdx.x & 31);
    const bool act = iv < nvec;
    int idx[VEC];
    long long b = 0, px = 0;
    if (act) {
      b = iv / nvec_img;
      px = (iv - b * nvec_img) * VEC;
      const T* zp = logits + b * C * HW + px;
      float best[VEC];
      ld_vec<VEC>(zp, best);
#pragma unroll
      for (int j = 0; j < VEC; ++j) idx[j] = 0;
#pragma unroll 4
      for (int c = 1; c < C; ++c) {
        float v[VEC];
        ld_vec<VEC>(zp + (long long)c * HW, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          bool take = (v[j] > best[j]) || ((v[j] != v[j]) && (best[j] == best[j]));
          best[j] = take ? v[j] : best[j];
          idx[j] = take ? c : idx[j];
        }
      }
      if (mask64) {
        long long* mp = mask64 + b * HW + px;
        if constexpr (VEC == 4) {
          *reinterpret_cast<longlong2*>(mp) = make_longlong2(idx[0], idx[1]);
          *reinterpret_cast<longlong2*>(mp + 2) = make_longlong2(idx[2], idx[3]);
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j) mp[j] = idx[j];
        }
      }
      if (mask8) {
        unsigned char* mp = mask8 + b * HW + px;
        if constexpr (VEC == 4) {
          *reinterpret_cast<uchar4*>(mp) = make_uchar4(idx[0], idx[1], idx[2], idx[3]);
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j) mp[j] = (unsigned char)idx[j];
        }
      }
    }
    if (do_hist) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        bool ok = false;
        int bin = 0;
        if (act) {
          long long t = __ldg(target + b * HW + px + j);
          ok = (t >= 0) && (t < C) && !(has_ignore && t == ignore_index);
          bin = ok ? (int)t * C + idx[j] : 0;
        }
        if (ncopies > 0) hist_add(my, bin, ok);
        else if (ok) atomicAdd(hist + bin, 1ull);
      }
    }
  }
  (void)nbad;
  if (do_hist && ncopies > 0) {
    __syncthreads();
    for (int i = threadIdx.x; i < CC; i += blockDim.x) {
      unsigned long long s = 0;
      for (int k = 0; k < ncopies; ++k) s += sh_hist[k * CC + i];
      if (s) atomicAdd(hist + i, s);
    }
  }
  (void)bad;
}

// Confusion matrix of two index maps (SegmentationMetrics._fast_hist drop-in).  PT = pred element type.
template <typename PT>
__global__ void __launch_bounds__(kEvalThreads)
confmat_kernel(const PT* __restrict__ pred, const long long* __restrict__ target,
               unsigned long long* __restrict__ hist, unsigned long long* __restrict__ bad, long long n, int C,
               long long ignore_index, int has_ignore, int ncopies) {
  extern __shared__ unsigned int sh_hist[];
  const int CC = C * C;
  if (ncopies > 0) {
    for (int i = threadIdx.x; i < CC * ncopies; i += blockDim.x) sh_hist[i] = 0u;
    __syncthreads();
  }
  unsigned int* my = sh_hist + ((threadIdx.x >> 5) % (ncopies > 0 ? ncopies : 1)) * CC;
  unsigned long long nbad = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long base0 = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
  for (long long wbase = base0; wbase < n; wbase += stride) {
    const long long i = wbase + (threadIdx.x & 31);
    bool ok = false;
    int bin = 0;
    if (i < n) {
      long long t = __ldg(target + i);
      long long pr = (long long)__ldg(pred + i);
      ok = (t >= 0) && (t < C) && !(has_ignore && t == ignore_index);
      if (ok && (pr < 0 || pr >= C)) { ok = false; ++nbad; }  // torch.bincount would grow/raise here
      bin = ok ? (int)t * C + (int)pr : 0;
    }
    if (ncopies > 0) hist_add(my, bin, ok);
    else if (ok) atomicAdd(hist + bin, 1ull);
  }
  if (nbad && bad) atomicAdd(bad, nbad);
  if (ncopies > 0) {
    __syncthreads();
    for (int i = threadIdx.x; i < CC; i += blockDim.x) {
      unsigned long long s = 0;
      for (int k = 0; k < ncopies; ++k) s += sh_hist[k * CC + i];
      if (s) atomicAdd(hist + i, s);
    }
  }
}

inline int hist_copies(int C) {
  const size_t one = (size_t)C * C * sizeof(unsigned int);
  if (one > 40 * 1024) return 0;  // too many classes for smem privatisation: global atomics
  size_t k = (40 * 1024) / one;
  if (k > (size_t)(kEvalThreads / 32)) k = kEvalThreads / 32;
  return (int)k;
}

}  // namespace
}  // namespace uda

using namespace uda;

extern "C" int uda_argmax_confmat(const void* logits, int dtype, const long long* target, long long* mask_i64,
                                  unsigned char* mask_u8, long long* hist, int B, int C, long long HW,
                                  long long ignore_index, int has_ignore, int zero_hist, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(logits, UDA_ERR_BAD_ARG, "argmax_confmat: null logits");
  UDA_REQUIRE(dtype == UDA_F32 || dtype == UDA_BF16, UDA_ERR_BAD_ARG, "argmax_confmat: dtype %d", dtype);
  UDA_REQUIRE(B > 0 && C > 0 && HW > 0, UDA_ERR_BAD_ARG, "argmax_confmat: empty shape");
  UDA_REQUIRE(!mask_u8 || C <= 256, UDA_ERR_UNSUPPORTED, "argmax_confmat: u8 mask needs C <= 256");
  UDA_REQUIRE((hist != nullptr) == (target != nullptr) || hist == nullptr, UDA_ERR_BAD_ARG,
              "argmax_confmat: hist requires targets");
  UDA_REQUIRE(mask_i64 || mask_u8 || hist, UDA_ERR_BAD_ARG, "argmax_confmat: no output requested");
  if (hist && zero_hist) UDA_CUDA_OK(cudaMemsetAsync(hist, 0, (size_t)C * C * sizeof(long long), st));
  const int ncopies = hist ? hist_copies(C) : 0;
  const size_t smem = (size_t)ncopies * C * C * sizeof(unsigned int);
  const bool bf = dtype == UDA_BF16;
  const size_t es = bf ? 2 : 4;
  const bool v4 = (HW % 4 == 0) && (reinterpret_cast<uintptr_t>(logits) % (4 * es) == 0) &&
                  (!mask_i64 || reinterpret_cast<uintptr_t>(mask_i64) % 16 == 0) &&
                  (!mask_u8 || reinterpret_cast<uintptr_t>(mask_u8) % 4 == 0);
  const long long nvec = (long long)B * (HW / (v4 ? 4 : 1));
  long long blocks = (nvec + kEvalThreads - 1) / kEvalThreads;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  unsigned long long* h = reinterpret_cast<unsigned long long*>(hist);
#define UDA_AC(T, V)                                                                                   \
  argmax_confmat_kernel<T, V><<<(unsigned)blocks, kEvalThreads, smem, st>>>(                           \
      (const T*)logits, target, mask_i64, mask_u8, h, nullptr, B, C, HW, ignore_index, has_ignore, ncopies)
  if (bf) { if (v4) UDA_AC(bf16, 4); else UDA_AC(bf16, 1); }
  else    { if (v4) UDA_AC(float, 4); else UDA_AC(float, 1); }
#undef UDA_AC
  UDA_LAUNCH_OK("argmax_confmat_kernel");
  return UDA_OK;
}

extern "C" int uda_confmat(const void* pred, int pred_dtype, const long long* target, long long* hist,
                           long long* bad_count, long long n, int C, long long ignore_index, int has_ignore,
                           int zero_hist, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(hist && C > 0 && n >= 0, UDA_ERR_BAD_ARG, "confmat: bad argument");
  UDA_REQUIRE(pred_dtype == UDA_I64 || pred_dtype == UDA_U8, UDA_ERR_BAD_ARG, "confmat: pred dtype %d", pred_dtype);
  if (zero_hist) UDA_CUDA_OK(cudaMemsetAsync(hist, 0, (size_t)C * C * sizeof(long long), st));
  if (bad_count) UDA_CUDA_OK(cudaMemsetAsync(bad_count, 0, sizeof(long long), st));
  if (n == 0) return UDA_OK;
  UDA_REQUIRE(pred && target, UDA_ERR_BAD_ARG, "confmat: null input");
  const int ncopies = hist_copies(C);
  const size_t smem = (size_t)ncopies * C * C * sizeof(unsigned int);
  long long blocks = (n + kEvalThreads - 1) / kEvalThreads;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  unsigned long long* h = reinterpret_cast<unsigned long long*>(hist);
  unsigned long long* bad = reinterpret_cast<unsigned long long*>(bad_count);
  if (pred_dtype == UDA_I64)
    confmat_kernel<long long><<<(unsigned)blocks, kEvalThreads, smem, st>>>(
        (const long long*)pred, target, h, bad, n, C, ignore_index, has_ignore, ncopies);
  else
    confmat_kernel<unsigned char><<<(unsigned)blocks, kEvalThreads, smem, st>>>(
        (const unsigned char*)pred, target, h, bad, n, C, ignore_index, has_ignore, ncopies);
  UDA_LAUNCH_OK("confmat_kernel");
  return UDA_OK;
}

// ================================================================================================================
// Metrics from the device-resident confusion matrix (the reference derives 26 scalars per step
// with one `.item()` each, src/models/train.py:225-243; here they are derived on the device and read once per epoch).
//   out[0]          mean IoU of the in-tree metric: nanmean(diag / (row + col - diag + 1e-7))   (src/analysis/metrics.py:29-42)
//   out[1]          pixel accuracy: sum(diag) / sum(hist)   (0 when the matrix is empty)         (train.py:232)
//   out[2]          macro Jaccard with torchmetrics semantics (train.py:209-212,231): mean over the classes that occur
//                   in the target or the prediction of diag / (row + col - diag); 0 when no class occurs
//   out[3]          number of pixels counted
//   out[4 .. 4+C)   per-class IoU of the in-tree metric
//   out[4+C .. 4+2C) per-class binary Jaccard (train.py:236-241): tp / (tp + fp + fn), 0 for 0/0
// One CTA; float64 like the numpy formulas.
// ================================================================================================================
namespace uda {
namespace {
__global__ void __launch_bounds__(256) metrics_from_hist_kernel(const long long* __restrict__ hist, int C, double* __restrict__ out) {
  extern __shared__ double sm[];      // row[C], col[C], diag[C]
  double* row = sm; double* col = sm + C; double* dg = sm + 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    long long r = 0, k = 0;
    for (int j = 0; j < C; ++j) { r += hist[(long long)c * C + j]; k += hist[(long long)j * C + c]; }
    row[c] = (double)r; col[c] = (double)k; dg[c] = (double)hist[(long long)c * C + c];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double uni = row[c] + col[c] - dg[c];
    out[4 + c] = dg[c] / (uni + 1e-7);
    out[4 + C + c] = uni > 0.0 ? dg[c] / uni : 0.0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s_iou = 0.0, s_jac = 0.0, correct = 0.0, total = 0.0;
    int n_iou = 0, n_jac = 0;
    for (int c = 0; c < C; ++c) {
      const double v = out[4 + c];
      if (v == v) { s_iou += v; ++n_iou; }                          // nanmean (the 1e-7 keeps every term finite)
      if (row[c] + col[c] > 0.0) { s_jac += out[4 + C + c]; ++n_jac; }
      correct += dg[c]; total += row[c];
    }
    out[0] = n_iou ? s_iou / n_iou : 0.0;
    out[1] = total > 0.0 ? correct / total : 0.0;
    out[2] = n_jac ? s_jac / n_jac : 0.0;
    out[3] = total;
  }
}

// ---- sliding-window evaluation glue (BASELINE config 5; reference preprocessing src/models/predict.py:93-97:
//      ToTensor (u8 HWC -> float CHW / 255) + Normalize(mean, std)) ------------------------------------------------
// windows [first, first + count) of a [H, W, 3] uint8 tile (row-major window grid, window = stride = win) as normalised
// fp32 NCHW network input
__global__ void __launch_bounds__(256)
gather_windows_u8_kernel(const unsigned char* __restrict__ tile, float* __restrict__ out, int H, int W, int win, int first,
                         int count, float m0, float m1, float m2, float i0, float i1, float i2) {
  const long long n = (long long)count * win * win;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int wpr = W / win;
  const int k = (int)(i / ((long long)win * win));
  const int r = (int)(i - (long long)k * win * win);
  const int y = r / win, x = r - y * win;
  const int wi = first + k, wy = wi / wpr, wx = wi - wy * wpr;
  const unsigned char* p = tile + ((long long)(wy * win + y) * W + (wx * win + x)) * 3;
  float* o = out + (long long)k * 3 * win * win + r;
  const long long plane = (long long)win * win;
  o[0] = ((float)p[0] * (1.f / 255.f) - m0) * i0;
  o[plane] = ((float)p[1] * (1.f / 255.f) - m1) * i1;
  o[2 * plane] = ((float)p[2] * (1.f / 255.f) - m2) * i2;
}
// the same windows of an int64 / uint8 [H, W] label tile as int64 [count, win, win] (confusion-matrix targets)
template <typename T>
__global__ void __launch_bounds__(256)
gather_label_windows_kernel(const T* __restrict__ tile, long long* __restrict__ out, int W, int win, int first, int count) {
  const long long n = (long long)count * win * win;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int wpr = W / win;
  const int k = (int)(i / ((long long)win * win));
  const int r = (int)(i - (long long)k * win * win);
  const int y = r / win, x = r - y * win;
  const int wi = first + k, wy = wi / wpr, wx = wi - wy * wpr;
  out[i] = (long long)tile[(long long)(wy * win + y) * W + (wx * win + x)];
}
// uint8 window masks [count, win, win] back into the [H, W] uint8 tile mask
__global__ void __launch_bounds__(256)
scatter_window_masks_kernel(const unsigned char* __restrict__ masks, unsigned char* __restrict__ tile, int W, int win,
                            int first, int count) {
  const long long n = (long long)count * win * win / 4;     // 4 pixels per thread (win % 4 == 0)
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int wpr = W / win;
  const long long e = i * 4;
  const int k = (int)(e / ((long long)win * win));
  const int r = (int)(e - (long long)k * win * win);
  const int y = r / win, x = r - y * win;
  const int wi = first + k, wy = wi / wpr, wx = wi - wy * wpr;
  *reinterpret_cast<uchar4*>(tile + (long long)(wy * win + y) * W + (wx * win + x)) =
      *reinterpret_cast<const uchar4*>(masks + e);
}
}  // namespace
}  // namespace uda

extern "C" int uda_metrics_from_hist(const long long* hist, int C, double* out, void* stream) {
  UDA_REQUIRE(hist && out && C > 0 && C <= 1024, UDA_ERR_BAD_ARG, "metrics_from_hist: bad argument");
  uda::metrics_from_hist_kernel<<<1, 256, 3 * C * sizeof(double), (cudaStream_t)stream>>>(hist, C, out);
  UDA_LAUNCH_OK("metrics_from_hist_kernel");
  return UDA_OK;
}

static int check_windows(int H, int W, int win, int first, int count, const char* who) {
  UDA_REQUIRE(H > 0 && W > 0 && win > 0 && H % win == 0 && W % win == 0, UDA_ERR_BAD_ARG,
              "%s: tile %dx%d is not a multiple of the window %d", who, H, W, win);
  UDA_REQUIRE(first >= 0 && count > 0 && first + count <= (H / win) * (W / win), UDA_ERR_BAD_ARG,
              "%s: windows [%d, %d) outside the %d windows of the tile", who, first, first + count, (H / win) * (W / win));
  return UDA_OK;
}

extern "C" int uda_gather_windows_u8(const unsigned char* tile_hwc, float* out_nchw, int H, int W, int win, int first,
                                     int count, const float* mean3, const float* std3, void* stream) {
  UDA_REQUIRE(tile_hwc && out_nchw && mean3 && std3, UDA_ERR_BAD_ARG, "gather_windows_u8: null pointer");
  if (int rc = check_windows(H, W, win, first, count, "gather_windows_u8")) return rc;
  const long long n = (long long)count * win * win;
  uda::gather_windows_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      tile_hwc, out_nchw, H, W, win, first, count, mean3[0], mean3[1], mean3[2], 1.f / std3[0], 1.f / std3[1], 1.f / std3[2]);
  UDA_LAUNCH_OK("gather_windows_u8_kernel");
  return UDA_OK;
}

extern "C" int uda_gather_label_windows(const void* tile, int dtype, long long* out, int H, int W, int win, int first,
                                        int count, void* stream) {
  UDA_REQUIRE(tile && out, UDA_ERR_BAD_ARG, "gather_label_windows: null pointer");
  UDA_REQUIRE(dtype == UDA_I64 || dtype == UDA_U8, UDA_ERR_BAD_ARG, "gather_label_windows: dtype must be UDA_I64 or UDA_U8");
  if (int rc = check_windows(H, W, win, first, count, "gather_label_windows")) return rc;
  const long long n = (long long)count * win * win;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (dtype == UDA_I64)
    uda::gather_label_windows_kernel<long long><<<grid, 256, 0, (cudaStream_t)stream>>>((const long long*)tile, out, W, win, first, count);
  else
    uda::gather_label_windows_kernel<unsigned char><<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)tile, out, W, win, first, count);
  UDA_LAUNCH_OK("gather_label_windows_kernel");
  return UDA_OK;
}

extern "C" int uda_scatter_window_masks(const unsigned char* masks, unsigned char* tile_mask, int H, int W, int win,
                                        int first, int count, void* stream) {
  UDA_REQUIRE(masks && tile_mask, UDA_ERR_BAD_ARG, "scatter_window_masks: null pointer");
  UDA_REQUIRE(win % 4 == 0 && W % 4 == 0 && uda::aligned<unsigned char>(masks, 4) && uda::aligned<unsigned char>(tile_mask, 4),
              UDA_ERR_UNSUPPORTED, "scatter_window_masks: window / tile width must be multiples of 4, buffers 4-byte aligned");
  if (int rc = check_windows(H, W, win, first, count, "scatter_window_masks")) return rc;
  const long long n = (long long)count * win * win / 4;
  uda::scatter_window_masks_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(masks, tile_mask, W, win, first, count);
  UDA_LAUNCH_OK("scatter_window_masks_kernel");
  return UDA_OK;
}
