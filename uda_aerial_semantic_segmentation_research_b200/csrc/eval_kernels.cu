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
