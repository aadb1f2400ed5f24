// Output-space adversarial path (north-star config 3: output-space discriminator behind a gradient-reversal layer),
// sm_100a.  The discriminator reads softmax(logits); the reference defines the gradient-reversal layer
// (src/models/uda.py:99-112: identity forward, -alpha * grad backward) and the discriminator
// (src/models/discriminator.py:4-55) but never wires them together — this is that wiring, as kernels:
//
//   uda_softmax_nchw_to_nhwc      fp32 NCHW logits -> channels-last bf16 probabilities padded to Cpad channels
//                                 (the tensor-core operand layout of the discriminator's first convolution)
//   uda_softmax_bwd_grl           dlogits (fp32 NCHW) (+)= scale * p_c * (dp_c - sum_k p_k dp_k): softmax backward with
//                                 the gradient-reversal factor (scale = -alpha) folded in, optionally accumulated
//                                 straight into the logit-gradient buffer the segmentation loss produced
//   uda_scale_f32 / uda_scale_bf16  y = scale * x: the stand-alone GradientReverseFunction backward (one pass)
//   uda_pad_channels / uda_unpad_channels_add   [rows][c] <-> [rows][cpad] glue for a first-layer weight whose input
//                                 channel count (the class count) is not a tensor-core channel atom
// All HBM-bound, one thread per pixel, class loop in registers (C <= 32).
#include "common.cuh"

namespace uda {
namespace {

constexpr int kMaxC = 32;

template <int CPAD>
__global__ void __launch_bounds__(256)
softmax_nchw_to_nhwc_kernel(const float* __restrict__ z, bf16* __restrict__ p, int C, long long HW, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // pixel index over B*HW
  if (i >= total) return;
  const long long b = i / HW, hw = i - b * HW;
  const float* zp = z + b * C * HW + hw;
  float v[kMaxC];
  float m = -INFINITY;
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) {
    v[c] = c < C ? __ldg(zp + (long long)c * HW) : -INFINITY;
    m = fmaxf(m, v[c]);
  }
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxC; ++c) { v[c] = c < C ? __expf(v[c] - m) : 0.f; s += v[c]; }
  const float inv = 1.f / s;
  bf16* out = p + i * CPAD;
#pragma unroll
  for (int c = 0; c < CPAD; c += 8) {
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = v[c + e] * inv;
    st_vec<8>(out + c, o);
  }
}

template <int CPAD>
__global__ void __launch_bounds__(256)
softmax_bwd_grl_kernel(const bf16* __restrict__ p, const bf16* __restrict__ dp, float* __restrict__ dz, float scale,
                       int accumulate, int C, long long HW, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / HW, hw = i - b * HW;
  float pv[CPAD], gv[CPAD];
#pragma unroll
  for (int c = 0; c < CPAD; c += 8) {
    float a[8], g[8];
    ld_vec<8>(p + i * CPAD + c, a);
    ld_vec<8>(dp + i * CPAD + c, g);
#pragma unroll
    for (int e = 0; e < 8; ++e) { pv[c + e] = a[e]; gv[c + e] = g[e]; }
  }
  float dot = 0.f;
#pragma unroll
  for (int c = 0; c < CPAD; ++c) dot = fmaf(pv[c], gv[c], dot);
  float* out = dz + b * C * HW + hw;
#pragma unroll
  for (int c = 0; c < CPAD; ++c) {
    if (c < C) {
      const float g = scale * pv[c] * (gv[c] - dot);
      float* q = out + (long long)c * HW;
      *q = accumulate ? *q + g : g;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) scale_kernel(const T* __restrict__ x, T* __restrict__ y, float s, long long n,
                                                    int vec_ok) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  if (vec_ok && i + 8 <= n) {
    float v[8];
    ld_vec<8>(x + i, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= s;
    st_vec<8>(y + i, v);
  } else {
    const long long e = i + 8 < n ? i + 8 : n;
    for (long long j = i; j < e; ++j) y[j] = from_f<T>(to_f(x[j]) * s);
  }
}

__global__ void pad_channels_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, long long rows, int c, int cpad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cpad) return;
  const long long r = i / cpad;
  const int k = (int)(i - r * cpad);
  dst[i] = k < c ? src[r * c + k] : __float2bfloat16_rn(0.f);
}
__global__ void unpad_channels_add_kernel(const float* __restrict__ src, float* __restrict__ dst, long long rows, int c,
                                          int cpad) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * c) return;
  const long long r = i / c;
  const int k = (int)(i - r * c);
  dst[i] += src[r * cpad + k];
}

}  // namespace
}  // namespace uda

using namespace uda;

extern "C" int uda_softmax_nchw_to_nhwc(const float* logits, void* probs, int B, int C, int Cpad, long long HW,
                                        void* stream) {
  UDA_REQUIRE(logits && probs && B > 0 && HW > 0, UDA_ERR_BAD_ARG, "softmax_nchw_to_nhwc: bad argument");
  UDA_REQUIRE(C >= 1 && C <= kMaxC && (Cpad == 8 || Cpad == 16 || Cpad == 32) && Cpad >= C, UDA_ERR_UNSUPPORTED,
              "softmax_nchw_to_nhwc: C = %d (<= 32) with Cpad in {8, 16, 32} >= C expected", C);
  UDA_REQUIRE(aligned<bf16>(probs, 16), UDA_ERR_BAD_ARG, "softmax_nchw_to_nhwc: output must be 16-byte aligned");
  const long long total = (long long)B * HW;
  const unsigned grid = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (Cpad == 32) softmax_nchw_to_nhwc_kernel<32><<<grid, 256, 0, st>>>(logits, (bf16*)probs, C, HW, total);
  else if (Cpad == 16) softmax_nchw_to_nhwc_kernel<16><<<grid, 256, 0, st>>>(logits, (bf16*)probs, C, HW, total);
  else softmax_nchw_to_nhwc_kernel<8><<<grid, 256, 0, st>>>(logits, (bf16*)probs, C, HW, total);
  UDA_LAUNCH_OK("softmax_nchw_to_nhwc_kernel");
  return UDA_OK;
}

extern "C" int uda_softmax_bwd_grl(const void* probs, const void* dprobs, float* dlogits, float scale, int accumulate,
                                   int B, int C, int Cpad, long long HW, void* stream) {
  UDA_REQUIRE(probs && dprobs && dlogits && B > 0 && HW > 0, UDA_ERR_BAD_ARG, "softmax_bwd_grl: bad argument");
  UDA_REQUIRE(C >= 1 && C <= kMaxC && (Cpad == 8 || Cpad == 16 || Cpad == 32) && Cpad >= C, UDA_ERR_UNSUPPORTED,
              "softmax_bwd_grl: C = %d (<= 32) with Cpad in {8, 16, 32} >= C expected", C);
  UDA_REQUIRE(aligned<bf16>(probs, 16) && aligned<bf16>(dprobs, 16), UDA_ERR_BAD_ARG,
              "softmax_bwd_grl: probs / dprobs must be 16-byte aligned");
  const long long total = (long long)B * HW;
  const unsigned grid = (unsigned)((total + 255) / 256);
  cudaStream_t st = (cudaStream_t)stream;
  const bf16 *p = (const bf16*)probs, *dp = (const bf16*)dprobs;
  if (Cpad == 32) softmax_bwd_grl_kernel<32><<<grid, 256, 0, st>>>(p, dp, dlogits, scale, accumulate, C, HW, total);
  else if (Cpad == 16) softmax_bwd_grl_kernel<16><<<grid, 256, 0, st>>>(p, dp, dlogits, scale, accumulate, C, HW, total);
  else softmax_bwd_grl_kernel<8><<<grid, 256, 0, st>>>(p, dp, dlogits, scale, accumulate, C, HW, total);
  UDA_LAUNCH_OK("softmax_bwd_grl_kernel");
  return UDA_OK;
}

extern "C" int uda_scale(const void* x, void* y, int dtype, float scale, long long n, void* stream) {
  UDA_REQUIRE(x && y && n >= 0, UDA_ERR_BAD_ARG, "scale: bad argument");
  if (n == 0) return UDA_OK;
  const unsigned grid = (unsigned)((n + 2047) / 2048);
  cudaStream_t st = (cudaStream_t)stream;
  const int vec_ok = aligned<float>(x, 32) && aligned<float>(y, 32);   // 8 elements per access (16 / 32 bytes)
  if (dtype == UDA_F32) {
    scale_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (float*)y, scale, n, vec_ok);
  } else if (dtype == UDA_BF16) {
    scale_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, (bf16*)y, scale, n, vec_ok);
  } else {
    return set_error(UDA_ERR_BAD_ARG, "scale: dtype must be UDA_F32 or UDA_BF16");
  }
  UDA_LAUNCH_OK("scale_kernel");
  return UDA_OK;
}

extern "C" int uda_pad_channels(const void* src, void* dst, long long rows, int c, int cpad, void* stream) {
  UDA_REQUIRE(src && dst && rows > 0 && c > 0 && cpad >= c, UDA_ERR_BAD_ARG, "pad_channels: bad argument");
  const long long n = rows * cpad;
  pad_channels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)src, (bf16*)dst, rows, c, cpad);
  UDA_LAUNCH_OK("pad_channels_kernel");
  return UDA_OK;
}

extern "C" int uda_unpad_channels_add(const float* src, float* dst, long long rows, int c, int cpad, void* stream) {
  UDA_REQUIRE(src && dst && rows > 0 && c > 0 && cpad >= c, UDA_ERR_BAD_ARG, "unpad_channels_add: bad argument");
  const long long n = rows * c;
  unpad_channels_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, dst, rows, c, cpad);
  UDA_LAUNCH_OK("unpad_channels_add_kernel");
  return UDA_OK;
}
