// BatchNorm per-channel finalize helpers shared by nn_kernels.cu and stream_kernels.cu.
#pragma once
#include "common.cuh"

namespace uda {
namespace bn {

// mean/var -> scale/shift (+ running statistics update, momentum, unbiased running var)
struct BnFwdFinal {
  const float* gamma; const float* beta; float* running_mean; float* running_var;
  float* mean_out; float* rstd_out; float* scale_out; float* shift_out;
  long long M; float eps, momentum;
  unsigned int* counter;   // zero on entry, zero again on exit
};
// B200's FP64 pipe is ~1/64 of FP32 and a DP divide / sqrt is a long software sequence: a per-channel finalize
// written in double costs ~20 us per launch.  Only the cancellation-prone part (var = E[x^2] - mean^2) stays
// in double (three DP instructions); everything else is fp32.
__device__ __forceinline__ void bn_fwd_finalize_channel(const BnFwdFinal& f, double s1, double s2, int c) {
  const float inv_mf = 1.f / (float)f.M;
  const double inv_m = (double)inv_mf;   // exact enough: M is a pixel count (power-of-two multiples in practice)
  const double mean_d = s1 * inv_m;
  const float mean = (float)mean_d;
  const float var = fmaxf((float)(s2 * inv_m - mean_d * mean_d), 0.f);
  const float rstd = rsqrtf(var + f.eps);
  const float g = f.gamma ? f.gamma[c] : 1.f, b = f.beta ? f.beta[c] : 0.f;
  f.mean_out[c] = mean;
  f.rstd_out[c] = rstd;
  f.scale_out[c] = g * rstd;
  f.shift_out[c] = b - mean * g * rstd;
  if (f.running_mean) {
    const float unb = (f.M > 1) ? var * ((float)f.M / (float)(f.M - 1)) : var;
    f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * mean;
    f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * unb;
  }
}

// "last block done": returns true in every thread of the block that finished last (all partial sums of
// all blocks are then visible).  Classic threadfence reduction pattern.
__device__ __forceinline__ bool last_block_done(unsigned int* counter) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int total = gridDim.x * gridDim.y * gridDim.z;
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == total - 1);
  }
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// dgamma = sums[C+c], dbeta = sums[c] (accumulated into fp32 grads), and the per-channel
// coefficients of the apply pass:  dx = k0*g - k1 - k2*xhat
struct BnBwdFinal {
  const float* gamma; const float* mean; const float* rstd; float* dgamma; float* dbeta; float* coef;
  long long M; int accumulate;
  unsigned int* counter;
};
__device__ __forceinline__ void bn_bwd_finalize_channel(const BnBwdFinal& f, double s1d, double s2d, int c, int C) {
  // dx = k0*g - k1 - k2*xhat  with  k1 = k0*s1/M, k2 = k0*s2/M, xhat = (x-mean)*rstd
  //    = A*g + Bc*x + Cc     (three per-channel coefficients for the apply pass); fp32 throughout (see above)
  const float s1 = (float)s1d, s2 = (float)s2d;
  const float inv_m = 1.f / (float)f.M;
  const float g = f.gamma ? f.gamma[c] : 1.f;
  const float rs = f.rstd[c], mu = f.mean[c];
  const float k0 = g * rs;
  const float k1 = k0 * s1 * inv_m, k2 = k0 * s2 * inv_m;
  if (f.dgamma) f.dgamma[c] = (f.accumulate ? f.dgamma[c] : 0.f) + s2;
  if (f.dbeta) f.dbeta[c] = (f.accumulate ? f.dbeta[c] : 0.f) + s1;
  f.coef[c] = k0;
  f.coef[C + c] = -k2 * rs;
  f.coef[2 * C + c] = -k1 + k2 * rs * mu;
}

}  // namespace bn

// stream_kernels.cu: bulk-copy streaming fast paths (bf16, C a power of two <= 2048, >= 1 MB tensors)
bool bn_stream_ok(int dtype, long long M, int C);
int bn_apply_stream(const void* x, const void* residual, void* y, const float* scale, const float* shift,
                    const double* sums, const bn::BnFwdFinal& fin, long long M, int C, float slope, cudaStream_t st);
int bn_apply_maxpool_stream(const void* x, void* a, void* y, unsigned char* idx, const double* sums,
                            const bn::BnFwdFinal& fin, int B, int H, int W, int C, float slope, cudaStream_t st);
int bn_bwd_stream(const void* dy, const void* x, const void* a, const float* mean, const float* rstd,
                  const float* scale, const float* shift, void* dx, void* dres, int dres_accumulate, double* sums,
                  float* coef, const bn::BnBwdFinal& fin, long long M, int C, float slope, cudaStream_t st);
int bn_bwd_apply_fused_stream(const void* dy, const void* x, const void* a, const double* sums, int v_is_z,
                              const float* gamma, const float* beta, const float* mean, const float* rstd,
                              const float* scale, const float* shift, void* dx, void* dres, int dres_accumulate,
                              float* dgamma, float* dbeta, int param_accumulate, long long M, int C, float slope,
                              cudaStream_t st);
}  // namespace uda
