// HBM-bound layer kernels around the convolutions (all activations NHWC, fp32 or bf16):
//   layout conversion NCHW fp32 <-> NHWC T, BatchNorm (batch statistics / apply / backward),
//   max-pool 3x3 s2 p1, nearest x2 upsample + skip concat, global average pool, fused Adam.
//
// Reference semantics: the U-Net the reference builds with smp.Unet (src/models/train.py:572-577) —
// BatchNorm2d eps=1e-5, momentum=0.1, ReLU, MaxPool2d(3,2,1), nearest x2 upsample + channel concat
//  — and the discriminator's BatchNorm + LeakyReLU(0.2) + AdaptiveAvgPool
// (src/models/discriminator.py:15-42); torch.optim.Adam defaults (src/models/train.py:461).
#include "bn_common.cuh"
#include <stdlib.h>

namespace uda {
namespace {

using namespace bn;

constexpr int kThreads = 256;

inline bool use_stream() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UDA_B200_BN_STREAM");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

inline unsigned grid_for(long long work_items, int per_sm = 8) {
  long long blocks = (work_items + kThreads - 1) / kThreads;
  long long cap = (long long)num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

// CTAs of a per-channel reduction: enough to fill the machine for big tensors, few for small ones (every CTA
// ends with 2*C atomics onto the same addresses)
inline long long reduce_grid_cap(long long bytes, int slabs) {
  long long want = bytes / (256 * 1024);           // >= 256 KB of input per CTA
  const long long hi = 2LL * num_sms();
  if (want > hi) want = hi;
  if (want < 1) want = 1;
  return (want + slabs - 1) / slabs;
}

// grid whose total thread count is a multiple of `period` (= C/VEC channel vectors), so that a
// grid-stride loop keeps every thread on the same channel vector
inline unsigned grid_for_channels(long long work_items, int period, int per_sm = 8) {
  long long blocks = (work_items + kThreads - 1) / kThreads;
  long long cap = (long long)num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  // kThreads * blocks must be a multiple of period
  long long need = 1;
  while ((need * kThreads) % period) ++need;   // period <= 512 in practice: tiny loop
  blocks = (blocks + need - 1) / need * need;
  return (unsigned)blocks;
}

// ---------------------------------------------------------------------------------------------
// Layout: per image, transpose a [C][HW] fp32 matrix <-> [HW][Cpad] T matrix through a 32x33 tile.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int C, int Cpad,
                                    long long HW) {
  __shared__ float tile[32][33];
  const long long b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const float* s = src + b * C * HW;
  T* d = dst + b * HW * Cpad;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i;
    long long p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? s[(long long)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    long long p = p0 + i;
    int c = c0 + threadIdx.x;
    if (p < HW && c < Cpad) d[p * Cpad + c] = from_f<T>(tile[threadIdx.x][i]);
  }
}

// Small channel counts (images: C=3, logit gradients: C<=32): one thread per pixel reads C planes
// (coalesced across the warp) and writes its Cpad contiguous channels with 16-byte stores.
template <typename T>
__global__ void __launch_bounds__(kThreads)
nchw_to_nhwc_smallc_kernel(const float* __restrict__ src, T* __restrict__ dst, int B, int C, int Cpad, long long HW) {
  const long long total = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    const float* s = src + b * C * HW + p;
    float v[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = (c < C) ? __ldg(s + (long long)c * HW) : 0.f;
    T* d = dst + i * Cpad;
    if (Cpad % 8 == 0) {
#pragma unroll
      for (int c = 0; c < 32; c += 8) {
        if (c < Cpad) {
          float o[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = v[c + k];
          st_vec<8>(d + c, o);
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (c < Cpad) d[c] = from_f<T>(v[c]);
    }
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int C, int Cpad,
                                    long long HW) {
  __shared__ float tile[32][33];
  const long long b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const T* s = src + b * HW * Cpad;
  float* d = dst + b * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    long long p = p0 + i;
    int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < HW && c < C) ? to_f(s[p * Cpad + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i;
    long long p = p0 + threadIdx.x;
    if (c < C && p < HW) d[(long long)c * HW + p] = tile[threadIdx.x][i];
  }
}

template <typename T>
__global__ void cast_f32_kernel(const float* __restrict__ src, T* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = from_f<T>(src[i]);
}

// ---------------------------------------------------------------------------------------------
// BatchNorm.  x is [M][C] (M = B*H*W rows, C innermost).  VEC channels per thread.
// ---------------------------------------------------------------------------------------------
// sums[c] += sum_rows x, sums[C+c] += sum_rows x^2   (double accumulators, zeroed by the host)
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
bn_stats_kernel(const T* __restrict__ x, double* sums, long long M, int Ctot, int C, const BnFwdFinal fin) {
  // blockIdx.y selects a slab of C channels out of Ctot (C <= blockDim.x * VEC).  Each thread owns one
  // channel vector and walks rows with 4 independent 16-byte loads in flight.
  extern __shared__ float sh[];  // [groups][C][2]
  const int c_off = blockIdx.y * C;
  x += c_off;
  const int cv = C / VEC;                // channel vectors per row
  const int groups = blockDim.x / cv;    // row groups per block (host guarantees >= 1)
  const int g = threadIdx.x / cv, v = threadIdx.x % cv;
  float s1[VEC], s2[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  if (g < groups) {
    const long long stride = (long long)gridDim.x * groups;
    const T* px = x + v * VEC;
    for (long long r = (long long)blockIdx.x * groups + g; r < M; r += 8 * stride) {
      float xv[8][VEC];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long rr = r + u * stride;
        if (rr < M) {
          ld_vec<VEC>(px + rr * Ctot, xv[u]);
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j) xv[u][j] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
#pragma unroll
        for (int j = 0; j < VEC; ++j) { s1[j] += xv[u][j]; s2[j] += xv[u][j] * xv[u][j]; }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      sh[(g * C + v * VEC + j) * 2 + 0] = s1[j];
      sh[(g * C + v * VEC + j) * 2 + 1] = s2[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    int c = i >> 1, k = i & 1;
    double a = 0.0;
    for (int gg = 0; gg < groups; ++gg) a += (double)sh[(gg * C + c) * 2 + k];
    atomicAdd(sums + k * Ctot + c_off + c, a);
  }
  if (last_block_done(fin.counter)) {   // fused finalize; leaves the workspace zeroed for the next call
    for (int c = threadIdx.x; c < Ctot; c += blockDim.x) {
      const double s1 = __ldcg(sums + c), s2 = __ldcg(sums + Ctot + c);
      bn_fwd_finalize_channel(fin, s1, s2, c);
      sums[c] = 0.0; sums[Ctot + c] = 0.0;
    }
    if (threadIdx.x == 0) *fin.counter = 0u;
  }
}

// eval mode: scale/shift from running statistics
__global__ void bn_eval_coeffs_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                      float* __restrict__ scale_out, float* __restrict__ shift_out, int C, float eps) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float rstd = rsqrtf(running_var[c] + eps);
  float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale_out[c] = g * rstd;
  shift_out[c] = b - running_mean[c] * g * rstd;
}

__device__ __forceinline__ float act_fwd(float v, float slope) { return v > 0.f ? v : v * slope; }

// y = act(x*scale[c] + shift[c] (+ residual)),  act: slope=1 -> identity, 0 -> ReLU, 0.2 -> LeakyReLU
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads, 2)
bn_apply_kernel(const T* __restrict__ x, const T* __restrict__ residual, T* __restrict__ y,
                const float* __restrict__ scale, const float* __restrict__ shift, long long n, int C, float slope,
                const double* __restrict__ sums, const BnFwdFinal fin) {
  // per-channel coefficients live in shared memory; a CTA streams 4 x 256 consecutive 16-byte vectors per
  // iteration (4 independent loads in flight per thread, DRAM-page friendly).
  // sums != null: the batch statistics come straight from the producing convolution's epilogue (sum and sum
  // of squares per channel); every CTA derives scale/shift from them, CTA 0 also publishes mean / rstd for
  // the backward pass and updates the running statistics (the separate finalize launch disappears).
  extern __shared__ float sp[];   // [2][C]
  if (sums) {
    const double inv_m = 1.0 / (double)fin.M;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const double mean = sums[c] * inv_m;
      const float var = fmaxf((float)(sums[C + c] * inv_m - mean * mean), 0.f);
      const float rstd = rsqrtf(var + fin.eps);
      const float g = fin.gamma ? fin.gamma[c] : 1.f, b = fin.beta ? fin.beta[c] : 0.f;
      sp[c] = g * rstd;
      sp[C + c] = b - (float)mean * g * rstd;
      if (blockIdx.x == 0) bn_fwd_finalize_channel(fin, sums[c], sums[C + c], c);
    }
  } else {
    for (int c = threadIdx.x; c < C; c += blockDim.x) { sp[c] = scale[c]; sp[C + c] = shift[c]; }
  }
  __syncthreads();
  const long long nv = n / VEC;
  const bool pow2 = (C & (C - 1)) == 0;
  const long long chunk = 4LL * blockDim.x;
  // (VEC * blockDim) % C == 0: a thread always meets the same channels -> coefficients live in registers
  const bool fixed_c = ((VEC * (int)blockDim.x) % C) == 0;
  float sc[VEC], sf[VEC];
  if (fixed_c) {
    const int c = (threadIdx.x * VEC) % C;
#pragma unroll
    for (int j = 0; j < VEC; ++j) { sc[j] = sp[c + j]; sf[j] = sp[C + c + j]; }
  }
  for (long long base = (long long)blockIdx.x * chunk; base < nv; base += (long long)gridDim.x * chunk) {
    float xv[4][VEC], rv[4][VEC];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long ii = base + u * blockDim.x + threadIdx.x;
      if (ii < nv) {
        ld_vec<VEC>(x + ii * VEC, xv[u]);
        if (residual) ld_vec<VEC>(residual + ii * VEC, rv[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long ii = base + u * blockDim.x + threadIdx.x;
      if (ii < nv) {
        if (!fixed_c) {
          const int c = pow2 ? ((int)ii * VEC) & (C - 1) : (int)((ii * VEC) % C);
#pragma unroll
          for (int j = 0; j < VEC; ++j) { sc[j] = sp[c + j]; sf[j] = sp[C + c + j]; }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float v = xv[u][j] * sc[j] + sf[j];
          if (residual) v += rv[u][j];
          xv[u][j] = act_fwd(v, slope);
        }
        st_vec<VEC>(y + ii * VEC, xv[u]);
      }
    }
  }
}

// Backward reduce: g = dy * act'(a);  sums[c] += sum g,  sums[C+c] += sum g * xhat,
// xhat = (x - mean) * rstd.  `a` is the saved post-activation output (sign decides act').
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads, 2)
bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ a,
                     const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     double* sums, long long M, int Ctot, int C, float slope, const BnBwdFinal fin) {
  // activation mask: from the saved output `a` when given, else (scale/shift given) recomputed from the
  // pre-activation z*scale+shift — saves reading `a` for every non-residual layer
  extern __shared__ float sh[];
  const int c_off = blockIdx.y * C;
  dy += c_off; x += c_off; if (a) a += c_off;
  mean += c_off; rstd += c_off;
  const bool zmask = (a == nullptr) && (scale != nullptr);
  const int cv = C / VEC;
  const int groups = blockDim.x / cv;
  const int g = threadIdx.x / cv, v = threadIdx.x % cv;
  float s1[VEC], s2[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
  if (g < groups) {
    float mu[VEC], rs[VEC], sc[VEC], sf[VEC];
    ld_vec<VEC>(mean + v * VEC, mu);
    ld_vec<VEC>(rstd + v * VEC, rs);
    if (zmask) { ld_vec<VEC>(scale + c_off + v * VEC, sc); ld_vec<VEC>(shift + c_off + v * VEC, sf); }
    const long long stride = (long long)gridDim.x * groups;
    const long long co = v * VEC;
    for (long long r = (long long)blockIdx.x * groups + g; r < M; r += 2 * stride) {
      float dv[2][VEC], xv[2][VEC], av[2][VEC];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const long long rr = r + u * stride;
        if (rr < M) {
          ld_vec<VEC>(dy + rr * Ctot + co, dv[u]);
          ld_vec<VEC>(x + rr * Ctot + co, xv[u]);
          if (a) ld_vec<VEC>(a + rr * Ctot + co, av[u]);
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j) { dv[u][j] = 0.f; xv[u][j] = 0.f; av[u][j] = 1.f; }
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float gg = dv[u][j];
          if (a) gg *= (av[u][j] > 0.f) ? 1.f : slope;
          else if (zmask) gg *= (xv[u][j] * sc[j] + sf[j] > 0.f) ? 1.f : slope;
          s1[j] += gg;
          s2[j] += gg * (xv[u][j] - mu[j]) * rs[j];
        }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      sh[(g * C + v * VEC + j) * 2 + 0] = s1[j];
      sh[(g * C + v * VEC + j) * 2 + 1] = s2[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    int c = i >> 1, k = i & 1;
    double acc = 0.0;
    for (int gg = 0; gg < groups; ++gg) acc += (double)sh[(gg * C + c) * 2 + k];
    atomicAdd(sums + k * Ctot + c_off + c, acc);
  }
  if (last_block_done(fin.counter)) {
    for (int c = threadIdx.x; c < Ctot; c += blockDim.x) {
      const double s1 = __ldcg(sums + c), s2 = __ldcg(sums + Ctot + c);
      bn_bwd_finalize_channel(fin, s1, s2, c, Ctot);
      sums[c] = 0.0; sums[Ctot + c] = 0.0;
    }
    if (threadIdx.x == 0) *fin.counter = 0u;
  }
}

// dx = k0*g - k1 - k2*xhat ;  optionally d_residual (+)= g  (identity branch of a residual block)
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads, 2)
bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ a,
                    const float* __restrict__ coef, const float* __restrict__ scale,
                    const float* __restrict__ shift, T* __restrict__ dx, T* dres, int dres_accumulate,
                    long long n, int C, float slope) {
  // dx = A*g + Bc*x + Cc with g = dy * act'(.) ; per-channel A, Bc, Cc from the reduce pass' last CTA.
  // A thread meets the same channels in every iteration ((VEC*blockDim) % C == 0): coefficients in registers.
  extern __shared__ float sp[];   // [5][C]: A, Bc, Cc, scale, shift
  const bool zmask = (a == nullptr) && (scale != nullptr);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    sp[c] = coef[c]; sp[C + c] = coef[C + c]; sp[2 * C + c] = coef[2 * C + c];
    sp[3 * C + c] = zmask ? scale[c] : 0.f; sp[4 * C + c] = zmask ? shift[c] : 0.f;
  }
  __syncthreads();
  const long long nv = n / VEC;
  const bool pow2 = (C & (C - 1)) == 0;
  const bool racc = dres && dres_accumulate;
  const long long chunk = 2LL * blockDim.x;
  const bool fixed_c = ((VEC * (int)blockDim.x) % C) == 0;
  float kA[VEC], kB[VEC], kC[VEC], sc[VEC], sf[VEC];
  if (fixed_c) {
    const int c = (threadIdx.x * VEC) % C;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      kA[j] = sp[c + j]; kB[j] = sp[C + c + j]; kC[j] = sp[2 * C + c + j]; sc[j] = sp[3 * C + c + j]; sf[j] = sp[4 * C + c + j];
    }
  }
  for (long long base = (long long)blockIdx.x * chunk; base < nv; base += (long long)gridDim.x * chunk) {
    float dv[2][VEC], xv[2][VEC], av[2][VEC], rv[2][VEC];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long ii = base + u * blockDim.x + threadIdx.x;
      if (ii < nv) {
        ld_vec<VEC>(dy + ii * VEC, dv[u]);
        ld_vec<VEC>(x + ii * VEC, xv[u]);
        if (a) ld_vec<VEC>(a + ii * VEC, av[u]);
        if (racc) ld_vec<VEC>(dres + ii * VEC, rv[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long ii = base + u * blockDim.x + threadIdx.x;
      if (ii < nv) {
        if (!fixed_c) {
          const int c = pow2 ? ((int)ii * VEC) & (C - 1) : (int)((ii * VEC) % C);
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            kA[j] = sp[c + j]; kB[j] = sp[C + c + j]; kC[j] = sp[2 * C + c + j]; sc[j] = sp[3 * C + c + j]; sf[j] = sp[4 * C + c + j];
          }
        }
        float ov[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          float g = dv[u][j];
          if (a) g *= (av[u][j] > 0.f) ? 1.f : slope;
          else if (zmask) g *= (xv[u][j] * sc[j] + sf[j] > 0.f) ? 1.f : slope;
          ov[j] = kA[j] * g + kB[j] * xv[u][j] + kC[j];
          dv[u][j] = racc ? rv[u][j] + g : g;
        }
        st_vec<VEC>(dx + ii * VEC, ov);
        if (dres) st_vec<VEC>(dres + ii * VEC, dv[u]);
      }
    }
  }
}

// Plain activation backward (no BN): dx = dy * act'(a)     (discriminator layer 1: conv + LeakyReLU)
template <typename T, int VEC>
__global__ void act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ a, T* __restrict__ dx,
                               long long n, float slope) {
  const long long nv = n / VEC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    float dv[VEC], av[VEC];
    ld_vec<VEC>(dy + i * VEC, dv);
    ld_vec<VEC>(a + i * VEC, av);
#pragma unroll
    for (int j = 0; j < VEC; ++j) dv[j] *= (av[j] > 0.f) ? 1.f : slope;
    st_vec<VEC>(dx + i * VEC, dv);
  }
}

// y = act(x + bias[c])  (in place allowed)
template <typename T, int VEC>
__global__ void bias_act_kernel(const T* __restrict__ x, const float* __restrict__ bias, T* __restrict__ y,
                                long long n, int C, float slope) {
  const long long nv = n / VEC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((i * VEC) % C);
    float xv[VEC], bv[VEC];
    ld_vec<VEC>(x + i * VEC, xv);
    if (bias) ld_vec<VEC>(bias + c, bv);
#pragma unroll
    for (int j = 0; j < VEC; ++j) xv[j] = act_fwd(xv[j] + (bias ? bv[j] : 0.f), slope);
    st_vec<VEC>(y + i * VEC, xv);
  }
}

// out[c] (+)= sum_rows x[r][c]      (bias gradients; global-average-pool backward helper)
template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads)
colsum_kernel(const T* __restrict__ x, double* __restrict__ sums, long long M, int Ctot, int C) {
  extern __shared__ float sh[];
  const int c_off = blockIdx.y * C;
  x += c_off;
  const int cv = C / VEC;
  const int groups = blockDim.x / cv;
  const int g = threadIdx.x / cv, v = threadIdx.x % cv;
  float s1[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) s1[j] = 0.f;
  if (g < groups) {
    for (long long r = (long long)blockIdx.x * groups + g; r < M; r += (long long)gridDim.x * groups) {
      float xv[VEC];
      ld_vec<VEC>(x + r * Ctot + v * VEC, xv);
#pragma unroll
      for (int j = 0; j < VEC; ++j) s1[j] += xv[j];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) sh[g * C + v * VEC + j] = s1[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double acc = 0.0;
    for (int gg = 0; gg < groups; ++gg) acc += (double)sh[gg * C + c];
    atomicAdd(sums + c_off + c, acc);
  }
}
__global__ void sums_to_f32_kernel(const double* __restrict__ sums, float* __restrict__ out, int n, float scale,
                                   int accumulate) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (accumulate ? out[i] : 0.f) + (float)(sums[i] * (double)scale);
}

// ---------------------------------------------------------------------------------------------
// MaxPool 3x3 stride 2 pad 1 (NHWC).  idx stores the winning tap (kh*3+kw), first max in scan order
// and NaN-propagating like torch.
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, typename IT>
__global__ void __launch_bounds__(kThreads)
maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, unsigned char* __restrict__ idx, int B, int H,
                   int W, int C, int Ho, int Wo) {
  const int cv = C / VEC;
  const IT total = (IT)B * Ho * Wo * cv;
  for (IT i = (IT)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (IT)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    IT p = i / cv;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho);
    const int b = (int)(p / Ho);
    float best[VEC];
    int bi[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { best[j] = -INFINITY; bi[j] = -1; }
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int h = ho * 2 - 1 + kh;
      if (h < 0 || h >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int w = wo * 2 - 1 + kw;
        if (w < 0 || w >= W) continue;
        float xv[VEC];
        ld_vec<VEC>(x + (((long long)b * H + h) * W + w) * C + v * VEC, xv);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          // at::max_pool2d: index starts at the first valid tap; (val > max) || isnan(val) replaces
          if (bi[j] < 0) bi[j] = kh * 3 + kw;
          if ((xv[j] > best[j]) || (xv[j] != xv[j])) { best[j] = xv[j]; bi[j] = kh * 3 + kw; }
        }
      }
    }
    const long long o = (((long long)b * Ho + ho) * Wo + wo) * C + v * VEC;
    st_vec<VEC>(y + o, best);
    if (idx) {
      if constexpr (VEC == 8) {
        uint2 iv;
        iv.x = (unsigned)bi[0] | ((unsigned)bi[1] << 8) | ((unsigned)bi[2] << 16) | ((unsigned)bi[3] << 24);
        iv.y = (unsigned)bi[4] | ((unsigned)bi[5] << 8) | ((unsigned)bi[6] << 16) | ((unsigned)bi[7] << 24);
        *reinterpret_cast<uint2*>(idx + o) = iv;
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) idx[o + j] = (unsigned char)bi[j];
      }
    }
  }
}

template <typename T, int VEC, typename IT>
__global__ void __launch_bounds__(kThreads)
maxpool_bwd_kernel(const T* __restrict__ dy, const unsigned char* __restrict__ idx, const T* addend, T* dx, int B,
                   int H, int W, int C, int Ho, int Wo) {
  const int cv = C / VEC;
  const IT total = (IT)B * H * W * cv;
  for (IT i = (IT)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (IT)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    IT p = i / cv;
    const int w = (int)(p % W); p /= W;
    const int h = (int)(p % H);
    const int b = (int)(p / H);
    float acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
    if (addend) ld_vec<VEC>(addend + (long long)i * VEC, acc);  // may alias dx (same elements, same thread)
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int t = h + 1 - kh;
      if (t < 0 || (t & 1)) continue;
      const int ho = t >> 1;
      if (ho >= Ho) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int s = w + 1 - kw;
        if (s < 0 || (s & 1)) continue;
        const int wo = s >> 1;
        if (wo >= Wo) continue;
        const long long o = (((long long)b * Ho + ho) * Wo + wo) * C + v * VEC;
        float dv[VEC];
        ld_vec<VEC>(dy + o, dv);
        if constexpr (VEC == 8) {   // the eight winning-tap bytes in one 8-byte load
          const uint2 iv = *reinterpret_cast<const uint2*>(idx + o);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const unsigned int b = ((j < 4 ? iv.x : iv.y) >> (8 * (j & 3))) & 0xffu;
            if (b == (unsigned int)(kh * 3 + kw)) acc[j] += dv[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j)
            if (idx[o + j] == kh * 3 + kw) acc[j] += dv[j];
        }
      }
    }
    st_vec<VEC>(dx + (long long)i * VEC, acc);
  }
}

// Even H, W: one thread per 2x2 INPUT block (and 8 channels).  The four pixels of a block can only have won in the four
// windows (i,j), (i,j+1), (i+1,j), (i+1,j+1) — the even/even pixel only as the centre tap of window (i,j) — so every
// thread does the same work (no parity divergence), with four dy / winner-tap loads and four dx stores in flight
// (the per-pixel gather above spends its time in divergent parity tests).
template <typename T, typename IT>
__global__ void __launch_bounds__(kThreads)
maxpool_bwd_block_kernel(const T* __restrict__ dy, const unsigned char* __restrict__ idx, const T* addend, T* dx, int B,
                         int H, int W, int C, int Ho, int Wo) {
  constexpr int VEC = 8;
  const int cv = C / VEC;
  const IT total = (IT)B * Ho * Wo * cv;
  for (IT i = (IT)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (IT)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    IT p = i / cv;
    const int wo = (int)(p % Wo); p /= Wo;
    const int ho = (int)(p % Ho);
    const int b = (int)(p / Ho);
    float g[4][VEC];
    uint2 win[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int oh = ho + (q >> 1), ow = wo + (q & 1);
      if (oh < Ho && ow < Wo) {
        const long long o = (((long long)b * Ho + oh) * Wo + ow) * C + v * VEC;
        ld_vec<VEC>(dy + o, g[q]);
        win[q] = *reinterpret_cast<const uint2*>(idx + o);
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) g[q][j] = 0.f;
        win[q] = make_uint2(0xffffffffu, 0xffffffffu);
      }
    }
    // window q = (i + qh, j + qw) covers input rows 2(i+qh)-1 .. 2(i+qh)+1: block row dh is its tap kh = dh + 1 - 2 qh
#pragma unroll
    for (int dh = 0; dh < 2; ++dh) {
#pragma unroll
      for (int dw = 0; dw < 2; ++dw) {
        const long long o = (((long long)b * H + 2 * ho + dh) * W + 2 * wo + dw) * C + v * VEC;
        float acc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
        if (addend) ld_vec<VEC>(addend + o, acc);     // may alias dx (same elements, same thread)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int kh = dh + 1 - 2 * (q >> 1), kw = dw + 1 - 2 * (q & 1);
          if (kh < 0 || kw < 0) continue;             // compile-time: this window does not reach the pixel
          const unsigned int tap = (unsigned int)(kh * 3 + kw);
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            const unsigned int wj = ((j < 4 ? win[q].x : win[q].y) >> (8 * (j & 3))) & 0xffu;
            if (wj == tap) acc[j] += g[q][j];
          }
        }
        st_vec<VEC>(dx + o, acc);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Decoder operand: out[b,h,w,:] = concat(x[b,h/2,w/2,:C1], skip[b,h,w,:C2])  (nearest x2)
// ---------------------------------------------------------------------------------------------
template <typename T, int VEC, typename IT>
__global__ void __launch_bounds__(kThreads)
upcat_fwd_kernel(const T* __restrict__ x, const T* __restrict__ skip, T* __restrict__ out, int B, int H, int W,
                 int C1, int C2) {
  // Work items: one low-resolution vector of x (-> its 2x2 children: one load, four stores), then groups of four
  // skip vectors of one pixel row segment (four independent loads in flight before the stores).  The first
  // version moved one vector per loop trip with a dependent load -> store and sat at ~50 % of the HBM roofline.
  const int Ct = C1 + C2, cv1 = C1 / VEC, cv2 = C2 / VEC;
  const int H2 = H / 2, W2 = W / 2;
  const IT n1 = (IT)B * H2 * W2 * cv1;
  const IT n2 = ((IT)B * H * W * cv2 + 3) / 4;
  const IT nskip = (IT)B * H * W * cv2;
  for (IT i = (IT)blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2; i += (IT)gridDim.x * blockDim.x) {
    if (i < n1) {
      const int c = (int)(i % cv1) * VEC;
      IT p = i / cv1;
      const int w2 = (int)(p % W2); p /= W2;
      const int h2 = (int)(p % H2);
      const int b = (int)(p / H2);
      float v[VEC];
      ld_vec<VEC>(x + (long long)i * VEC, v);
      T* o = out + (((long long)b * H + 2 * h2) * W + 2 * w2) * Ct + c;
      st_vec<VEC>(o, v);
      st_vec<VEC>(o + Ct, v);
      st_vec<VEC>(o + (long long)W * Ct, v);
      st_vec<VEC>(o + (long long)W * Ct + Ct, v);
    } else {
      const IT k0 = (i - n1) * 4;
      float v[4][VEC];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (k0 + u < nskip) ld_vec<VEC>(skip + (long long)(k0 + u) * VEC, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const IT k = k0 + u;
        if (k < nskip) {
          const int c = (int)(k % cv2) * VEC;
          const IT p = k / cv2;
          st_vec<VEC>(out + (long long)p * Ct + C1 + c, v[u]);
        }
      }
    }
  }
}
// dx[b,h2,w2,c] = sum of the 2x2 children of dout[..., c<C1];  dskip = dout[..., C1:]
template <typename T, int VEC, typename IT>
__global__ void __launch_bounds__(kThreads)
upcat_bwd_kernel(const T* __restrict__ dout, T* __restrict__ dx, T* __restrict__ dskip, int B, int H, int W,
                 int C1, int C2) {
  const int Ct = C1 + C2;
  const int cv1 = C1 / VEC, cv2 = C2 / VEC;
  const IT n1 = (IT)B * (H / 2) * (W / 2) * cv1;
  const IT n2 = (IT)B * H * W * cv2;
  for (IT i = (IT)blockIdx.x * blockDim.x + threadIdx.x; i < n1 + n2;
       i += (IT)gridDim.x * blockDim.x) {
    if (i < n1) {
      const int c = (int)(i % cv1) * VEC;
      IT p = i / cv1;
      const int w2 = (int)(p % (W / 2)); p /= (W / 2);
      const int h2 = (int)(p % (H / 2));
      const int b = (int)(p / (H / 2));
      float acc[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
#pragma unroll
      for (int dh = 0; dh < 2; ++dh)
#pragma unroll
        for (int dw = 0; dw < 2; ++dw) {
          float v[VEC];
          ld_vec<VEC>(dout + (((long long)b * H + 2 * h2 + dh) * W + 2 * w2 + dw) * Ct + c, v);
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[j] += v[j];
        }
      st_vec<VEC>(dx + (long long)i * VEC, acc);
    } else {
      const IT k = i - n1;
      const int c = (int)(k % cv2) * VEC;
      const IT p = k / cv2;
      float v[VEC];
      ld_vec<VEC>(dout + (long long)p * Ct + C1 + c, v);
      st_vec<VEC>(dskip + (long long)k * VEC, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Discriminator tail: global average pool + Linear(C,1) + sigmoid, and its backward.
// ---------------------------------------------------------------------------------------------
// One CTA per image: out[b] = sigmoid(bias + sum_c w[c] * mean_hw x[b,hw,c]);  pooled[b][c] saved.
template <typename T>
__global__ void gap_linear_sigmoid_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                          const float* __restrict__ bias, float* __restrict__ pooled,
                                          float* __restrict__ out, long long HW, int C) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const T* xb = x + (long long)b * HW * C;
  float dot = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (long long p = 0; p < HW; ++p) s += to_f(xb[p * C + c]);
    s /= (float)HW;
    pooled[(long long)b * C + c] = s;
    dot += s * w[c];
  }
  float v[1] = {dot}, o[1];
  block_sum<1>(v, red, o);
  if (threadIdx.x == 0) out[b] = 1.f / (1.f + expf(-(o[0] + bias[0])));
}
// dz = dout*y*(1-y); dw[c] += sum_b dz_b pooled[b][c]; dbias += sum dz; dx[b,hw,c] = dz_b*w[c]/HW
template <typename T>
__global__ void gap_linear_sigmoid_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ y,
                                              const float* __restrict__ pooled, const float* __restrict__ w,
                                              float* __restrict__ dw, float* __restrict__ dbias,
                                              T* __restrict__ dx, int B, long long HW, int C, int accumulate) {
  // grid.x = B (dx), plus block 0 also produces dw/dbias
  const int b = blockIdx.x;
  const float dz = dout[b] * y[b] * (1.f - y[b]);
  T* dxb = dx + (long long)b * HW * C;
  const float k = dz / (float)HW;
  for (long long i = threadIdx.x; i < HW * C; i += blockDim.x) dxb[i] = from_f<T>(k * w[i % C]);
  if (b == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float s = 0.f;
      for (int bb = 0; bb < B; ++bb) s += dout[bb] * y[bb] * (1.f - y[bb]) * pooled[(long long)bb * C + c];
      dw[c] = (accumulate ? dw[c] : 0.f) + s;
    }
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int bb = 0; bb < B; ++bb) s += dout[bb] * y[bb] * (1.f - y[bb]);
      dbias[0] = (accumulate ? dbias[0] : 0.f) + s;
    }
  }
}

// Wide versions of the two kernels above for bf16 NHWC features with C % 8 == 0 and 256 % (C/8) == 0 (the
// discriminator tail: [B,32,32,512] at 512x512 inputs).  The one-CTA-per-image kernels walk HW pixels serially per
// channel (1.1 ms + 0.8 ms per adversarial step at B=8, ncu launch list); here the pooling is spread over
// (image, pixel slice) CTAs with 16-byte loads, the last slice CTA of an image finishes Linear + sigmoid, and the
// backward is a plain coalesced fill.
__global__ void __launch_bounds__(256)
gap_partial_kernel(const bf16* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ pooled, float* __restrict__ out, unsigned int* __restrict__ counters,
                   int HW, int C) {
  __shared__ float part[256 * 8];
  __shared__ bool is_last;
  __shared__ float red[32];
  const int b = blockIdx.y, cvn = C / 8, rows = 256 / cvn;
  const int cv = threadIdx.x % cvn, row = threadIdx.x / cvn;
  const int per = (HW + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(HW, p0 + per);
  const bf16* xb = x + (long long)b * HW * C + cv * 8;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int p = p0 + row; p < p1; p += rows) {
    float v[8];
    ld_vec<8>(xb + (long long)p * C, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) part[threadIdx.x * 8 + j] = acc[j];
  __syncthreads();
  const float inv = 1.f / (float)HW;
  for (int c = threadIdx.x; c < C; c += 256) {
    float sum = 0.f;
    for (int r = 0; r < rows; ++r) sum += part[(r * cvn + c / 8) * 8 + (c & 7)];
    atomicAdd(pooled + (long long)b * C + c, sum * inv);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counters + b, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float dot = 0.f;
  for (int c = threadIdx.x; c < C; c += 256) dot += __ldcg(pooled + (long long)b * C + c) * w[c];
  float v[1] = {dot}, o[1];
  block_sum<1>(v, red, o);
  if (threadIdx.x == 0) {
    out[b] = 1.f / (1.f + expf(-(o[0] + bias[0])));
    counters[b] = 0u;
  }
}

__global__ void __launch_bounds__(256)
gap_bwd_fill_kernel(const float* __restrict__ dout, const float* __restrict__ y, const float* __restrict__ pooled,
                    const float* __restrict__ w, float* __restrict__ dw, float* __restrict__ dbias,
                    bf16* __restrict__ dx, int B, int HW, int C, int accumulate) {
  const int b = blockIdx.y, cvn = C / 8;
  const float dz = dout[b] * y[b] * (1.f - y[b]);
  const float k = dz / (float)HW;
  const int cv = threadIdx.x % cvn;   // 256 % cvn == 0: the channel vector of a thread is fixed
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = k * w[cv * 8 + j];
  bf16* dxb = dx + (long long)b * HW * C;
  const int nvec = HW * cvn;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < nvec; i += gridDim.x * 256) st_vec<8>(dxb + (long long)i * 8, v);
  if (b == 0 && blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += 256) {
      float sacc = 0.f;
      for (int bb = 0; bb < B; ++bb) sacc += dout[bb] * y[bb] * (1.f - y[bb]) * pooled[(long long)bb * C + c];
      dw[c] = (accumulate ? dw[c] : 0.f) + sacc;
    }
    if (threadIdx.x == 0) {
      float sacc = 0.f;
      for (int bb = 0; bb < B; ++bb) sacc += dout[bb] * y[bb] * (1.f - y[bb]);
      dbias[0] = (accumulate ? dbias[0] : 0.f) + sacc;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused Adam over a flat fp32 parameter buffer (torch.optim.Adam defaults semantics, no amsgrad);
// also refreshes the bf16 shadow copy the tensor-core convolutions read.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            bf16* __restrict__ shadow, long long n, float lr, float beta1, float beta2, float eps,
            float weight_decay, float bc1, float bc2_sqrt, float grad_scale, const float* __restrict__ dev_clip,
            const int* __restrict__ dev_step) {
  const float clip = dev_clip ? *dev_clip : 1.f;
  if (dev_step) {   // graph-capturable mode: the step count lives on the device
    const double t = (double)*dev_step;
    bc1 = (float)(1.0 - pow((double)beta1, t));
    bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, t));
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale * clip;
    float pi = p[i];
    if (weight_decay != 0.f) gi += weight_decay * pi;
    float mi = beta1 * m[i] + (1.f - beta1) * gi;
    float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi; v[i] = vi;
    // torch: denom = sqrt(v)/sqrt(bc2) + eps ; p -= (lr/bc1) * m / denom
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= (lr / bc1) * mi / denom;
    p[i] = pi;
    if (shadow) shadow[i] = __float2bfloat16_rn(pi);
  }
}

__global__ void step_increment_kernel(int* step) { *step += 1; }

// sum of squares of a flat fp32 buffer -> out[0] (double); clip coefficient kernel for clip_grad_norm_
__global__ void __launch_bounds__(kThreads) sumsq_kernel(const float* __restrict__ x, double* __restrict__ out, long long n) {
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) { float t = x[i]; s += t * t; }
  float v[1] = {s}, o[1];
  block_sum<1>(v, red, o);
  if (threadIdx.x == 0) atomicAdd(out, (double)o[0]);
}
// clip_grad_norm_ (src/models/unsupervised_trainer.py:144): coef = min(1, max_norm/(norm+1e-6))
__global__ void clip_coef_kernel(const double* __restrict__ sumsq, float* __restrict__ coef, float* __restrict__ norm_out,
                                 float max_norm, float pre_scale) {
  if (threadIdx.x == 0) {
    float norm = sqrtf((float)*sumsq) * pre_scale;
    float c = max_norm / (norm + 1e-6f);
    coef[0] = c < 1.f ? c : 1.f;
    if (norm_out) norm_out[0] = norm;
  }
}

}  // namespace
}  // namespace uda

// =================================================================================================
// C ABI
// =================================================================================================
using namespace uda;

#define UDA_DT(dtype, FN, ...)                                                   \
  do {                                                                           \
    if ((dtype) == UDA_BF16) { FN(bf16, __VA_ARGS__); }                          \
    else if ((dtype) == UDA_F32) { FN(float, __VA_ARGS__); }                     \
    else return set_error(UDA_ERR_BAD_ARG, "unsupported dtype %d", (dtype));     \
  } while (0)

static inline int vec_for(int dtype, int C, const void* a, const void* b = nullptr, const void* c = nullptr,
                          const void* d = nullptr) {
  // widest channel vector (16 bytes) every pointer and C allow
  int v = dtype == UDA_BF16 ? 8 : 4;
  const size_t es = dtype == UDA_BF16 ? 2 : 4;
  auto ok = [&](int vv) {
    if (C % vv) return false;
    const void* ps[4] = {a, b, c, d};
    for (auto p : ps)
      if (p && reinterpret_cast<uintptr_t>(p) % (vv * es)) return false;
    return true;
  };
  while (v > 1 && !ok(v)) v >>= 1;
  return v;
}

extern "C" int uda_nchw_f32_to_nhwc(const float* src, void* dst, int dtype, int B, int C, int Cpad,
                                    long long HW, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(src && dst && B > 0 && C > 0 && Cpad >= C && HW > 0, UDA_ERR_BAD_ARG, "nchw_to_nhwc: bad argument");
  if (Cpad <= 32) {
    const bool v8 = reinterpret_cast<uintptr_t>(dst) % 16 == 0;
    UDA_REQUIRE(v8 || Cpad % 8, UDA_ERR_BAD_ARG, "nchw_to_nhwc: destination must be 16-byte aligned");
#define K(T, ...) nchw_to_nhwc_smallc_kernel<T><<<grid_for((long long)B * HW), kThreads, 0, st>>>(src, (T*)dst, B, C, Cpad, HW)
    UDA_DT(dtype, K, 0);
#undef K
    UDA_LAUNCH_OK("nchw_to_nhwc_smallc_kernel");
    return UDA_OK;
  }
  dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((Cpad + 31) / 32), (unsigned)B), block(32, 8);
  UDA_REQUIRE(B <= 65535 && grid.y <= 65535, UDA_ERR_UNSUPPORTED, "nchw_to_nhwc: shape too large");
#define K(T, ...) nchw_to_nhwc_kernel<T><<<grid, block, 0, st>>>(src, (T*)dst, C, Cpad, HW)
  UDA_DT(dtype, K, 0);
#undef K
  UDA_LAUNCH_OK("nchw_to_nhwc_kernel");
  return UDA_OK;
}

extern "C" int uda_nhwc_to_nchw_f32(const void* src, int dtype, float* dst, int B, int C, int Cpad,
                                    long long HW, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(src && dst && B > 0 && C > 0 && Cpad >= C && HW > 0, UDA_ERR_BAD_ARG, "nhwc_to_nchw: bad argument");
  dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)B), block(32, 8);
  UDA_REQUIRE(B <= 65535 && grid.y <= 65535, UDA_ERR_UNSUPPORTED, "nhwc_to_nchw: shape too large");
#define K(T, ...) nhwc_to_nchw_kernel<T><<<grid, block, 0, st>>>((const T*)src, dst, C, Cpad, HW)
  UDA_DT(dtype, K, 0);
#undef K
  UDA_LAUNCH_OK("nhwc_to_nchw_kernel");
  return UDA_OK;
}

extern "C" int uda_cast_f32(const float* src, void* dst, int dtype, long long n, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(src && dst && n >= 0, UDA_ERR_BAD_ARG, "cast: bad argument");
  if (n == 0) return UDA_OK;
#define K(T, ...) cast_f32_kernel<T><<<grid_for(n), kThreads, 0, st>>>(src, (T*)dst, n)
  UDA_DT(dtype, K, 0);
#undef K
  UDA_LAUNCH_OK("cast_f32_kernel");
  return UDA_OK;
}

// workspace: 2*C doubles
extern "C" int uda_bn_stats(const void* x, int dtype, long long M, int C, const float* gamma, const float* beta,
                            float* running_mean, float* running_var, float* mean, float* rstd, float* scale,
                            float* shift, float eps, float momentum, void* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(x && mean && rstd && scale && shift && workspace, UDA_ERR_BAD_ARG, "bn_stats: null pointer");
  UDA_REQUIRE(M > 0 && C > 0 && C <= 4096, UDA_ERR_BAD_ARG, "bn_stats: bad shape M=%lld C=%d", M, C);
  // workspace layout (fixed offsets, C <= 4096): [counter: 16 B][sums: 2*4096 doubles][bwd coefficients]
  // counter + sums are zero on entry (contract) and zero again on exit
  double* sums = (double*)workspace + 2;
  BnFwdFinal fin{gamma, beta, running_mean, running_var, mean, rstd, scale, shift, M, eps, momentum,
                 reinterpret_cast<unsigned int*>(workspace)};
  const int vec = vec_for(dtype, C, x);
  int slabs = 1;
  while ((C / slabs) / vec > kThreads || (C % slabs)) ++slabs;  // channel slabs of <= 256*vec channels
  const int Cs = C / slabs;
  UDA_REQUIRE(Cs % vec == 0, UDA_ERR_UNSUPPORTED, "bn_stats: C=%d not supported", C);
  const int groups = kThreads / (Cs / vec);
  long long blocks = (M + groups - 1) / groups;
  long long cap = reduce_grid_cap(M * (long long)C * (dtype == UDA_BF16 ? 2 : 4), slabs);
  if (blocks > cap) blocks = cap;
  size_t smem = (size_t)groups * Cs * 2 * sizeof(float);
  UDA_REQUIRE(smem <= 48 * 1024, UDA_ERR_UNSUPPORTED, "bn_stats: C=%d needs too much shared memory", C);
#define K(T, V) bn_stats_kernel<T, V><<<dim3((unsigned)blocks, slabs), kThreads, smem, st>>>((const T*)x, sums, M, C, Cs, fin)
#define KV(T, ...) do { if (vec == 8) K(T, 8); else if (vec == 4) K(T, 4); else if (vec == 2) K(T, 2); else K(T, 1); } while (0)
  UDA_DT(dtype, KV, 0);
#undef KV
#undef K
  UDA_LAUNCH_OK("bn_stats_kernel");
  return UDA_OK;
}

extern "C" int uda_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean,
                                  const float* running_var, float* scale, float* shift, int C, float eps,
                                  void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(running_mean && running_var && scale && shift && C > 0, UDA_ERR_BAD_ARG, "bn_eval_coeffs: bad argument");
  bn_eval_coeffs_kernel<<<(C + 127) / 128, 128, 0, st>>>(gamma, beta, running_mean, running_var, scale, shift, C, eps);
  UDA_LAUNCH_OK("bn_eval_coeffs_kernel");
  return UDA_OK;
}

extern "C" int uda_bn_apply(const void* x, const void* residual, void* y, int dtype, const float* scale,
                            const float* shift, long long M, int C, float slope, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(x && y && scale && shift && M > 0 && C > 0, UDA_ERR_BAD_ARG, "bn_apply: bad argument");
  if (use_stream() && bn_stream_ok(dtype, M, C) && vec_for(dtype, C, x, residual, y) == 8)
    return bn_apply_stream(x, residual, y, scale, shift, nullptr, BnFwdFinal{}, M, C, slope, st);
  int vec = vec_for(dtype, C, x, residual, y);
  if (reinterpret_cast<uintptr_t>(scale) % 16 || reinterpret_cast<uintptr_t>(shift) % 16) vec = 1;
  const long long n = M * C;
#define K(T, V) bn_apply_kernel<T, V><<<grid_for(n / V / 4), kThreads, 2 * C * sizeof(float), st>>>((const T*)x, (const T*)residual, (T*)y, scale, shift, n, C, slope, nullptr, BnFwdFinal{})
#define KV(T, ...) do { if (vec == 8) K(T, 8); else if (vec == 4) K(T, 4); else if (vec == 2) K(T, 2); else K(T, 1); } while (0)
  UDA_DT(dtype, KV, 0);
#undef KV
#undef K
  UDA_LAUNCH_OK("bn_apply_kernel");
  return UDA_OK;
}

// The ResNet stem tail as one pass: uda_bn_apply_fused followed by uda_maxpool3x3s2_fwd of its output, without reading
// the normalised tensor back (stream_kernels.cu: bn_apply_maxpool_stream_kernel).  Declines (nothing launched) unless
// bf16, H and W even, C a power of two in [8, 2048] and four rows of x fit the shared-memory ring.
extern "C" int uda_bn_apply_maxpool_fused(const void* x, void* a, void* y, unsigned char* idx, int dtype,
                                          const double* sums, const float* gamma, const float* beta,
                                          float* running_mean, float* running_var, float* mean, float* rstd,
                                          float* scale, float* shift, int B, int H, int W, int C, float eps,
                                          float momentum, float slope, void* stream) {
  UDA_REQUIRE(x && a && y && idx && sums && mean && rstd && scale && shift && B > 0 && H > 0 && W > 0 && C > 0,
              UDA_ERR_BAD_ARG, "bn_apply_maxpool_fused: bad argument");
  if (dtype != UDA_BF16 || !use_stream() || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(a) |
                                              reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(idx)) & 15))
    return UDA_ERR_UNSUPPORTED;
  const long long M = (long long)B * H * W;
  BnFwdFinal fin{gamma, beta, running_mean, running_var, mean, rstd, scale, shift, M, eps, momentum, nullptr};
  return bn_apply_maxpool_stream(x, a, y, idx, sums, fin, B, H, W, C, slope, (cudaStream_t)stream);
}

// BatchNorm apply with the statistics taken from the producing convolution's epilogue (sums = [sum | sum of squares],
// double[2*C]); scale/shift (float[C], scratch outputs kept for API symmetry) may be NULL.
extern "C" int uda_bn_apply_fused(const void* x, const void* residual, void* y, int dtype, const double* sums,
                                  const float* gamma, const float* beta, float* running_mean, float* running_var,
                                  float* mean, float* rstd, float* scale, float* shift, long long M, int C, float eps,
                                  float momentum, float slope, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(x && y && sums && mean && rstd && scale && shift && M > 0 && C > 0 && C <= 4096, UDA_ERR_BAD_ARG,
              "bn_apply_fused: bad argument");
  int vec = vec_for(dtype, C, x, residual, y);
  const long long n = M * C;
  BnFwdFinal fin{gamma, beta, running_mean, running_var, mean, rstd, scale, shift, M, eps, momentum, nullptr};
  if (use_stream() && bn_stream_ok(dtype, M, C) && vec == 8)
    return bn_apply_stream(x, residual, y, nullptr, nullptr, sums, fin, M, C, slope, st);
#define K(T, V) bn_apply_kernel<T, V><<<grid_for(n / V / 4), kThreads, 2 * C * sizeof(float), st>>>((const T*)x, (const T*)residual, (T*)y, nullptr, nullptr, n, C, slope, sums, fin)
#define KV(T, ...) do { if (vec == 8) K(T, 8); else if (vec == 4) K(T, 4); else if (vec == 2) K(T, 2); else K(T, 1); } while (0)
  UDA_DT(dtype, KV, 0);
#undef KV
#undef K
  UDA_LAUNCH_OK("bn_apply_kernel<fused>");
  return UDA_OK;
}

// workspace: 2*C doubles + 3*C floats
extern "C" int uda_bn_bwd_fused_supported(int dtype, long long M, int C) {
  return (use_stream() && bn_stream_ok(dtype, M, C)) ? 1 : 0;
}

// BatchNorm backward apply with the statistics from uda_conv2d_tc_dgrad_bnstats (include/uda_b200.h)
extern "C" int uda_bn_bwd_apply_fused(const void* dy, const void* x, const void* a, int dtype, const double* sums,
                                      int v_is_z, const float* gamma, const float* beta, const float* mean,
                                      const float* rstd, const float* scale, const float* shift, void* dx, void* dres,
                                      int dres_accumulate, float* dgamma, float* dbeta, int param_accumulate,
                                      long long M, int C, float slope, void* stream) {
  UDA_REQUIRE(dy && x && sums && mean && rstd && dx, UDA_ERR_BAD_ARG, "bn_bwd_apply_fused: null pointer");
  UDA_REQUIRE((scale == nullptr) == (shift == nullptr), UDA_ERR_BAD_ARG, "bn_bwd_apply_fused: scale and shift go together");
  UDA_REQUIRE(a || scale || slope == 1.f, UDA_ERR_BAD_ARG, "bn_bwd_apply_fused: the activation mask needs a or scale/shift");
  UDA_REQUIRE(uda_bn_bwd_fused_supported(dtype, M, C), UDA_ERR_UNSUPPORTED,
              "bn_bwd_apply_fused: bf16, power-of-two C <= 2048 and >= 1 MiB tensors only (M=%lld C=%d)", M, C);
  return bn_bwd_apply_fused_stream(dy, x, a, sums, v_is_z, gamma, beta, mean, rstd, scale, shift, dx, dres,
                                   dres_accumulate, dgamma, dbeta, param_accumulate, M, C, slope, (cudaStream_t)stream);
}

extern "C" int uda_bn_bwd(const void* dy, const void* x, const void* a, int dtype, const float* gamma,
                          const float* mean, const float* rstd, const float* scale, const float* shift, void* dx,
                          void* dres, int dres_accumulate,
                          float* dgamma, float* dbeta, int param_accumulate, long long M, int C, float slope,
                          void* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(dy && x && mean && rstd && dx && workspace, UDA_ERR_BAD_ARG, "bn_bwd: null pointer");
  UDA_REQUIRE(M > 0 && C > 0 && C <= 4096, UDA_ERR_BAD_ARG, "bn_bwd: bad shape");
  double* sums = (double*)workspace + 2;                  // see uda_bn_stats for the layout
  unsigned int* counter = reinterpret_cast<unsigned int*>(workspace);
  float* coef = (float*)((double*)workspace + 2 + 2 * 4096);
  BnBwdFinal fin{gamma, mean, rstd, dgamma, dbeta, coef, M, param_accumulate, counter};
  int vec = vec_for(dtype, C, dy, x, a, dx);
  if (dres) vec = vec < vec_for(dtype, C, dres) ? vec : vec_for(dtype, C, dres);
  UDA_REQUIRE((scale == nullptr) == (shift == nullptr), UDA_ERR_BAD_ARG, "bn_bwd: scale and shift go together");
  if (use_stream() && bn_stream_ok(dtype, M, C) && vec == 8)
    return bn_bwd_stream(dy, x, a, mean, rstd, scale, shift, dx, dres, dres_accumulate, sums, coef, fin, M, C, slope, st);
  if (vec > 1 && (reinterpret_cast<uintptr_t>(mean) % 16 || reinterpret_cast<uintptr_t>(rstd) % 16 ||
                  reinterpret_cast<uintptr_t>(coef) % 16 || C % 4 ||
                  (scale && (reinterpret_cast<uintptr_t>(scale) % 16 || reinterpret_cast<uintptr_t>(shift) % 16)))) vec = 1;
  UDA_REQUIRE((scale == nullptr) == (shift == nullptr), UDA_ERR_BAD_ARG, "bn_bwd: scale and shift go together");
  const int rvec = vec;
  int slabs = 1;
  while ((C / slabs) / rvec > kThreads || (C % slabs)) ++slabs;
  const int Cs = C / slabs;
  UDA_REQUIRE(Cs % rvec == 0, UDA_ERR_UNSUPPORTED, "bn_bwd: C=%d not supported", C);
  const int groups = kThreads / (Cs / rvec);
  long long blocks = (M + groups - 1) / groups;
  long long cap = reduce_grid_cap(M * (long long)C * (dtype == UDA_BF16 ? 2 : 4), slabs);
  if (blocks > cap) blocks = cap;
  size_t smem = (size_t)groups * Cs * 2 * sizeof(float);
  UDA_REQUIRE(smem <= 48 * 1024, UDA_ERR_UNSUPPORTED, "bn_bwd: C=%d needs too much shared memory", C);
#define K(T, V) bn_bwd_reduce_kernel<T, V><<<dim3((unsigned)blocks, slabs), kThreads, smem, st>>>((const T*)dy, (const T*)x, (const T*)a, mean, rstd, scale, shift, sums, M, C, Cs, slope, fin)
#define KV(T, ...) do { if (rvec == 8) K(T, 8); else if (rvec == 4) K(T, 4); else if (rvec == 2) K(T, 2); else K(T, 1); } while (0)
  UDA_DT(dtype, KV, 0);
#undef KV
#undef K
  UDA_LAUNCH_OK("bn_bwd_reduce_kernel");
  const long long n = M * C;
#define K(T, V) bn_bwd_apply_kernel<T, V><<<grid_for(n / V / 2), kThreads, 5 * C * sizeof(float), st>>>((const T*)dy, (const T*)x, (const T*)a, coef, scale, shift, (T*)dx, (T*)dres, dres_accumulate, n, C, slope)
#define KV(T, ...) do { if (vec == 8) K(T, 8); else if (vec == 4) K(T, 4); else if (vec == 2) K(T, 2); else K(T, 1); } while (0)
  UDA_DT(dtype, KV, 0);
#undef KV
#undef K
  UDA_LAUNCH_OK("bn_bwd_apply_kernel");
  return UDA_OK;
}

extern "C" int uda_act_bwd(const void* dy, const void* a, void* dx, int dtype, long long n, float slope, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(dy && a && dx && n > 0, UDA_ERR_BAD_ARG, "act_bwd: bad argument");
  int vec = vec_for(dtype, (int)(n % 8 == 0 ? 8 : 1), dy, a, dx);
#define K(T, V) act_bwd_kernel<T, V><<<grid_for(n / V), kThreads, 0, st>>>((const T*)dy, (const T*)a, (T*)dx, n, slope)
#define KV(T, ...) do { if (vec == 8) K(T, 8); else if (vec == 4) K(T, 4); else if (vec == 2) K(T, 2); else K(T, 1); } while (0)
  UDA_DT(dtype, KV, 0);
#undef KV
#undef K
  UDA_LAUNCH_OK("act_bwd_kernel");
  return UDA_OK;
}

extern "C" int uda_bias_act(const void* x, const float* bias, void* y, int dtype, long long M, int C, float slope,
                            void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(x && y && M > 0 && C > 0, UDA_ERR_BAD_ARG, "bias_act: bad argument");
  int vec = vec_for(dtype, C, x, y);
  if (vec > 1 && bias && (reinterpret_cast<uintptr_t>(bias) % 16 || C % 4)) vec = 1;
  const long long n = M * C;
#define K(T, V) bias_act_kernel<T, V><<<grid_for(n / V), kThreads, 0, st>>>((const T*)x, bias, (T*)y, n, C, slope)
#define KV(T, ...) do { if (vec == 8) K(T, 8); else if (vec == 4) K(T, 4); else if (vec == 2) K(T, 2); else K(T, 1); } while (0)
  UDA_DT(dtype, KV, 0);
#undef KV
#undef K
  UDA_LAUNCH_OK("bias_act_kernel");
  return UDA_OK;
}

// out[c] (+)= scale * sum_rows x[r][c];  workspace: C doubles
extern "C" int uda_colsum(const void* x, int dtype, float* out, long long M, int C, float scale, int accumulate,
                          void* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(x && out && workspace && M > 0 && C > 0 && C <= 4096, UDA_ERR_BAD_ARG, "colsum: bad argument");
  double* sums = (double*)workspace;
  UDA_CUDA_OK(cudaMemsetAsync(sums, 0, C * sizeof(double), st));
  const int vec = vec_for(dtype, C, x);
  int slabs = 1;
  while ((C / slabs) / vec > kThreads || (C % slabs)) ++slabs;
  const int Cs = C / slabs;
  UDA_REQUIRE(Cs % vec == 0, UDA_ERR_UNSUPPORTED, "colsum: C=%d not supported", C);
  const int groups = kThreads / (Cs / vec);
  long long blocks = (M + groups - 1) / groups;
  long long cap = reduce_grid_cap(M * (long long)C * (dtype == UDA_BF16 ? 2 : 4), slabs);
  if (blocks > cap) blocks = cap;
  size_t smem = (size_t)groups * Cs * sizeof(float);
#define K(T, V) colsum_kernel<T, V><<<dim3((unsigned)blocks, slabs), kThreads, smem, st>>>((const T*)x, sums, M, C, Cs)
#define KV(T, ...) do { if (vec == 8) K(T, 8); else if (vec == 4) K(T, 4); else if (vec == 2) K(T, 2); else K(T, 1); } while (0)
  UDA_DT(dtype, KV, 0);
#undef KV
#undef K
  UDA_LAUNCH_OK("colsum_kernel");
  sums_to_f32_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, out, C, scale, accumulate);
  UDA_LAUNCH_OK("sums_to_f32_kernel");
  return UDA_OK;
}

extern "C" int uda_maxpool3x3s2_fwd(const void* x, void* y, unsigned char* idx, int dtype, int B, int H, int W,
                                    int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(x && y && B > 0 && H > 0 && W > 0 && C > 0, UDA_ERR_BAD_ARG, "maxpool_fwd: bad argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  int vec = vec_for(dtype, C, x, y);
  const long long total = (long long)B * Ho * Wo * (C / vec);
#define K(T, V) do { if (total * 8 < (1LL << 31)) maxpool_fwd_kernel<T, V, int><<<grid_for(total), kThreads, 0, st>>>((const T*)x, (T*)y, idx, B, H, W, C, Ho, Wo); else maxpool_fwd_kernel<T, V, long long><<<grid_for(total), kThreads, 0, st>>>((const T*)x, (T*)y, idx, B, H, W, C, Ho, Wo); } while (0)
#define KV(T, ...) do { if (vec == 8) K(T, 8); else if (vec == 4) K(T, 4); else if (vec == 2) K(T, 2); else K(T, 1); } while (0)
  UDA_DT(dtype, KV, 0);
#undef KV
#undef K
  UDA_LAUNCH_OK("maxpool_fwd_kernel");
  return UDA_OK;
}

extern "C" int uda_maxpool3x3s2_bwd(const void* dy, const unsigned char* idx, const void* addend, void* dx, int dtype,
                                    int B, int H, int W, int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(dy && idx && dx && B > 0 && H > 0 && W > 0 && C > 0, UDA_ERR_BAD_ARG, "maxpool_bwd: bad argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  int vec = vec_for(dtype, C, dy, dx, addend);
  if (vec == 8 && H % 2 == 0 && W % 2 == 0 && aligned<unsigned char>(idx, 8)) {   // 2x2-block form (no divergence)
    const long long items = (long long)B * Ho * Wo * (C / 8);
#define KB(T, ...) do { if (items * 32 < (1LL << 31)) maxpool_bwd_block_kernel<T, int><<<grid_for(items), kThreads, 0, st>>>((const T*)dy, idx, (const T*)addend, (T*)dx, B, H, W, C, Ho, Wo); else maxpool_bwd_block_kernel<T, long long><<<grid_for(items), kThreads, 0, st>>>((const T*)dy, idx, (const T*)addend, (T*)dx, B, H, W, C, Ho, Wo); } while (0)
    UDA_DT(dtype, KB, 0);
#undef KB
    UDA_LAUNCH_OK("maxpool_bwd_block_kernel");
    return UDA_OK;
  }
  const long long total = (long long)B * H * W * (C / vec);
#define K(T, V) do { if (total * 8 < (1LL << 31)) maxpool_bwd_kernel<T, V, int><<<grid_for(total), kThreads, 0, st>>>((const T*)dy, idx, (const T*)addend, (T*)dx, B, H, W, C, Ho, Wo); else maxpool_bwd_kernel<T, V, long long><<<grid_for(total), kThreads, 0, st>>>((const T*)dy, idx, (const T*)addend, (T*)dx, B, H, W, C, Ho, Wo); } while (0)
#define KV(T, ...) do { if (vec == 8) K(T, 8); else if (vec == 4) K(T, 4); else if (vec == 2) K(T, 2); else K(T, 1); } while (0)
  UDA_DT(dtype, KV, 0);
#undef KV
#undef K
  UDA_LAUNCH_OK("maxpool_bwd_kernel");
  return UDA_OK;
}

// H, W are the OUTPUT (upsampled) spatial size; x is [B,H/2,W/2,C1], skip [B,H,W,C2] (C2 may be 0)
extern "C" int uda_upsample2x_concat_fwd(const void* x, const void* skip, void* out, int dtype, int B, int H, int W,
                                         int C1, int C2, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(x && out && (skip || C2 == 0) && B > 0 && H > 0 && W > 0 && C1 > 0 && C2 >= 0, UDA_ERR_BAD_ARG,
              "upcat_fwd: bad argument");
  UDA_REQUIRE(H % 2 == 0 && W % 2 == 0, UDA_ERR_BAD_ARG, "upcat_fwd: output size must be even");
  int vec = vec_for(dtype, C1, x, out);
  if (C2) { int v2 = vec_for(dtype, C2, skip); vec = vec < v2 ? vec : v2; }
  const long long total = (long long)B * H * W * ((C1 + C2) / vec);   // bound for the index type; work items are fewer
#define K(T, V) do { if (total * 8 < (1LL << 31)) upcat_fwd_kernel<T, V, int><<<grid_for(total), kThreads, 0, st>>>((const T*)x, (const T*)skip, (T*)out, B, H, W, C1, C2); else upcat_fwd_kernel<T, V, long long><<<grid_for(total), kThreads, 0, st>>>((const T*)x, (const T*)skip, (T*)out, B, H, W, C1, C2); } while (0)
#define KV(T, ...) do { if (vec == 8) K(T, 8); else if (vec == 4) K(T, 4); else if (vec == 2) K(T, 2); else K(T, 1); } while (0)
  UDA_DT(dtype, KV, 0);
#undef KV
#undef K
  UDA_LAUNCH_OK("upcat_fwd_kernel");
  return UDA_OK;
}

extern "C" int uda_upsample2x_concat_bwd(const void* dout, void* dx, void* dskip, int dtype, int B, int H, int W,
                                         int C1, int C2, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(dout && dx && (dskip || C2 == 0) && B > 0 && H > 0 && W > 0 && C1 > 0 && C2 >= 0, UDA_ERR_BAD_ARG,
              "upcat_bwd: bad argument");
  int vec = vec_for(dtype, C1, dout, dx);
  if (C2) { int v2 = vec_for(dtype, C2, dskip); vec = vec < v2 ? vec : v2; }
  const long long total = (long long)B * (H / 2) * (W / 2) * (C1 / vec) + (long long)B * H * W * (C2 / vec);
#define K(T, V) do { if (total * 8 < (1LL << 31)) upcat_bwd_kernel<T, V, int><<<grid_for(total), kThreads, 0, st>>>((const T*)dout, (T*)dx, (T*)dskip, B, H, W, C1, C2); else upcat_bwd_kernel<T, V, long long><<<grid_for(total), kThreads, 0, st>>>((const T*)dout, (T*)dx, (T*)dskip, B, H, W, C1, C2); } while (0)
#define KV(T, ...) do { if (vec == 8) K(T, 8); else if (vec == 4) K(T, 4); else if (vec == 2) K(T, 2); else K(T, 1); } while (0)
  UDA_DT(dtype, KV, 0);
#undef KV
#undef K
  UDA_LAUNCH_OK("upcat_bwd_kernel");
  return UDA_OK;
}

extern "C" int uda_gap_linear_sigmoid_fwd(const void* x, int dtype, const float* w, const float* bias, float* pooled,
                                          float* out, int B, long long HW, int C, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(x && w && bias && pooled && out && B > 0 && HW > 0 && C > 0, UDA_ERR_BAD_ARG, "gap_linear: bad argument");
  if (dtype == UDA_BF16 && C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0 && HW >= 64 && HW < (1 << 24) && B <= 1024 &&
      aligned<bf16>(x, 16)) {
    // one "slices done" counter per image, left at zero by every launch; one array per device (allocated on first
    // use, which GraphedStep / GraphedFn / GraphedPhases guarantee to happen in their eager warm-up, not under capture)
    static unsigned int* counters_of[64] = {};
    int devid = 0;
    UDA_CUDA_OK(cudaGetDevice(&devid));
    UDA_REQUIRE(devid >= 0 && devid < 64, UDA_ERR_UNSUPPORTED, "gap_linear: device index %d", devid);
    if (!counters_of[devid]) {
      UDA_CUDA_OK(cudaMalloc(&counters_of[devid], 1024 * sizeof(unsigned int)));
      UDA_CUDA_OK(cudaMemset(counters_of[devid], 0, 1024 * sizeof(unsigned int)));
    }
    unsigned int* counters = counters_of[devid];
    UDA_CUDA_OK(cudaMemsetAsync(pooled, 0, (size_t)B * C * sizeof(float), st));
    int slices = (int)(HW / 32);
    const int cap = (2 * num_sms() + B - 1) / B;
    if (slices > cap) slices = cap;
    if (slices < 1) slices = 1;
    gap_partial_kernel<<<dim3((unsigned)slices, (unsigned)B), 256, 0, st>>>((const bf16*)x, w, bias, pooled, out, counters,
                                                                           (int)HW, C);
    UDA_LAUNCH_OK("gap_partial_kernel");
    return UDA_OK;
  }
#define K(T, ...) gap_linear_sigmoid_kernel<T><<<B, 256, 0, st>>>((const T*)x, w, bias, pooled, out, HW, C)
  UDA_DT(dtype, K, 0);
#undef K
  UDA_LAUNCH_OK("gap_linear_sigmoid_kernel");
  return UDA_OK;
}

extern "C" int uda_gap_linear_sigmoid_bwd(const float* dout, const float* y, const float* pooled, const float* w,
                                          float* dw, float* dbias, void* dx, int dtype, int B, long long HW, int C,
                                          int accumulate, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(dout && y && pooled && w && dw && dbias && dx && B > 0, UDA_ERR_BAD_ARG, "gap_linear_bwd: bad argument");
  if (dtype == UDA_BF16 && C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0 && HW * (C / 8) < (1LL << 30) &&
      aligned<bf16>(dx, 16)) {
    long long blocks = (HW * (C / 8) + 255) / 256;
    const long long cap = (8LL * num_sms() + B - 1) / B;
    if (blocks > cap) blocks = cap;
    gap_bwd_fill_kernel<<<dim3((unsigned)blocks, (unsigned)B), 256, 0, st>>>(dout, y, pooled, w, dw, dbias, (bf16*)dx, B,
                                                                            (int)HW, C, accumulate);
    UDA_LAUNCH_OK("gap_bwd_fill_kernel");
    return UDA_OK;
  }
#define K(T, ...) gap_linear_sigmoid_bwd_kernel<T><<<B, 256, 0, st>>>(dout, y, pooled, w, dw, dbias, (T*)dx, B, HW, C, accumulate)
  UDA_DT(dtype, K, 0);
#undef K
  UDA_LAUNCH_OK("gap_linear_sigmoid_bwd_kernel");
  return UDA_OK;
}

extern "C" int uda_adam_step(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                             const float* dev_clip_coef, int* dev_step, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(p && g && m && v && n >= 0 && (dev_step || step >= 1), UDA_ERR_BAD_ARG, "adam: bad argument");
  if (n == 0) return UDA_OK;
  if (dev_step) {
    step_increment_kernel<<<1, 1, 0, st>>>(dev_step);
    UDA_LAUNCH_OK("step_increment_kernel");
    step = 1;
  }
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<grid_for(n), kThreads, 0, st>>>(p, g, m, v, (bf16*)bf16_shadow, n, lr, beta1, beta2, eps,
                                                weight_decay, (float)bc1, (float)sqrt(bc2), grad_scale, dev_clip_coef,
                                                dev_step);
  UDA_LAUNCH_OK("adam_kernel");
  return UDA_OK;
}

// coef[0] = min(1, max_norm / (pre_scale*||g||_2 + 1e-6)); norm_out[0] = pre_scale*||g||_2.  workspace: 1 double
extern "C" int uda_grad_clip_coef(const float* g, long long n, float max_norm, float pre_scale, float* coef,
                                  float* norm_out, void* workspace, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(g && coef && workspace && n > 0, UDA_ERR_BAD_ARG, "grad_clip: bad argument");
  double* acc = (double*)workspace;
  UDA_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(double), st));
  sumsq_kernel<<<grid_for(n, 4), kThreads, 0, st>>>(g, acc, n);
  UDA_LAUNCH_OK("sumsq_kernel");
  clip_coef_kernel<<<1, 32, 0, st>>>(acc, coef, norm_out, max_norm, pre_scale);
  UDA_LAUNCH_OK("clip_coef_kernel");
  return UDA_OK;
}
