// Shared helpers for the uda_b200 sm_100a kernels (error plumbing, typed loads, reductions).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>

#include "../../include/uda_b200.h"

namespace uda {

// ---- error plumbing (thread-local message, negative return codes; include/uda_b200.h) -------
int set_error(int code, const char* fmt, ...);

#define UDA_REQUIRE(cond, code, ...)                                  \
  do {                                                                \
    if (!(cond)) return ::uda::set_error((code), __VA_ARGS__);        \
  } while (0)

#define UDA_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ::uda::set_error(UDA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                 \
                              cudaGetErrorString(_e), __FILE__, __LINE__);                  \
  } while (0)

#define UDA_LAUNCH_OK(name)                                                                 \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess)                                                                  \
      return ::uda::set_error(UDA_ERR_CUDA, "launch of %s failed: %s", name,                \
                              cudaGetErrorString(_e));                                      \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T>
inline bool aligned(const void* p, size_t a = sizeof(T)) {
  return (reinterpret_cast<uintptr_t>(p) % a) == 0;
}

// ---- typed scalar / vector access ------------------------------------------------------------
using bf16 = __nv_bfloat16;

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// Load / store VEC consecutive elements (VEC in {1,2,4,8}); pointer must be VEC*sizeof(T) aligned.
template <int VEC> __device__ __forceinline__ void ld_vec(const float* p, float (&v)[VEC]) {
  if constexpr (VEC == 1) {
    v[0] = __ldg(p);
  } else if constexpr (VEC == 2) {
    float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  } else {
#pragma unroll
    for (int i = 0; i < VEC; i += 4) {
      float4 t = __ldg(reinterpret_cast<const float4*>(p + i));
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
  }
}
template <int VEC> __device__ __forceinline__ void ld_vec(const bf16* p, float (&v)[VEC]) {
  if constexpr (VEC == 1) {
    v[0] = __bfloat162float(*p);
  } else if constexpr (VEC == 2) {
    unsigned int t = __ldg(reinterpret_cast<const unsigned int*>(p));
    v[0] = __uint_as_float(t << 16); v[1] = __uint_as_float(t & 0xffff0000u);
  } else if constexpr (VEC == 4) {
    uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; i += 8) {
      uint4 t = __ldg(reinterpret_cast<const uint4*>(p + i));
      v[i + 0] = __uint_as_float(t.x << 16); v[i + 1] = __uint_as_float(t.x & 0xffff0000u);
      v[i + 2] = __uint_as_float(t.y << 16); v[i + 3] = __uint_as_float(t.y & 0xffff0000u);
      v[i + 4] = __uint_as_float(t.z << 16); v[i + 5] = __uint_as_float(t.z & 0xffff0000u);
      v[i + 6] = __uint_as_float(t.w << 16); v[i + 7] = __uint_as_float(t.w & 0xffff0000u);
    }
  }
}
__device__ __forceinline__ unsigned int pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<unsigned int*>(&t);
}
template <int VEC> __device__ __forceinline__ void st_vec(float* p, const float (&v)[VEC]) {
  if constexpr (VEC == 1) {
    *p = v[0];
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
  } else {
#pragma unroll
    for (int i = 0; i < VEC; i += 4)
      *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
}
template <int VEC> __device__ __forceinline__ void st_vec(bf16* p, const float (&v)[VEC]) {
  if constexpr (VEC == 1) {
    *p = __float2bfloat16_rn(v[0]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<unsigned int*>(p) = pack_bf16x2(v[0], v[1]);
  } else if constexpr (VEC == 4) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  } else {
#pragma unroll
    for (int i = 0; i < VEC; i += 8)
      *reinterpret_cast<uint4*>(p + i) =
          make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]),
                     pack_bf16x2(v[i + 4], v[i + 5]), pack_bf16x2(v[i + 6], v[i + 7]));
  }
}

// ---- programmatic dependent launch (PDL) ----------------------------------------------------------
// A training step is ~300 dependent launches of 20-60 us each; with plain stream order every launch pays the drain of
// its predecessor, the launch latency and its own prologue (barrier init, TMEM allocation, descriptor prefetch).
// Kernels launched through launch_pdl() may start as soon as every CTA of the predecessor has executed
// pdl_launch_dependents() (they do so first thing) and an SM is free; pdl_wait() — executed by EVERY CTA before it
// touches global memory and before it exits, so that completion stays transitive along the stream — blocks until
// the predecessor grid has completed and its writes are visible.  Without the launch attribute both are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UDA_B200_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- reductions ------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum of N per-thread floats; result valid in thread 0 (as doubles in `out`).
// `smem` must hold N * (blockDim.x/32) floats.
template <int N>
__device__ __forceinline__ void block_sum(float (&v)[N], float* smem, float (&out)[N]) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float s = warp_sum(v[i]);
    if (lane == 0) smem[i * nw + wid] = s;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float s = (lane < nw) ? smem[i * nw + lane] : 0.f;
      out[i] = warp_sum(s);
    }
  }
  __syncthreads();
}

}  // namespace uda
