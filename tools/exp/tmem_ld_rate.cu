// TMEM -> register read rate of tcgen05.ld on B200 (sm_100a), per SM, as a function of the instruction shape and the
// number of reading warps.  The conv kernels' epilogues drain 128 x BN fp32 accumulators through tcgen05.ld.32x32b.x32;
// in-graph role traces (profiles/r02_trace_step*.txt) show the epilogue, not the MMAs, bounding the few-channel
// high-resolution layers and the exposed tail of every single-tile launch, and doubling the epilogue warps changed
// nothing — this measures whether the TMEM read path itself is the limit.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_ld_rate tmem_ld_rate.cu && ./tmem_ld_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int X> struct Ld;
template <> struct Ld<8> {
  static __device__ __forceinline__ uint32_t go(uint32_t taddr) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= v[i];
    return s;
  }
};
template <> struct Ld<16> {
  static __device__ __forceinline__ uint32_t go(uint32_t taddr) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s ^= v[i];
    return s;
  }
};
template <> struct Ld<32> {
  static __device__ __forceinline__ uint32_t go(uint32_t taddr) {
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr)
                 : "memory");
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s ^= v[i];
    return s;
  }
};

// NW reading warps (warp w reads lane quadrant w % 4), each issues `iters` loads of X columns walking over the 512
// allocated columns; DEPTH loads are in flight before each tcgen05.wait::ld.
template <int X, int DEPTH>
__global__ void __launch_bounds__(512, 1) ld_rate_kernel(int iters, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i += DEPTH) {
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) acc ^= Ld<X>::go(base + (uint32_t)(((i + d) * X) & 511 & ~(X - 1)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_slot), "r"(512) : "memory");
}

template <int X, int DEPTH>
void run(int nw, long long* dout, uint32_t* sink) {
  const int iters = 4096;
  ld_rate_kernel<X, DEPTH><<<148, nw * 32>>>(iters, dout, sink);
  ld_rate_kernel<X, DEPTH><<<148, nw * 32>>>(iters, dout, sink);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("x%d depth %d warps %d: %s\n", X, DEPTH, nw, cudaGetErrorString(e)); return; }
  long long h[148];
  cudaMemcpy(h, dout, sizeof(h), cudaMemcpyDeviceToHost);
  double clk = 0;
  for (int i = 0; i < 148; ++i) clk += (double)h[i];
  clk /= 148;
  const double bytes = (double)nw * iters * 32.0 * X * 4.0;    // per SM
  printf("32x32b.x%-3d in flight %d  warps %2d : %7.1f clk per load per warp, %6.1f B/clk/SM\n", X, DEPTH, nw,
         clk / iters, bytes / clk);
}

int main() {
  long long* dout; uint32_t* sink;
  cudaMalloc(&dout, 148 * sizeof(long long));
  cudaMalloc(&sink, 4);
  for (int nw : {1, 4, 8, 16}) {
    run<32, 1>(nw, dout, sink);
    run<32, 2>(nw, dout, sink);
    run<32, 4>(nw, dout, sink);
    run<16, 1>(nw, dout, sink);
    run<16, 4>(nw, dout, sink);
    run<8, 4>(nw, dout, sink);
  }
  return 0;
}
