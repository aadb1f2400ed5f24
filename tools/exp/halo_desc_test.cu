// Experiment: can a tcgen05 K-major smem descriptor start at an address that is NOT aligned to the swizzle
// repeat (shifted window into a TMA-loaded halo tile)?  Variants: base_offset = 0 vs (addr>>7)&7.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I<csrc> -o halo_desc_test halo_desc_test.cu
#include "tc_common.cuh"
#include <vector>
#include <cmath>
#include <cstdlib>
namespace uda { int set_error(int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); printf("\n"); return code; } }
using namespace uda; using namespace uda::tc;

template <int KC>
__global__ void __launch_bounds__(128) halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                   float* out, int WP, int mode) {
  constexpr int ROWB = KC * 2;                  // bytes per pixel row
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  const int halo_bytes = 3 * WP * ROWB;
  const int halo_pad = (halo_bytes + 1023) / 1024 * 1024;
  const int b_bytes = 32 * ROWB;                // one tap's weights [32 x KC]
  uint64_t* bars = (uint64_t*)(smem + halo_pad + 9 * 1024 * ((b_bytes + 1023) / 1024));
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);
  const uint32_t a_base = smem_u32(smem), b_base = a_base + halo_pad, bar = smem_u32(bars), dbar = bar + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    if (lane == 0) { mbar_init(bar, 1); mbar_init(dbar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 32); tmem_relinquish();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int b_stride = 1024 * ((b_bytes + 1023) / 1024);
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, halo_bytes + 9 * b_bytes);
    tma_load_4d(a_base, &map_a, bar, 0, 0, 0, 0);                  // box {KC, WP, 3, 1}
    for (int t = 0; t < 9; ++t) tma_load_2d(b_base + t * b_stride, &map_b, bar, t * KC, 0);
    mbar_wait(bar, 0); tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(128, 32);
    for (int t = 0; t < 9; ++t) {
      const int kh = t / 3, kw = t % 3;
      const uint32_t a_addr = a_base + (kh * WP + kw) * ROWB;
      uint64_t adesc = make_kmajor_desc(a_addr, ROWB);
      if (mode == 1) adesc |= (uint64_t)((a_addr >> 7) & 7) << 49;
      const uint64_t bdesc = make_kmajor_desc(b_base + t * b_stride, ROWB);
      for (int k = 0; k < KC / 16; ++k) umma_bf16(tmem, adesc + 2ull * k, bdesc + 2ull * k, idesc, (t | k) ? 1u : 0u);
    }
    umma_commit(dbar);
  }
  mbar_wait(dbar, 0); tc_fence_after();
  uint32_t v[32];
  tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), v); tmem_ld_wait();
  for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(v[j]);
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 32); }
}

static float bf(float x) { __nv_bfloat16 b = __float2bfloat16(x); return __bfloat162float(b); }

template <int KC> int run(int mode) {
  const int WP = 130, N = 32;
  std::vector<__nv_bfloat16> hx(3 * WP * KC), hw(N * 9 * KC);
  std::vector<float> fx(hx.size()), fw(hw.size());
  srand(1);
  for (size_t i = 0; i < hx.size(); ++i) { fx[i] = bf((rand() % 2001 - 1000) / 1000.f); hx[i] = __float2bfloat16(fx[i]); }
  for (size_t i = 0; i < hw.size(); ++i) { fw[i] = bf((rand() % 2001 - 1000) / 4000.f); hw[i] = __float2bfloat16(fw[i]); }
  __nv_bfloat16 *dx, *dw; float* dout;
  cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dw, hw.size() * 2); cudaMalloc(&dout, 128 * 32 * 4);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma, mb;
  { uint64_t dims[4] = {KC, WP, 3, 1}; uint64_t str[3] = {KC * 2, (uint64_t)WP * KC * 2, (uint64_t)3 * WP * KC * 2}; uint32_t box[4] = {KC, WP, 3, 1};
    if (make_tmap_bf16(&ma, dx, 4, dims, str, box, KC * 2)) return 1; }
  { uint64_t dims[2] = {9 * KC, N}; uint64_t str[1] = {9 * KC * 2}; uint32_t box[2] = {KC, N};
    if (make_tmap_bf16(&mb, dw, 2, dims, str, box, KC * 2)) return 1; }
  const int smem = 3 * WP * KC * 2 + 1024 + 9 * 4096 + 2048;
  cudaFuncSetAttribute(halo_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  halo_kernel<KC><<<1, 128, smem>>>(ma, mb, dout, WP, mode);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("KC=%d mode=%d: CUDA error %s\n", KC, mode, cudaGetErrorString(e)); return 2; }
  std::vector<float> ho(128 * 32);
  cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
    double acc = 0;
    for (int t = 0; t < 9; ++t) { int kh = t / 3, kw = t % 3;
      for (int c = 0; c < KC; ++c) acc += (double)fx[(kh * WP + m + kw) * KC + c] * fw[(n * 9 + t) * KC + c]; }
    maxerr = fmax(maxerr, fabs(acc - ho[m * 32 + n])); maxref = fmax(maxref, fabs(acc));
  }
  printf("KC=%d (swizzle %dB) base_offset_mode=%d: max_abs_err=%.5f (max_ref=%.3f) -> %s\n", KC, KC * 2, mode, maxerr, maxref,
         maxerr < 1e-2 * maxref ? "OK" : "MISMATCH");
  cudaFree(dx); cudaFree(dw); cudaFree(dout);
  return 0;
}
int main() {
  for (int mode = 0; mode < 2; ++mode) { run<16>(mode); run<32>(mode); run<64>(mode); }
  return 0;
}
