// Micro-benchmarks behind the CTA-pair convolution kernel, B200 / sm_100a:
//   check : one cluster of two CTAs computes D[256 x N] = A[256 x 64] * B[N x 64]^T with ONE tcgen05.mma.cta_group::2
//           stream (validates the operand split, the cross-CTA barrier protocol and the TMEM layout of tc_pair.cuh)
//   rate  : clocks per tcgen05.mma (K = 16) with both operands resident in shared memory, M = 128 (one CTA) or
//           M = 256 (CTA pair), N = 16 .. 256, optionally while a TMA stream writes into the same shared memory
//   tma   : bytes per clock per SM that TMA box loads of an L2-resident buffer deliver when every SM streams
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I<csrc> -o umma_rate umma_rate.cu
#include "tc_pair.cuh"
#include <vector>
#include <cmath>
#include <cstdlib>
#include <algorithm>
namespace uda { int set_error(int code, const char* fmt, ...) { va_list ap; va_start(ap, fmt); vprintf(fmt, ap); va_end(ap); printf("\n"); return code; } }
using namespace uda; using namespace uda::tc;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// ------------------------------------------------------------------------------------------------------------
// check: D = A * B^T through one CTA pair
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
pair_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* out, int N) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  const uint32_t a_smem = smem_u32(smem), b_smem = a_smem + 16384;
  uint64_t* bars = (uint64_t*)(smem + 16384 + 32768);
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);
  const uint32_t full_bar = smem_u32(bars), done_bar = full_bar + 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t r = cluster_ctarank();
  const uint32_t ncols = N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256));
  if (warp == 0) {
    if (lane == 0) { mbar_init(full_bar, 1); mbar_init(done_bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc_pair(smem_u32(tmem_slot), ncols); tmem_relinquish_pair();
  }
  tc_fence_before(); __syncthreads(); cluster_sync(); tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t leader_full = mapa(full_bar, 0);
    const uint32_t bytes = 16384 + (uint32_t)(N / 2) * 128;
    if (r == 0) mbar_expect_tx(full_bar, 2 * bytes);
    tma_load_2d_pair(a_smem, &map_a, leader_full, 0, (int)r * 128);
    tma_load_2d_pair(b_smem, &map_b, leader_full, 0, (int)r * (N / 2));
    if (r == 0) {
      mbar_wait(full_bar, 0); tc_fence_after();
      const uint32_t idesc = make_idesc_bf16(256, N);
      const uint64_t adesc = make_kmajor_desc(a_smem, 128), bdesc = make_kmajor_desc(b_smem, 128);
      for (int k = 0; k < 4; ++k) umma_bf16_pair(tmem, adesc + 2ull * k, bdesc + 2ull * k, idesc, k ? 1u : 0u);
      umma_commit_pair(done_bar);
    }
  }
  mbar_wait(done_bar, 0); tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v); tmem_ld_wait();
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) out[(size_t)(r * 128 + warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads(); cluster_sync();
  if (warp == 0) { tc_fence_after(); tmem_dealloc_pair(tmem, ncols); }
}

static int make_map2d(CUtensorMap* m, const void* p, int rows, int box_rows) {
  uint64_t dims[2] = {64, (uint64_t)rows}; uint64_t str[1] = {128}; uint32_t box[2] = {64, (uint32_t)box_rows};
  return make_tmap_bf16(m, p, 2, dims, str, box, 128);
}

static int run_check(int N) {
  std::vector<__nv_bfloat16> ha(256 * 64), hb(N * 64);
  for (auto& v : ha) v = __float2bfloat16((rand() % 17 - 8) / 8.f);
  for (auto& v : hb) v = __float2bfloat16((rand() % 13 - 6) / 4.f);
  __nv_bfloat16 *da, *db; float* dout;
  CK(cudaMalloc(&da, ha.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2)); CK(cudaMalloc(&dout, 256 * N * 4));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, 256 * N * 4));
  CUtensorMap ma, mb;
  if (make_map2d(&ma, da, 256, 128) || make_map2d(&mb, db, N, N / 2)) return 1;
  const size_t smem = 16384 + 32768 + 1024 + 256;
  CK(cudaFuncSetAttribute(pair_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(launch_pair(pair_gemm_kernel, dim3(2), dim3(128), smem, 0, false, ma, mb, dout, N));
  CK(cudaDeviceSynchronize());
  std::vector<float> ho(256 * N);
  CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < 64; ++k) ref += (double)__bfloat162float(ha[m * 64 + k]) * __bfloat162float(hb[n * 64 + k]);
      maxerr = std::max(maxerr, std::fabs(ref - ho[m * N + n]));
    }
  printf("check pair GEMM 256 x %3d x 64: max abs err %.3g %s\n", N, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return maxerr < 1e-3 ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------------------
// rate: MMA issue rate with resident operands (+ optional concurrent TMA stream into a 4-slot ring)
// ------------------------------------------------------------------------------------------------------------
struct RateOut { long long mma_clk, tma_clk; };

// CE / WE (compile-time powers of two, 0 = never): a tcgen05.commit every CE MMAs, an mbarrier wait + fence every WE MMAs
// — the issuing thread is ONE in-order thread: whatever it executes between MMAs (runtime divisions!) throttles the pipe
template <int CG, int CE, int WE>
__global__ void __launch_bounds__(128, 1)
rate_kernel(const __grid_constant__ CUtensorMap map_s, RateOut* out, int N, int n_mma, int n_loads, int nboxes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  const uint32_t a_smem = smem_u32(smem), b_smem = a_smem + 16384, ring = b_smem + 32768;
  uint64_t* bars = (uint64_t*)(smem + 16384 + 32768 + 4 * 16384);
  uint32_t* tmem_slot = (uint32_t*)(bars + 8);
  const uint32_t done_bar = smem_u32(bars), sbar0 = done_bar + 8, cbar = done_bar + 40, wbar = done_bar + 48;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t r = CG == 2 ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;  // small bf16 values
  fence_proxy_async();
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(done_bar, 1); for (int s = 0; s < 4; ++s) mbar_init(sbar0 + 8 * s, 1);
      mbar_init(cbar, 1); mbar_init(wbar, 1); mbar_arrive(wbar);   // wbar: phase 0 already complete
      fence_barrier_init();
    }
    __syncwarp();
    if (CG == 2) { tmem_alloc_pair(smem_u32(tmem_slot), 512); tmem_relinquish_pair(); }
    else { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  }
  tc_fence_before(); __syncthreads(); if (CG == 2) cluster_sync(); tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  long long mma_clk = 0, tma_clk = 0;
  if (warp == 1 && lane == 0 && n_mma > 0) {
    if (r == 0) {
      const uint32_t idesc = make_idesc_bf16(128 * CG, N);
      const uint64_t adesc = make_kmajor_desc(a_smem, 128), bdesc = make_kmajor_desc(b_smem, 128);
      const long long t0 = clock64();
      for (int i = 0; i < n_mma; i += 4) {
        if (WE && (i & (WE - 1)) == 0) { mbar_wait(wbar, 0); tc_fence_after(); }   // as in a real main loop
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          // two accumulators alternate so that consecutive MMAs are independent, as in a two-block tile
          const uint32_t acc = tmem + (uint32_t)(((i >> 2) & 1) * N);
          if (CG == 2) umma_bf16_pair(acc, adesc + 2ull * k, bdesc + 2ull * k, idesc, 1u);
          else umma_bf16(acc, adesc + 2ull * k, bdesc + 2ull * k, idesc, 1u);
        }
        // a per-stage "slot free" commit as the pipelined kernels issue it (nobody waits on cbar)
        if (CE && ((i + 4) & (CE - 1)) == 0) { if (CG == 2) umma_commit_pair(cbar); else umma_commit(cbar); }
      }
      if (CG == 2) umma_commit_pair(done_bar); else umma_commit(done_bar);
      mbar_wait(done_bar, 0);
      mma_clk = clock64() - t0;
    } else {
      mbar_wait(done_bar, 0);
    }
  }
  if (warp == 2 && lane == 0 && n_loads > 0) {
    const long long t0 = clock64();
    int box = (blockIdx.x * 37) % nboxes;
    for (int i = 0; i < n_loads; ++i) {
      const int s = i & 3;
      if (i >= 4) mbar_wait(sbar0 + 8 * s, ((i >> 2) - 1) & 1);
      mbar_expect_tx(sbar0 + 8 * s, 16384);
      tma_load_2d(ring + s * 16384, &map_s, sbar0 + 8 * s, 0, box * 128);
      box = box + 1 == nboxes ? 0 : box + 1;
    }
    for (int i = n_loads < 4 ? 0 : n_loads - 4; i < n_loads; ++i) mbar_wait(sbar0 + 8 * (i & 3), (i >> 2) & 1);
    tma_clk = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x == 32) out[blockIdx.x].mma_clk = mma_clk;
  if (threadIdx.x == 64) out[blockIdx.x].tma_clk = tma_clk;
  tc_fence_before(); __syncthreads(); if (CG == 2) cluster_sync();
  if (warp == 0) { tc_fence_after(); if (CG == 2) tmem_dealloc_pair(tmem, 512); else tmem_dealloc(tmem, 512); }
}

template <int CG, int CE = 0, int WE = 0>
static void run_rate(const CUtensorMap& ms, int nboxes, int N, int n_mma, int n_loads, int sms) {
  const int commit_every = CE, wait_every = WE;
  RateOut* dout; CK(cudaMalloc(&dout, sms * sizeof(RateOut))); CK(cudaMemset(dout, 0, sms * sizeof(RateOut)));
  const size_t smem = 16384 + 32768 + 4 * 16384 + 1024 + 256;
  CK(cudaFuncSetAttribute(rate_kernel<CG, CE, WE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = CG == 2 ? sms / 2 * 2 : sms;
  for (int rep = 0; rep < 2; ++rep) {
    if (CG == 2) CK(launch_pair(rate_kernel<CG, CE, WE>, dim3(grid), dim3(128), smem, 0, false, ms, dout, N, n_mma, n_loads, nboxes));
    else { rate_kernel<CG, CE, WE><<<grid, 128, smem>>>(ms, dout, N, n_mma, n_loads, nboxes); CK(cudaGetLastError()); }
    CK(cudaDeviceSynchronize());
  }
  std::vector<RateOut> h(sms);
  CK(cudaMemcpy(h.data(), dout, sms * sizeof(RateOut), cudaMemcpyDeviceToHost));
  double mma = 0, tma = 0; int nm = 0, nt = 0; long long tmax = 0, mmax = 0;
  for (int i = 0; i < grid; ++i) {
    if (h[i].mma_clk) { mma += h[i].mma_clk; ++nm; mmax = std::max(mmax, h[i].mma_clk); }
    if (h[i].tma_clk) { tma += h[i].tma_clk; ++nt; tmax = std::max(tmax, h[i].tma_clk); }
  }
  printf("rate M=%3d N=%3d mma=%5d loads=%4d commit/%-3d wait/%-3d:", 128 * CG, N, n_mma, n_loads, commit_every, wait_every);
  if (nm) {
    const double clk = mma / nm / n_mma;
    printf("  %.1f clk/MMA (max %.1f)  = %.0f%% of the %d-clk floor", clk, (double)mmax / n_mma, 100.0 * (N / 2.0) / clk, N / 2);
  }
  if (nt) printf("  | TMA %.1f B/clk/SM (slowest SM %.1f)", 16384.0 * n_loads / (tma / nt), 16384.0 * n_loads / tmax);
  printf("\n");
  cudaFree(dout);
}


// ------------------------------------------------------------------------------------------------------------
// issuers: is the ~55-clock floor per tcgen05.mma a limit of the issuing THREAD or of the SM's tensor front end?
// NW warps (one elected lane each) issue n_mma MMAs each into their own accumulators.
// ------------------------------------------------------------------------------------------------------------
template <int NW>
__global__ void __launch_bounds__(32 * (NW + 1), 1)
issuers_kernel(RateOut* out, int N, int n_mma) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~uintptr_t(1023));
  const uint32_t a_smem = smem_u32(smem), b_smem = a_smem + 16384;
  uint64_t* bars = (uint64_t*)(smem + 16384 + 32768);
  uint32_t* tmem_slot = (uint32_t*)(bars + 8);
  const uint32_t bar0 = smem_u32(bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (warp == 0) {
    if (lane == 0) { for (int w = 0; w < NW; ++w) mbar_init(bar0 + 8 * w, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish();
  }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  long long clk = 0;
  if (warp >= 1 && lane == 0) {
    const int w = warp - 1;
    const uint32_t idesc = make_idesc_bf16(128, N);
    const uint64_t adesc = make_kmajor_desc(a_smem, 128), bdesc = make_kmajor_desc(b_smem, 128);
    const uint32_t acc = tmem + (uint32_t)(w * (512 / NW));
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 4) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(acc + (uint32_t)(((i >> 2) & 1) * N), adesc + 2ull * k, bdesc + 2ull * k, idesc, 1u);
    }
    umma_commit(bar0 + 8 * w);
    mbar_wait(bar0 + 8 * w, 0);
    clk = clock64() - t0;
  }
  __syncthreads();
  if (threadIdx.x == 32) out[blockIdx.x].mma_clk = clk;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int NW>
static void run_issuers(int N, int n_mma, int sms) {
  RateOut* dout; CK(cudaMalloc(&dout, sms * sizeof(RateOut))); CK(cudaMemset(dout, 0, sms * sizeof(RateOut)));
  const size_t smem = 16384 + 32768 + 1024 + 256;
  CK(cudaFuncSetAttribute(issuers_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int rep = 0; rep < 2; ++rep) { issuers_kernel<NW><<<sms, 32 * (NW + 1), smem>>>(dout, N, n_mma); CK(cudaGetLastError()); CK(cudaDeviceSynchronize()); }
  std::vector<RateOut> h(sms);
  CK(cudaMemcpy(h.data(), dout, sms * sizeof(RateOut), cudaMemcpyDeviceToHost));
  double t = 0; for (int i = 0; i < sms; ++i) t += h[i].mma_clk;
  printf("issuers %d x %5d MMAs  M=128 N=%3d : %.1f clk per MMA of the SM (pipe floor %d)\n", NW, n_mma, N, t / sms / (NW * (double)n_mma), N / 2);
  cudaFree(dout);
}

int main(int argc, char** argv) {
  int dev = 0, sms = 0;
  CK(cudaGetDevice(&dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int bad = 0;
  for (int N : {256, 128, 64, 32}) bad |= run_check(N);
  // L2-resident stream source: 16 MB = 1024 boxes of 128 rows x 128 bytes
  const int nboxes = 1024;
  __nv_bfloat16* dsrc; CK(cudaMalloc(&dsrc, (size_t)nboxes * 16384)); CK(cudaMemset(dsrc, 0, (size_t)nboxes * 16384));
  CUtensorMap ms;
  if (make_map2d(&ms, dsrc, nboxes * 128, 128)) return 1;
  printf("--- MMA only (operands resident in shared memory)\n");
  for (int N : {16, 32, 64, 128, 256}) run_rate<1>(ms, nboxes, N, 4096, 0, sms);
  for (int N : {32, 64, 128, 256}) run_rate<2>(ms, nboxes, N, 4096, 0, sms);
  printf("--- MMA with a tcgen05.commit every c MMAs / an (already complete) mbarrier wait + fence every w MMAs\n");
  run_rate<1, 4, 0>(ms, nboxes, 128, 4096, 0, sms); run_rate<1, 8, 0>(ms, nboxes, 128, 4096, 0, sms);
  run_rate<1, 16, 0>(ms, nboxes, 128, 4096, 0, sms); run_rate<1, 32, 0>(ms, nboxes, 128, 4096, 0, sms);
  run_rate<1, 4, 0>(ms, nboxes, 256, 4096, 0, sms); run_rate<1, 8, 0>(ms, nboxes, 256, 4096, 0, sms);
  run_rate<1, 0, 4>(ms, nboxes, 128, 4096, 0, sms); run_rate<1, 0, 8>(ms, nboxes, 128, 4096, 0, sms);
  run_rate<1, 0, 16>(ms, nboxes, 128, 4096, 0, sms);
  run_rate<1, 4, 4>(ms, nboxes, 128, 4096, 0, sms); run_rate<1, 8, 8>(ms, nboxes, 128, 4096, 0, sms);
  run_rate<1, 16, 16>(ms, nboxes, 128, 4096, 0, sms);
  run_rate<1, 4, 4>(ms, nboxes, 64, 4096, 0, sms); run_rate<1, 8, 8>(ms, nboxes, 64, 4096, 0, sms);
  run_rate<2, 4, 4>(ms, nboxes, 128, 4096, 0, sms); run_rate<2, 8, 8>(ms, nboxes, 128, 4096, 0, sms);
  run_rate<2, 4, 4>(ms, nboxes, 256, 4096, 0, sms); run_rate<2, 8, 8>(ms, nboxes, 256, 4096, 0, sms);
  run_rate<1, 8, 8>(ms, nboxes, 128, 4096, 768, sms); run_rate<2, 8, 8>(ms, nboxes, 256, 2048, 768, sms);
  printf("--- one / two / four issuing warps\n");
  for (int N : {16, 32, 64, 128}) { run_issuers<1>(N, 4096, sms); run_issuers<2>(N, 4096, sms); if (N <= 64) run_issuers<4>(N, 4096, sms); }
  printf("--- TMA only (16 KB boxes of a 16 MB L2-resident buffer, 4 in flight per SM)\n");
  run_rate<1>(ms, nboxes, 128, 0, 1024, sms);
  printf("--- both\n");
  for (int N : {64, 128, 256}) run_rate<1>(ms, nboxes, N, 4096 * 128 / N, 768, sms);
  for (int N : {64, 128, 256}) run_rate<2>(ms, nboxes, N, 4096 * 128 / N, 768, sms);
  return bad;
}
