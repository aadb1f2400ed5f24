// BatchNorm apply / backward as bulk-copy streaming kernels (bf16 NHWC, C a power of two <= 2048).
// See stream_common.cuh for the pipeline; nn_kernels.cu keeps the generic (fp32 / odd C) versions and
// dispatches here when the fast-path conditions hold.
#include "stream_common.cuh"
#include "bn_common.cuh"
#include <stdlib.h>

namespace uda {
namespace {

using namespace stream;
using namespace bn;

constexpr int kMaxIn = 4, kMaxOut = 2;
// 16 compute warps + one IO warp per CTA, one CTA per SM, 16 KB of every input tensor per stage.  The compute
// warps only ever wait for a tile to land (full barrier) and never synchronise with each other; results are
// written IN PLACE over input tiles that are no longer needed and leave through cp.async.bulk stores issued by
// the IO warp, which also refills the stages (same scheme as loss_stream.cuh).  The first version of these
// kernels (8 KB tiles, two __syncthreads and a single thread doing all copies per tile) sat at ~55 % of the HBM
// roofline on the large backward tensors (ncu launch list, profiles/).
constexpr int kCompute = 512;
constexpr int kCta = kCompute + 32;
constexpr int kTile = 16384;                       // bytes per tensor per stage
constexpr int kVecs = kTile / 16 / kCompute;       // 16-byte vectors per compute thread per tile (2)
constexpr int kMaxStages = 4;

struct StreamIO {
  const uint8_t* in[kMaxIn];
  uint8_t* out[kMaxOut];
  int out_slot[kMaxOut];   // input slot whose shared-memory tile holds output o after the compute step
  long long nbytes;        // per tensor
  int nin, nout, stages;
};

struct Pipe {
  const StreamIO& io;
  uint8_t* smem;
  uint32_t full_base, done_base;
  long long total_tiles;
  int n_my;
  __device__ Pipe(const StreamIO& io_, uint8_t* smem_, uint64_t* bars) : io(io_), smem(smem_) {
    full_base = smem_u32(bars);
    done_base = full_base + 8u * kMaxStages;
    total_tiles = (io.nbytes + kTile - 1) / kTile;
    n_my = total_tiles > blockIdx.x ? (int)((total_tiles - blockIdx.x - 1) / gridDim.x + 1) : 0;
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
      for (int s = 0; s < io.stages; ++s) { mbar_init(full_base + 8u * s, 1); mbar_init(done_base + 8u * s, kCompute); }
      fence_barrier_init();
    }
    __syncthreads();
    pdl_wait();   // barrier setup overlapped the predecessor's tail; its outputs are visible from here on
  }
  __device__ __forceinline__ bool is_io() const { return threadIdx.x >= kCompute; }
  __device__ __forceinline__ long long tile_of(int k) const { return (long long)blockIdx.x + (long long)k * gridDim.x; }
  __device__ __forceinline__ uint32_t bytes_of(long long t) const {
    const long long r = io.nbytes - t * kTile;
    return (uint32_t)(r < kTile ? r : kTile);
  }
  __device__ __forceinline__ uint8_t* tile(int k, int i) const {
    return smem + ((size_t)(k % io.stages) * io.nin + i) * kTile;
  }
  // ---- compute threads ----
  __device__ __forceinline__ void wait(int k) const {
    mbar_wait(full_base + 8u * (k % io.stages), (uint32_t)((k / io.stages) & 1));
  }
  __device__ __forceinline__ void release(int k) const {   // inputs consumed, outputs (if any) written in place
    if (io.nout) fence_async_smem();
    mbar_arrive(done_base + 8u * (k % io.stages));
  }
  // ---- IO warp ----
  __device__ __forceinline__ void load(int k, int lane) const {
    const long long t = tile_of(k);
    const uint32_t nb = bytes_of(t);
    const uint32_t bar = full_base + 8u * (k % io.stages);
    if (lane == 0) mbar_expect_tx(bar, nb * io.nin);
    __syncwarp();
    if (lane < io.nin) bulk_load(smem_u32(tile(k, lane)), io.in[lane] + t * kTile, nb, bar);
  }
  __device__ void io_loop() const {
    const int lane = threadIdx.x - kCompute;
    const int S = io.stages, lag = S >= 4 ? 2 : 1;
    for (int k = 0; k < S && k < n_my; ++k) load(k, lane);
    for (int k = 0; k < n_my; ++k) {
      mbar_wait(done_base + 8u * (k % S), (uint32_t)((k / S) & 1));
      if (io.nout) {
        const long long t = tile_of(k);
        if (lane < io.nout) bulk_store(io.out[lane] + t * kTile, smem_u32(tile(k, io.out_slot[lane])), bytes_of(t));
        bulk_commit();   // bulk groups are per thread: every lane tracks what it stored
        // the stage of tile k-lag is free once ITS stores have read shared memory: refill it
        if (k >= lag && k - lag + S < n_my) {
          if (lag == 2) bulk_wait_read<2>(); else bulk_wait_read<1>();
          __syncwarp();
          load(k - lag + S, lane);
        }
      } else if (k + S < n_my) {
        load(k + S, lane);
      }
    }
    if (io.nout) bulk_wait_all<0>();
  }
};

// ------------------------------------------------------------------------------------------------
// y = act(x*scale + shift (+ residual)); scale/shift from arrays or (sums != null) from the conv-epilogue
// statistics, in which case CTA 0 also publishes mean/rstd/scale/shift and updates the running statistics.
// ------------------------------------------------------------------------------------------------
struct ApplyParams {
  StreamIO io;   // in: x (, residual)   out: y
  const float* scale; const float* shift; const double* sums; BnFwdFinal fin;
  int C; float slope;
};

__global__ void __launch_bounds__(kCta, 1) bn_apply_stream_kernel(const ApplyParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.io.stages * p.io.nin * kTile);
  Pipe pipe(p.io, smem, bars);
  const int C = p.C;
  if (pipe.is_io()) { pipe.io_loop(); return; }
  const int c = (threadIdx.x * 8) % C;   // (kCompute*8) % C == 0: fixed channels per thread
  float sc[8], sf[8];
  if (p.sums) {
    // Per-channel scale/shift once per CTA through shared memory (each channel by one thread), not once per
    // thread: the FP64 pipe of B200 is ~1/64 of FP32 and per-thread DP arithmetic in every CTA was a ~8 us floor
    // under every launch.  FP64 only where the cancellation lives (var = E[x^2] - mean^2).
    float* s_sc = reinterpret_cast<float*>(bars + 2 * kMaxStages);
    float* s_sf = s_sc + C;
    const double inv_m = 1.0 / (double)p.fin.M;
    for (int ch = threadIdx.x; ch < C; ch += kCompute) {
      const double s1 = p.sums[ch], s2 = p.sums[C + ch];
      const double mean = s1 * inv_m;
      const float var = fmaxf((float)(s2 * inv_m - mean * mean), 0.f);
      const float rstd = rsqrtf(var + p.fin.eps);
      const float g = p.fin.gamma ? p.fin.gamma[ch] : 1.f, b = p.fin.beta ? p.fin.beta[ch] : 0.f;
      s_sc[ch] = g * rstd;
      s_sf[ch] = b - (float)mean * g * rstd;
      if (blockIdx.x == 0) bn_fwd_finalize_channel(p.fin, s1, s2, ch);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kCompute) : "memory");   // compute warps only (the IO warp is streaming)
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = s_sc[c + j]; sf[j] = s_sf[c + j]; }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = p.scale[c + j]; sf[j] = p.shift[c + j]; }
  }
  const bool has_res = p.io.nin > 1;
  for (int k = 0; k < pipe.n_my; ++k) {
    pipe.wait(k);
    const uint32_t nb = pipe.bytes_of(pipe.tile_of(k));
    uint8_t* xin = pipe.tile(k, 0);          // y overwrites x in place
    const uint8_t* rin = pipe.tile(k, 1);
#pragma unroll
    for (int u = 0; u < kVecs; ++u) {
      const uint32_t off = (threadIdx.x + u * kCompute) * 16;
      if (off < nb) {
        float v[8], r[8];
        lds8(xin + off, v);
        if (has_res) lds8(rin + off, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float t = v[j] * sc[j] + sf[j];
          if (has_res) t += r[j];
          v[j] = t > 0.f ? t : t * p.slope;
        }
        sts8(xin + off, v);
      }
    }
    pipe.release(k);
  }
}

// ------------------------------------------------------------------------------------------------
// backward reduce: sums[c] += sum g, sums[C+c] += sum g*xhat (g = dy * act'), finalize in the last CTA
// ------------------------------------------------------------------------------------------------
struct ReduceParams {
  StreamIO io;   // in: dy, x (, a)
  const float* mean; const float* rstd; const float* scale; const float* shift;
  double* sums; BnBwdFinal fin;
  int C; float slope; int has_a;
};

__global__ void __launch_bounds__(kCta, 1) bn_bwd_reduce_stream_kernel(const ReduceParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.io.stages * p.io.nin * kTile);
  float* red = reinterpret_cast<float*>(bars + 2 * kMaxStages);   // [kCompute][16]
  Pipe pipe(p.io, smem, bars);
  const int C = p.C;
  if (pipe.is_io()) {
    pipe.io_loop();
  } else {
    const int c = (threadIdx.x * 8) % C;
    const bool zmask = !p.has_a && p.scale != nullptr;
    float nmr[8], rs[8], sc[8], sf[8], s1[8], s2[8];   // nmr = -mean*rstd:  xhat = fma(x, rstd, nmr)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      rs[j] = p.rstd[c + j]; nmr[j] = -p.mean[c + j] * rs[j];
      sc[j] = zmask ? p.scale[c + j] : 0.f; sf[j] = zmask ? p.shift[c + j] : 0.f;
      s1[j] = 0.f; s2[j] = 0.f;
    }
    for (int k = 0; k < pipe.n_my; ++k) {
      pipe.wait(k);
      const uint32_t nb = pipe.bytes_of(pipe.tile_of(k));
      const uint8_t* dyin = pipe.tile(k, 0);
      const uint8_t* xin = pipe.tile(k, 1);
      const uint8_t* ain = pipe.tile(k, 2);
#pragma unroll
      for (int u = 0; u < kVecs; ++u) {
        const uint32_t off = (threadIdx.x + u * kCompute) * 16;
        if (off < nb) {
          float d[8], x[8], a[8];
          lds8(dyin + off, d);
          lds8(xin + off, x);
          if (p.has_a) lds8(ain + off, a);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float g = d[j];
            if (p.has_a) g *= (a[j] > 0.f) ? 1.f : p.slope;
            else if (zmask) g *= (x[j] * sc[j] + sf[j] > 0.f) ? 1.f : p.slope;
            s1[j] += g;
            s2[j] = fmaf(g, fmaf(x[j], rs[j], nmr[j]), s2[j]);
          }
        }
      }
      pipe.release(k);
    }
    // threads with the same channel vector: tid = c/8 + m*(C/8)
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[threadIdx.x * 16 + j] = s1[j]; red[threadIdx.x * 16 + 8 + j] = s2[j]; }
  }
  __syncthreads();
  const int cv = C / 8;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int ch = i % C, k = i / C;
    float acc = 0.f;   // <= 64 fp32 partials; the cross-CTA accumulation is the double atomic
    for (int t = ch / 8; t < kCompute; t += cv) acc += red[t * 16 + k * 8 + (ch & 7)];
    atomicAdd(p.sums + k * C + ch, (double)acc);
  }
  if (last_block_done(p.fin.counter)) {
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      const double a1 = __ldcg(p.sums + ch), a2 = __ldcg(p.sums + C + ch);
      bn_bwd_finalize_channel(p.fin, a1, a2, ch, C);
      p.sums[ch] = 0.0; p.sums[C + ch] = 0.0;
    }
    if (threadIdx.x == 0) *p.fin.counter = 0u;
  }
}

// ------------------------------------------------------------------------------------------------
// backward apply: dx = A*g + Bc*x + Cc ; optional residual-branch gradient dres (=g, or += g)
// ------------------------------------------------------------------------------------------------
struct BwdApplyParams {
  StreamIO io;   // in: dy, x (, a) (, dres when accumulating)   out: dx (, dres)
  const float* coef; const float* scale; const float* shift;
  int C; float slope; int has_a, res_in;   // res_in: index of the dres input (accumulate) or -1
  // fused finalize (uda_bn_bwd_apply_fused): sums = [sum g | sum g*v] from the producing dgrad epilogue
  const double* sums; int v_is_z;
  const float* gamma; const float* beta; const float* mean; const float* rstd;
  float* dgamma; float* dbeta; int accumulate; long long M;
};

__global__ void __launch_bounds__(kCta, 1) bn_bwd_apply_stream_kernel(const BwdApplyParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.io.stages * p.io.nin * kTile);
  Pipe pipe(p.io, smem, bars);
  if (pipe.is_io()) { pipe.io_loop(); return; }
  const int C = p.C;
  const int c = (threadIdx.x * 8) % C;
  const bool zmask = !p.has_a && p.scale != nullptr;
  float kA[8], kB[8], kC[8], sc[8], sf[8];
  if (p.sums) {
    // finalize fused into the apply pass: the per-channel sums come from the epilogue of the dgrad that produced dy.
    //   s1 = sum g;   s2 = sum g*xhat = rstd*(S2 - mean*s1)   (v = z)      or  (S2 - beta*s1)/gamma   (v = the
    //   pre-activation recovered from a:  xhat = (v - beta)/gamma)
    //   dx = A*g + Bc*x + Cc  with  k0 = gamma*rstd,  A = k0,  Bc = -k0*s2/M*rstd,  Cc = -k0*s1/M - Bc*mean
    float* s_coef = reinterpret_cast<float*>(bars + 2 * kMaxStages);
    const double inv_m = 1.0 / (double)p.M;
    for (int ch = threadIdx.x; ch < C; ch += kCompute) {
      const double s1 = p.sums[ch], S2 = p.sums[C + ch];
      const float g = p.gamma ? p.gamma[ch] : 1.f, b = p.beta ? p.beta[ch] : 0.f;
      const float rs = p.rstd[ch], mu = p.mean[ch];
      double s2;
      if (p.v_is_z) s2 = (double)rs * (S2 - (double)mu * s1);
      else s2 = g != 0.f ? (S2 - (double)b * s1) / (double)g : 0.0;
      const float k0 = g * rs;
      const float k1 = k0 * (float)(s1 * inv_m), k2 = k0 * (float)(s2 * inv_m);
      s_coef[ch] = k0;
      s_coef[C + ch] = -k2 * rs;
      s_coef[2 * C + ch] = -k1 + k2 * rs * mu;
      if (blockIdx.x == 0) {
        if (p.dgamma) p.dgamma[ch] = (p.accumulate ? p.dgamma[ch] : 0.f) + (float)s2;
        if (p.dbeta) p.dbeta[ch] = (p.accumulate ? p.dbeta[ch] : 0.f) + (float)s1;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kCompute) : "memory");   // compute warps only
#pragma unroll
    for (int j = 0; j < 8; ++j) { kA[j] = s_coef[c + j]; kB[j] = s_coef[C + c + j]; kC[j] = s_coef[2 * C + c + j]; }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) { kA[j] = p.coef[c + j]; kB[j] = p.coef[C + c + j]; kC[j] = p.coef[2 * C + c + j]; }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = zmask ? p.scale[c + j] : 0.f; sf[j] = zmask ? p.shift[c + j] : 0.f; }
  const bool wres = p.io.nout > 1;
  for (int k = 0; k < pipe.n_my; ++k) {
    pipe.wait(k);
    const uint32_t nb = pipe.bytes_of(pipe.tile_of(k));
    uint8_t* dyin = pipe.tile(k, 0);     // dx overwrites dy in place
    uint8_t* xin = pipe.tile(k, 1);      // dres overwrites x in place
    const uint8_t* ain = pipe.tile(k, 2);
    const uint8_t* rin = pipe.tile(k, p.res_in >= 0 ? p.res_in : 0);
#pragma unroll
    for (int u = 0; u < kVecs; ++u) {
      const uint32_t off = (threadIdx.x + u * kCompute) * 16;
      if (off < nb) {
        float d[8], x[8], a[8], r[8], o[8];
        lds8(dyin + off, d);
        lds8(xin + off, x);
        if (p.has_a) lds8(ain + off, a);
        if (p.res_in >= 0) lds8(rin + off, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float g = d[j];
          if (p.has_a) g *= (a[j] > 0.f) ? 1.f : p.slope;
          else if (zmask) g *= (x[j] * sc[j] + sf[j] > 0.f) ? 1.f : p.slope;
          o[j] = kA[j] * g + kB[j] * x[j] + kC[j];
          d[j] = p.res_in >= 0 ? r[j] + g : g;
        }
        sts8(dyin + off, o);
        if (wres) sts8(xin + off, d);
      }
    }
    pipe.release(k);
  }
}

// ------------------------------------------------------------------------------------------------
// backward reduce + apply as ONE launch for tensors that fit the shared memory of one wave of CTAs (layer3 / layer4 /
// decoder block 0 at B=16, 512x512: <= 8 MB per tensor).  In-graph traces (profiles/r02_trace_step.txt) show the
// two-launch form costing ~25 us per layer on these L2-resident tensors — launch boundaries, prologues and a second
// read, not bandwidth.  Here every CTA loads its tiles ONCE (they stay in shared memory), reduces, meets the other
// CTAs at a grid-wide barrier (grid <= one CTA per SM, all co-resident), finalizes the per-channel coefficients from
// the completed sums and applies them in place; outputs leave through bulk stores.
// ------------------------------------------------------------------------------------------------
struct MergedParams {
  StreamIO io;   // in: dy, x (, a) (, dres when accumulating)   out: dx (, dres);  io.stages >= tiles per CTA
  const float* gamma; const float* mean; const float* rstd; const float* scale; const float* shift;
  double* sums;            // [2*C], zero on entry, zero again on exit
  unsigned int* counters;  // [0] grid-barrier arrivals, [1] CTAs that have read the sums; zero on entry / exit
  float* dgamma; float* dbeta; int accumulate;
  long long M; int C; float slope; int has_a, res_in;
};

__global__ void __launch_bounds__(kCta, 1) bn_bwd_merged_stream_kernel(const MergedParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.io.stages * p.io.nin * kTile);
  float* red = reinterpret_cast<float*>(bars + 2 * kMaxStages);   // [kCompute][16] partial sums, then [3][C] coefficients
  __shared__ bool is_last;
  Pipe pipe(p.io, smem, bars);
  const int C = p.C;
  const int c = (threadIdx.x * 8) % C;
  const bool zmask = !p.has_a && p.scale != nullptr;
  float sc[8], sf[8];
  if (pipe.is_io()) {
    const int lane = threadIdx.x - kCompute;
    for (int k = 0; k < pipe.n_my; ++k) pipe.load(k, lane);      // every tile of this CTA: they stay resident
  } else {
    float nmr[8], rs[8], s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      rs[j] = p.rstd[c + j]; nmr[j] = -p.mean[c + j] * rs[j];
      sc[j] = zmask ? p.scale[c + j] : 0.f; sf[j] = zmask ? p.shift[c + j] : 0.f;
      s1[j] = 0.f; s2[j] = 0.f;
    }
    for (int k = 0; k < pipe.n_my; ++k) {
      pipe.wait(k);
      const uint32_t nb = pipe.bytes_of(pipe.tile_of(k));
      const uint8_t* dyin = pipe.tile(k, 0);
      const uint8_t* xin = pipe.tile(k, 1);
      const uint8_t* ain = pipe.tile(k, 2);
#pragma unroll
      for (int u = 0; u < kVecs; ++u) {
        const uint32_t off = (threadIdx.x + u * kCompute) * 16;
        if (off < nb) {
          float d[8], x[8], a[8];
          lds8(dyin + off, d);
          lds8(xin + off, x);
          if (p.has_a) lds8(ain + off, a);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float g = d[j];
            if (p.has_a) g *= (a[j] > 0.f) ? 1.f : p.slope;
            else if (zmask) g *= (x[j] * sc[j] + sf[j] > 0.f) ? 1.f : p.slope;
            s1[j] += g;
            s2[j] = fmaf(g, fmaf(x[j], rs[j], nmr[j]), s2[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { red[threadIdx.x * 16 + j] = s1[j]; red[threadIdx.x * 16 + 8 + j] = s2[j]; }
  }
  __syncthreads();
  const int cv = C / 8;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int ch = i % C, k = i / C;
    float acc = 0.f;
    for (int t = ch / 8; t < kCompute; t += cv) acc += red[t * 16 + k * 8 + (ch & 7)];
    atomicAdd(p.sums + k * C + ch, (double)acc);
  }
  // ---- grid-wide barrier (all CTAs co-resident: grid <= SMs, one CTA per SM) ----
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.counters) : "memory");
    long long t0 = 0;
    for (unsigned int spin = 0;; ++spin) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.counters) : "memory");
      if (v >= gridDim.x) break;
      __nanosleep(32);
      if ((spin & 0x3ff) == 0x3ff) {
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000LL) { printf("uda_b200: bn_bwd grid barrier timed out (block %d)\n", blockIdx.x); __trap(); }
      }
    }
    __threadfence();
  }
  __syncthreads();
  // ---- per-channel coefficients  dx = kA*g + kB*x + kC  (bn_bwd_finalize_channel's arithmetic) ----
  {
    const float inv_m = 1.f / (float)p.M;
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      const float s1 = (float)__ldcg(p.sums + ch), s2 = (float)__ldcg(p.sums + C + ch);
      const float g = p.gamma ? p.gamma[ch] : 1.f;
      const float r = p.rstd[ch], mu = p.mean[ch];
      const float k0 = g * r;
      const float k1 = k0 * s1 * inv_m, k2 = k0 * s2 * inv_m;
      red[ch] = k0; red[C + ch] = -k2 * r; red[2 * C + ch] = -k1 + k2 * r * mu;
      if (blockIdx.x == 0) {
        if (p.dgamma) p.dgamma[ch] = (p.accumulate ? p.dgamma[ch] : 0.f) + s2;
        if (p.dbeta) p.dbeta[ch] = (p.accumulate ? p.dbeta[ch] : 0.f) + s1;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(p.counters + 1, 1u) == gridDim.x - 1;   // every CTA has read the sums
  if (pipe.is_io()) {
    const int lane = threadIdx.x - kCompute;
    for (int k = 0; k < pipe.n_my; ++k) {
      mbar_wait(pipe.done_base + 8u * (k % p.io.stages), 0u);
      const long long t = pipe.tile_of(k);
      if (lane < p.io.nout) bulk_store(p.io.out[lane] + t * kTile, smem_u32(pipe.tile(k, p.io.out_slot[lane])), pipe.bytes_of(t));
      bulk_commit();
    }
    bulk_wait_all<0>();
  } else {
    float kA[8], kB[8], kC[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { kA[j] = red[c + j]; kB[j] = red[C + c + j]; kC[j] = red[2 * C + c + j]; }
    const bool wres = p.io.nout > 1;
    for (int k = 0; k < pipe.n_my; ++k) {
      const uint32_t nb = pipe.bytes_of(pipe.tile_of(k));
      uint8_t* dyin = pipe.tile(k, 0);     // dx overwrites dy in place
      uint8_t* xin = pipe.tile(k, 1);      // dres overwrites x in place
      const uint8_t* ain = pipe.tile(k, 2);
      const uint8_t* rin = pipe.tile(k, p.res_in >= 0 ? p.res_in : 0);
#pragma unroll
      for (int u = 0; u < kVecs; ++u) {
        const uint32_t off = (threadIdx.x + u * kCompute) * 16;
        if (off < nb) {
          float d[8], x[8], a[8], r[8], o[8];
          lds8(dyin + off, d);
          lds8(xin + off, x);
          if (p.has_a) lds8(ain + off, a);
          if (p.res_in >= 0) lds8(rin + off, r);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float g = d[j];
            if (p.has_a) g *= (a[j] > 0.f) ? 1.f : p.slope;
            else if (zmask) g *= (x[j] * sc[j] + sf[j] > 0.f) ? 1.f : p.slope;
            o[j] = kA[j] * g + kB[j] * x[j] + kC[j];
            d[j] = p.res_in >= 0 ? r[j] + g : g;
          }
          sts8(dyin + off, o);
          if (wres) sts8(xin + off, d);
        }
      }
      pipe.release(k);
    }
  }
  __syncthreads();
  if (is_last) {   // leave the workspace zeroed for the next launch
    for (int ch = threadIdx.x; ch < 2 * C; ch += blockDim.x) p.sums[ch] = 0.0;
    if (threadIdx.x == 0) { p.counters[0] = 0u; p.counters[1] = 0u; }
  }
}

// one CTA per SM; as many stages (<= 4) as fit beside `extra_smem`; small tensors get >= 2 tiles per CTA
int stream_launch_geometry(StreamIO& io, size_t extra_smem, int* grid, size_t* smem) {
#ifdef UDA_B200_EXPERIMENTS
  {   // UDA_B200_BN_DEBUG=1: timing experiment — the kernels launch but move no data (results are WRONG).  Compiled
      // only into experiment builds (make EXPERIMENTS=1); the product library has no work-skipping switch.
    static const bool skip = [] {
      const char* e = getenv("UDA_B200_BN_DEBUG");
      const bool on = e && e[0] == '1';
      if (on) fprintf(stderr, "uda_b200: UDA_B200_BN_DEBUG=1 is a TIMING EXPERIMENT: BatchNorm results are WRONG\n");
      return on;
    }();
    if (skip) io.nbytes = 0;
  }
#endif
  const size_t stage = (size_t)io.nin * kTile;
  int st = (int)((208 * 1024 - extra_smem) / stage);
  if (st > kMaxStages) st = kMaxStages;
  if (st < 2) return set_error(UDA_ERR_UNSUPPORTED, "bn stream: shared memory budget");
  io.stages = st;
  *smem = st * stage + 128 /*align*/ + 64 /*barriers*/ + extra_smem;
  const long long total_tiles = (io.nbytes + kTile - 1) / kTile;
  long long g = num_sms();
  if (g > (total_tiles + 1) / 2) g = (total_tiles + 1) / 2;
  if (g < 1) g = 1;
  *grid = (int)g;
  return UDA_OK;
}

template <typename K>
int set_smem_attr(K kernel) {
  UDA_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
  return UDA_OK;
}

}  // namespace

// fast-path predicate shared with nn_kernels.cu
bool bn_stream_ok(int dtype, long long M, int C) {
  return dtype == UDA_BF16 && C >= 8 && C <= 2048 && (C & (C - 1)) == 0 && M * (long long)C * 2 >= (1 << 20);
}

int bn_apply_stream(const void* x, const void* residual, void* y, const float* scale, const float* shift,
                    const double* sums, const bn::BnFwdFinal& fin, long long M, int C, float slope, cudaStream_t st) {
  ApplyParams p{};
  p.io.in[0] = (const uint8_t*)x; p.io.in[1] = (const uint8_t*)residual; p.io.nin = residual ? 2 : 1;
  p.io.out[0] = (uint8_t*)y; p.io.out_slot[0] = 0; p.io.nout = 1; p.io.nbytes = M * (long long)C * 2;
  p.scale = scale; p.shift = shift; p.sums = sums; p.fin = fin; p.C = C; p.slope = slope;
  int grid; size_t smem;
  if (int rc = stream_launch_geometry(p.io, 2 * (size_t)C * sizeof(float), &grid, &smem)) return rc;
  static bool cfg = false;
  if (!cfg) { if (int rc = set_smem_attr(bn_apply_stream_kernel)) return rc; cfg = true; }
  UDA_CUDA_OK(launch_pdl(bn_apply_stream_kernel, dim3(grid), dim3(kCta), smem, st, p));
  UDA_LAUNCH_OK("bn_apply_stream_kernel");
  return UDA_OK;
}

int bn_bwd_stream(const void* dy, const void* x, const void* a, const float* mean, const float* rstd,
                  const float* scale, const float* shift, void* dx, void* dres, int dres_accumulate, double* sums,
                  float* coef, const bn::BnBwdFinal& fin, long long M, int C, float slope, cudaStream_t st) {
  const long long nbytes = M * (long long)C * 2;
  {   // one launch when every CTA's tiles fit its shared memory.  Opt-in (UDA_B200_BN_BWD_MERGED=1): parity-tested, but in
      // the captured step it measured no faster than the two launches (8.426 vs 8.422 ms; the backward is bound by the
      // sum of kernel work — the side-stream weight-gradient kernels fill every gap — not by launch boundaries)
    const char* const env_merged = getenv("UDA_B200_BN_BWD_MERGED");     // read on every call: the tests switch it
    const bool merged_on = env_merged && env_merged[0] == '1';
    const int nin = 2 + (a ? 1 : 0) + ((dres && dres_accumulate) ? 1 : 0);
    const long long total_tiles = (nbytes + kTile - 1) / kTile;
    const long long grid = total_tiles < num_sms() ? total_tiles : num_sms();
    const long long per_cta = (total_tiles + grid - 1) / grid;
    const size_t extra = kCompute * 16 * sizeof(float) > 3 * (size_t)C * sizeof(float) ? kCompute * 16 * sizeof(float)
                                                                                         : 3 * (size_t)C * sizeof(float);
    const size_t need = (size_t)per_cta * nin * kTile + 128 + 64 + extra;
    if (merged_on && per_cta <= kMaxStages && need <= 224 * 1024 && fin.counter) {
      MergedParams p{};
      int n = 0;
      p.io.in[n++] = (const uint8_t*)dy; p.io.in[n++] = (const uint8_t*)x;
      if (a) { p.io.in[2] = (const uint8_t*)a; n = 3; }
      p.res_in = -1;
      if (dres && dres_accumulate) { p.res_in = a ? 3 : 2; p.io.in[p.res_in] = (const uint8_t*)dres; n = p.res_in + 1; }
      p.io.nin = n;
      p.io.out[0] = (uint8_t*)dx; p.io.out[1] = (uint8_t*)dres; p.io.nout = dres ? 2 : 1;
      p.io.out_slot[0] = 0; p.io.out_slot[1] = 1;   // dx over the dy tile, dres over the x tile
      p.io.nbytes = nbytes; p.io.stages = (int)per_cta;
      p.gamma = fin.gamma; p.mean = mean; p.rstd = rstd; p.scale = scale; p.shift = shift;
      p.sums = sums; p.counters = fin.counter; p.dgamma = fin.dgamma; p.dbeta = fin.dbeta; p.accumulate = fin.accumulate;
      p.M = M; p.C = C; p.slope = slope; p.has_a = a ? 1 : 0;
      static bool cfg = false;
      if (!cfg) { if (int rc = set_smem_attr(bn_bwd_merged_stream_kernel)) return rc; cfg = true; }
      UDA_CUDA_OK(launch_pdl(bn_bwd_merged_stream_kernel, dim3((unsigned)grid), dim3(kCta), need, st, p));
      UDA_LAUNCH_OK("bn_bwd_merged_stream_kernel");
      return UDA_OK;
    }
  }
  {
    ReduceParams p{};
    p.io.in[0] = (const uint8_t*)dy; p.io.in[1] = (const uint8_t*)x; p.io.in[2] = (const uint8_t*)a;
    p.io.nin = a ? 3 : 2; p.io.nout = 0; p.io.nbytes = nbytes;
    p.mean = mean; p.rstd = rstd; p.scale = scale; p.shift = shift; p.sums = sums; p.fin = fin;
    p.C = C; p.slope = slope; p.has_a = a ? 1 : 0;
    int grid; size_t smem;
    if (int rc = stream_launch_geometry(p.io, kCompute * 16 * sizeof(float), &grid, &smem)) return rc;
    static bool cfg = false;
    if (!cfg) { if (int rc = set_smem_attr(bn_bwd_reduce_stream_kernel)) return rc; cfg = true; }
    UDA_CUDA_OK(launch_pdl(bn_bwd_reduce_stream_kernel, dim3(grid), dim3(kCta), smem, st, p));
    UDA_LAUNCH_OK("bn_bwd_reduce_stream_kernel");
  }
  {
    BwdApplyParams p{};
    int n = 0;
    p.io.in[n++] = (const uint8_t*)dy; p.io.in[n++] = (const uint8_t*)x;
    if (a) { p.io.in[2] = (const uint8_t*)a; n = 3; }
    p.res_in = -1;
    if (dres && dres_accumulate) { p.res_in = a ? 3 : 2; p.io.in[p.res_in] = (const uint8_t*)dres; n = p.res_in + 1; }
    p.io.nin = n;
    p.io.out[0] = (uint8_t*)dx; p.io.out[1] = (uint8_t*)dres; p.io.nout = dres ? 2 : 1;
    p.io.out_slot[0] = 0; p.io.out_slot[1] = 1;   // dx over the dy tile, dres over the x tile
    p.io.nbytes = nbytes;
    p.coef = coef; p.scale = scale; p.shift = shift; p.C = C; p.slope = slope; p.has_a = a ? 1 : 0;
    int grid; size_t smem;
    if (int rc = stream_launch_geometry(p.io, 0, &grid, &smem)) return rc;
    static bool cfg = false;
    if (!cfg) { if (int rc = set_smem_attr(bn_bwd_apply_stream_kernel)) return rc; cfg = true; }
    UDA_CUDA_OK(launch_pdl(bn_bwd_apply_stream_kernel, dim3(grid), dim3(kCta), smem, st, p));
    UDA_LAUNCH_OK("bn_bwd_apply_stream_kernel");
  }
  return UDA_OK;
}

// BatchNorm backward whose reduction already happened in the dgrad epilogue (GemmConv::st_sums): one apply launch
// with the per-channel finalize in its prologue.
int bn_bwd_apply_fused_stream(const void* dy, const void* x, const void* a, const double* sums, int v_is_z,
                              const float* gamma, const float* beta, const float* mean, const float* rstd,
                              const float* scale, const float* shift, void* dx, void* dres, int dres_accumulate,
                              float* dgamma, float* dbeta, int param_accumulate, long long M, int C, float slope,
                              cudaStream_t st) {
  BwdApplyParams p{};
  int n = 0;
  p.io.in[n++] = (const uint8_t*)dy; p.io.in[n++] = (const uint8_t*)x;
  if (a) { p.io.in[2] = (const uint8_t*)a; n = 3; }
  p.res_in = -1;
  if (dres && dres_accumulate) { p.res_in = a ? 3 : 2; p.io.in[p.res_in] = (const uint8_t*)dres; n = p.res_in + 1; }
  p.io.nin = n;
  p.io.out[0] = (uint8_t*)dx; p.io.out[1] = (uint8_t*)dres; p.io.nout = dres ? 2 : 1;
  p.io.out_slot[0] = 0; p.io.out_slot[1] = 1;
  p.io.nbytes = M * (long long)C * 2;
  p.coef = nullptr; p.scale = scale; p.shift = shift; p.C = C; p.slope = slope; p.has_a = a ? 1 : 0;
  p.sums = sums; p.v_is_z = v_is_z; p.gamma = gamma; p.beta = beta; p.mean = mean; p.rstd = rstd;
  p.dgamma = dgamma; p.dbeta = dbeta; p.accumulate = param_accumulate; p.M = M;
  int grid; size_t smem;
  if (int rc = stream_launch_geometry(p.io, 3 * (size_t)C * sizeof(float), &grid, &smem)) return rc;
  static bool cfg = false;
  if (!cfg) { if (int rc = set_smem_attr(bn_bwd_apply_stream_kernel)) return rc; cfg = true; }
  UDA_CUDA_OK(launch_pdl(bn_bwd_apply_stream_kernel, dim3(grid), dim3(kCta), smem, st, p));
  UDA_LAUNCH_OK("bn_bwd_apply_stream_kernel");
  return UDA_OK;
}

// ------------------------------------------------------------------------------------------------
// a = act(x*scale + shift) AND (y, idx) = maxpool3x3 s2 p1 of a, ONE pass over x (the ResNet stem tail bn1 -> relu ->
// maxpool).  Separate launches read `a` back for the pooling (x 1 + a 2 + y 0.25 tensor sizes of traffic); here every
// CTA owns a strip of pooled rows and streams the 2*rows (+1 halo) input rows it needs through a shared-memory ring,
// one cp.async.bulk per row: the compute warps normalise a row in place, the IO thread stores it as `a`, and when the
// bottom row of a pooling window has been normalised the window's three rows are still in the ring and the pooled row
// is taken from shared memory (x 1 + a 1 + y 0.25).  Values and indices are those of bn_apply_stream_kernel followed by
// maxpool_fwd_kernel (first maximum in scan order, NaN propagating), bit for bit.
// ------------------------------------------------------------------------------------------------
constexpr int kPoolMaxStages = 8;
struct PoolParams {
  const uint8_t* x; uint8_t* a; uint8_t* y; unsigned char* idx;
  const double* sums; bn::BnFwdFinal fin;
  int B, H, W, C, stages;
  uint32_t row_bytes;
  float slope;
};

namespace {

__global__ void __launch_bounds__(kCta, 1) bn_apply_maxpool_stream_kernel(const PoolParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int S = p.stages, C = p.C;
  const uint32_t RB = p.row_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * RB);
  const uint32_t full_base = smem_u32(bars);                       // row landed (bulk load)
  const uint32_t norm_base = full_base + 8u * kPoolMaxStages;      // row normalised in place by every compute thread
  const uint32_t done_base = norm_base + 8u * kPoolMaxStages;      // row no longer needed by the pooling
  float* s_sc = reinterpret_cast<float*>(bars + 3 * kPoolMaxStages);
  float* s_sf = s_sc + C;
  const int Ho = p.H >> 1, Wo = p.W >> 1;
  // pooled rows r = b*Ho + ho need input rows 2r, 2r+1 (and 2r-1 inside an image): a strip is a contiguous row range
  const long long NR = (long long)p.B * Ho;
  const long long r0 = NR * blockIdx.x / gridDim.x, r1 = NR * (blockIdx.x + 1) / gridDim.x;
  const int halo = (r0 % Ho) > 0 ? 1 : 0;
  const long long z0 = 2 * r0 - halo;                  // first input row (b*H + h) of the strip
  const int n = (int)(2 * (r1 - r0)) + halo;           // rows streamed (host: grid <= NR, so n >= 2)
  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_base + 8u * s, 1); mbar_init(norm_base + 8u * s, kCompute); mbar_init(done_base + 8u * s, kCompute);
    }
    fence_barrier_init();
  }
  __syncthreads();
  pdl_wait();
  if (threadIdx.x >= kCompute) {
    if (threadIdx.x != kCompute) return;
    auto load = [&](int k) {
      const uint32_t bar = full_base + 8u * (k % S);
      mbar_expect_tx(bar, RB);
      bulk_load(smem_u32(smem + (size_t)(k % S) * RB), p.x + (z0 + k) * (long long)RB, RB, bar);
    };
    for (int k = 0; k < S && k < n; ++k) load(k);
    int jr = 0;   // next row whose slot is refilled (with row jr + S)
    for (int k = 0; k < n; ++k) {
      mbar_wait(norm_base + 8u * (k % S), (uint32_t)((k / S) & 1));
      if (k >= halo) bulk_store(p.a + (z0 + k) * (long long)RB, smem_u32(smem + (size_t)(k % S) * RB), RB);
      bulk_commit();   // one group per row (empty for the halo row, which the strip above stores)
      // rows <= k-2 are released by the pooling step that row k completes at the latest; their stores are among all
      // but the two newest groups
      while (jr + S < n && jr <= k - 2) {
        mbar_wait(done_base + 8u * (jr % S), (uint32_t)((jr / S) & 1));
        bulk_wait_read<2>();
        load(jr + S);
        ++jr;
      }
    }
    bulk_wait_all<0>();
    return;
  }
  // ---- compute warps ----
  {
    const double inv_m = 1.0 / (double)p.fin.M;     // same arithmetic as bn_apply_stream_kernel
    for (int ch = threadIdx.x; ch < C; ch += kCompute) {
      const double s1 = p.sums[ch], s2 = p.sums[C + ch];
      const double mean = s1 * inv_m;
      const float var = fmaxf((float)(s2 * inv_m - mean * mean), 0.f);
      const float rstd = rsqrtf(var + p.fin.eps);
      const float g = p.fin.gamma ? p.fin.gamma[ch] : 1.f, b = p.fin.beta ? p.fin.beta[ch] : 0.f;
      s_sc[ch] = g * rstd;
      s_sf[ch] = b - (float)mean * g * rstd;
      if (blockIdx.x == 0) bn_fwd_finalize_channel(p.fin, s1, s2, ch);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kCompute) : "memory");
  }
  const int c = (threadIdx.x * 8) % C;   // (kCompute*8) % C == 0: fixed channels per thread
  float sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = s_sc[c + j]; sf[j] = s_sf[c + j]; }
  const int row_vecs = (int)(RB >> 4);
  const int cv = C >> 3, items = Wo * cv;
  for (int k = 0; k < n; ++k) {
    mbar_wait(full_base + 8u * (k % S), (uint32_t)((k / S) & 1));
    uint8_t* row = smem + (size_t)(k % S) * RB;
    for (int v = threadIdx.x; v < row_vecs; v += kCompute) {
      float t8[8];
      lds8(row + (size_t)v * 16, t8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = t8[j] * sc[j] + sf[j];
        t8[j] = t > 0.f ? t : t * p.slope;
      }
      sts8(row + (size_t)v * 16, t8);
    }
    fence_async_smem();
    mbar_arrive(norm_base + 8u * (k % S));
    const long long g = z0 + k;
    const int h = (int)(g % p.H);
    if (!(h & 1) || (halo && k == 0)) continue;
    // bottom row of pooled row g/2: its window rows k-2 (inside an image), k-1, k are normalised once everyone is here
    asm volatile("bar.sync 1, %0;" ::"n"(kCompute) : "memory");
    const int ho = h >> 1;
    const bool has_top = ho > 0;
    const uint8_t* const row_t = smem + (size_t)((k + S - 2) % S) * RB;
    const uint8_t* const row_m = smem + (size_t)((k + S - 1) % S) * RB;
    const long long out_row = (g >> 1) * (long long)Wo;
    for (int it = threadIdx.x; it < items; it += kCompute) {
      const int v = it % cv, wo = it / cv;
      float best[8];
      int bi[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = -1; }
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        if (kh == 0 && !has_top) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int w = wo * 2 - 1 + kw;
          if (w < 0) continue;                   // w < W always: W is even
          float xv[8];
          lds8((kh == 0 ? row_t : kh == 1 ? row_m : row) + ((size_t)w * C + v * 8) * 2, xv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (bi[j] < 0) bi[j] = kh * 3 + kw;
            if ((xv[j] > best[j]) || (xv[j] != xv[j])) { best[j] = xv[j]; bi[j] = kh * 3 + kw; }
          }
        }
      }
      const long long o = (out_row + wo) * C + v * 8;
      *reinterpret_cast<uint4*>(p.y + o * 2) = make_uint4(pack_bf16x2(best[0], best[1]), pack_bf16x2(best[2], best[3]),
                                                          pack_bf16x2(best[4], best[5]), pack_bf16x2(best[6], best[7]));
      uint2 iv;
      iv.x = (unsigned)bi[0] | ((unsigned)bi[1] << 8) | ((unsigned)bi[2] << 16) | ((unsigned)bi[3] << 24);
      iv.y = (unsigned)bi[4] | ((unsigned)bi[5] << 8) | ((unsigned)bi[6] << 16) | ((unsigned)bi[7] << 24);
      *reinterpret_cast<uint2*>(p.idx + o) = iv;
    }
    // the top and middle rows are finished; the bottom row is the next window's top row unless the image or the strip ends
    if (has_top) mbar_arrive(done_base + 8u * ((k + S - 2) % S));
    mbar_arrive(done_base + 8u * ((k + S - 1) % S));
    if (ho == Ho - 1 || k == n - 1) mbar_arrive(done_base + 8u * (k % S));
  }
}

}  // namespace

int bn_apply_maxpool_stream(const void* x, void* a, void* y, unsigned char* idx, const double* sums,
                            const bn::BnFwdFinal& fin, int B, int H, int W, int C, float slope, cudaStream_t st) {
  const long long rb = (long long)W * C * 2;
  if ((H & 1) || (W & 1) || C < 8 || C > 2048 || (C & (C - 1)) || rb > 64 * 1024) return UDA_ERR_UNSUPPORTED;
  PoolParams p{};
  p.x = (const uint8_t*)x; p.a = (uint8_t*)a; p.y = (uint8_t*)y; p.idx = idx; p.sums = sums; p.fin = fin;
  p.B = B; p.H = H; p.W = W; p.C = C; p.row_bytes = (uint32_t)rb; p.slope = slope;
  const size_t extra = 3 * 8 * kPoolMaxStages + 2 * (size_t)C * sizeof(float) + 128;
  int S = (int)((208 * 1024 - extra) / (size_t)rb);
  if (S > kPoolMaxStages) S = kPoolMaxStages;
  if (S < 4) return UDA_ERR_UNSUPPORTED;
  p.stages = S;
  const long long NR = (long long)B * (H / 2);
  const int grid = (int)(NR < num_sms() ? NR : num_sms());
  static bool cfg = false;
  if (!cfg) { if (int rc = set_smem_attr(bn_apply_maxpool_stream_kernel)) return rc; cfg = true; }
  UDA_CUDA_OK(launch_pdl(bn_apply_maxpool_stream_kernel, dim3(grid), dim3(kCta), (size_t)S * rb + extra, st, p));
  UDA_LAUNCH_OK("bn_apply_maxpool_stream_kernel");
  return UDA_OK;
}

}  // namespace uda
