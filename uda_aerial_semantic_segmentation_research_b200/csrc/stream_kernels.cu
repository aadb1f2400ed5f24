// BatchNorm apply / backward as bulk-copy streaming kernels (bf16 NHWC, C a power of two <= 2048).
// See stream_common.cuh for the pipeline; nn_kernels.cu keeps the generic (fp32 / odd C) versions and
// dispatches here when the fast-path conditions hold.
#include "stream_common.cuh"
#include "bn_common.cuh"

namespace uda {
namespace {

using namespace stream;
using namespace bn;

constexpr int kMaxIn = 4, kMaxOut = 2;

struct StreamIO {
  const uint8_t* in[kMaxIn];
  uint8_t* out[kMaxOut];
  long long nbytes;     // per tensor
  int nin, nout, stages;
};

struct Pipe {
  const StreamIO& io;
  uint8_t* smem;
  uint32_t bar_base;
  long long total_tiles;
  int n_my;
  __device__ Pipe(const StreamIO& io_, uint8_t* smem_, uint64_t* bars) : io(io_), smem(smem_) {
    bar_base = smem_u32(bars);
    total_tiles = (io.nbytes + kTileBytes - 1) / kTileBytes;
    n_my = total_tiles > blockIdx.x ? (int)((total_tiles - blockIdx.x - 1) / gridDim.x + 1) : 0;
    if (threadIdx.x == 0) {
      for (int s = 0; s < io.stages; ++s) mbar_init(bar_base + 8u * s, 1);
      fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (int k = 0; k < io.stages && k < n_my; ++k) load(k);
  }
  __device__ __forceinline__ long long tile_of(int k) const { return (long long)blockIdx.x + (long long)k * gridDim.x; }
  __device__ __forceinline__ uint32_t bytes_of(long long t) const {
    const long long r = io.nbytes - t * kTileBytes;
    return (uint32_t)(r < kTileBytes ? r : kTileBytes);
  }
  __device__ __forceinline__ uint8_t* in_tile(int s, int i) const { return smem + ((size_t)s * io.nin + i) * kTileBytes; }
  __device__ __forceinline__ uint8_t* out_tile(int o, int i) const {
    return smem + ((size_t)io.stages * io.nin + (size_t)o * io.nout + i) * kTileBytes;
  }
  __device__ __forceinline__ void load(int k) const {   // one thread
    const int s = k % io.stages;
    const long long t = tile_of(k);
    const uint32_t nb = bytes_of(t);
    const uint32_t bar = bar_base + 8u * s;
    mbar_expect_tx(bar, nb * io.nin);
    for (int i = 0; i < io.nin; ++i) bulk_load(smem_u32(in_tile(s, i)), io.in[i] + t * kTileBytes, nb, bar);
  }
  // all threads: wait for tile k's inputs; when there are outputs, also make sure out buffer (k&1) is free
  __device__ __forceinline__ void acquire(int k) const {
    mbar_wait(bar_base + 8u * (k % io.stages), (uint32_t)((k / io.stages) & 1));
    if (io.nout) {
      if (threadIdx.x == 0) bulk_wait_read<1>();
      __syncthreads();
    }
  }
  // all threads: inputs of tile k consumed (and outputs written to smem): store outputs, refill the stage
  __device__ __forceinline__ void release(int k) const {
    if (io.nout) fence_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      if (io.nout) {
        const long long t = tile_of(k);
        const uint32_t nb = bytes_of(t);
        for (int i = 0; i < io.nout; ++i) bulk_store(io.out[i] + t * kTileBytes, smem_u32(out_tile(k & 1, i)), nb);
        bulk_commit();
      }
      if (k + io.stages < n_my) load(k + io.stages);
    }
  }
  __device__ __forceinline__ void finish() const {
    if (io.nout && threadIdx.x == 0) bulk_wait_all<0>();
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

__global__ void __launch_bounds__(kThreads) bn_apply_stream_kernel(const ApplyParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ((size_t)p.io.stages * p.io.nin + 2 * p.io.nout) * kTileBytes);
  Pipe pipe(p.io, smem, bars);
  const int C = p.C;
  const int c = (threadIdx.x * 8) % C;   // (kThreads*8) % C == 0: fixed channels per thread
  float sc[8], sf[8];
  if (p.sums) {
    // FP64 only where the cancellation lives (var = E[x^2] - mean^2: three DP instructions per channel);
    // everything else in fp32 — B200's FP64 rate makes a DP divide/sqrt per CTA per channel cost tens of us
    const double inv_m = 1.0 / (double)p.fin.M;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double s1 = p.sums[c + j], s2 = p.sums[C + c + j];
      const double mean = s1 * inv_m;
      const float var = fmaxf((float)(s2 * inv_m - mean * mean), 0.f);
      const float rstd = rsqrtf(var + p.fin.eps);
      const float g = p.fin.gamma ? p.fin.gamma[c + j] : 1.f, b = p.fin.beta ? p.fin.beta[c + j] : 0.f;
      sc[j] = g * rstd;
      sf[j] = b - (float)mean * g * rstd;
      if (blockIdx.x == 0 && threadIdx.x < C / 8) bn_fwd_finalize_channel(p.fin, s1, s2, c + j);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = p.scale[c + j]; sf[j] = p.shift[c + j]; }
  }
  const bool has_res = p.io.nin > 1;
  for (int k = 0; k < pipe.n_my; ++k) {
    pipe.acquire(k);
    const int s = k % p.io.stages;
    const uint32_t nb = pipe.bytes_of(pipe.tile_of(k));
    const uint8_t* xin = pipe.in_tile(s, 0);
    const uint8_t* rin = pipe.in_tile(s, 1);
    uint8_t* yout = pipe.out_tile(k & 1, 0);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t off = (threadIdx.x + u * kThreads) * 16;
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
        sts8(yout + off, v);
      }
    }
    pipe.release(k);
  }
  pipe.finish();
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

__global__ void __launch_bounds__(kThreads) bn_bwd_reduce_stream_kernel(const ReduceParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.io.stages * p.io.nin * kTileBytes);
  float* red = reinterpret_cast<float*>(bars + 8);   // [kThreads][16]
  Pipe pipe(p.io, smem, bars);
  const int C = p.C;
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
    pipe.acquire(k);
    const int s = k % p.io.stages;
    const uint32_t nb = pipe.bytes_of(pipe.tile_of(k));
    const uint8_t* dyin = pipe.in_tile(s, 0);
    const uint8_t* xin = pipe.in_tile(s, 1);
    const uint8_t* ain = pipe.in_tile(s, 2);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t off = (threadIdx.x + u * kThreads) * 16;
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
  __syncthreads();
  const int cv = C / 8;
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    const int ch = i % C, k = i / C;
    double acc = 0.0;
    for (int t = ch / 8; t < kThreads; t += cv) acc += (double)red[t * 16 + k * 8 + (ch & 7)];
    atomicAdd(p.sums + k * C + ch, acc);
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
};

__global__ void __launch_bounds__(kThreads) bn_bwd_apply_stream_kernel(const BwdApplyParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ((size_t)p.io.stages * p.io.nin + 2 * p.io.nout) * kTileBytes);
  Pipe pipe(p.io, smem, bars);
  const int C = p.C;
  const int c = (threadIdx.x * 8) % C;
  const bool zmask = !p.has_a && p.scale != nullptr;
  float kA[8], kB[8], kC[8], sc[8], sf[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    kA[j] = p.coef[c + j]; kB[j] = p.coef[C + c + j]; kC[j] = p.coef[2 * C + c + j];
    sc[j] = zmask ? p.scale[c + j] : 0.f; sf[j] = zmask ? p.shift[c + j] : 0.f;
  }
  const bool wres = p.io.nout > 1;
  for (int k = 0; k < pipe.n_my; ++k) {
    pipe.acquire(k);
    const int s = k % p.io.stages;
    const uint32_t nb = pipe.bytes_of(pipe.tile_of(k));
    const uint8_t* dyin = pipe.in_tile(s, 0);
    const uint8_t* xin = pipe.in_tile(s, 1);
    const uint8_t* ain = pipe.in_tile(s, 2);
    const uint8_t* rin = pipe.in_tile(s, p.res_in >= 0 ? p.res_in : 0);
    uint8_t* dxo = pipe.out_tile(k & 1, 0);
    uint8_t* dro = pipe.out_tile(k & 1, 1);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const uint32_t off = (threadIdx.x + u * kThreads) * 16;
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
        sts8(dxo + off, o);
        if (wres) sts8(dro + off, d);
      }
    }
    pipe.release(k);
  }
  pipe.finish();
}

int stream_launch_geometry(const StreamIO& io, size_t extra_smem, int* grid, size_t* smem) {
  const size_t tiles = ((size_t)io.stages * io.nin + 2 * (size_t)io.nout) * kTileBytes;
  *smem = tiles + 128 /*align*/ + 64 /*barriers*/ + extra_smem;
  int per_sm = (int)((200 * 1024) / *smem);
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  const long long total_tiles = (io.nbytes + kTileBytes - 1) / kTileBytes;
  long long g = (long long)num_sms() * per_sm;
  if (g > total_tiles) g = total_tiles;
  if (g < 1) g = 1;
  *grid = (int)g;
  return UDA_OK;
}

template <typename K>
int set_smem_attr(K kernel) {
  UDA_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
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
  p.io.out[0] = (uint8_t*)y; p.io.nout = 1; p.io.nbytes = M * (long long)C * 2; p.io.stages = 4;
  p.scale = scale; p.shift = shift; p.sums = sums; p.fin = fin; p.C = C; p.slope = slope;
  int grid; size_t smem;
  stream_launch_geometry(p.io, 0, &grid, &smem);
  static bool cfg = false;
  if (!cfg) { if (int rc = set_smem_attr(bn_apply_stream_kernel)) return rc; cfg = true; }
  bn_apply_stream_kernel<<<grid, kThreads, smem, st>>>(p);
  UDA_LAUNCH_OK("bn_apply_stream_kernel");
  return UDA_OK;
}

int bn_bwd_stream(const void* dy, const void* x, const void* a, const float* mean, const float* rstd,
                  const float* scale, const float* shift, void* dx, void* dres, int dres_accumulate, double* sums,
                  float* coef, const bn::BnBwdFinal& fin, long long M, int C, float slope, cudaStream_t st) {
  const long long nbytes = M * (long long)C * 2;
  {
    ReduceParams p{};
    p.io.in[0] = (const uint8_t*)dy; p.io.in[1] = (const uint8_t*)x; p.io.in[2] = (const uint8_t*)a;
    p.io.nin = a ? 3 : 2; p.io.nout = 0; p.io.nbytes = nbytes; p.io.stages = 4;
    p.mean = mean; p.rstd = rstd; p.scale = scale; p.shift = shift; p.sums = sums; p.fin = fin;
    p.C = C; p.slope = slope; p.has_a = a ? 1 : 0;
    int grid; size_t smem;
    stream_launch_geometry(p.io, kThreads * 16 * sizeof(float), &grid, &smem);
    // the reduction ends with 2*C atomics per CTA: keep the CTA count moderate
    const int cap = 2 * num_sms();
    if (grid > cap) grid = cap;
    static bool cfg = false;
    if (!cfg) { if (int rc = set_smem_attr(bn_bwd_reduce_stream_kernel)) return rc; cfg = true; }
    bn_bwd_reduce_stream_kernel<<<grid, kThreads, smem, st>>>(p);
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
    p.io.nbytes = nbytes; p.io.stages = 4;
    p.coef = coef; p.scale = scale; p.shift = shift; p.C = C; p.slope = slope; p.has_a = a ? 1 : 0;
    int grid; size_t smem;
    stream_launch_geometry(p.io, 0, &grid, &smem);
    static bool cfg = false;
    if (!cfg) { if (int rc = set_smem_attr(bn_bwd_apply_stream_kernel)) return rc; cfg = true; }
    bn_bwd_apply_stream_kernel<<<grid, kThreads, smem, st>>>(p);
    UDA_LAUNCH_OK("bn_bwd_apply_stream_kernel");
  }
  return UDA_OK;
}

}  // namespace uda
