// Internal interface between the tensor-core convolution translation units.
#pragma once
#include "tc_common.cuh"
#include "bn_common.cuh"

namespace uda {
namespace tcconv {

constexpr int kMaxTaps = 16;
constexpr int kMaxClasses = 4;

// Warp roles of the persistent / halo / pitched-halo kernels: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// then kEpiWarps epilogue warps.  A warp can only read the TMEM lane quadrant (warp % 4), so kEpiSplit = kEpiWarps / 4
// warps share a quadrant and split its work items — (sub-tile, 32-column chunk) pairs — between them.  In-graph role
// traces (tools/trace_step.py, profiles/r02_trace_step.txt) showed the 4-warp epilogue as the exposed tail of every
// single-tile launch (~8 us of a 21 us layer3 convolution) and as what the MMA thread waits for in the few-channel
// high-resolution layers (decoder blocks 3-4, head: more than half of the kernel).  -DUDA_EPI_WARPS=4 restores it.
#ifndef UDA_EPI_WARPS
#define UDA_EPI_WARPS 8
#endif
constexpr int kEpiWarps = UDA_EPI_WARPS;
constexpr int kEpiSplit = kEpiWarps / 4;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kConvThreads = 64 + kEpiThreads;
static_assert(kEpiWarps == 4 || kEpiWarps == 8, "epilogue warps: one or two per TMEM lane quadrant");

struct TilePlan { int TW, TH, NB; bool ok; };

// Tile of `target` (128 or 256) pixels = NB images x TH rows x TW columns of a [B, MH, MW] pixel grid.
inline TilePlan plan_tiles(int B, int MH, int MW, int target = 128) {
  TilePlan t{0, 0, 0, false};
  if (MW <= 0 || MH <= 0) return t;
  t.TW = MW < target ? MW : target;
  if (target % t.TW || MW % t.TW) return t;
  int rows = target / t.TW;
  t.TH = rows < MH ? rows : MH;
  if (rows % t.TH || MH % t.TH) return t;
  t.NB = rows / t.TH;
  if (B % t.NB) return t;
  if (t.TW > 256 || t.TH > 256 || t.NB > 256) return t;
  t.ok = true;
  return t;
}

// One tap class: a set of taps that all write the same output sub-grid (oh, ow).  A forward convolution or
// a stride-1 dgrad has one class; a stride-2 dgrad has one class per output parity.
struct TapClass {
  int ntaps;
  int dh[kMaxTaps], dw[kMaxTaps], ph[kMaxTaps], pw[kMaxTaps], wtap[kMaxTaps];
  int oh, ow;
};

// Training-mode BatchNorm + activation (+ residual) of the convolution's output applied in the SAME launch (north-star
// "BatchNorm+ReLU fused in the epilogue").  Batch statistics couple every output pixel, so the launch must hold ALL its
// accumulators in TMEM across a grid-wide barrier: pass 1 of the epilogue reduces sum / sum of squares of the
// bf16-rounded outputs (GemmConv::bn_sums), every CTA arrives on `counter` and spins until the whole grid has (the grid
// is <= one CTA per SM, all co-resident), then pass 2 re-reads TMEM, stores z (saved for the backward) and
// a = act(z*scale + shift (+ residual)).  Replaces the bn_apply launch (one kernel boundary, one read of z) for every
// layer whose output tile set fits the TMEM of one wave: layer2-4 and decoder blocks 0-1 at B=16, 512x512.
struct BnFuse {
  void* a_out;              // bf16, shape of `out`; nullptr = not fused
  const void* residual;     // optional bf16, shape of `out`
  const float* gamma; const float* beta; float* running_mean; float* running_var;
  float* mean; float* rstd; float* scale; float* shift;     // per-channel outputs (saved for the backward)
  long long M;              // pixels per channel (B*OH*OW)
  double inv_m;             // 1.0 / M computed on the host (a DP divide on the device is a long software sequence)
  float eps, momentum, slope;
  unsigned int* counter;    // zero on entry; arrival counter of the grid barrier
};

// out[b, i*os+oh, j*os+ow, :] = sum_taps src[b, (i,j)+tap, :] * wmat[:, wtap, :]   (+bias, +addend)
//   src  : [B,SH,SW,Cred] bf16 read through a stride-1 4-D map (src_s2 = 0; M-grid = SHxSW) or the stride-2
//          space-to-depth 5-D map (src_s2 = 1; M-grid = SH/2 x SW/2);   wmat : [Cout][wtaps][Cred] bf16
struct GemmConv {
  const void* src; int B, SH, SW, Cred;
  int src_s2;
  const void* wmat; int Cout, wtaps;
  int ncls; TapClass cls[kMaxClasses];
  int OH, OW, os;
  const float* bias; const void* addend; void* out; float* out_nchw;
  // optional activation applied LAST in the epilogue (after bias and addend): v = v > 0 ? v : act_slope * v; act = 0: none.
  // Eval-mode inference folds BatchNorm into the weights / bias, so conv + BN (+ residual) + ReLU is ONE launch.
  int act; float act_slope;
  double* bn_sums;   // optional [2*Cout]: += per-channel sum / sum of squares of the bf16-rounded outputs
  // optional BatchNorm-BACKWARD statistics of the tensor this launch produces (it is the dgrad of the convolution
  // that consumed a = act(BN(z) (+res)), so `out` is dL/da):  st_sums[c] += sum g,  st_sums[Cout+c] += sum g*v  over
  // all output pixels, with g = out * (a > 0 ? 1 : slope) and v = z when st_z is given, else v = the pre-activation
  // recovered from a (a > 0 ? a : a/slope) — see uda_bn_bwd_apply_fused for how the finalize turns them into
  // sum g*xhat.  Every output pixel must be written by exactly one epilogue row (not 1x1 stride-2 dgrads).
  const void* st_a; const void* st_z; float st_slope; double* st_sums;
  // optional: caller-built A-operand tensor map (rank-5 coordinate form {c, w, ph, h, b}) over an M-grid of
  // MH x MW pixels — used by the Cin=3 stem, whose A operand is a padded 4-channel row view of the image
  const CUtensorMap* a_map; int a_MH, a_MW, a_kc;
  const BnFuse* fuse;   // optional (needs bn_sums and out): see BnFuse; UDA_ERR_UNSUPPORTED when the tiles are not TMEM-resident
};

inline int pick_kc(int c) {
  if (c % 64 == 0) return 64;
  if (c % 32 == 0) return 32;
  if (c % 16 == 0) return 16;
  if (c % 8 == 0 && c < 32) return c <= 16 ? 16 : 32;
  return 0;
}
inline int pick_bn(int cout) { return cout > 64 ? 128 : (cout > 32 ? 64 : 32); }

// ---- experiment builds (make EXPERIMENTS=1): per-CTA phase / wait-time trace of the warp roles --------------------
// One record of 16 int64 clocks per CTA, written to the buffer given to uda_exp_set_trace():
//   0 entry  1 after griddepcontrol.wait  2 producer: clocks waiting for free ring slots  3 producer: last TMA issued
//   4 MMA thread: clocks waiting for operands (full barriers)  5 MMA thread: clocks waiting for a drained accumulator
//   6 first MMA issued  7 last commit issued  8 epilogue warp 2: clocks waiting for accumulators  9 epilogue: busy clocks
//   10 epilogue done  11 CTA exit  12 tiles of this CTA     (all times relative to 0)
//   14 %globaltimer (ns) at entry  15 kernel kind | Cout << 8 | Cred << 24 | tiles << 40   (kind: 1 persist, 2 halo, 3 phalo, 4 wgrad_big)
// With uda_exp_set_trace_series(buf, n) every traced launch takes the NEXT slice of 148 x 16 int64 (n slices), so the
// launches of a whole captured training step can be laid on one time axis (tools/trace_step.py).
#ifdef UDA_B200_EXPERIMENTS
extern long long* g_trace_buf;
extern int g_trace_series_left;
inline long long* take_trace_slice() {
  long long* b = g_trace_buf;
  if (b && g_trace_series_left > 0) {
    g_trace_buf += 148 * 16;
    if (--g_trace_series_left == 0) g_trace_buf = nullptr;
  }
  return b;
}
#ifdef __CUDACC__
__device__ __forceinline__ long long trace_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif
#define UDA_TR(...) __VA_ARGS__
#define UDA_TR_WAIT(acc, ...) { const long long _t0 = clock64(); __VA_ARGS__; acc += clock64() - _t0; }
#else
#define UDA_TR(...)
#define UDA_TR_WAIT(acc, ...) { __VA_ARGS__; }
#endif

#ifdef __CUDACC__
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// Grid-wide barrier for the `nthreads` epilogue threads of every CTA (named barrier `id`); `leader` = one of them.
// Everything the calling threads wrote (the statistics atomics) is visible to every thread of the grid afterwards.
// All CTAs of the grid must be co-resident (grid <= SMs at one CTA per SM).  A protocol bug traps instead of hanging.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int expected, int id, int nthreads,
                                             bool leader) {
  bar_sync(id, nthreads);      // CTA-scope: the other threads' global atomics happen-before the leader's release
  if (leader) {
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    long long t0 = 0;
    for (unsigned int spin = 0;; ++spin) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v >= expected) break;
      __nanosleep(32);
      if ((spin & 0x3ff) == 0x3ff) {
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        else if (now - t0 > 4000000000LL) {
          printf("uda_b200: grid barrier timed out (block %d: %u of %u arrived)\n", blockIdx.x, v, expected);
          __trap();
        }
      }
    }
    __threadfence();
  }
  bar_sync(id, nthreads);
}
// per-channel scale / shift of a fused BatchNorm from the completed statistics (same arithmetic as bn_apply_stream_kernel)
__device__ __forceinline__ void bn_fuse_coeffs(const BnFuse& f, const double* sums, int C, int ch, float& sc, float& sf) {
  const double s1 = __ldcg(sums + ch), s2 = __ldcg(sums + C + ch);
  const double inv_m = f.inv_m;
  const double mean = s1 * inv_m;
  const float var = fmaxf((float)(s2 * inv_m - mean * mean), 0.f);
  const float rstd = rsqrtf(var + f.eps);
  const float g = f.gamma ? f.gamma[ch] : 1.f, b = f.beta ? f.beta[ch] : 0.f;
  sc = g * rstd;
  sf = b - (float)mean * g * rstd;
}
// the ONE thread that owns channel `ch` publishes mean / rstd / scale / shift and updates the running statistics
__device__ __forceinline__ void bn_fuse_publish(const BnFuse& f, const double* sums, int C, int ch) {
  const bn::BnFwdFinal fin{f.gamma, f.beta, f.running_mean, f.running_var, f.mean, f.rstd, f.scale, f.shift,
                           f.M, f.eps, f.momentum, nullptr};
  bn::bn_fwd_finalize_channel(fin, __ldcg(sums + ch), __ldcg(sums + C + ch), ch);
}
// Column sums over the 32 rows a warp holds (one row per lane, 32 values per lane): recursive halving,
// 31 shuffles; afterwards lane L holds the sum of column L in v[0].
__device__ __forceinline__ float warp_column_sums(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}
// BatchNorm batch statistics of one 32-column chunk held by an epilogue warp: the values are rounded to bf16
// first (exactly what the separate bn_stats pass would read back), then summed over the warp's 32 rows.
__device__ __forceinline__ void bn_chunk_stats(const float (&f)[32], int lane, float& s, float& q) {
  float t[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) t[k] = __bfloat162float(__float2bfloat16_rn(f[k]));
  float u[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) u[k] = t[k] * t[k];
  s += warp_column_sums(t, lane);
  q += warp_column_sums(u, lane);
}
// the same when `f` is dead afterwards (called after the stores): rounds in place, 64 live values instead of 96
__device__ __forceinline__ void bn_chunk_stats_clobber(float (&f)[32], int lane, float& s, float& q) {
  float u[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) { f[k] = __bfloat162float(__float2bfloat16_rn(f[k])); u[k] = f[k] * f[k]; }
  q += warp_column_sums(u, lane);
  s += warp_column_sums(f, lane);
}
// BatchNorm-backward partial sums of one 32-column chunk (GemmConv::st_*): o = the final output values of this
// thread's row (after the addend), a_row / z_row = the row's 32 channels of a / z (z_row may be null).
__device__ __forceinline__ void bn_bwd_chunk_terms(const float (&o)[32], const bf16* a_row, const bf16* z_row,
                                                   float slope, float inv_slope, int ncols, float (&g)[32],
                                                   float (&gv)[32]) {
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    float av[8], zv[8];
    if (k < ncols) {
      ld_vec<8>(a_row + k, av);
      if (z_row) ld_vec<8>(z_row + k, zv);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float gg = 0.f, vv = 0.f;
      if (k < ncols) {
        const float ob = __bfloat162float(__float2bfloat16_rn(o[k + e]));
        const bool pos = av[e] > 0.f;
        gg = pos ? ob : ob * slope;
        vv = z_row ? zv[e] : (pos ? av[e] : av[e] * inv_slope);
      }
      g[k + e] = gg;
      gv[k + e] = gg * vv;
    }
  }
}
#endif

// conv_tc_persist.cu: persistent, TMEM-double-buffered kernel (all classes in one launch)
int run_gemm_conv_persistent(const GemmConv& g, cudaStream_t st);
// conv_tc_halo.cu: halo-tile kernel for 3x3 stride-1 convolutions on wide images; returns UDA_ERR_UNSUPPORTED
// (no message) when the shape does not qualify
int run_gemm_conv_halo(const GemmConv& g, cudaStream_t st);
// conv_tc_phalo.cu: pitched-halo kernel for 3x3 stride-1 convolutions on narrow images (W = 32 / 64) with 64-multiple
// channel counts; UDA_ERR_UNSUPPORTED (no message) otherwise
int run_gemm_conv_phalo(const GemmConv& g, cudaStream_t st);
// conv_tc_uphalo.cu: conv_transpose4x4_s2_p1 on wide images with <= 32 output channels (decoder conv1, x half);
// UDA_ERR_UNSUPPORTED (no message) otherwise
int run_upconv_halo(const void* x, const void* wx_ft, void* y, double* bn_sums, int B, int h, int w, int C1, int O,
                    cudaStream_t st);
// conv_tc_downhalo.cu: 4x4 stride-2 pad-1 convolution of a wide 16-channel tensor into <= 32 channels (dx of the transposed
// half of decoder conv1 in block 4); UDA_ERR_UNSUPPORTED (no message) otherwise
int run_downconv_halo(const void* x, const void* wmat, void* y, int B, int H, int W, int Cin, int Cout, cudaStream_t st);
int run_wgrad_downhalo(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin, int Cout, cudaStream_t st);
// conv_tc_wgrad_halo.cu: halo-tile wgrad (3x3 stride 1 pad 1, W % 128 == 0); UDA_ERR_UNSUPPORTED otherwise
int run_wgrad_halo(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin, int Cout, cudaStream_t st);
// conv_tc_wgrad_big.cu: multi-accumulator wgrad sharing one dY tile (Cin, Cout multiples of 64); UDA_ERR_UNSUPPORTED otherwise
int run_wgrad_big(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin, int Cout, int KH, int KW,
                  int stride, int pad, cudaStream_t st);

}  // namespace tcconv
}  // namespace uda
