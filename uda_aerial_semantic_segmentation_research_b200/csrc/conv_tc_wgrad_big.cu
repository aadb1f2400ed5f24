// Large-tile tcgen05 wgrad for the wide layers (Cin % 64 == 0, Cout % 64 == 0), sm_100a.
//
//   dW[co][tap][ci] += sum over pixels  dY[pix][co] * X[pix + tap][ci]
//
// conv_tc_wgrad_kernel (conv_tc.cu) gives every CTA one 128 x 128 accumulator: each 128-pixel reduction step
// moves 64 KB through L2 for 512 tensor-pipe cycles (128 B/clk/SM, twice what an SM can pull), and the grid
// (m-tiles x n-tiles x pixel splits) spills into a second wave.  Here a CTA owns up to four accumulators that
// share ONE dY tile: `apg` (tap, channel-chunk) atoms of 64 rows = ceil(apg/2) accumulators of 128 rows, times
// BN <= 256 output channels, NACC x BN <= 512 TMEM columns.  Per 64-pixel reduction step it loads apg X atoms
// + BN/64 dY atoms (8 KB each) for apg x 64 x BN x 64 MACs — 64 B/clk/SM for the 256 x 256 tile — and the grid
// is sized to at most one wave (groups x n-tiles x pixel splits <= SMs).  Operands are MN-major tiles straight
// out of the NHWC tensors, exactly as in conv_tc_wgrad_kernel; partial sums leave through fp32 atomics.
#include "conv_tc_internal.cuh"
#include <stdlib.h>

namespace uda {
namespace tcconv {
namespace {

using namespace tc;

constexpr int kThreads = 192;   // warp 0 TMA, warp 1 MMA + TMEM owner, warps 2..5 epilogue
constexpr int kSmemRing = 200 * 1024;
constexpr int kMaxStages = 8;

struct WBParams {
  int TW, TH, NB, tiles_w, tiles_h;     // KP-pixel boxes over the OUTPUT grid (Ho x Wo)
  int KP;                               // pixels per reduction step (32 or 64)
  int Cin, Cout, ntaps, cchunks;        // cchunks = Cin / 64
  int rank5;                            // input read through the stride-2 space-to-depth view
  signed char dh[kMaxTaps], dw[kMaxTaps], ph[kMaxTaps], pw[kMaxTaps];
  int total_atoms, apg, nacc;           // atoms per group (<= 8), accumulators per CTA
  int n_steps, steps_per_split, stages;
  float* dw_out;                        // [Cout][ntaps][Cin] fp32
  long long* trace;                     // experiment builds only: per-CTA trace records (conv_tc_internal.cuh)
};

__device__ __forceinline__ uint64_t mn_desc128(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;   // LBO: next 64-channel atom
  d |= (uint64_t)(1024u >> 4) << 32;                    // SBO: 8 pixel rows of 128 bytes
  d |= (uint64_t)1 << 46;
  d |= 2ull << 61;                                      // SWIZZLE_128B
  return d;
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_wgrad_big_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                         const WBParams p) {
  constexpr int kBAtoms = BN / 64;
  constexpr uint32_t kTmemCols = 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  const uint32_t atom_bytes = (uint32_t)p.KP * 128u;
  const uint32_t a_bytes = (uint32_t)(2 * p.nacc) * atom_bytes;    // room for an even number of atoms
  const uint32_t stage_bytes = a_bytes + kBAtoms * atom_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kMaxStages);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int group = blockIdx.x;
  const int n0 = blockIdx.y * BN;
  const int atom0 = group * p.apg;
  int valid_atoms = p.total_atoms - atom0;
  if (valid_atoms > p.apg) valid_atoms = p.apg;
  const int st_begin = blockIdx.z * p.steps_per_split;
  int st_end = st_begin + p.steps_per_split;
  if (st_end > p.n_steps) st_end = p.n_steps;
  const int n_iters = st_end - st_begin;   // >= 1 by construction of the grid
  const int tiles_per_group = p.tiles_w * p.tiles_h;

  UDA_TR(const long long tr0 = clock64(); const long long tr_g0 = trace_globaltimer();
         const int tr_cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
         long long* const trp = (p.trace && tr_cta < 148) ? p.trace + (size_t)tr_cta * 16 : nullptr;)
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_dy); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      mbar_init(tmem_full_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the predecessor's tail; its outputs are visible from here on
  UDA_TR(if (trp && threadIdx.x == 0) { trp[0] = tr0; trp[1] = clock64() - tr0; trp[12] = n_iters; trp[14] = tr_g0;
                                        trp[15] = 4LL | ((long long)gridDim.x * gridDim.y * gridDim.z << 40); })

  if (warp == 0) {
    if (elect_one()) {
      UDA_TR(long long tr_w = 0;)
      // The producer is ONE thread issuing up to 16 + 4 TMA instructions per 64-pixel step: anything it computes per
      // instruction is on the critical path (measured: with two integer divisions per atom and four per step the MMA
      // thread waited for operands half of the time).  Atom coordinates are decoded once, the pixel-tile position is a
      // set of counters.
      constexpr int kMaxAtoms = 16;
      int at_c[kMaxAtoms], at_sh[kMaxAtoms];      // channel coordinate; packed (dw, dh, ph) shifts
#pragma unroll
      for (int a = 0; a < kMaxAtoms; ++a) {
        const int gidx = atom0 + (a < valid_atoms ? a : 0);
        const int tap = gidx / p.cchunks, c0 = (gidx - tap * p.cchunks) * 64;
        at_c[a] = p.rank5 ? p.pw[tap] * p.Cin + c0 : c0;
        at_sh[a] = (p.dw[tap] & 0xff) | ((p.dh[tap] & 0xff) << 8) | ((p.ph[tap] & 0xff) << 16);
      }
      int grp = st_begin / tiles_per_group, tin = st_begin - grp * tiles_per_group;
      int th = tin / p.tiles_w, tw = tin - th * p.tiles_w;
      int s = 0; uint32_t phs = 0;
      for (int it = 0; it < n_iters; ++it) {
        UDA_TR_WAIT(tr_w, mbar_wait(empty_bar(s), phs ^ 1))
        const int b0 = grp * p.NB, h0 = th * p.TH, w0 = tw * p.TW;
        const uint32_t a_dst = smem_base + s * stage_bytes;
        const uint32_t b_dst = a_dst + a_bytes;
        mbar_expect_tx(full_bar(s), (uint32_t)(valid_atoms + kBAtoms) * atom_bytes);
#pragma unroll
        for (int j = 0; j < kBAtoms; ++j)
          tma_load_4d(b_dst + j * atom_bytes, &map_dy, full_bar(s), n0 + j * 64, w0, h0, b0);
#pragma unroll
        for (int a = 0; a < kMaxAtoms; ++a) {
          if (a < valid_atoms) {
            const int dw = (int)(signed char)(at_sh[a] & 0xff), dh = (int)(signed char)((at_sh[a] >> 8) & 0xff);
            if (p.rank5)
              tma_load_5d(a_dst + a * atom_bytes, &map_x, full_bar(s), at_c[a], w0 + dw, (at_sh[a] >> 16) & 0xff, h0 + dh, b0);
            else
              tma_load_4d(a_dst + a * atom_bytes, &map_x, full_bar(s), at_c[a], w0 + dw, h0 + dh, b0);
          }
        }
        if (++s == S) { s = 0; phs ^= 1; }
        if (++tw == p.tiles_w) { tw = 0; if (++th == p.tiles_h) { th = 0; ++grp; } }
      }
      UDA_TR(if (trp) { trp[2] = tr_w; trp[3] = clock64() - tr0; })
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // M = 128 rows (two (tap, ci-chunk) atoms), N = BN output channels, both operands MN-major
      constexpr uint32_t idesc = make_idesc_bf16(128, BN) | (1u << 15) | (1u << 16);
      const int nacc = (valid_atoms + 1) >> 1;
      const int ksteps = p.KP >> 4;   // 16 pixels per MMA
      // lean issue loop (one in-order thread feeds the tensor pipe): ring slot / phase as counters, descriptors as a
      // constant high word plus a low word that is only added to
      const uint64_t d0 = mn_desc128(0, atom_bytes);
      const uint32_t dhi = (uint32_t)(d0 >> 32), dlo0 = (uint32_t)d0;      // start-address field zero
      const uint32_t stage16 = stage_bytes >> 4, a16 = a_bytes >> 4, atom16 = atom_bytes >> 4;
      uint32_t st_lo = (smem_base & 0x3FFFFu) >> 4;
      const uint32_t ring_lo = st_lo;
      int s = 0; uint32_t phs = 0;
      UDA_TR(long long tr_wf = 0, tr_first = 0;)
      for (int it = 0; it < n_iters; ++it) {
        UDA_TR_WAIT(tr_wf, mbar_wait(full_bar(s), phs))
        UDA_TR(if (!tr_first) tr_first = clock64() - tr0;)
        tc_fence_after();
        const uint32_t b_lo = dlo0 + st_lo + a16;
        for (int acc = 0; acc < nacc; ++acc) {
          const uint32_t a_lo = dlo0 + st_lo + (uint32_t)(2 * acc) * atom16;
#pragma unroll 4
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(tmem_base + (uint32_t)(acc * BN), desc64(a_lo + k * 128, dhi), desc64(b_lo + k * 128, dhi), idesc,
                      (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
        if (++s == S) { s = 0; phs ^= 1; st_lo = ring_lo; } else { st_lo += stage16; }
      }
      umma_commit(tmem_full_bar);
      UDA_TR(if (trp) { trp[4] = tr_wf; trp[6] = tr_first; trp[7] = clock64() - tr0; })
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int nacc = (valid_atoms + 1) >> 1;
    UDA_TR(long long tr_wt = 0;)
    UDA_TR_WAIT(tr_wt, mbar_wait(tmem_full_bar, 0))
    UDA_TR(const long long tr_b0 = clock64();)
    tc_fence_after();
#pragma unroll 1
    for (int acc = 0; acc < nacc; ++acc) {
      const int a = 2 * acc + (r >> 6);
      const bool row_ok = a < valid_atoms;
      const int gidx = atom0 + (row_ok ? a : 0);
      const int tap = gidx / p.cchunks;
      const int ci = (gidx % p.cchunks) * 64 + (r & 63);
      float* dst = p.dw_out + ((long long)n0 * p.ntaps + tap) * p.Cin + ci;
      const long long co_stride = (long long)p.ntaps * p.Cin;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c), v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int k = 0; k < 32; ++k) atomicAdd(dst + (c + k) * co_stride, __uint_as_float(v[k]));
        }
      }
    }
    tc_fence_before();
    UDA_TR(if (trp && warp == 2 && lane == 0) { trp[8] = tr_wt; trp[9] = clock64() - tr_b0; trp[10] = clock64() - tr0; trp[11] = trp[10]; })
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

bool big_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UDA_B200_WGRAD_BIG");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

template <int BN>
int launch_big(const CUtensorMap& mx, const CUtensorMap& mdy, const WBParams& p, int groups, int n_tiles,
               int splits, size_t smem, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    UDA_CUDA_OK(cudaFuncSetAttribute(conv_tc_wgrad_big_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
    configured = true;
  }
  dim3 grid((unsigned)groups, (unsigned)n_tiles, (unsigned)splits);
  UDA_CUDA_OK(launch_pdl(conv_tc_wgrad_big_kernel<BN>, grid, dim3(kThreads), smem, st, mx, mdy, p));
  UDA_LAUNCH_OK("conv_tc_wgrad_big_kernel");
  return UDA_OK;
}

}  // namespace

// Returns UDA_ERR_UNSUPPORTED (no message) when the shape does not qualify; the caller falls back to
// conv_tc_wgrad_kernel.  Shape checks common to all wgrad kernels are the caller's (wgrad_shape_ok).
int run_wgrad_big(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin, int Cout, int KH, int KW,
                  int stride, int pad, cudaStream_t st) {
  if (!big_enabled() || Cin % 64 || Cout % 64) return UDA_ERR_UNSUPPORTED;
  const int BN = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
  const int Ho = stride == 1 ? H : H / 2, Wo = stride == 1 ? W : W / 2;
  WBParams p{};
  p.Cin = Cin; p.Cout = Cout; p.ntaps = KH * KW; p.cchunks = Cin / 64; p.rank5 = stride == 2;
  p.total_atoms = p.ntaps * p.cchunks;
  int max_atoms = 2 * (512 / BN);                           // TMEM: NACC x BN <= 512 columns
  // ... and three stages of 64-pixel steps must fit: (2*NACC + BN/64) atoms of 8 KB <= ring/3.  Measured: with
  // 32-pixel steps (twice the barrier round trips per byte) dec1.c1 takes 131 us, with six atoms per group 84 us.
  if (plan_tiles(B, Ho, Wo, 64).ok) {
    const int cap = 2 * ((kSmemRing / 3 / (64 * 128) - BN / 64) / 2);
    if (cap >= 2 && cap < max_atoms) max_atoms = cap;
  }
  if (const char* e = getenv("UDA_B200_WGRAD_APG")) {       // experiment hook: cap the atoms per group
    const int v = atoi(e);
    if (v >= 2 && v < max_atoms) max_atoms = v;
  }
  const int groups = (p.total_atoms + max_atoms - 1) / max_atoms;
  p.apg = (p.total_atoms + groups - 1) / groups;
  p.nacc = (p.apg + 1) / 2;
  const int n_tiles = Cout / BN;
  // reduction step: 64 pixels when at least three stages fit, else 32
  TilePlan tp{};
  for (int kp = 64; kp >= 32; kp >>= 1) {
    const int stage_bytes = (2 * p.nacc + BN / 64) * kp * 128;
    int s = kSmemRing / stage_bytes;
    if (s > kMaxStages) s = kMaxStages;
    tp = plan_tiles(B, Ho, Wo, kp);
    if (!tp.ok) continue;
    p.KP = kp; p.stages = s;
    if (s >= 3) break;
  }
  if (p.KP == 0 || p.stages < 2) return UDA_ERR_UNSUPPORTED;
  tp = plan_tiles(B, Ho, Wo, p.KP);
  p.TW = tp.TW; p.TH = tp.TH; p.NB = tp.NB; p.tiles_w = Wo / tp.TW; p.tiles_h = Ho / tp.TH;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      const int t = kh * KW + kw, oh = kh - pad, ow = kw - pad;
      if (stride == 1) {
        p.dh[t] = (signed char)oh; p.dw[t] = (signed char)ow; p.ph[t] = p.pw[t] = 0;
      } else {
        const int ah = oh >= 0 ? oh / 2 : -((-oh + 1) / 2), aw = ow >= 0 ? ow / 2 : -((-ow + 1) / 2);
        p.dh[t] = (signed char)ah; p.dw[t] = (signed char)aw;
        p.ph[t] = (signed char)(oh - 2 * ah); p.pw[t] = (signed char)(ow - 2 * aw);
      }
    }
  p.n_steps = (B / tp.NB) * p.tiles_w * p.tiles_h;
  p.dw_out = dw;
  UDA_TR(p.trace = take_trace_slice();)
  // at most one wave: groups x n_tiles x splits <= SMs
  int splits = num_sms() / (groups * n_tiles);
  if (splits < 1) splits = 1;
  if (splits > p.n_steps) splits = p.n_steps;
  p.steps_per_split = (p.n_steps + splits - 1) / splits;
  splits = (p.n_steps + p.steps_per_split - 1) / p.steps_per_split;

  CUtensorMap mx, mdy;
  const uint64_t C = (uint64_t)Cin;
  if (stride == 1) {
    uint64_t dims[4] = {C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {64, (uint32_t)tp.TW, (uint32_t)tp.TH, (uint32_t)tp.NB};
    if (int rc = make_tmap_bf16(&mx, x, 4, dims, str, box, 128)) return rc;
  } else {
    uint64_t dims[5] = {2 * C, (uint64_t)W / 2, 2, (uint64_t)H / 2, (uint64_t)B};
    uint64_t str[4] = {2 * C * 2, (uint64_t)W * C * 2, 2 * (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[5] = {64, (uint32_t)tp.TW, 1, (uint32_t)tp.TH, (uint32_t)tp.NB};
    if (int rc = make_tmap_bf16(&mx, x, 5, dims, str, box, 128)) return rc;
  }
  {
    const uint64_t Co = (uint64_t)Cout;
    uint64_t dims[4] = {Co, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
    uint64_t str[3] = {Co * 2, (uint64_t)Wo * Co * 2, (uint64_t)Ho * Wo * Co * 2};
    uint32_t box[4] = {64, (uint32_t)tp.TW, (uint32_t)tp.TH, (uint32_t)tp.NB};
    if (int rc = make_tmap_bf16(&mdy, dy, 4, dims, str, box, 128)) return rc;
  }
  const size_t smem = (size_t)p.stages * (2 * p.nacc + BN / 64) * p.KP * 128 + 1024 + 256;
  if (BN == 256) return launch_big<256>(mx, mdy, p, groups, n_tiles, splits, smem, st);
  if (BN == 128) return launch_big<128>(mx, mdy, p, groups, n_tiles, splits, smem, st);
  return launch_big<64>(mx, mdy, p, groups, n_tiles, splits, smem, st);
}

}  // namespace tcconv
}  // namespace uda
