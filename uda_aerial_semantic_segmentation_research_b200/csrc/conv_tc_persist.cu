// Persistent tcgen05 implicit-GEMM convolution (forward / dgrad), sm_100a.
//
// Same operand scheme as conv_tc.cu (TMA boxes of the NHWC tensor as the A operand, OHWI weights as the
// K-major B operand, accumulators in TMEM) with the per-CTA costs amortised:
//   * grid = number of SMs; every CTA walks a static round-robin list of output tiles
//     (tap class x pixel tile x channel tile), so barrier init, TMEM allocation and descriptor prefetch
//     happen once per SM instead of once per 128 pixels;
//   * the TMA producer streams across tile boundaries (the smem ring never drains);
//   * TMEM holds TWO accumulator sets: the epilogue warps drain tile j (tcgen05.ld -> bf16 NHWC /
//     fp32 NCHW stores) while the MMA thread already accumulates tile j+1;
//   * MT = 2: a tile is 256 pixels = two 128-row MMAs that share every weight tile (halves the weight
//     traffic per FLOP, one TMA instruction per 256 pixels);
//   * weight-stationary mode: when the layer's whole weight matrix fits (<= 96 KB, Cout <= 128) it is loaded
//     into shared memory once per CTA and the main loop streams activations only — the 16/32/64-channel
//     high-resolution layers (decoder blocks 2-4, head, layer1) are HBM/issue bound, not FLOP bound;
//   * stride-2 dgrad: the four output-parity classes are tiles of ONE launch.
#include "conv_tc_internal.cuh"
#include "stream_common.cuh"
#include <stdlib.h>

namespace uda {
namespace tcconv {
namespace {

using namespace tc;

constexpr int kThreads = kConvThreads;
constexpr int kMaxStages = 12;

struct PClass {
  int ntaps;
  signed char dh[kMaxTaps], dw[kMaxTaps], ph[kMaxTaps], pw[kMaxTaps];
  unsigned char wtap[kMaxTaps];
  short oh, ow;
};

struct PParams {
  int TW, TH, NB, tiles_w, tiles_h;   // pixel tile (128*MT pixels) and tiles per image group
  int m_tiles, n_tiles, ncls;
  int MH, MW, OH, OW, os;
  int Cout, Cred, kchunks, rank5, wtaps;
  int stages, ws;
  PClass cls[kMaxClasses];
  bf16* out; float* out_nchw; const float* bias; const bf16* addend;
  double* bn_sums;
  int act; float act_slope;            // GemmConv::act
  const bf16* st_a; const bf16* st_z; float st_slope; double* st_sums;   // GemmConv::st_*
  BnFuse fuse;        // fuse.a_out != nullptr: BatchNorm + activation in this launch (FUSE instances: every CTA's tiles stay in TMEM)
  long long* trace;   // experiment builds only: per-CTA trace records (conv_tc_internal.cuh)
  int debug;   // experiment builds only (UDA_B200_TC_DEBUG bit mask): 1 = no epilogue stores, 2 = no MMAs, 4 = no TMA loads
};
#ifdef UDA_B200_EXPERIMENTS
#define UDA_TC_DBG(p, bit) ((p).debug & (bit))
#else
#define UDA_TC_DBG(p, bit) false
#endif

template <int KC, int BN, int MT, bool FUSE>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_persist_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const PParams p) {
  constexpr int kABytes = MT * 128 * KC * 2;
  constexpr int kBBytes = BN * KC * 2;
  constexpr uint32_t kAccCols = MT * BN;                      // one accumulator set
  constexpr int kSets = 2 * kAccCols <= 512 ? 2 : 1;          // 256 x 256 tiles fill TMEM: single-buffered
  constexpr uint32_t kTmemCols = kSets * kAccCols < 32 ? 32 : kSets * kAccCols;   // power of two for every (MT,BN)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  const int ws_bytes = p.ws ? p.wtaps * p.kchunks * kBBytes : 0;
  const int stage_bytes = kABytes + (p.ws ? 0 : kBBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ws_bytes + S * stage_bytes);
  // bars: full[kMaxStages], empty[kMaxStages], tmem_full[2], tmem_empty[2], ws_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 5);
  const uint32_t ws_base = smem_u32(smem);
  const uint32_t ring_base = ws_base + ws_bytes;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int q) { return bar_base + 8u * (2 * kMaxStages + q); };
  auto tempty_bar = [&](int q) { return bar_base + 8u * (2 * kMaxStages + 2 + q); };
  const uint32_t ws_bar = bar_base + 8u * (2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_cls = p.m_tiles * p.n_tiles;
  const int total_tiles = p.ncls * tiles_per_cls;
  const int tiles_per_group = p.tiles_w * p.tiles_h;

  UDA_TR(const long long tr0 = clock64(); const long long tr_g0 = trace_globaltimer();
         long long* const trp = p.trace ? p.trace + (size_t)blockIdx.x * 16 : nullptr;)
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (int q = 0; q < 2; ++q) { mbar_init(tfull_bar(q), 1); mbar_init(tempty_bar(q), kEpiWarps); }
      mbar_init(ws_bar, 1);
      if (FUSE) mbar_init(bar_base + 8u * (2 * kMaxStages + 6), 1);   // residual rows landed (fused BatchNorm pass 2)
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
  UDA_TR(if (trp && threadIdx.x == 0) { trp[0] = tr0; trp[1] = clock64() - tr0; trp[14] = tr_g0;
                                        trp[15] = (FUSE ? 17LL : 1LL) | ((long long)p.Cout << 8) | ((long long)p.Cred << 24) | ((long long)total_tiles << 40); })

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      UDA_TR(long long tr_w = 0;)
      if (p.ws) {
        mbar_expect_tx(ws_bar, ws_bytes);
        for (int wt = 0; wt < p.wtaps; ++wt)
          for (int kc = 0; kc < p.kchunks; ++kc)
            tma_load_2d(ws_base + (wt * p.kchunks + kc) * kBBytes, &map_b, ws_bar, wt * p.Cred + kc * KC, 0);
      }
      // ring position kept as (slot, phase) counters: no integer division in the per-stage loop
      int s = 0; uint32_t phs = 0;
      uint32_t a_dst = ring_base;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int ci = t / tiles_per_cls, rem = t % tiles_per_cls;
        const int mt = rem % p.m_tiles, n0 = (rem / p.m_tiles) * BN;
        const int grp = mt / tiles_per_group, tin = mt % tiles_per_group;
        const int b0 = grp * p.NB, h0 = (tin / p.tiles_w) * p.TH, w0 = (tin % p.tiles_w) * p.TW;
        const PClass& c = p.cls[ci];
        for (int tap = 0; tap < c.ntaps; ++tap) {
          const int cw = c.pw[tap] * p.Cred, xw = w0 + c.dw[tap], xh = h0 + c.dh[tap], ph = c.ph[tap];
          const int bw = c.wtap[tap] * p.Cred;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            UDA_TR_WAIT(tr_w, mbar_wait(empty_bar(s), phs ^ 1))
            if (UDA_TC_DBG(p, 4)) {
              mbar_arrive(full_bar(s));
            } else {
              mbar_expect_tx(full_bar(s), stage_bytes);
              if (p.rank5)
                tma_load_5d(a_dst, &map_a, full_bar(s), cw + kc * KC, xw, ph, xh, b0);
              else
                tma_load_4d(a_dst, &map_a, full_bar(s), kc * KC, xw, xh, b0);
              if (!p.ws) tma_load_2d(a_dst + kABytes, &map_b, full_bar(s), bw + kc * KC, n0);
            }
            if (++s == S) { s = 0; phs ^= 1; a_dst = ring_base; } else { a_dst += stage_bytes; }
          }
        }
      }
      UDA_TR(if (trp) { trp[2] = tr_w; trp[3] = clock64() - tr0; })
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN < 16 ? 16 : BN);
      UDA_TR(long long tr_wf = 0, tr_we = 0, tr_first = 0;)
      if (p.ws) { UDA_TR_WAIT(tr_wf, mbar_wait(ws_bar, 0)) tc_fence_after(); }
      // The issuing thread is ONE in-order thread: at N = 128 a tcgen05.mma occupies the tensor pipe for 64 clocks and
      // the thread needs ~55 clocks to issue one, so every extra instruction per stage (integer divisions for the ring
      // index, descriptor construction, parameter loads) used to throttle the pipe (measured: ~100-140 clk per MMA,
      // tools/exp/umma_rate.cu, profiles/r02_trace_conv.txt).  Ring slot / phase are counters, descriptors are a
      // constant high word plus a low word that is only ever added to.
      const uint32_t dhi = kmajor_desc_hi(KC * 2);
      const uint32_t ring_lo = kmajor_desc_lo(ring_base), ws_lo = kmajor_desc_lo(ws_base);
      const uint32_t stage_step = (uint32_t)stage_bytes >> 4;
      int s = 0, j = 0; uint32_t phs = 0;
      uint32_t a_lo = ring_lo;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
        const int ci = t / tiles_per_cls;
        const PClass& c = p.cls[ci];
        const int q = j % kSets;
        UDA_TR_WAIT(tr_we, mbar_wait(tempty_bar(q), ((j / kSets) & 1) ^ 1))   // epilogue has drained this accumulator set
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)q * kAccCols;
        for (int tap = 0; tap < c.ntaps; ++tap) {
          const uint32_t wt_lo = ws_lo + (((uint32_t)c.wtap[tap] * p.kchunks * kBBytes) >> 4);
          for (int kc = 0; kc < p.kchunks; ++kc) {
            UDA_TR_WAIT(tr_wf, mbar_wait(full_bar(s), phs))
            UDA_TR(if (!tr_first) tr_first = clock64() - tr0;)
            tc_fence_after();
            const uint32_t b_lo = p.ws ? wt_lo + (uint32_t)kc * (kBBytes >> 4) : a_lo + (kABytes >> 4);
            if (!UDA_TC_DBG(p, 2)) {
#pragma unroll
              for (int sub = 0; sub < MT; ++sub) {
#pragma unroll
                for (int k = 0; k < KC / 16; ++k)
                  umma_bf16(acc + (uint32_t)sub * BN, desc64(a_lo + sub * ((128 * KC * 2) >> 4) + 2 * k, dhi),
                            desc64(b_lo + 2 * k, dhi), idesc, (tap > 0 || kc > 0 || k > 0) ? 1u : 0u);
              }
            }
            umma_commit(empty_bar(s));
            if (++s == S) { s = 0; phs ^= 1; a_lo = ring_lo; } else { a_lo += stage_step; }
          }
        }
        umma_commit(tfull_bar(q));
      }
      UDA_TR(if (trp) { trp[4] = tr_wf; trp[5] = tr_we; trp[6] = tr_first; trp[7] = clock64() - tr0; trp[12] = j; })
    }
  } else {
    // ===================== epilogue (kEpiWarps warps) =====================
    UDA_TR(long long tr_wt = 0, tr_busy = 0;)
    const int qw = warp & 3;
    const int eh = (warp - 2) >> 2;          // which of the kEpiSplit warps of this lane quadrant
    constexpr int kChunks = (BN + 31) / 32;
    constexpr bool kSplitCols = (kChunks % kEpiSplit) == 0;   // else (BN = 32) the quadrant's warps take alternate sub-tiles
    float bn_s[kChunks], bn_q[kChunks];
#pragma unroll
    for (int cc = 0; cc < kChunks; ++cc) { bn_s[cc] = 0.f; bn_q[cc] = 0.f; }
    // BN <= 32 (the HBM-bound few-channel layers): per-thread running sums over all rows this thread ever
    // owns, ONE cross-lane reduction at the very end of the CTA instead of 62 shuffles per 32x32 chunk
    constexpr bool kLate = (kChunks == 1);
    float late_s[kLate ? 32 : 1], late_q[kLate ? 32 : 1];
#pragma unroll
    for (int k = 0; k < (kLate ? 32 : 1); ++k) { late_s[k] = 0.f; late_q[k] = 0.f; }
    int bn_n0 = -1;
    double* const sums_out = p.bn_sums ? p.bn_sums : p.st_sums;   // forward statistics or BN-backward statistics
    const float inv_slope = p.st_slope != 0.f ? 1.f / p.st_slope : 0.f;
    int j = 0;
    if constexpr (FUSE) {
      // ---- conv + BatchNorm + activation in one launch (BnFuse): this CTA's (<= kSets) tiles stay in TMEM across
      // the grid barrier; one tap class, BN >= 64 ----
      float* const s_tab = reinterpret_cast<float*>(bars + 2 * kMaxStages + 8);   // [2][BN] scale | shift of the current tile
      const int et = threadIdx.x - 64;   // index among the epilogue threads
      // pass 1: statistics of the bf16-rounded outputs; the accumulator sets are NOT released
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
        const int mt = t % p.m_tiles, n0 = (t / p.m_tiles) * BN;
        const int grp = mt / tiles_per_group, tin = mt % tiles_per_group;
        const int b0 = grp * p.NB, h0 = (tin / p.tiles_w) * p.TH, w0 = (tin % p.tiles_w) * p.TW;
        if (n0 != bn_n0) {
          if (bn_n0 >= 0) {
#pragma unroll
            for (int cc = 0; cc < kChunks; ++cc) {
              if ((cc % kEpiSplit) != eh) continue;
              const int col = bn_n0 + cc * 32 + lane;
              atomicAdd(p.bn_sums + col, (double)bn_s[cc]); atomicAdd(p.bn_sums + p.Cout + col, (double)bn_q[cc]);
              bn_s[cc] = 0.f; bn_q[cc] = 0.f;
            }
          }
          bn_n0 = n0;
        }
        UDA_TR_WAIT(tr_wt, mbar_wait(tfull_bar(j), 0))
        tc_fence_after();
#pragma unroll 1
        for (int sub = 0; sub < MT; ++sub) {
          const int r = sub * 128 + qw * 32 + lane;
          const int nb = r / (p.TH * p.TW);
          const int th = (r / p.TW) % p.TH, tw = r % p.TW;
          const long long pix = ((long long)(b0 + nb) * p.OH + (h0 + th)) * p.OW + (w0 + tw);
          const uint32_t tbase = tmem_base + ((uint32_t)(qw * 32) << 16) + (uint32_t)j * kAccCols + (uint32_t)sub * BN;
#pragma unroll
          for (int c0 = 0; c0 < BN; c0 += 32) {
            if (((c0 / 32) % kEpiSplit) != eh) continue;
            uint32_t v[32];
            tmem_ld_32x32(tbase + (uint32_t)c0, v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) f[k] = __uint_as_float(v[k]);
            if (p.addend) {
              const bf16* add = p.addend + pix * p.Cout + n0 + c0;
#pragma unroll
              for (int k = 0; k < 32; k += 8) {
                float a8[8];
                ld_vec<8>(add + k, a8);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[k + e] += a8[e];
              }
            }
            bn_chunk_stats(f, lane, bn_s[c0 / 32], bn_q[c0 / 32]);
            {   // z is saved for the backward: stored here, under the remaining main loops of the grid
              bf16* dst = p.out + pix * p.Cout + n0 + c0;
#pragma unroll
              for (int k = 0; k < 32; k += 8) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = f[k + e];
                st_vec<8>(dst + k, o);
              }
            }
          }
        }
      }
      if (bn_n0 >= 0) {
#pragma unroll
        for (int cc = 0; cc < kChunks; ++cc) {
          if ((cc % kEpiSplit) != eh) continue;
          const int col = bn_n0 + cc * 32 + lane;
          atomicAdd(p.bn_sums + col, (double)bn_s[cc]); atomicAdd(p.bn_sums + p.Cout + col, (double)bn_q[cc]);
        }
      }
      UDA_TR(if (trp && warp == 2 && lane == 0) trp[5] = clock64() - tr0;)      // pass 1 done
      // Every MMA of this CTA has completed (the loop above waited for its last accumulator): the operand ring is free.
      // The residual rows of a tile arrive there through bulk copies (the first tile's are issued BEFORE the barrier),
      // pass 2 reads them conflict-free (row pitch + 16 bytes) and stages a = act(z*scale + shift (+ residual)) for one
      // bulk store per row — the register-file version of this pass (row-strided 16-byte global loads and stores from
      // 8 warps) took 8 us per launch without and 14-20 us with a residual (profiles/r02_trace_step.txt).
      using stream::bulk_load; using stream::bulk_store; using stream::bulk_commit; using stream::fence_async_smem;
      constexpr int kRowB = BN * 2, kPitch = kRowB + 16, kRows = MT * 128;
      static_assert(kRows <= kEpiThreads, "one epilogue thread per tile row");
      uint8_t* const st_r = smem + ws_bytes;               // [kRows][kPitch] residual rows of the current tile
      uint8_t* const st_a = st_r + kRows * kPitch;         // [kRows][kPitch] a of the current tile
      const uint32_t rbar = bar_base + 8u * (2 * kMaxStages + 6);
      const bf16* const res = (const bf16*)p.fuse.residual;
      bf16* const aout = (bf16*)p.fuse.a_out;
      const float slope = p.fuse.slope;
      const int n_mine = j;
      auto tile_pix0 = [&](int t) {          // a tile's 128*MT pixels are contiguous (plan_tiles)
        const int mt = t % p.m_tiles;
        const int grp = mt / tiles_per_group, tin = mt % tiles_per_group;
        const int b0 = grp * p.NB, h0 = (tin / p.tiles_w) * p.TH, w0 = (tin % p.tiles_w) * p.TW;
        return ((long long)b0 * p.OH + h0) * p.OW + w0;
      };
      auto load_res = [&](int t) {
        if (!res) return;
        if (et == 0) mbar_expect_tx(rbar, (uint32_t)kRows * kRowB);
        if (et < kRows)
          bulk_load(smem_u32(st_r + et * kPitch), res + (tile_pix0(t) + et) * p.Cout + (t / p.m_tiles) * BN, kRowB, rbar);
      };
      if (n_mine > 0) load_res(blockIdx.x);
      grid_barrier(p.fuse.counter, gridDim.x, 2, kEpiThreads, et == 0);
      UDA_TR(if (trp && warp == 2 && lane == 0) trp[13] = clock64() - tr0;)     // barrier passed
      j = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
        const int mt = t % p.m_tiles, n0 = (t / p.m_tiles) * BN;
        const long long pix0 = tile_pix0(t);
        float* const sc = s_tab;
        for (int ch = et; ch < BN; ch += kEpiThreads) {
          bn_fuse_coeffs(p.fuse, p.bn_sums, p.Cout, n0 + ch, sc[ch], sc[BN + ch]);
          if (mt == 0) bn_fuse_publish(p.fuse, p.bn_sums, p.Cout, n0 + ch);
        }
        if (j > 0 && et < kRows) stream::bulk_wait_read<0>();      // the previous tile's a rows have left st_a
        bar_sync(2, kEpiThreads);
        if (res) mbar_wait(rbar, (uint32_t)(j & 1));
#pragma unroll 1
        for (int sub = 0; sub < MT; ++sub) {
          const int r = sub * 128 + qw * 32 + lane;
          const uint32_t tbase = tmem_base + ((uint32_t)(qw * 32) << 16) + (uint32_t)j * kAccCols + (uint32_t)sub * BN;
#pragma unroll
          for (int c0 = 0; c0 < BN; c0 += 32) {
            if (((c0 / 32) % kEpiSplit) != eh) continue;
            uint32_t v[32];
            tmem_ld_32x32(tbase + (uint32_t)c0, v);
            tmem_ld_wait();
            const long long off = (pix0 + r) * p.Cout + n0 + c0;
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
              float z8[8], a8[8], r8[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) z8[e] = __uint_as_float(v[k + e]);
              if (p.addend) {
                ld_vec<8>(p.addend + off + k, a8);
#pragma unroll
                for (int e = 0; e < 8; ++e) z8[e] += a8[e];
              }
              if (res) stream::lds8(st_r + r * kPitch + (c0 + k) * 2, r8);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float zr = __bfloat162float(__float2bfloat16_rn(z8[e]));
                float y = zr * sc[c0 + k + e] + sc[BN + c0 + k + e];
                if (res) y += r8[e];
                a8[e] = y > 0.f ? y : y * slope;
              }
              stream::sts8(st_a + r * kPitch + (c0 + k) * 2, a8);
            }
          }
        }
        fence_async_smem();
        bar_sync(2, kEpiThreads);                 // a staged, residual rows and coefficient table consumed
        if (t + (int)gridDim.x < total_tiles) load_res(t + gridDim.x);
        if (et < kRows) {
          bulk_store(aout + (pix0 + et) * p.Cout + n0, smem_u32(st_a + et * kPitch), kRowB);
          bulk_commit();
        }
      }
      if (et < kRows) stream::bulk_wait_read<0>();   // shared memory must outlive the copies' reads
      UDA_TR(if (trp && warp == 2 && lane == 0) { trp[8] = tr_wt; trp[10] = clock64() - tr0; })
    } else {
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
      const int ci = t / tiles_per_cls, rem = t % tiles_per_cls;
      const int mt = rem % p.m_tiles, n0 = (rem / p.m_tiles) * BN;
      const int grp = mt / tiles_per_group, tin = mt % tiles_per_group;
      const int b0 = grp * p.NB, h0 = (tin / p.tiles_w) * p.TH, w0 = (tin % p.tiles_w) * p.TW;
      const int oh = p.cls[ci].oh, ow = p.cls[ci].ow;
      const int q = j % kSets;
      if (sums_out && n0 != bn_n0) {   // channel tile changed: flush the partial statistics
        if (bn_n0 >= 0) {
#pragma unroll
          for (int cc = 0; cc < kChunks; ++cc) {
            const int col = bn_n0 + cc * 32 + lane;
            if (col < p.Cout) { atomicAdd(sums_out + col, (double)bn_s[cc]); atomicAdd(sums_out + p.Cout + col, (double)bn_q[cc]); }
            bn_s[cc] = 0.f; bn_q[cc] = 0.f;
          }
        }
        bn_n0 = n0;
      }
      UDA_TR_WAIT(tr_wt, mbar_wait(tfull_bar(q), (j / kSets) & 1))
      UDA_TR(const long long tr_b0 = clock64();)
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < MT; ++sub) {
        if (!kSplitCols && (sub % kEpiSplit) != eh) continue;
        const int r = sub * 128 + qw * 32 + lane;
        const int nb = r / (p.TH * p.TW);
        const int th = (r / p.TW) % p.TH, tw = r % p.TW;
        const int b = b0 + nb, h = (h0 + th) * p.os + oh, w = (w0 + tw) * p.os + ow;
        const long long pix = ((long long)b * p.OH + h) * p.OW + w;
        const uint32_t tbase = tmem_base + ((uint32_t)(qw * 32) << 16) + (uint32_t)q * kAccCols + (uint32_t)sub * BN;
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          const int nbase = n0 + c0;
          if (nbase >= p.Cout) break;   // warp-uniform
          if (kSplitCols && ((c0 / 32) % kEpiSplit) != eh) continue;
          uint32_t v[32];
          tmem_ld_32x32(tbase + (uint32_t)c0, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) f[k] = __uint_as_float(v[k]);
          if (p.bias) {
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (nbase + k < p.Cout) f[k] += __ldg(p.bias + nbase + k);
          }
          if (p.addend) {
            const bf16* add = p.addend + pix * p.Cout + nbase;
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
              if (nbase + k < p.Cout) {
                float a8[8];
                ld_vec<8>(add + k, a8);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[k + e] += a8[e];
              }
            }
          }
          if (p.act) {
#pragma unroll
            for (int k = 0; k < 32; ++k) f[k] = f[k] > 0.f ? f[k] : f[k] * p.act_slope;
          }
          if (p.bn_sums) {
            if constexpr (kLate) {
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                const float r = __bfloat162float(__float2bfloat16_rn(f[k]));
                late_s[k] += r;
                late_q[k] = fmaf(r, r, late_q[k]);
              }
            } else {
              bn_chunk_stats(f, lane, bn_s[c0 / 32], bn_q[c0 / 32]);
            }
          } else if (!kLate && p.st_sums) {   // (the BN = 32 instances do not carry the BatchNorm-backward statistics)
            float g[32], gv[32];
            const long long off = pix * p.Cout + nbase;
            bn_bwd_chunk_terms(f, p.st_a + off, p.st_z ? p.st_z + off : nullptr, p.st_slope, inv_slope,
                               p.Cout - nbase, g, gv);
            if constexpr (kLate) {
#pragma unroll
              for (int k = 0; k < 32; ++k) { late_s[k] += g[k]; late_q[k] += gv[k]; }
            } else {
              bn_s[c0 / 32] += warp_column_sums(g, lane);
              bn_q[c0 / 32] += warp_column_sums(gv, lane);
            }
          }
          if (p.out && !UDA_TC_DBG(p, 1)) {
            bf16* dst = p.out + pix * p.Cout + nbase;
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
              if (nbase + k < p.Cout) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = f[k + e];
                st_vec<8>(dst + k, o);
              }
            }
          }
          if (p.out_nchw) {
            const long long hw = (long long)p.OH * p.OW;
            float* dst = p.out_nchw + ((long long)b * p.Cout + nbase) * hw + (long long)h * p.OW + w;
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (nbase + k < p.Cout) dst[(long long)k * hw] = f[k];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(q));
      UDA_TR(tr_busy += clock64() - tr_b0;)
    }
    UDA_TR(if (trp && warp == 2 && lane == 0) { trp[8] = tr_wt; trp[9] = tr_busy; trp[10] = clock64() - tr0; })
    if (sums_out && bn_n0 >= 0) {
      if constexpr (kLate) {   // BN = 32: a single channel tile, nothing was flushed before
        float ts[32], tq[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) { ts[k] = late_s[k]; tq[k] = late_q[k]; }
        bn_s[0] = warp_column_sums(ts, lane);
        bn_q[0] = warp_column_sums(tq, lane);
      }
#pragma unroll
      for (int cc = 0; cc < kChunks; ++cc) {
        const int col = bn_n0 + cc * 32 + lane;
        if (col < p.Cout) { atomicAdd(sums_out + col, (double)bn_s[cc]); atomicAdd(sums_out + p.Cout + col, (double)bn_q[cc]); }
      }
    }
    }   // !FUSE
  }
  tc_fence_before();
  __syncthreads();
  UDA_TR(if (trp && threadIdx.x == 0) trp[11] = clock64() - tr0;)
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int KC, int BN, int MT, bool FUSE = false>
int launch_persist(const CUtensorMap& ma, const CUtensorMap& mb, PParams& p, int total_tiles, cudaStream_t st) {
  constexpr int kABytes = MT * 128 * KC * 2, kBBytes = BN * KC * 2;
  constexpr int kSetsH = 2 * MT * BN <= 512 ? 2 : 1;
  const int budget = 200 * 1024;
  const int ws_bytes = p.wtaps * p.kchunks * kBBytes;
  p.ws = (p.n_tiles == 1 && ws_bytes <= 96 * 1024) ? 1 : 0;
  const int stage_bytes = kABytes + (p.ws ? 0 : kBBytes);
  int S = (budget - (p.ws ? ws_bytes : 0)) / stage_bytes;
  if (S > kMaxStages) S = kMaxStages;
  UDA_REQUIRE(S >= 2, UDA_ERR_UNSUPPORTED, "conv_tc_persist: not enough shared memory for a 2-stage ring");
  p.stages = S;
  const int smem = (p.ws ? ws_bytes : 0) + S * stage_bytes + 1024 + 512 + (FUSE ? 2 * BN * 4 + 64 : 0);
  static int configured = 0;
  if (configured < smem) {
    UDA_CUDA_OK(cudaFuncSetAttribute(conv_tc_persist_kernel<KC, BN, MT, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
    configured = 227 * 1024;
  }
  int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  // fused BatchNorm: all tiles of a CTA must still be in TMEM at the grid barrier, and pass 2 stages one tile's residual
  // and output rows in the (then idle) operand ring
  if (FUSE && ((total_tiles + grid - 1) / grid > kSetsH || S * stage_bytes < 2 * MT * 128 * (BN * 2 + 16)))
    return UDA_ERR_UNSUPPORTED;
  UDA_CUDA_OK(launch_pdl(conv_tc_persist_kernel<KC, BN, MT, FUSE>, dim3(grid), dim3(kThreads), smem, st, ma, mb, p));
  UDA_LAUNCH_OK("conv_tc_persist_kernel");
  return UDA_OK;
}

bool big_tiles_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UDA_B200_TC_BIGTILES");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

}  // namespace

int run_gemm_conv_persistent(const GemmConv& g, cudaStream_t st) {
  const int MH = g.a_map ? g.a_MH : (g.src_s2 ? g.SH / 2 : g.SH), MW = g.a_map ? g.a_MW : (g.src_s2 ? g.SW / 2 : g.SW);
  const int KC = g.a_map ? g.a_kc : pick_kc(g.Cred);
  int BN = pick_bn(g.Cout);
  UDA_REQUIRE(KC > 0 && g.ncls >= 1 && g.ncls <= kMaxClasses && g.Cout % 8 == 0, UDA_ERR_UNSUPPORTED,
              "conv_tc_persist: shape not covered (Cred=%d Cout=%d)", g.Cred, g.Cout);
  int n_tiles = (g.Cout + BN - 1) / BN;
  // 256-pixel tiles when that still leaves at least two tiles per SM, else 128-pixel tiles
  TilePlan tp = plan_tiles(g.B, MH, MW, 256);
  int MT = 2;
  if (g.a_map) tp.ok = false;   // caller-built maps use 128-pixel boxes
  const bool can256 = tp.ok;
  if (tp.ok) {
    const long long t2 = (long long)g.ncls * n_tiles * ((long long)g.B * MH * MW / 256);
    if (t2 < 2LL * num_sms()) tp.ok = false;
  }
  if (!tp.ok) { tp = plan_tiles(g.B, MH, MW, 128); MT = 1; }
  // Wide layers whose weights cannot stay resident (layer2-4, decoder conv1/conv2 of blocks 0-1) stream
  // (MT*128 + BN) operand rows per MT*128 x BN tile.  Measured on B200 (tools/conv_bench.py): whatever the tile
  // shape, an SM takes in only ~30 B/clk of operands (3-4 ring slots against the slot round-trip latency), so these layers are bound
  // by the operand bytes of the BUSIEST SM, not by the tensor pipe: pick the tile shape that minimises
  // rounds x bytes per tile (vs the tensor-pipe time), among 128/256 pixels x 128/256 channels.
  if (KC == 64 && g.Cout >= 128 && g.Cout % 128 == 0 && !g.a_map && big_tiles_enabled()) {
    long long taps = 0;
    for (int c = 0; c < g.ncls; ++c) taps += g.cls[c].ntaps;
    const double K = (double)taps * g.Cred / g.ncls;   // average reduction length of a tile
    const long long pixels = (long long)g.B * MH * MW;
    double best = 1e30;
    int best_mt = MT, best_bn = BN;
    for (int mt = 1; mt <= 2; ++mt) {
      if (mt == 2 && !can256) continue;
      for (int bn = 128; bn <= 256; bn += 128) {
        if (g.Cout % bn) continue;
        if (pixels % (128 * mt)) continue;
        const long long tiles = (pixels / (128 * mt)) * (g.Cout / bn) * g.ncls;
        const int sms = num_sms();
        const double rounds = (double)((tiles + sms - 1) / sms);
        const double t_tma = rounds * K * 2.0 * (128 * mt + bn) / 30.0;
        const double t_mma = rounds * (128.0 * mt) * bn * K / 4096.0;
        const double t_epi = (2 * mt * bn > 512 ? rounds : 1.0) * mt * bn * 12.0;   // exposed epilogue
        const double t = (t_tma > t_mma ? t_tma : t_mma) + t_epi + 1e-3 * mt * bn;  // ties: smaller tile
        if (t < best) { best = t; best_mt = mt; best_bn = bn; }
      }
    }
    // test hook: UDA_B200_TC_TILE="<MT>,<BN>" forces a tile shape (read on every call)
    if (const char* e = getenv("UDA_B200_TC_TILE")) {
      int fm = 0, fb = 0;
      if (sscanf(e, "%d,%d", &fm, &fb) == 2 && (fm == 1 || (fm == 2 && can256)) && (fb == 128 || fb == 256) &&
          g.Cout % fb == 0 && pixels % (128 * fm) == 0) {
        best_mt = fm; best_bn = fb;
      }
    }
    MT = best_mt; BN = best_bn;
    if (g.fuse) {
      // fused BatchNorm: every tile of the launch must still sit in TMEM at the grid barrier (<= kSets tiles per CTA);
      // keep the cost model's shape when it is resident, else the first resident one
      auto resident = [&](int mt, int bn) {
        if ((mt == 2 && !can256) || g.Cout % bn || pixels % (128 * mt)) return false;
        const long long tiles = (pixels / (128 * mt)) * (g.Cout / bn);
        const long long grid = tiles < num_sms() ? tiles : num_sms();
        const int sets = 2 * mt * bn <= 512 ? 2 : 1;
        return (tiles + grid - 1) / grid <= sets;
      };
      if (!resident(MT, BN)) {
        const int cand[4][2] = {{2, 128}, {1, 256}, {1, 128}, {2, 256}};
        bool found = false;
        for (int i = 0; i < 4 && !found; ++i)
          if (resident(cand[i][0], cand[i][1])) { MT = cand[i][0]; BN = cand[i][1]; found = true; }
        if (!found) return UDA_ERR_UNSUPPORTED;
      }
    }
    n_tiles = g.Cout / BN;
    tp = plan_tiles(g.B, MH, MW, 128 * MT);
  } else if (g.fuse) {
    return UDA_ERR_UNSUPPORTED;   // fused BatchNorm: 64-channel chunks and 128-multiple output channels only
  }
  UDA_REQUIRE(tp.ok, UDA_ERR_UNSUPPORTED, "conv_tc_persist: pixel grid %dx%dx%d cannot be tiled", g.B, MH, MW);
  UDA_REQUIRE(aligned<bf16>(g.src, 16) && aligned<bf16>(g.wmat, 16) && (!g.out || aligned<bf16>(g.out, 16)) &&
                  (!g.addend || aligned<bf16>(g.addend, 16)),
              UDA_ERR_BAD_ARG, "conv_tc_persist: pointers must be 16-byte aligned");
  PParams p{};
  p.TW = tp.TW; p.TH = tp.TH; p.NB = tp.NB; p.tiles_w = MW / tp.TW; p.tiles_h = MH / tp.TH;
  p.m_tiles = (g.B / tp.NB) * p.tiles_w * p.tiles_h; p.n_tiles = n_tiles; p.ncls = g.ncls;
  p.MH = MH; p.MW = MW; p.OH = g.OH; p.OW = g.OW; p.os = g.os;
  p.Cout = g.Cout; p.Cred = g.Cred; p.kchunks = (g.Cred + KC - 1) / KC; p.rank5 = (g.src_s2 || g.a_map) ? 1 : 0;
  p.wtaps = g.wtaps;
  for (int c = 0; c < g.ncls; ++c) {
    const TapClass& s = g.cls[c];
    UDA_REQUIRE(s.ntaps >= 1 && s.ntaps <= kMaxTaps, UDA_ERR_BAD_ARG, "conv_tc_persist: bad tap class");
    PClass& d = p.cls[c];
    d.ntaps = s.ntaps; d.oh = (short)s.oh; d.ow = (short)s.ow;
    for (int t = 0; t < s.ntaps; ++t) {
      d.dh[t] = (signed char)s.dh[t]; d.dw[t] = (signed char)s.dw[t];
      d.ph[t] = (signed char)s.ph[t]; d.pw[t] = (signed char)s.pw[t];
      d.wtap[t] = (unsigned char)s.wtap[t];
    }
  }
  p.out = (bf16*)g.out; p.out_nchw = g.out_nchw; p.bias = g.bias; p.addend = (const bf16*)g.addend;
  p.bn_sums = g.bn_sums;
  p.act = g.act; p.act_slope = g.act_slope;
  p.debug = 0;
  p.trace = nullptr;
  UDA_TR(p.trace = take_trace_slice();)
#ifdef UDA_B200_EXPERIMENTS
  {   // timing experiments that SKIP WORK: compiled only into experiment builds (make EXPERIMENTS=1)
    const char* e = getenv("UDA_B200_TC_DEBUG");
    p.debug = e ? atoi(e) : 0;
    static bool warned = false;
    if (p.debug && !warned) {
      fprintf(stderr, "uda_b200: UDA_B200_TC_DEBUG=%d is a TIMING EXPERIMENT (work is skipped): convolution results are WRONG\n", p.debug);
      warned = true;
    }
  }
#endif
  p.st_a = (const bf16*)g.st_a; p.st_z = (const bf16*)g.st_z; p.st_slope = g.st_slope; p.st_sums = g.st_sums;
  UDA_REQUIRE(!(g.bn_sums && g.st_sums), UDA_ERR_BAD_ARG, "conv_tc_persist: forward and backward statistics are exclusive");
  UDA_REQUIRE(!g.st_sums || (g.st_a && g.out), UDA_ERR_BAD_ARG, "conv_tc_persist: backward statistics need `a` and an NHWC output");
  UDA_REQUIRE(!g.st_sums || BN >= 64, UDA_ERR_UNSUPPORTED, "conv_tc_persist: backward statistics need more than 32 output channels");

  CUtensorMap ma, mb;
  const uint64_t C = (uint64_t)g.Cred, H = (uint64_t)g.SH, W = (uint64_t)g.SW;
  if (g.a_map) {
    ma = *g.a_map;
  } else if (!g.src_s2) {
    uint64_t dims[4] = {C, W, H, (uint64_t)g.B};
    uint64_t str[3] = {C * 2, W * C * 2, H * W * C * 2};
    uint32_t box[4] = {(uint32_t)KC, (uint32_t)tp.TW, (uint32_t)tp.TH, (uint32_t)tp.NB};
    if (int rc = make_tmap_bf16(&ma, g.src, 4, dims, str, box, KC * 2)) return rc;
  } else {
    uint64_t dims[5] = {2 * C, W / 2, 2, H / 2, (uint64_t)g.B};
    uint64_t str[4] = {2 * C * 2, W * C * 2, 2 * W * C * 2, H * W * C * 2};
    uint32_t box[5] = {(uint32_t)KC, (uint32_t)tp.TW, 1, (uint32_t)tp.TH, (uint32_t)tp.NB};
    if (int rc = make_tmap_bf16(&ma, g.src, 5, dims, str, box, KC * 2)) return rc;
  }
  {
    const uint64_t Kt = (uint64_t)g.wtaps * g.Cred;
    uint64_t dims[2] = {Kt, (uint64_t)g.Cout};
    uint64_t str[1] = {Kt * 2};
    uint32_t box[2] = {(uint32_t)KC, (uint32_t)BN};
    if (int rc = make_tmap_bf16(&mb, g.wmat, 2, dims, str, box, KC * 2)) return rc;
  }
  const int total_tiles = p.ncls * p.m_tiles * p.n_tiles;
  if (g.fuse) {
    if (!g.bn_sums || !g.out || g.out_nchw || g.bias || g.act || g.st_sums || g.ncls != 1 || g.os != 1 || KC != 64)
      return UDA_ERR_UNSUPPORTED;
    p.fuse = *g.fuse;
    if (BN == 128) return MT == 2 ? launch_persist<64, 128, 2, true>(ma, mb, p, total_tiles, st)
                                  : launch_persist<64, 128, 1, true>(ma, mb, p, total_tiles, st);
    if (BN == 256) return MT == 2 ? launch_persist<64, 256, 2, true>(ma, mb, p, total_tiles, st)
                                  : launch_persist<64, 256, 1, true>(ma, mb, p, total_tiles, st);
    return UDA_ERR_UNSUPPORTED;
  }
#define UDA_P(KCv, BNv)                                                                            \
  if (KC == KCv && BN == BNv)                                                                      \
    return MT == 2 ? launch_persist<KCv, BNv, 2>(ma, mb, p, total_tiles, st)                       \
                   : launch_persist<KCv, BNv, 1>(ma, mb, p, total_tiles, st);
  UDA_P(64, 256) UDA_P(64, 128) UDA_P(64, 64) UDA_P(64, 32)
  UDA_P(32, 128) UDA_P(32, 64) UDA_P(32, 32)
  UDA_P(16, 128) UDA_P(16, 64) UDA_P(16, 32)
#undef UDA_P
  return set_error(UDA_ERR_UNSUPPORTED, "conv_tc_persist: no kernel instance for KC=%d BN=%d", KC, BN);
}

}  // namespace tcconv
}  // namespace uda
