// Halo-tile tcgen05 convolution for 3x3 stride-1 "same" convolutions on wide images (W % 128 == 0), sm_100a.
//
// conv_tc_persist.cu loads every input pixel nine times (one TMA box per tap).  For the high-resolution,
// few-channel layers (decoder blocks 2-4, head, layer1 and their dgrads) that makes the kernel bound by
// the TMA row rate / L2->SM traffic instead of HBM.  Here a tile is R image rows x 128 pixels and its
// (R+2) x 130 pixel halo is loaded ONCE per channel chunk with a single TMA box (zero-filled at the image
// border); the nine taps are nine tcgen05.mma descriptors that start at shifted addresses inside that
// halo buffer — (sub+kh) rows and kw pixels further — which is legal because the hardware applies the
// shared-memory swizzle to absolute address bits (verified on B200: tools/exp/halo_desc_test.cu).
// Each image row of 128 pixels is one M=128 accumulator; all weights stay resident in shared memory;
// two accumulator sets in TMEM overlap the epilogue of tile j with the MMAs of tile j+1.
#include "conv_tc_internal.cuh"
#include <stdlib.h>

namespace uda {
namespace tcconv {
namespace {

using namespace tc;

constexpr int kThreads = kConvThreads;
constexpr int kHaloW = 130;   // 128 pixels + one halo pixel on each side
constexpr int kSmemBudget = 222 * 1024;

struct HParams {
  int H, W, B, tiles_w, tiles_h;     // image size; tiles per image (W/128, H/R)
  int total_tiles;
  int Cout, Cred, kchunks;
  int stages;
  bf16* out; float* out_nchw; const float* bias; const bf16* addend;
  double* bn_sums;
  int act; float act_slope;            // GemmConv::act
  const bf16* st_a; const bf16* st_z; float st_slope; double* st_sums;   // GemmConv::st_*
  long long* trace;   // experiment builds only: per-CTA trace records (conv_tc_internal.cuh)
  int debug;          // experiment builds only (UDA_B200_TC_DEBUG bit mask): 1 = no epilogue stores, 8 = no statistics
};
#ifdef UDA_B200_EXPERIMENTS
#define UDA_H_DBG(p, bit) ((p).debug & (bit))
#else
#define UDA_H_DBG(p, bit) false
#endif

// NP ("N-packed", BN = 32 instances): at N <= 64 a tcgen05.mma holds the issue slot for a flat ~55 clocks whatever N is,
// and the 16- / 32-channel layers were bound by exactly that — 72 instructions per 1024-pixel tile (experiment build:
// dec4.c2 still takes 110k clocks per CTA with the stores AND the statistics switched off = 2016 instructions x 55).
// With NP the N dimension also carries the kernel row: for ONE halo row rx of X and one kw, the weight rows of
// kh = 2, 1, 0 are three N atoms (three consecutive 32-row weight boxes in shared memory) that accumulate into the
// three consecutive output rows rx-2, rx-1, rx of the TMEM accumulator — (R+2) x 3 instructions per channel chunk and
// tile instead of R x 9 (dec4.c2: 40 instead of 72).  Output row rx is written for the first time by halo row rx, so
// that instruction is issued apart from the two rows that already hold data.
template <int KC, int BN, int R, bool NP>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const HParams p) {
  constexpr int kRowB = KC * 2;
  constexpr int kHaloBytes = (R + 2) * kHaloW * kRowB;
  constexpr int kHaloStride = (kHaloBytes + 1023) / 1024 * 1024;
  constexpr int kBBytes = BN * KC * 2;
  constexpr uint32_t kAccCols = R * BN;
  constexpr uint32_t kTmemCols = 2 * kAccCols < 32 ? 32 : 2 * kAccCols;
  static_assert(2 * R * BN <= 512, "TMEM holds two accumulator sets of R x BN columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  const int ws_bytes = 9 * p.kchunks * kBBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ws_bytes + S * kHaloStride);
  // bars: full[4], empty[4], tmem_full[2], tmem_empty[2], ws_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  const uint32_t ws_base = smem_u32(smem);
  const uint32_t ring_base = ws_base + ws_bytes;
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  auto tfull_bar = [&](int q) { return bar_base + 8u * (8 + q); };
  auto tempty_bar = [&](int q) { return bar_base + 8u * (10 + q); };
  const uint32_t ws_bar = bar_base + 8u * 12;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_w * p.tiles_h;

  UDA_TR(const long long tr0 = clock64(); const long long tr_g0 = trace_globaltimer();
         long long* const trp = p.trace ? p.trace + (size_t)blockIdx.x * 16 : nullptr;)
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      for (int q = 0; q < 2; ++q) { mbar_init(tfull_bar(q), 1); mbar_init(tempty_bar(q), kEpiWarps); }
      mbar_init(ws_bar, 1);
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
                                        trp[15] = 2LL | ((long long)p.Cout << 8) | ((long long)p.Cred << 24) | ((long long)p.total_tiles << 40); })

  if (warp == 0) {
    // ===================== TMA producer: weights once, then one halo box per (tile, channel chunk) ==========
    if (elect_one()) {
      UDA_TR(long long tr_w = 0;)
      mbar_expect_tx(ws_bar, ws_bytes);
      if constexpr (NP) {      // [kw][channel chunk][kh = 2, 1, 0]: the three kernel rows of a (kw, chunk) are adjacent N atoms
        for (int kw = 0; kw < 3; ++kw)
          for (int kc = 0; kc < p.kchunks; ++kc)
            for (int jj = 0; jj < 3; ++jj)
              tma_load_2d(ws_base + ((kw * p.kchunks + kc) * 3 + jj) * kBBytes, &map_b, ws_bar,
                          ((2 - jj) * 3 + kw) * p.Cred + kc * KC, 0);
      } else {
        for (int t = 0; t < 9; ++t)
          for (int kc = 0; kc < p.kchunks; ++kc)
            tma_load_2d(ws_base + (t * p.kchunks + kc) * kBBytes, &map_b, ws_bar, t * p.Cred + kc * KC, 0);
      }
      int it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int b = t / tiles_per_img, tin = t % tiles_per_img;
        const int h0 = (tin / p.tiles_w) * R, w0 = (tin % p.tiles_w) * 128;
        for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
          const int s = it % S;
          UDA_TR_WAIT(tr_w, mbar_wait(empty_bar(s), ((it / S) & 1) ^ 1))
          mbar_expect_tx(full_bar(s), kHaloBytes);
          tma_load_4d(ring_base + s * kHaloStride, &map_a, full_bar(s), kc * KC, w0 - 1, h0 - 1, b);
        }
      }
      UDA_TR(if (trp) { trp[2] = tr_w; trp[3] = clock64() - tr0; })
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN < 16 ? 16 : BN);
      UDA_TR(long long tr_wf = 0, tr_we = 0, tr_first = 0;)
      UDA_TR_WAIT(tr_wf, mbar_wait(ws_bar, 0))
      tc_fence_after();
      int it = 0, j = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++j) {
        const int q = j & 1;
        UDA_TR_WAIT(tr_we, mbar_wait(tempty_bar(q), ((j >> 1) & 1) ^ 1))
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)q * kAccCols;
        for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
          const int s = it % S;
          UDA_TR_WAIT(tr_wf, mbar_wait(full_bar(s), (it / S) & 1))
          UDA_TR(if (!tr_first) tr_first = clock64() - tr0;)
          tc_fence_after();
          const uint32_t halo = ring_base + s * kHaloStride;
          if constexpr (NP) {
            constexpr uint32_t id1 = make_idesc_bf16(128, BN), id2 = make_idesc_bf16(128, 2 * BN), id3 = make_idesc_bf16(128, 3 * BN);
#pragma unroll 1
            for (int rx = 0; rx < R + 2; ++rx) {
              // N atom j <-> output row rx-2+j (kh = 2-j); only rows inside the tile
              const int j_lo = rx < 2 ? 2 - rx : 0, j_hi = (R + 1 - rx) < 2 ? (R + 1 - rx) : 2;
              const bool fresh = kc == 0 && rx < R;        // output row rx has not been written in this tile yet
              const uint32_t d0 = acc + (uint32_t)(rx - 2 + j_lo) * BN;
              const int n = j_hi - j_lo + 1;
              const uint32_t idn = n == 3 ? id3 : (n == 2 ? id2 : id1);
#pragma unroll
              for (int kw = 0; kw < 3; ++kw) {
                const uint64_t adesc = make_kmajor_desc(halo + (rx * kHaloW + kw) * kRowB, kRowB);
                const uint32_t wrow = ws_base + ((kw * p.kchunks + kc) * 3) * kBBytes;
                const uint64_t bdesc = make_kmajor_desc(wrow + j_lo * kBBytes, kRowB);
#pragma unroll
                for (int k = 0; k < KC / 16; ++k) {
                  if (fresh && kw == 0 && k == 0) {
                    if (j_lo < 2) umma_bf16(d0, adesc, bdesc, j_lo == 0 ? id2 : id1, 1u);              // rows rx-2 / rx-1
                    umma_bf16(acc + (uint32_t)rx * BN, adesc, make_kmajor_desc(wrow + 2 * kBBytes, kRowB), id1, 0u);   // row rx
                  } else {
                    umma_bf16(d0, adesc + 2ull * k, bdesc + 2ull * k, idn, 1u);
                  }
                }
              }
            }
          } else {
#pragma unroll 1
          for (int sub = 0; sub < R; ++sub) {
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const int kh = tap / 3, kw = tap % 3;
              // window of image row (sub) for this tap: starts (sub+kh) halo rows down, kw pixels right
              const uint64_t adesc = make_kmajor_desc(halo + ((sub + kh) * kHaloW + kw) * kRowB, kRowB);
              const uint64_t bdesc = make_kmajor_desc(ws_base + (tap * p.kchunks + kc) * kBBytes, kRowB);
#pragma unroll
              for (int k = 0; k < KC / 16; ++k)
                umma_bf16(acc + (uint32_t)sub * BN, adesc + 2ull * k, bdesc + 2ull * k, idesc,
                          (kc > 0 || tap > 0 || k > 0) ? 1u : 0u);
            }
          }
          }   // !NP
          umma_commit(empty_bar(s));
        }
        umma_commit(tfull_bar(q));
      }
      UDA_TR(if (trp) { trp[4] = tr_wf; trp[5] = tr_we; trp[6] = tr_first; trp[7] = clock64() - tr0; trp[12] = j; })
    }
  } else {
    // ===================== epilogue (kEpiWarps warps): one image row of 128 pixels per sub-tile ===============
    UDA_TR(long long tr_wt = 0, tr_busy = 0;)
    const int qw = warp & 3;
    const int eh = (warp - 2) >> 2;          // which of the kEpiSplit warps of this lane quadrant
    constexpr int kChunks = (BN + 31) / 32;
    constexpr bool kSplitCols = (kChunks % kEpiSplit) == 0;   // else the quadrant's warps take alternate rows
    float bn_s[kChunks], bn_q[kChunks];
#pragma unroll
    for (int cc = 0; cc < kChunks; ++cc) { bn_s[cc] = 0.f; bn_q[cc] = 0.f; }
    // BN <= 32 (the HBM-bound few-channel layers): per-thread running sums over all rows this thread ever
    // owns, ONE cross-lane reduction at the very end of the CTA instead of 62 shuffles per 32x32 chunk
    constexpr bool kLate = (kChunks == 1);
    float late_s[kLate ? 32 : 1], late_q[kLate ? 32 : 1];
#pragma unroll
    for (int k = 0; k < (kLate ? 32 : 1); ++k) { late_s[k] = 0.f; late_q[k] = 0.f; }
    double* const sums_out = p.bn_sums ? p.bn_sums : p.st_sums;   // forward statistics or BN-backward statistics
    const float inv_slope = p.st_slope != 0.f ? 1.f / p.st_slope : 0.f;
    int j = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++j) {
      const int b = t / tiles_per_img, tin = t % tiles_per_img;
      const int h0 = (tin / p.tiles_w) * R, w0 = (tin % p.tiles_w) * 128;
      const int q = j & 1;
      UDA_TR_WAIT(tr_wt, mbar_wait(tfull_bar(q), (j >> 1) & 1))
      UDA_TR(const long long tr_b0 = clock64();)
      tc_fence_after();
      const int w = w0 + qw * 32 + lane;
#pragma unroll 1
      for (int sub = 0; sub < R; ++sub) {
        if (!kSplitCols && (sub % kEpiSplit) != eh) continue;
        const int h = h0 + sub;
        const long long pix = ((long long)b * p.H + h) * p.W + w;
        const uint32_t tbase = tmem_base + ((uint32_t)(qw * 32) << 16) + (uint32_t)q * kAccCols + (uint32_t)sub * BN;
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          if (c0 >= p.Cout) break;
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
              if (c0 + k < p.Cout) f[k] += __ldg(p.bias + c0 + k);
          }
          if (p.addend) {
            const bf16* add = p.addend + pix * p.Cout + c0;
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
              if (c0 + k < p.Cout) {
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
          if (p.bn_sums && !UDA_H_DBG(p, 8)) {
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
            const long long off = pix * p.Cout + c0;
            bn_bwd_chunk_terms(f, p.st_a + off, p.st_z ? p.st_z + off : nullptr, p.st_slope, inv_slope, p.Cout - c0, g, gv);
            if constexpr (kLate) {
#pragma unroll
              for (int k = 0; k < 32; ++k) { late_s[k] += g[k]; late_q[k] += gv[k]; }
            } else {
              bn_s[c0 / 32] += warp_column_sums(g, lane);
              bn_q[c0 / 32] += warp_column_sums(gv, lane);
            }
          }
          if (p.out && !UDA_H_DBG(p, 1)) {
            bf16* dst = p.out + pix * p.Cout + c0;
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
              if (c0 + k < p.Cout) {
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = f[k + e];
                st_vec<8>(dst + k, o);
              }
            }
          }
          if (p.out_nchw && !UDA_H_DBG(p, 1)) {
            const long long hw = (long long)p.H * p.W;
            float* dst = p.out_nchw + ((long long)b * p.Cout + c0) * hw + (long long)h * p.W + w;
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (c0 + k < p.Cout) dst[(long long)k * hw] = f[k];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(q));
      UDA_TR(tr_busy += clock64() - tr_b0;)
    }
    UDA_TR(if (trp && warp == 2 && lane == 0) { trp[8] = tr_wt; trp[9] = tr_busy; trp[10] = clock64() - tr0; })
    if (sums_out) {
      if constexpr (kLate) {
        float ts[32], tq[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) { ts[k] = late_s[k]; tq[k] = late_q[k]; }
        bn_s[0] = warp_column_sums(ts, lane);
        bn_q[0] = warp_column_sums(tq, lane);
      }
#pragma unroll
      for (int cc = 0; cc < kChunks; ++cc) {
        const int col = cc * 32 + lane;
        if (col < p.Cout) { atomicAdd(sums_out + col, (double)bn_s[cc]); atomicAdd(sums_out + p.Cout + col, (double)bn_q[cc]); }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  UDA_TR(if (trp && threadIdx.x == 0) trp[11] = clock64() - tr0;)
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int KC, int BN, int R, bool NP = false>
int launch_halo(const CUtensorMap& ma, const CUtensorMap& mb, HParams& p, cudaStream_t st) {
  constexpr int kHaloBytes = (R + 2) * kHaloW * KC * 2;
  constexpr int kHaloStride = (kHaloBytes + 1023) / 1024 * 1024;
  const int ws_bytes = 9 * p.kchunks * BN * KC * 2;
  int S = (kSmemBudget - ws_bytes) / kHaloStride;
  if (S > 4) S = 4;
  if (S < 2) return UDA_ERR_UNSUPPORTED;   // caller falls back
  p.stages = S;
  const int smem = ws_bytes + S * kHaloStride + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    UDA_CUDA_OK(cudaFuncSetAttribute(conv_tc_halo_kernel<KC, BN, R, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
    configured = true;
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  UDA_CUDA_OK(launch_pdl(conv_tc_halo_kernel<KC, BN, R, NP>, dim3(grid), dim3(kThreads), smem, st, ma, mb, p));
  UDA_LAUNCH_OK("conv_tc_halo_kernel");
  return UDA_OK;
}

}  // namespace

// Returns UDA_ERR_UNSUPPORTED (without setting an error message the caller would surface) when the shape is
// not a 3x3 stride-1 "same" convolution on a 128-multiple-wide image with resident weights.
int run_gemm_conv_halo(const GemmConv& g, cudaStream_t st) {
  if (g.fuse) return UDA_ERR_UNSUPPORTED;   // wide images: the tile set does not fit TMEM, BatchNorm stays a separate pass
  if (g.ncls != 1 || g.src_s2 || g.a_map || g.os != 1 || g.wtaps != 9 || g.cls[0].ntaps != 9) return UDA_ERR_UNSUPPORTED;
  const TapClass& c = g.cls[0];
  if (c.oh != 0 || c.ow != 0) return UDA_ERR_UNSUPPORTED;
  for (int t = 0; t < 9; ++t)
    if (c.dh[t] != t / 3 - 1 || c.dw[t] != t % 3 - 1 || c.wtap[t] != t) return UDA_ERR_UNSUPPORTED;
  const int H = g.SH, W = g.SW;
  if (W % 128 || g.OH != H || g.OW != W || g.Cout > 128 || g.Cout % 8) return UDA_ERR_UNSUPPORTED;
  const int KC = pick_kc(g.Cred), BN = pick_bn(g.Cout);
  if (KC == 0) return UDA_ERR_UNSUPPORTED;
  const int kchunks = (g.Cred + KC - 1) / KC;
  const int ws_bytes = 9 * kchunks * BN * KC * 2;
  if (ws_bytes > 100 * 1024) return UDA_ERR_UNSUPPORTED;
  // rows per tile: as many as TMEM (2 sets x R x BN <= 512 columns) and shared memory (>= 2 halo stages) allow
  int R = 256 / BN;
  if (R > 8) R = 8;
  while (R > 1 && (H % R || (kSmemBudget - ws_bytes) / (((R + 2) * kHaloW * KC * 2 + 1023) / 1024 * 1024) < 2)) R >>= 1;
  if (R < 2) return UDA_ERR_UNSUPPORTED;
  if (!(aligned<bf16>(g.src, 16) && aligned<bf16>(g.wmat, 16) && (!g.out || aligned<bf16>(g.out, 16)) &&
        (!g.addend || aligned<bf16>(g.addend, 16))))
    return UDA_ERR_UNSUPPORTED;
  HParams p{};
  p.H = H; p.W = W; p.B = g.B; p.tiles_w = W / 128; p.tiles_h = H / R;
  p.total_tiles = g.B * p.tiles_w * p.tiles_h;
  p.Cout = g.Cout; p.Cred = g.Cred; p.kchunks = kchunks;
  p.out = (bf16*)g.out; p.out_nchw = g.out_nchw; p.bias = g.bias; p.addend = (const bf16*)g.addend;
  p.bn_sums = g.bn_sums;
  p.act = g.act; p.act_slope = g.act_slope;
  p.st_a = (const bf16*)g.st_a; p.st_z = (const bf16*)g.st_z; p.st_slope = g.st_slope; p.st_sums = g.st_sums;
  if (g.st_sums && (!g.st_a || !g.out || g.bn_sums || BN < 64)) return UDA_ERR_UNSUPPORTED;
  UDA_TR(p.trace = take_trace_slice();)
#ifdef UDA_B200_EXPERIMENTS
  { const char* e = getenv("UDA_B200_TC_DEBUG"); p.debug = e ? atoi(e) : 0; }   // timing experiments that SKIP WORK
#endif
  CUtensorMap ma, mb;
  {
    const uint64_t C = (uint64_t)g.Cred;
    uint64_t dims[4] = {C, (uint64_t)W, (uint64_t)H, (uint64_t)g.B};
    uint64_t str[3] = {C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {(uint32_t)KC, (uint32_t)kHaloW, (uint32_t)(R + 2), 1};
    if (int rc = make_tmap_bf16(&ma, g.src, 4, dims, str, box, KC * 2)) return rc;
  }
  {
    const uint64_t Kt = (uint64_t)9 * g.Cred;
    uint64_t dims[2] = {Kt, (uint64_t)g.Cout};
    uint64_t str[1] = {Kt * 2};
    uint32_t box[2] = {(uint32_t)KC, (uint32_t)BN};
    if (int rc = make_tmap_bf16(&mb, g.wmat, 2, dims, str, box, KC * 2)) return rc;
  }
  // UDA_B200_HALO_NPACK=0 keeps one instruction group per tap for the BN = 32 instances (read on every call: A/B, tests)
  const char* const env_np = getenv("UDA_B200_HALO_NPACK");
  const bool npack = BN == 32 && !(env_np && env_np[0] == '0');
#define UDA_H(KCv, BNv, Rv)                                                                  \
  if (KC == KCv && BN == BNv && R == Rv) {                                                   \
    if constexpr (BNv == 32) { if (npack) return launch_halo<KCv, BNv, Rv, true>(ma, mb, p, st); } \
    return launch_halo<KCv, BNv, Rv>(ma, mb, p, st);                                         \
  }
#define UDA_HK(KCv) \
  UDA_H(KCv, 32, 8) UDA_H(KCv, 32, 4) UDA_H(KCv, 32, 2) UDA_H(KCv, 64, 4) UDA_H(KCv, 64, 2) UDA_H(KCv, 128, 2)
  UDA_HK(16) UDA_HK(32) UDA_HK(64)
#undef UDA_HK
#undef UDA_H
  return UDA_ERR_UNSUPPORTED;
}

}  // namespace tcconv
}  // namespace uda
