// Pitched-halo tcgen05 convolution for 3x3 stride-1 "same" convolutions on NARROW images (W <= 64) with many
// channels (layer2 / layer3, decoder blocks 0-1, and their dgrads), sm_100a.
//
// conv_tc_persist.cu loads every input pixel nine times (one TMA box per tap) and, for these layers, is bound by
// the operand bytes an SM can take in (~30 B/clk with 3-4 ring slots against the slot latency), not by the tensor pipe.
// conv_tc_halo.cu removes the nine-fold re-read for wide images, where one image row is one 128-row MMA block.
// Here the same trick works for narrow images by making the GEMM-M index a PITCHED position: an image of H x W
// pixels is walked as H rows of P = W + 2 positions (the two extra positions per row are junk outputs that are
// never written), 128 consecutive positions = one MMA block.  The halo of a block — the image rows it touches
// plus one above / below, each with one extra pixel left / right — is ONE zero-filled TMA box {KC, P, HR, 1};
// in it, tap (kh, kw) of position m sits exactly (kh*P + kw) rows after the block's first row, so the nine taps
// are nine MMA descriptors at shifted start addresses inside the same shared-memory tile (the swizzle is a
// function of absolute shared-memory address bits: tools/exp/halo_desc_test.cu).  A-operand rows per tile drop
// ~5x (W = 32) / ~3.5x (W = 64); the weight tiles are streamed per (channel chunk, tap) through a second ring.
// A CTA tile is NBLK consecutive blocks (of the whole batch: blocks never straddle images, tiles may) x BN
// output channels; TMEM is double-buffered when 2*NBLK*BN <= 512 columns.
#include "conv_tc_internal.cuh"
#include "stream_common.cuh"
#include <stdlib.h>

namespace uda {
namespace tcconv {
namespace {

using namespace tc;

constexpr int kThreads = kConvThreads;
constexpr int kSmemBudget = 218 * 1024;
constexpr int kAStages = 2;
constexpr int kMaxBStages = 8;

struct PHParams {
  int H, W, P, B;
  int nb_img, total_blocks, m_tiles, n_tiles;
  int Cout, Cred, kchunks;
  int HR, halo_bytes, halo_stride, b_stages;
  bf16* out; const bf16* addend; double* bn_sums;
  const float* bias; int act; float act_slope;   // GemmConv::bias / act
  BnFuse fuse;        // fuse.a_out != nullptr: BatchNorm + activation in this launch (FUSE instances, one tile per CTA)
  long long* trace;   // experiment builds only: per-CTA trace records (conv_tc_internal.cuh)
};

template <int KC, int BN, int NBLK, bool FUSE>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_phalo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const PHParams p) {
  constexpr int kRowB = KC * 2;
  constexpr int kBBytes = BN * KC * 2;
  constexpr uint32_t kAccCols = NBLK * BN;
  constexpr int kSets = 2 * kAccCols <= 512 ? 2 : 1;
  constexpr uint32_t kTmemCols = kSets * kAccCols < 32 ? 32 : kSets * kAccCols;
  static_assert(kAccCols <= 512, "accumulators exceed TMEM");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int SB = p.b_stages;
  const int a_stage_bytes = NBLK * p.halo_stride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAStages * a_stage_bytes + SB * kBBytes);
  // bars: a_full[2], a_empty[2], b_full[8], b_empty[8], tmem_full[2], tmem_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  const uint32_t a_base = smem_u32(smem);
  const uint32_t b_base = a_base + kAStages * a_stage_bytes;
  const uint32_t bar_base = smem_u32(bars);
  auto afull = [&](int s) { return bar_base + 8u * s; };
  auto aempty = [&](int s) { return bar_base + 8u * (2 + s); };
  auto bfull = [&](int s) { return bar_base + 8u * (4 + s); };
  auto bempty = [&](int s) { return bar_base + 8u * (12 + s); };
  auto tfull = [&](int q) { return bar_base + 8u * (20 + q); };
  auto tempty = [&](int q) { return bar_base + 8u * (22 + q); };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.m_tiles * p.n_tiles;

  UDA_TR(const long long tr0 = clock64(); const long long tr_g0 = trace_globaltimer();
         long long* const trp = p.trace ? p.trace + (size_t)blockIdx.x * 16 : nullptr;)
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kAStages; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
      for (int s = 0; s < SB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
      for (int q = 0; q < 2; ++q) { mbar_init(tfull(q), 1); mbar_init(tempty(q), kEpiWarps); }
      if (FUSE) mbar_init(bar_base + 8u * 25, 1);     // residual rows landed (fused BatchNorm epilogue)
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
                                        trp[15] = (FUSE ? 19LL : 3LL) | ((long long)p.Cout << 8) | ((long long)p.Cred << 24) | ((long long)total_tiles << 40); })

  if (warp == 0) {
    // ===================== TMA producer: per channel chunk one halo box per block, then nine weight tiles =====
    if (elect_one()) {
      UDA_TR(long long tr_w = 0;)
      // ring positions are (slot, phase) counters: no integer division in the per-stage loops
      int sa = 0, sb = 0; uint32_t pha = 0, phb = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int mt = t % p.m_tiles, n0 = (t / p.m_tiles) * BN;
        const int gb0 = mt * NBLK;
        const int nblk = min(NBLK, p.total_blocks - gb0);
        int hb[NBLK], hrow[NBLK];   // image / first halo row of each block
#pragma unroll
        for (int j = 0; j < NBLK; ++j) {
          const int gb = gb0 + j, b = gb / p.nb_img, m0 = (gb - b * p.nb_img) * 128;
          hb[j] = b; hrow[j] = m0 / p.P - 1;
        }
        for (int kc = 0; kc < p.kchunks; ++kc) {
          UDA_TR_WAIT(tr_w, mbar_wait(aempty(sa), pha ^ 1))
          mbar_expect_tx(afull(sa), (uint32_t)nblk * p.halo_bytes);
#pragma unroll
          for (int j = 0; j < NBLK; ++j)
            if (j < nblk)
              tma_load_4d(a_base + sa * a_stage_bytes + j * p.halo_stride, &map_a, afull(sa), kc * KC, -1, hrow[j], hb[j]);
          if (++sa == kAStages) { sa = 0; pha ^= 1; }
          for (int tap = 0; tap < 9; ++tap) {
            UDA_TR_WAIT(tr_w, mbar_wait(bempty(sb), phb ^ 1))
            mbar_expect_tx(bfull(sb), kBBytes);
            tma_load_2d(b_base + sb * kBBytes, &map_b, bfull(sb), tap * p.Cred + kc * KC, n0);
            if (++sb == SB) { sb = 0; phb ^= 1; }
          }
        }
      }
      UDA_TR(if (trp) { trp[2] = tr_w; trp[3] = clock64() - tr0; })
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN);
      UDA_TR(long long tr_wf = 0, tr_we = 0, tr_first = 0, tr_wa = 0;)
      // one in-order thread issues every MMA: keep its per-stage instruction count minimal — ring
      // positions are counters, descriptors a constant high word plus a low word that is only added to
      const uint32_t dhi = kmajor_desc_hi(kRowB);
      const uint32_t a_lo0 = kmajor_desc_lo(a_base), b_lo0 = kmajor_desc_lo(b_base);
      const uint32_t row16 = kRowB >> 4;                    // one pixel row in descriptor address units
      int sa = 0, sb = 0, j = 0; uint32_t pha = 0, phb = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
        const int mt = t % p.m_tiles;
        const int gb0 = mt * NBLK;
        const int nblk = min(NBLK, p.total_blocks - gb0);
        uint32_t off[NBLK];   // descriptor offset of each block's tap (0,0) window inside its halo
#pragma unroll
        for (int i = 0; i < NBLK; ++i) {
          const int m0 = ((gb0 + i) % p.nb_img) * 128;
          off[i] = (uint32_t)(i * p.halo_stride) / 16u + (uint32_t)(m0 - (m0 / p.P) * p.P) * row16;
        }
        const int q = j % kSets;
        UDA_TR_WAIT(tr_we, mbar_wait(tempty(q), ((j / kSets) & 1) ^ 1))
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)q * kAccCols;
        for (int kc = 0; kc < p.kchunks; ++kc) {
          UDA_TR_WAIT(tr_wa, mbar_wait(afull(sa), pha))
          tc_fence_after();
          const uint32_t halo_lo = a_lo0 + (uint32_t)(sa * a_stage_bytes) / 16u;
          uint32_t tap_lo = 0;                              // (kh * P + kw) rows
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            UDA_TR_WAIT(tr_wf, mbar_wait(bfull(sb), phb))
            UDA_TR(if (!tr_first) tr_first = clock64() - tr0;)
            tc_fence_after();
            const uint32_t b_lo = b_lo0 + (uint32_t)sb * (kBBytes >> 4);
#pragma unroll
            for (int i = 0; i < NBLK; ++i) {
              if (i < nblk) {
                const uint32_t a_lo = halo_lo + off[i] + tap_lo;
#pragma unroll
                for (int k = 0; k < KC / 16; ++k)
                  umma_bf16(acc + (uint32_t)i * BN, desc64(a_lo + 2 * k, dhi), desc64(b_lo + 2 * k, dhi), idesc,
                            (kc > 0 || tap > 0 || k > 0) ? 1u : 0u);
              }
            }
            umma_commit(bempty(sb));
            if (++sb == SB) { sb = 0; phb ^= 1; }
            tap_lo += (tap == 2 || tap == 5) ? (uint32_t)(p.P - 2) * row16 : row16;   // next kw, or next kh row
          }
          umma_commit(aempty(sa));
          if (++sa == kAStages) { sa = 0; pha ^= 1; }
        }
        umma_commit(tfull(q));
      }
      UDA_TR(if (trp) { trp[4] = tr_wf; trp[5] = tr_we; trp[6] = tr_first; trp[7] = clock64() - tr0; trp[12] = j; trp[13] = tr_wa; })
    }
  } else {
    // ===================== epilogue (4 warps): 128 pitched positions per block, junk columns skipped ==========
    UDA_TR(long long tr_wt = 0, tr_busy = 0;)
    const int qw = warp & 3;
    const int eh = (warp - 2) >> 2;          // which of the kEpiSplit warps of this lane quadrant: alternate column chunks
    constexpr int kChunks = (BN + 31) / 32;
    static_assert(kChunks % kEpiSplit == 0, "BN >= 64");
    float bn_s[kChunks], bn_q[kChunks];
#pragma unroll
    for (int cc = 0; cc < kChunks; ++cc) { bn_s[cc] = 0.f; bn_q[cc] = 0.f; }
    if constexpr (FUSE) {
      // ---- conv + BatchNorm + activation in one launch (BnFuse): exactly one tile per CTA.  When the epilogue starts
      // every MMA of the CTA has completed, so the operand rings are free: the bf16 z tile is STAGED there (rows of
      // kPitch bytes: conflict-free 16-byte accesses), leaves through one bulk copy per row (full lines instead of the
      // 32 row-strided 16-byte pieces of a register-file store), survives the grid barrier, and pass 2 turns it into
      // a = act(z*scale + shift (+ residual)) IN PLACE with all epilogue threads on consecutive vectors (no TMEM lane
      // constraint).  The residual rows arrive through bulk copies issued before pass 1. ----
      using stream::bulk_load; using stream::bulk_store; using stream::bulk_commit; using stream::fence_async_smem;
      constexpr int kPitch = BN * 2 + 16;
      constexpr int kVecs = BN / 8;                       // 16-byte vectors per row
      const int t = blockIdx.x;
      const int mt = t % p.m_tiles, n0 = (t / p.m_tiles) * BN;
      const int gb0 = mt * NBLK;
      const int nblk = min(NBLK, p.total_blocks - gb0);
      float* const s_sc = reinterpret_cast<float*>(bars + 26);
      float* const s_sf = s_sc + BN;
      float* const s_red = s_sf + BN;                     // [2][BN] CTA-level partial statistics
      uint8_t* const st_out = smem;                       // [NBLK*128][kPitch] z
      uint8_t* const st_res = smem + NBLK * 128 * kPitch; // [NBLK*128][kPitch] residual rows
      uint8_t* const st_a = st_res + NBLK * 128 * kPitch; // [NBLK*128][kPitch] a (the z rows may still be leaving)
      const uint32_t rbar = bar_base + 8u * 25;
      const int et = threadIdx.x - 64;   // index among the epilogue threads
      const bf16* const res = (const bf16*)p.fuse.residual;
      bf16* const aout = (bf16*)p.fuse.a_out;
      for (int ch = et; ch < 2 * BN; ch += kEpiThreads) s_red[ch] = 0.f;
      bar_sync(2, kEpiThreads);
      UDA_TR_WAIT(tr_wt, mbar_wait(tfull(0), 0))
      UDA_TR(const long long tr_b0 = clock64();)
      tc_fence_after();
      // this thread's row of each block (two warps share a row: alternate 32-column chunks)
      long long pix[NBLK]; bool valid[NBLK];
#pragma unroll
      for (int i = 0; i < NBLK; ++i) {
        const int gb = gb0 + i, b = gb / p.nb_img;
        const int m = (gb % p.nb_img) * 128 + qw * 32 + lane;
        const int r = m / p.P, c = m - r * p.P;
        valid[i] = i < nblk && c < p.W && r < p.H;
        pix[i] = valid[i] ? ((long long)b * p.H + r) * p.W + c : (long long)b * p.H * p.W;   // junk rows: any valid pixel
      }
      if (res) {   // residual rows -> shared memory, under pass 1
        if (et == 0) mbar_expect_tx(rbar, (uint32_t)nblk * 128u * (uint32_t)(BN * 2));
        if (eh == 0) {
#pragma unroll
          for (int i = 0; i < NBLK; ++i)
            if (i < nblk)
              bulk_load(smem_u32(st_res + (i * 128 + qw * 32 + lane) * kPitch), res + pix[i] * p.Cout + n0, BN * 2, rbar);
        }
      }
      // pass 1: statistics of the bf16-rounded outputs (junk rows contribute zeros); z staged for its bulk store
#pragma unroll
      for (int i = 0; i < NBLK; ++i) {
        if (i >= nblk) break;
        const uint32_t tbase = tmem_base + ((uint32_t)(qw * 32) << 16) + (uint32_t)i * BN;
        uint8_t* const srow = st_out + (i * 128 + qw * 32 + lane) * kPitch;
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          if (((c0 / 32) % kEpiSplit) != eh) continue;
          uint32_t v[32];
          tmem_ld_32x32(tbase + (uint32_t)c0, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) f[k] = valid[i] ? __uint_as_float(v[k]) : 0.f;
          if (p.addend && valid[i]) {
            const bf16* add = p.addend + pix[i] * p.Cout + n0 + c0;
#pragma unroll
            for (int k = 0; k < 32; k += 8) {
              float a8[8];
              ld_vec<8>(add + k, a8);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[k + e] += a8[e];
            }
          }
#pragma unroll
          for (int k = 0; k < 32; k += 8)
            *reinterpret_cast<uint4*>(srow + (c0 + k) * 2) =
                make_uint4(pack_bf16x2(f[k], f[k + 1]), pack_bf16x2(f[k + 2], f[k + 3]),
                           pack_bf16x2(f[k + 4], f[k + 5]), pack_bf16x2(f[k + 6], f[k + 7]));
        }
      }
      fence_async_smem();
      bar_sync(2, kEpiThreads);
      // statistics of the staged (bf16) tile: a thread owns one channel pair over a slice of the rows — consecutive
      // lanes read consecutive 4-byte words (no bank conflicts), ~3x fewer instructions than the per-warp shuffle trees
      {
        constexpr int kPairs = BN / 2, kSlices = kEpiThreads / kPairs;
        const int cp = et % kPairs, sl = et / kPairs;
        const int rows = nblk * 128;
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        for (int row = sl; row < rows; row += kSlices) {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(st_out + row * kPitch + cp * 4);
          const float x0 = __uint_as_float(w << 16), x1 = __uint_as_float(w & 0xffff0000u);
          s0 += x0; s1 += x1; q0 = fmaf(x0, x0, q0); q1 = fmaf(x1, x1, q1);
        }
        atomicAdd(s_red + 2 * cp, s0); atomicAdd(s_red + 2 * cp + 1, s1);
        atomicAdd(s_red + BN + 2 * cp, q0); atomicAdd(s_red + BN + 2 * cp + 1, q1);
      }
      bar_sync(2, kEpiThreads);
      if (et < 2 * BN)    // one fp64 atomic per channel and CTA
        atomicAdd(p.bn_sums + (et < BN ? n0 + et : p.Cout + n0 + et - BN), (double)s_red[et]);
      if (eh == 0) {      // z leaves: one bulk copy per valid row
#pragma unroll
        for (int i = 0; i < NBLK; ++i)
          if (valid[i])
            bulk_store(p.out + pix[i] * p.Cout + n0, smem_u32(st_out + (i * 128 + qw * 32 + lane) * kPitch), BN * 2);
        bulk_commit();
      }
      UDA_TR(if (trp && warp == 2 && lane == 0) trp[5] = clock64() - tr0;)      // pass 1 done
      grid_barrier(p.fuse.counter, gridDim.x, 2, kEpiThreads, et == 0);
      for (int ch = et; ch < BN; ch += kEpiThreads) {
        bn_fuse_coeffs(p.fuse, p.bn_sums, p.Cout, n0 + ch, s_sc[ch], s_sf[ch]);
        if (mt == 0) bn_fuse_publish(p.fuse, p.bn_sums, p.Cout, n0 + ch);
      }
      bar_sync(2, kEpiThreads);
      UDA_TR(if (trp && warp == 2 && lane == 0) trp[13] = clock64() - tr0;)     // barrier passed, coefficients ready
      // pass 2: a = act(z*scale + shift (+ residual)), consecutive threads on consecutive 16-byte vectors
      if (res) mbar_wait(rbar, 0);
      const float slope = p.fuse.slope;
      for (int vi = et; vi < nblk * 128 * kVecs; vi += kEpiThreads) {
        const int row = vi / kVecs, pc = vi - row * kVecs;
        uint8_t* const zp = st_out + row * kPitch + pc * 16;
        float z8[8], r8[8];
        stream::lds8(zp, z8);
        if (res) stream::lds8(st_res + row * kPitch + pc * 16, r8);
        const float4 sa = *reinterpret_cast<const float4*>(s_sc + pc * 8), sb = *reinterpret_cast<const float4*>(s_sc + pc * 8 + 4);
        const float4 fa = *reinterpret_cast<const float4*>(s_sf + pc * 8), fb = *reinterpret_cast<const float4*>(s_sf + pc * 8 + 4);
        const float sc8[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
        const float sf8[8] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float y = z8[e] * sc8[e] + sf8[e];
          if (res) y += r8[e];
          z8[e] = y > 0.f ? y : y * slope;
        }
        stream::sts8(st_a + row * kPitch + pc * 16, z8);
      }
      fence_async_smem();
      bar_sync(2, kEpiThreads);
      if (eh == 0) {
#pragma unroll
        for (int i = 0; i < NBLK; ++i)
          if (valid[i])
            bulk_store(aout + pix[i] * p.Cout + n0, smem_u32(st_a + (i * 128 + qw * 32 + lane) * kPitch), BN * 2);
        bulk_commit();
        stream::bulk_wait_read<0>();     // shared memory must outlive the copies' reads
      }
      UDA_TR(tr_busy += clock64() - tr_b0;)
      UDA_TR(if (trp && warp == 2 && lane == 0) { trp[8] = tr_wt; trp[9] = tr_busy; trp[10] = clock64() - tr0; })
    } else {
    int bn_n0 = -1;
    int j = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++j) {
      const int mt = t % p.m_tiles, n0 = (t / p.m_tiles) * BN;
      const int gb0 = mt * NBLK;
      const int nblk = min(NBLK, p.total_blocks - gb0);
      const int q = j % kSets;
      if (p.bn_sums && n0 != bn_n0) {   // channel tile changed: flush the partial statistics
        if (bn_n0 >= 0) {
#pragma unroll
          for (int cc = 0; cc < kChunks; ++cc) {
            const int col = bn_n0 + cc * 32 + lane;
            if (col < p.Cout) { atomicAdd(p.bn_sums + col, (double)bn_s[cc]); atomicAdd(p.bn_sums + p.Cout + col, (double)bn_q[cc]); }
            bn_s[cc] = 0.f; bn_q[cc] = 0.f;
          }
        }
        bn_n0 = n0;
      }
      UDA_TR_WAIT(tr_wt, mbar_wait(tfull(q), (j / kSets) & 1))
      UDA_TR(const long long tr_b0 = clock64();)
      tc_fence_after();
#pragma unroll 1
      for (int i = 0; i < nblk; ++i) {
        const int gb = gb0 + i, b = gb / p.nb_img;
        const int m = (gb % p.nb_img) * 128 + qw * 32 + lane;
        const int r = m / p.P, c = m - r * p.P;
        const bool valid = c < p.W && r < p.H;
        const long long pix = ((long long)b * p.H + r) * p.W + c;
        const uint32_t tbase = tmem_base + ((uint32_t)(qw * 32) << 16) + (uint32_t)q * kAccCols + (uint32_t)i * BN;
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          const int nbase = n0 + c0;
          if (nbase >= p.Cout) break;   // warp-uniform
          if (((c0 / 32) % kEpiSplit) != eh) continue;
          uint32_t v[32];
          tmem_ld_32x32(tbase + (uint32_t)c0, v);
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) f[k] = valid ? __uint_as_float(v[k]) : 0.f;
          if (p.bias && valid) {
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if (nbase + k < p.Cout) f[k] += __ldg(p.bias + nbase + k);
          }
          if (p.addend && valid) {
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
          if (p.bn_sums) bn_chunk_stats(f, lane, bn_s[c0 / 32], bn_q[c0 / 32]);   // junk rows contribute zeros
          if (valid) {
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
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(q));
      UDA_TR(tr_busy += clock64() - tr_b0;)
    }
    UDA_TR(if (trp && warp == 2 && lane == 0) { trp[8] = tr_wt; trp[9] = tr_busy; trp[10] = clock64() - tr0; })
    if (p.bn_sums && bn_n0 >= 0) {
#pragma unroll
      for (int cc = 0; cc < kChunks; ++cc) {
        const int col = bn_n0 + cc * 32 + lane;
        if (col < p.Cout) { atomicAdd(p.bn_sums + col, (double)bn_s[cc]); atomicAdd(p.bn_sums + p.Cout + col, (double)bn_q[cc]); }
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

// UDA_B200_TC_PHALO: 0 = off, 1 (default) = on for layers with at least half a wave of tiles, 2 = always (tests)
int phalo_mode() {
  const char* e = getenv("UDA_B200_TC_PHALO");   // read on every call: the tests switch it
  return e ? atoi(e) : 1;
}

template <int KC, int BN, int NBLK, bool FUSE>
int launch_phalo(const CUtensorMap& ma, const CUtensorMap& mb, PHParams& p, cudaStream_t st) {
  constexpr int kBBytes = BN * KC * 2;
  p.halo_bytes = p.HR * p.P * KC * 2;
  p.halo_stride = (p.halo_bytes + 1023) / 1024 * 1024;
  const int a_bytes = kAStages * NBLK * p.halo_stride;
  int SB = (kSmemBudget - a_bytes) / kBBytes;
  if (SB > kMaxBStages) SB = kMaxBStages;
  if (SB < 3) return UDA_ERR_UNSUPPORTED;   // caller falls back
  p.b_stages = SB;
  p.m_tiles = (p.total_blocks + NBLK - 1) / NBLK;
  const int smem = a_bytes + SB * kBBytes + 1024 + 256 + (FUSE ? 4 * BN * 4 : 0);   // + scale / shift / statistics tables
  if (FUSE && a_bytes + SB * kBBytes < 3 * NBLK * 128 * (BN * 2 + 16)) return UDA_ERR_UNSUPPORTED;   // z / residual / a staging
  static bool configured = false;
  if (!configured) {
    UDA_CUDA_OK(cudaFuncSetAttribute(conv_tc_phalo_kernel<KC, BN, NBLK, FUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
    configured = true;
  }
  const int total_tiles = p.m_tiles * p.n_tiles;
  // fused BatchNorm: every tile must be resident in TMEM at the grid barrier -> one tile per CTA, one CTA per SM
  if (FUSE && total_tiles > num_sms()) return UDA_ERR_UNSUPPORTED;
  const int grid = total_tiles < num_sms() ? total_tiles : num_sms();
  UDA_CUDA_OK(launch_pdl(conv_tc_phalo_kernel<KC, BN, NBLK, FUSE>, dim3(grid), dim3(kThreads), smem, st, ma, mb, p));
  UDA_LAUNCH_OK("conv_tc_phalo_kernel");
  return UDA_OK;
}

}  // namespace

// Returns UDA_ERR_UNSUPPORTED (without an error message the caller would surface) when the shape is not a 3x3
// stride-1 "same" convolution on a narrow image with 64-multiple channel counts.
int run_gemm_conv_phalo(const GemmConv& g, cudaStream_t st) {
  const int mode = phalo_mode();
  if (mode == 0) return UDA_ERR_UNSUPPORTED;
  if (g.ncls != 1 || g.src_s2 || g.a_map || g.os != 1 || g.wtaps != 9 || g.cls[0].ntaps != 9) return UDA_ERR_UNSUPPORTED;
  const TapClass& c = g.cls[0];
  if (c.oh != 0 || c.ow != 0) return UDA_ERR_UNSUPPORTED;
  for (int t = 0; t < 9; ++t)
    if (c.dh[t] != t / 3 - 1 || c.dw[t] != t % 3 - 1 || c.wtap[t] != t) return UDA_ERR_UNSUPPORTED;
  const int H = g.SH, W = g.SW;
  if (g.OH != H || g.OW != W || g.out_nchw || !g.out || g.st_sums) return UDA_ERR_UNSUPPORTED;
  if (g.Cred % 64 || g.Cout % 64 || g.Cout < 64) return UDA_ERR_UNSUPPORTED;
  // W = 64 (330-row halos) is implemented and tested, but measured no faster than the persistent kernel's
  // 256 x 128 tiles (30.8 vs 30.7 us on layer2 at B=16, although it moves half the operand bytes): default on for
  // W = 32 only
  if (!(W == 32 || (W == 64 && mode == 2)) || H < 8) return UDA_ERR_UNSUPPORTED;
  if (!(aligned<bf16>(g.src, 16) && aligned<bf16>(g.wmat, 16) && aligned<bf16>(g.out, 16) &&
        (!g.addend || aligned<bf16>(g.addend, 16))))
    return UDA_ERR_UNSUPPORTED;
  const int BN = g.Cout % 128 == 0 ? 128 : 64;
  const int KC = 64;
  PHParams p{};
  p.H = H; p.W = W; p.P = W + 2; p.B = g.B;
  p.nb_img = (H * p.P + 127) / 128;
  p.total_blocks = g.B * p.nb_img;
  p.n_tiles = g.Cout / BN;
  p.Cout = g.Cout; p.Cred = g.Cred; p.kchunks = g.Cred / KC;
  p.HR = (128 + p.P - 1) / p.P + 3;
  p.out = (bf16*)g.out; p.addend = (const bf16*)g.addend; p.bn_sums = g.bn_sums;
  p.bias = g.bias; p.act = g.act; p.act_slope = g.act_slope;
  UDA_TR(p.trace = take_trace_slice();)
  // enough work for a full wave, else the persistent kernel's smaller tiles are the better fit
  const int nblk = 2;
  if (mode != 2 && (long long)((p.total_blocks + nblk - 1) / nblk) * p.n_tiles < num_sms() / 2) return UDA_ERR_UNSUPPORTED;
  CUtensorMap ma, mb;
  {
    const uint64_t C = (uint64_t)g.Cred;
    uint64_t dims[4] = {C, (uint64_t)W, (uint64_t)H, (uint64_t)g.B};
    uint64_t str[3] = {C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {(uint32_t)KC, (uint32_t)p.P, (uint32_t)p.HR, 1};
    if (int rc = make_tmap_bf16(&ma, g.src, 4, dims, str, box, KC * 2)) return rc;
  }
  {
    const uint64_t Kt = (uint64_t)9 * g.Cred;
    uint64_t dims[2] = {Kt, (uint64_t)g.Cout};
    uint64_t str[1] = {Kt * 2};
    uint32_t box[2] = {(uint32_t)KC, (uint32_t)BN};
    if (int rc = make_tmap_bf16(&mb, g.wmat, 2, dims, str, box, KC * 2)) return rc;
  }
  if (g.fuse) {
    if (!g.bn_sums || g.bias || g.act) return UDA_ERR_UNSUPPORTED;
    p.fuse = *g.fuse;
    if (BN == 128) return launch_phalo<64, 128, 2, true>(ma, mb, p, st);
    return launch_phalo<64, 64, 2, true>(ma, mb, p, st);
  }
  if (BN == 128) return launch_phalo<64, 128, 2, false>(ma, mb, p, st);
  return launch_phalo<64, 64, 2, false>(ma, mb, p, st);
}

}  // namespace tcconv
}  // namespace uda
