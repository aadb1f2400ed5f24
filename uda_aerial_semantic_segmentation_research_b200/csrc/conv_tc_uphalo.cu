// Halo-tile tcgen05 kernel for the x half of decoder conv1: y = conv_transpose4x4_s2_p1(x, wx_ft)
// (== conv3x3(upsample2x(x), Wx), see uda_upconv_* in include/uda_b200.h) on wide images (w % 128 == 0) with few
// output channels (O <= 32): decoder blocks 3 and 4 of the U-Net at 512 x 512.
//
// On the persistent kernel the four output parities are four tap classes of one launch, each with its own four TMA
// boxes: every x pixel is loaded 16 times, and with 32- / 64-byte rows the launch is bound by the TMA row rate
// (in-graph trace: dec4 180 us — 113k operand rows per SM — against a 31 us HBM floor).  Here
//   * the (R+2) x 130 pixel halo of an R-row x 128-pixel tile of x is loaded ONCE per channel chunk (one TMA box);
//   * all four parities of the tile accumulate in ONE TMEM accumulator set laid out [c][i][a][32 columns]
//     (c = column parity, i = x row of the tile, a = row parity);
//   * the N dimension of an MMA carries the (row, row-parity) pairs a halo row contributes to: halo row rx with
//     column shift dw feeds (i, a) = (rx-2, 1), (rx-1, 0), (rx-1, 1), (rx, 0) — four CONSECUTIVE 32-column blocks of
//     the accumulator for the column parities that use dw — so one instruction of N <= 128 replaces four of N = 32
//     (an instruction costs ~55 clocks up to N = 64, 64 at N = 128): 4 instructions per halo row and K step;
//   * an epilogue thread owns one x pixel and writes, per output row, its two output pixels (both column parities)
//     as one contiguous run of 2*O channels.
#include "conv_tc_internal.cuh"
#include <stdlib.h>

namespace uda {
namespace tcconv {
namespace {

using namespace tc;

constexpr int kThreads = kConvThreads;
constexpr int kHaloW = 130;
constexpr int kSmemBudget = 222 * 1024;
constexpr int kBN = 32;          // accumulator block: one (c, i, a) triple
constexpr int kR = 2;            // x rows per tile: 2 * kR * 2 * kBN = 256 columns per accumulator set, two sets

struct UHParams {
  int h, w, B, tiles_w, tiles_h, total_tiles;     // low-resolution image; tiles per image
  int O, C1, kchunks, stages;
  bf16* out;                                      // [B, 2h, 2w, O]
  double* bn_sums;                                // optional [2*O]
};

// (column shift dw, column parity c) combinations in issue order: (-1, 0), (0, 0), (0, 1), (+1, 1); combinations 0 and 2
// are the first to touch their parity's half of the accumulator set.

template <int KC>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_uphalo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const UHParams p) {
  constexpr int R = kR, BN = kBN;
  constexpr int kRowB = KC * 2;
  constexpr int kHaloBytes = (R + 2) * kHaloW * kRowB;
  constexpr int kHaloStride = (kHaloBytes + 1023) / 1024 * 1024;
  constexpr int kBBytes = BN * KC * 2;
  constexpr uint32_t kAccCols = 2 * R * 2 * BN;     // 256
  constexpr uint32_t kTmemCols = 2 * kAccCols;      // 512
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  const int ws_bytes = 16 * p.kchunks * kBBytes;
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

  if (warp == 0) {
    // ===================== TMA producer: weights once, then one halo box per (tile, channel chunk) ==========
    if (elect_one()) {
      // weight blocks [combination][channel chunk][b]: block b of a combination = weight row 3-b, column wcol
      mbar_expect_tx(ws_bar, ws_bytes);
      const int wcol[4] = {0, 2, 1, 3};
      for (int cmb = 0; cmb < 4; ++cmb)
        for (int kc = 0; kc < p.kchunks; ++kc)
          for (int b = 0; b < 4; ++b)
            tma_load_2d(ws_base + ((cmb * p.kchunks + kc) * 4 + b) * kBBytes, &map_b, ws_bar,
                        ((3 - b) * 4 + wcol[cmb]) * p.C1 + kc * KC, 0);
      int it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int b = t / tiles_per_img, tin = t % tiles_per_img;
        const int i0 = (tin / p.tiles_w) * R, j0 = (tin % p.tiles_w) * 128;
        for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
          const int s = it % S;
          mbar_wait(empty_bar(s), ((it / S) & 1) ^ 1);
          mbar_expect_tx(full_bar(s), kHaloBytes);
          tma_load_4d(ring_base + s * kHaloStride, &map_a, full_bar(s), kc * KC, j0 - 1, i0 - 1, b);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      mbar_wait(ws_bar, 0);
      tc_fence_after();
      int it = 0, j = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++j) {
        const int q = j & 1;
        mbar_wait(tempty_bar(q), ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)q * kAccCols;
        for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
          const int s = it % S;
          mbar_wait(full_bar(s), (it / S) & 1);
          tc_fence_after();
          const uint32_t halo = ring_base + s * kHaloStride;
#pragma unroll 1
          for (int rx = 0; rx < R + 2; ++rx) {
            // block b of this halo row <-> accumulator block (i, a) with 2*i + a = 2*rx - 3 + b; blocks 2, 3 are written
            // for the first time by this halo row, blocks 0, 1 already hold the previous rows' contributions
            const int b_lo = 3 - 2 * rx > 0 ? 3 - 2 * rx : 0;
            const int b_hi = 2 * R + 2 - 2 * rx < 3 ? 2 * R + 2 - 2 * rx : 3;
#pragma unroll
            for (int cmb = 0; cmb < 4; ++cmb) {
              const int dwi = cmb == 0 ? 0 : (cmb == 3 ? 2 : 1), cpar = cmb >> 1;
              const bool first = kc == 0 && (cmb == 0 || cmb == 2);
              const uint64_t adesc = make_kmajor_desc(halo + (rx * kHaloW + dwi) * kRowB, kRowB);
              const uint32_t wblk = ws_base + ((cmb * p.kchunks + kc) * 4) * kBBytes;
              const uint32_t dcol = acc + (uint32_t)((cpar * R * 2 + 2 * rx - 3) * BN);
#pragma unroll
              for (int k = 0; k < KC / 16; ++k) {
                if (first && k == 0) {
                  const int a_hi = b_hi < 1 ? b_hi : 1;        // blocks that already hold data
                  if (b_lo <= a_hi)
                    umma_bf16(dcol + (uint32_t)b_lo * BN, adesc, make_kmajor_desc(wblk + b_lo * kBBytes, kRowB),
                              make_idesc_bf16(128, (a_hi - b_lo + 1) * BN), 1u);
                  const int f_lo = b_lo > 2 ? b_lo : 2;        // blocks written for the first time
                  if (f_lo <= b_hi)
                    umma_bf16(dcol + (uint32_t)f_lo * BN, adesc, make_kmajor_desc(wblk + f_lo * kBBytes, kRowB),
                              make_idesc_bf16(128, (b_hi - f_lo + 1) * BN), 0u);
                } else {
                  umma_bf16(dcol + (uint32_t)b_lo * BN, adesc + 2ull * k,
                            make_kmajor_desc(wblk + b_lo * kBBytes, kRowB) + 2ull * k,
                            make_idesc_bf16(128, (b_hi - b_lo + 1) * BN), 1u);
                }
              }
            }
          }
          umma_commit(empty_bar(s));
        }
        umma_commit(tfull_bar(q));
      }
    }
  } else {
    // ===================== epilogue: thread = one x pixel; per (i, a) two output pixels of one output row =========
    const int qw = warp & 3;
    const int eh = (warp - 2) >> 2;
    float late_s[32], late_q[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) { late_s[k] = 0.f; late_q[k] = 0.f; }
    const int OH = 2 * p.h, OW = 2 * p.w, O = p.O;
    int j = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++j) {
      const int b = t / tiles_per_img, tin = t % tiles_per_img;
      const int i0 = (tin / p.tiles_w) * R, j0 = (tin % p.tiles_w) * 128;
      const int q = j & 1;
      mbar_wait(tfull_bar(q), (j >> 1) & 1);
      tc_fence_after();
      const int m = qw * 32 + lane;
      const uint32_t tlane = tmem_base + ((uint32_t)(qw * 32) << 16) + (uint32_t)q * kAccCols;
#pragma unroll 1
      for (int ia = 0; ia < 2 * R; ++ia) {
        if ((ia % kEpiSplit) != eh) continue;
        const int orow = 2 * i0 + ia;      // 2*(i0 + i) + a
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(tlane + (uint32_t)(ia * BN), v0);                     // column parity 0
        tmem_ld_32x32(tlane + (uint32_t)((2 * R + ia) * BN), v1);           // column parity 1
        tmem_ld_wait();
        bf16* dst = p.out + (((long long)b * OH + orow) * OW + 2 * (j0 + m)) * O;
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          if (k < O) {
            float o0[8], o1[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { o0[e] = __uint_as_float(v0[k + e]); o1[e] = __uint_as_float(v1[k + e]); }
            st_vec<8>(dst + k, o0);
            st_vec<8>(dst + O + k, o1);
            if (p.bn_sums) {
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const float r0 = __bfloat162float(__float2bfloat16_rn(o0[e]));
                const float r1 = __bfloat162float(__float2bfloat16_rn(o1[e]));
                late_s[k + e] += r0 + r1;
                late_q[k + e] = fmaf(r0, r0, fmaf(r1, r1, late_q[k + e]));
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(q));
    }
    if (p.bn_sums) {
      const float s = warp_column_sums(late_s, lane), qq = warp_column_sums(late_q, lane);
      if (lane < O) { atomicAdd(p.bn_sums + lane, (double)s); atomicAdd(p.bn_sums + O + lane, (double)qq); }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int KC>
int launch_uphalo(const CUtensorMap& ma, const CUtensorMap& mb, UHParams& p, cudaStream_t st) {
  constexpr int kHaloBytes = (kR + 2) * kHaloW * KC * 2;
  constexpr int kHaloStride = (kHaloBytes + 1023) / 1024 * 1024;
  const int ws_bytes = 16 * p.kchunks * kBN * KC * 2;
  int S = (kSmemBudget - ws_bytes) / kHaloStride;
  if (S > 4) S = 4;
  if (S < 2) return UDA_ERR_UNSUPPORTED;
  p.stages = S;
  const int smem = ws_bytes + S * kHaloStride + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    UDA_CUDA_OK(cudaFuncSetAttribute(conv_tc_uphalo_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  UDA_CUDA_OK(launch_pdl(conv_tc_uphalo_kernel<KC>, dim3(grid), dim3(kThreads), smem, st, ma, mb, p));
  UDA_LAUNCH_OK("conv_tc_uphalo_kernel");
  return UDA_OK;
}

}  // namespace

// y[B, 2h, 2w, O] = conv_transpose4x4_s2_p1(x[B, h, w, C1], wx_ft[O][4][4][C1]) (+ BatchNorm statistics of y).
// UDA_ERR_UNSUPPORTED (no message) unless w % 128 == 0, h % 2 == 0, O <= 32, C1 a multiple of 32 with all sixteen
// weight blocks resident; UDA_B200_UPHALO=0 switches the path off (read on every call: A/B, tests).
int run_upconv_halo(const void* x, const void* wx_ft, void* y, double* bn_sums, int B, int h, int w, int C1, int O,
                    cudaStream_t st) {
  const char* e = getenv("UDA_B200_UPHALO");
  if (e && e[0] == '0') return UDA_ERR_UNSUPPORTED;
  if (w % 128 || h % kR || O > 32 || O % 8 || O < 8 || C1 % 32 || C1 > 128) return UDA_ERR_UNSUPPORTED;
  if (!(aligned<bf16>(x, 16) && aligned<bf16>(wx_ft, 16) && aligned<bf16>(y, 16))) return UDA_ERR_UNSUPPORTED;
  const int KC = C1 % 64 == 0 ? 64 : 32;
  UHParams p{};
  p.h = h; p.w = w; p.B = B; p.tiles_w = w / 128; p.tiles_h = h / kR; p.total_tiles = B * p.tiles_w * p.tiles_h;
  p.O = O; p.C1 = C1; p.kchunks = C1 / KC; p.out = (bf16*)y; p.bn_sums = bn_sums;
  CUtensorMap ma, mb;
  {
    const uint64_t C = (uint64_t)C1;
    uint64_t dims[4] = {C, (uint64_t)w, (uint64_t)h, (uint64_t)B};
    uint64_t str[3] = {C * 2, (uint64_t)w * C * 2, (uint64_t)h * w * C * 2};
    uint32_t box[4] = {(uint32_t)KC, (uint32_t)kHaloW, (uint32_t)(kR + 2), 1};
    if (int rc = make_tmap_bf16(&ma, x, 4, dims, str, box, KC * 2)) return rc;
  }
  {
    const uint64_t Kt = (uint64_t)16 * C1;
    uint64_t dims[2] = {Kt, (uint64_t)O};
    uint64_t str[1] = {Kt * 2};
    uint32_t box[2] = {(uint32_t)KC, (uint32_t)kBN};
    if (int rc = make_tmap_bf16(&mb, wx_ft, 2, dims, str, box, KC * 2)) return rc;
  }
  if (KC == 64) return launch_uphalo<64>(ma, mb, p, st);
  return launch_uphalo<32>(ma, mb, p, st);
}

}  // namespace tcconv
}  // namespace uda
