// Halo-tile tcgen05 wgrad for 3x3 stride-1 "same" convolutions on wide images (W % 128 == 0), sm_100a.
//
//   dW[co][kh][kw][ci] += sum over pixels  dY[pix][co] * X[pix + (kh-1, kw-1)][ci]
//
// The reduction runs over pixels, so both operands are MN-major tiles straight out of the NHWC tensors
// (see conv_tc.cu).  conv_tc_wgrad_kernel re-loads X once per tap; here the (R+2) x 130 pixel halo of an
// R-row x 128-pixel tile is loaded ONCE and every tap is a descriptor that starts at a shifted address
// inside it.  The 128 rows of one accumulator are channel atoms spaced LBO apart:
//   * Cin <= 64 (one channel chunk): the atoms of one MMA are the SAME image row shifted by 0,1,2,... pixels
//     (LBO = one pixel), i.e. the taps kw = 0,1,2 of one kernel row kh; atoms beyond kw = 2 are junk shifts
//     whose rows are never written back (the tensor pipe is idle anyway: these layers are HBM-bound);
//   * Cin = 128 (two chunks): one accumulator per tap, its two atoms are the two channel-chunk halos.
// Each CTA owns a contiguous range of pixel tiles, keeps all accumulators (<= 512 TMEM columns) for its
// whole life and adds them to the fp32 gradient with one atomic per element at the end.
#include "conv_tc_internal.cuh"
#include <stdlib.h>

namespace uda {
namespace tcconv {
namespace {

using namespace tc;

constexpr int kThreads = 192;
constexpr int kHaloW = 130;
constexpr int kSmemBudget = 222 * 1024;

struct WHParams {
  int H, W, B, tiles_w, tiles_h, total_tiles, tiles_per_cta;
  int Cin, Cout, stages;
  float* dw_out;   // [Cout][9][Cin] fp32
};

__device__ __forceinline__ uint64_t mn_desc(uint32_t addr, uint32_t row_bytes, uint32_t lbo_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((8u * row_bytes) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

template <int AA, int CA, int AB, int BN, int R>
struct WHCfg {
  static constexpr int kRowA = AA * 2, kRowB = AB * 2;
  static constexpr int kHaloBytes = (R + 2) * kHaloW * kRowA;
  static constexpr int kHaloStride = (kHaloBytes + 8 * kRowA + 1023) / 1024 * 1024;   // slack: junk shifts over-read
  static constexpr int kBAtoms = BN / AB;
  static constexpr int kDyAtomBytes = R * 128 * kRowB;
  static constexpr int kDyStride = (kDyAtomBytes + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = CA * kHaloStride + kBAtoms * kDyStride;
  static constexpr int kAtomsPerMma = 128 / AA;                                       // 8, 4, 2
  static constexpr int kGroupsPerKh = CA == 1 ? (3 + kAtomsPerMma - 1) / kAtomsPerMma : 3;   // CA==2: one per kw
  static constexpr int kGroups = 3 * kGroupsPerKh;
  static constexpr int kCols = kGroups * BN;
  static constexpr uint32_t kTmemCols = kCols <= 32 ? 32 : kCols <= 64 ? 64 : kCols <= 128 ? 128 : kCols <= 256 ? 256 : 512;
  static_assert(kCols <= 512, "accumulators exceed TMEM");
  static_assert(CA == 1 || AA == 64, "two channel chunks only with 64-channel atoms");
};

template <int AA, int CA, int AB, int BN, int R>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                          const WHParams p) {
  using Cf = WHCfg<AA, CA, AB, BN, R>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cf::kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);   // full[4], empty[4], done
  const uint32_t ring_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  const uint32_t done_bar = bar_base + 8u * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  int t_end = t_begin + p.tiles_per_cta;
  if (t_end > p.total_tiles) t_end = p.total_tiles;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_dy); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      mbar_init(done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), Cf::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the predecessor's tail; its outputs are visible from here on

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int b = t / tiles_per_img, tin = t % tiles_per_img;
        const int h0 = (tin / p.tiles_w) * R, w0 = (tin % p.tiles_w) * 128;
        const int s = it % S;
        mbar_wait(empty_bar(s), ((it / S) & 1) ^ 1);
        const uint32_t st = ring_base + s * Cf::kStageBytes;
        mbar_expect_tx(full_bar(s), CA * Cf::kHaloBytes + Cf::kBAtoms * Cf::kDyAtomBytes);
#pragma unroll
        for (int c = 0; c < CA; ++c)
          tma_load_4d(st + c * Cf::kHaloStride, &map_x, full_bar(s), c * AA, w0 - 1, h0 - 1, b);
#pragma unroll
        for (int j = 0; j < Cf::kBAtoms; ++j)
          tma_load_4d(st + CA * Cf::kHaloStride + j * Cf::kDyStride, &map_dy, full_bar(s), j * AB, w0, h0, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN < 16 ? 16 : BN) | (1u << 15) | (1u << 16);
      int it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int s = it % S;
        mbar_wait(full_bar(s), (it / S) & 1);
        tc_fence_after();
        const uint32_t st = ring_base + s * Cf::kStageBytes;
        const uint32_t dy0 = st + CA * Cf::kHaloStride;
#pragma unroll 1
        for (int sub = 0; sub < R; ++sub) {
#pragma unroll
          for (int g = 0; g < Cf::kGroups; ++g) {
            const int kh = g / Cf::kGroupsPerKh;
            const int kw0 = CA == 1 ? (g % Cf::kGroupsPerKh) * Cf::kAtomsPerMma : (g % 3);
            const uint32_t a0 = st + ((sub + kh) * kHaloW + kw0) * Cf::kRowA;
            const uint32_t lbo_a = CA == 1 ? (uint32_t)Cf::kRowA : (uint32_t)Cf::kHaloStride;
            const uint32_t b0 = dy0 + sub * 128 * Cf::kRowB;
#pragma unroll
            for (int k = 0; k < 8; ++k) {   // 16 pixels per MMA
              const uint64_t adesc = mn_desc(a0 + k * 16 * Cf::kRowA, Cf::kRowA, lbo_a);
              const uint64_t bdesc = mn_desc(b0 + k * 16 * Cf::kRowB, Cf::kRowB, Cf::kDyStride);
              umma_bf16(tmem_base + (uint32_t)g * BN, adesc, bdesc, idesc, (it > 0 || sub > 0 || k > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(done_bar);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    if (t_end > t_begin) {
#pragma unroll 1
      for (int g = 0; g < Cf::kGroups; ++g) {
        const int kh = g / Cf::kGroupsPerKh;
        int kw, ci;
        if (CA == 1) { kw = (g % Cf::kGroupsPerKh) * Cf::kAtomsPerMma + r / AA; ci = r % AA; }
        else { kw = g % 3; ci = r; }
        const bool row_ok = kw < 3 && ci < p.Cin;
        const int tap = kh * 3 + kw;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * BN + (BN >= 32 ? c0 : 0)), v);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int co = c0 + k;
              if (co < BN && co < p.Cout)
                atomicAdd(p.dw_out + ((long long)co * 9 + tap) * p.Cin + ci, __uint_as_float(v[k]));
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cf::kTmemCols);
  }
}

template <int AA, int CA, int AB, int BN, int R>
int launch_wh(const void* x, const void* dy, WHParams& p, cudaStream_t st) {
  using Cf = WHCfg<AA, CA, AB, BN, R>;
  int S = kSmemBudget / Cf::kStageBytes;
  if (S > 4) S = 4;
  if (S < 2) return UDA_ERR_UNSUPPORTED;
  p.stages = S;
  p.tiles_h = p.H / R;
  p.total_tiles = p.B * p.tiles_w * p.tiles_h;
  int ctas = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  p.tiles_per_cta = (p.total_tiles + ctas - 1) / ctas;
  ctas = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  CUtensorMap mx, mdy;
  {
    const uint64_t C = (uint64_t)p.Cin;
    uint64_t dims[4] = {C, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
    uint64_t str[3] = {C * 2, (uint64_t)p.W * C * 2, (uint64_t)p.H * p.W * C * 2};
    uint32_t box[4] = {(uint32_t)AA, (uint32_t)kHaloW, (uint32_t)(R + 2), 1};
    if (int rc = make_tmap_bf16(&mx, x, 4, dims, str, box, AA * 2)) return rc;
  }
  {
    const uint64_t Co = (uint64_t)p.Cout;
    uint64_t dims[4] = {Co, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
    uint64_t str[3] = {Co * 2, (uint64_t)p.W * Co * 2, (uint64_t)p.H * p.W * Co * 2};
    uint32_t box[4] = {(uint32_t)AB, 128, (uint32_t)R, 1};
    if (int rc = make_tmap_bf16(&mdy, dy, 4, dims, str, box, AB * 2)) return rc;
  }
  const int smem = S * Cf::kStageBytes + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    UDA_CUDA_OK(cudaFuncSetAttribute(conv_tc_wgrad_halo_kernel<AA, CA, AB, BN, R>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  UDA_CUDA_OK(launch_pdl(conv_tc_wgrad_halo_kernel<AA, CA, AB, BN, R>, dim3(ctas), dim3(kThreads), smem, st, mx, mdy, p));
  UDA_LAUNCH_OK("conv_tc_wgrad_halo_kernel");
  return UDA_OK;
}

// largest R in {8,4,2,1} that divides H and leaves >= 2 stages
template <int AA, int CA, int AB, int BN>
int pick_r_and_launch(const void* x, const void* dy, WHParams& p, cudaStream_t st) {
#define UDA_TRY(Rv)                                                                                   \
  if (p.H % Rv == 0 && kSmemBudget / WHCfg<AA, CA, AB, BN, Rv>::kStageBytes >= 2)                      \
    return launch_wh<AA, CA, AB, BN, Rv>(x, dy, p, st);
  UDA_TRY(8) UDA_TRY(4) UDA_TRY(2) UDA_TRY(1)
#undef UDA_TRY
  return UDA_ERR_UNSUPPORTED;
}


// ------------------------------------------------------------------------------------------------
// Three kernel rows per MMA.  At N <= 64 a tcgen05.mma occupies the issue slot for a flat ~55 clocks whatever N is
// (tools/exp/umma_rate.cu), and the kernel above issues one 8-instruction group per kernel row kh with N = Cout = 16 / 32:
// the 16- and 32-channel layers of decoder blocks 3-4 and the head were bound by that, not by HBM (dec4.c2: 145 us
// against a 41 us HBM floor).  Here the N dimension also carries the kernel row: for ONE halo row rx of X the three dY
// rows rx-2, rx-1, rx (kh = 2, 1, 0) are three N atoms spaced one dY row (LBO) apart, so one group per HALO row
// ((R+2)/R x 8 instructions per image row instead of 24) covers all nine taps.  dY rows outside the tile must
// contribute nothing: each stage's dY buffer has two zero rows above and below the R rows the TMA box fills.
// Needs a single dY channel atom (BN == AB <= 64; 24 channels travel as a zero-filled 32-channel box).
// ------------------------------------------------------------------------------------------------
template <int AA, int AB, int R>
struct WH3Cfg {
  static constexpr int kRowA = AA * 2, kRowB = AB * 2;
  static constexpr int kHaloBytes = (R + 2) * kHaloW * kRowA;
  static constexpr int kHaloStride = (kHaloBytes + 8 * kRowA + 1023) / 1024 * 1024;   // slack: junk shifts over-read
  static constexpr int kDyRowBytes = 128 * kRowB;
  static constexpr int kDyBytes = R * kDyRowBytes;
  static constexpr int kStageBytes = kHaloStride + kDyBytes;
  static constexpr int kAtomsPerMma = 128 / AA;                    // 8, 4, 2 pixel shifts per accumulator
  static constexpr int kGroups = (3 + kAtomsPerMma - 1) / kAtomsPerMma;   // 1, 1, 2
  static constexpr int kN = 3 * AB;                                // (kh, co)
  static constexpr int kCols = kGroups * kN;
  static constexpr uint32_t kTmemCols = kCols <= 32 ? 32 : kCols <= 64 ? 64 : kCols <= 128 ? 128 : kCols <= 256 ? 256 : 512;
  static_assert(kN <= 256 && kCols <= 512, "accumulators exceed the MMA / TMEM limits");
};

template <int AA, int AB, int R>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_wgrad_halo3_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                           const WHParams p) {
  using Cf = WH3Cfg<AA, AB, R>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * Cf::kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);   // full[4], empty[4], done
  const uint32_t ring_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (4 + s); };
  const uint32_t done_bar = bar_base + 8u * 8;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_img = p.tiles_w * p.tiles_h;
  const int t_begin = blockIdx.x * p.tiles_per_cta;
  int t_end = t_begin + p.tiles_per_cta;
  if (t_end > p.total_tiles) t_end = p.total_tiles;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_dy); }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
      mbar_init(done_bar, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(smem_u32(tmem_slot), Cf::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the predecessor's tail; its outputs are visible from here on

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int b = t / tiles_per_img, tin = t % tiles_per_img;
        const int h0 = (tin / p.tiles_w) * R, w0 = (tin % p.tiles_w) * 128;
        const int s = it % S;
        mbar_wait(empty_bar(s), ((it / S) & 1) ^ 1);
        const uint32_t st = ring_base + s * Cf::kStageBytes;
        mbar_expect_tx(full_bar(s), Cf::kHaloBytes + Cf::kDyBytes);
        tma_load_4d(st, &map_x, full_bar(s), 0, w0 - 1, h0 - 1, b);
        tma_load_4d(st + Cf::kHaloStride, &map_dy, full_bar(s), 0, w0, h0, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      int it = 0;
      // one group of 8 instructions: X halo row rx against the `nat` dY rows (N atoms) starting at dY row d0, into the
      // TMEM columns of N atom j0 (kh = 2 - j); `fresh`: these columns have never been written (first tile of the CTA)
      auto issue = [&](uint32_t st, int rx, int d0, int j0, int nat, bool fresh) {
        const uint32_t idesc = make_idesc_bf16(128, nat * AB) | (1u << 15) | (1u << 16);
#pragma unroll
        for (int g = 0; g < Cf::kGroups; ++g) {
          const uint32_t a0 = st + (rx * kHaloW + g * Cf::kAtomsPerMma) * Cf::kRowA;
          const uint32_t b0 = st + Cf::kHaloStride + d0 * Cf::kDyRowBytes;
#pragma unroll
          for (int k = 0; k < 8; ++k) {   // 16 pixels per MMA
            const uint64_t adesc = mn_desc(a0 + k * 16 * Cf::kRowA, Cf::kRowA, Cf::kRowA);
            const uint64_t bdesc = mn_desc(b0 + k * 16 * Cf::kRowB, Cf::kRowB, Cf::kDyRowBytes);
            umma_bf16(tmem_base + (uint32_t)(g * Cf::kN + j0 * AB), adesc, bdesc, idesc, (!fresh || k > 0) ? 1u : 0u);
          }
        }
      };
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int s = it % S;
        mbar_wait(full_bar(s), (it / S) & 1);
        tc_fence_after();
        const uint32_t st = ring_base + s * Cf::kStageBytes;
        const bool first = it == 0;
        // halo row rx of X pairs with the dY rows rx-2+j (j = 0,1,2 <-> kh = 2,1,0) that lie inside the tile; on the
        // CTA's first tile N atom j is written for the first time at rx = 2-j (its dY row 0): that instruction group
        // must not accumulate, so it is issued apart from the atoms that already hold data
#pragma unroll 1
        for (int rx = 0; rx < R + 2; ++rx) {
          const int j_lo = rx < 2 ? 2 - rx : 0, j_hi = (R + 1 - rx) < 2 ? (R + 1 - rx) : 2;
          if (first && rx <= 2) {
            issue(st, rx, rx - 2 + j_lo, j_lo, 1, true);
            if (j_hi > j_lo) issue(st, rx, rx - 1 + j_lo, j_lo + 1, j_hi - j_lo, false);
          } else {
            issue(st, rx, rx - 2 + j_lo, j_lo, j_hi - j_lo + 1, false);
          }
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(done_bar);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    if (t_end > t_begin) {
#pragma unroll 1
      for (int g = 0; g < Cf::kGroups; ++g) {
        const int kw = g * Cf::kAtomsPerMma + r / AA, ci = r % AA;
        const bool row_ok = kw < 3 && ci < p.Cin;
#pragma unroll 1
        for (int c0 = 0; c0 < Cf::kN; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * Cf::kN + c0), v);
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int col = c0 + k, j = col / AB, co = col % AB;     // N atom j = dY row rx-2+j  <->  kh = 2 - j
              if (j < 3 && co < p.Cout)
                atomicAdd(p.dw_out + ((long long)co * 9 + (2 - j) * 3 + kw) * p.Cin + ci, __uint_as_float(v[k]));
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cf::kTmemCols);
  }
}

template <int AA, int AB, int R>
int launch_wh3(const void* x, const void* dy, WHParams& p, cudaStream_t st) {
  using Cf = WH3Cfg<AA, AB, R>;
  int S = kSmemBudget / Cf::kStageBytes;
  if (S > 4) S = 4;
  if (S < 2) return UDA_ERR_UNSUPPORTED;
  p.stages = S;
  p.tiles_h = p.H / R;
  p.total_tiles = p.B * p.tiles_w * p.tiles_h;
  int ctas = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  p.tiles_per_cta = (p.total_tiles + ctas - 1) / ctas;
  ctas = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  CUtensorMap mx, mdy;
  {
    const uint64_t C = (uint64_t)p.Cin;
    uint64_t dims[4] = {C, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
    uint64_t str[3] = {C * 2, (uint64_t)p.W * C * 2, (uint64_t)p.H * p.W * C * 2};
    uint32_t box[4] = {(uint32_t)AA, (uint32_t)kHaloW, (uint32_t)(R + 2), 1};
    if (int rc = make_tmap_bf16(&mx, x, 4, dims, str, box, AA * 2)) return rc;
  }
  {
    const uint64_t Co = (uint64_t)p.Cout;     // may be smaller than the AB-channel box: the rest reads as zero
    uint64_t dims[4] = {Co, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)p.B};
    uint64_t str[3] = {Co * 2, (uint64_t)p.W * Co * 2, (uint64_t)p.H * p.W * Co * 2};
    uint32_t box[4] = {(uint32_t)AB, 128, (uint32_t)R, 1};
    if (int rc = make_tmap_bf16(&mdy, dy, 4, dims, str, box, AB * 2)) return rc;
  }
  const int smem = S * Cf::kStageBytes + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    UDA_CUDA_OK(cudaFuncSetAttribute(conv_tc_wgrad_halo3_kernel<AA, AB, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     227 * 1024));
    configured = true;
  }
  UDA_CUDA_OK(launch_pdl(conv_tc_wgrad_halo3_kernel<AA, AB, R>, dim3(ctas), dim3(kThreads), smem, st, mx, mdy, p));
  UDA_LAUNCH_OK("conv_tc_wgrad_halo3_kernel");
  return UDA_OK;
}

template <int AA, int AB>
int pick_r_and_launch3(const void* x, const void* dy, WHParams& p, cudaStream_t st) {
#define UDA_TRY(Rv)                                                                                   \
  if (p.H % Rv == 0 && kSmemBudget / WH3Cfg<AA, AB, Rv>::kStageBytes >= 2)                             \
    return launch_wh3<AA, AB, Rv>(x, dy, p, st);
  UDA_TRY(8) UDA_TRY(4) UDA_TRY(2)
#undef UDA_TRY
  return UDA_ERR_UNSUPPORTED;
}

// UDA_B200_WGRAD_HALO3=0 keeps the one-kernel-row-per-group kernel (read on every call: the tests compare the two)
bool halo3_enabled() {
  const char* e = getenv("UDA_B200_WGRAD_HALO3");
  return !(e && e[0] == '0');
}

}  // namespace

// 3x3 stride-1 pad-1 wgrad on W % 128 == 0 images, Cin in {16,32,64,128}, Cout <= 128 (accumulators must fit
// 512 TMEM columns).  Returns UDA_ERR_UNSUPPORTED (no message) otherwise.
int run_wgrad_halo(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin, int Cout,
                   cudaStream_t st) {
  if (W % 128 || Cout % 8 || Cout < 8 || Cout > 128) return UDA_ERR_UNSUPPORTED;
  if (!(aligned<bf16>(dy, 16) && aligned<bf16>(x, 16))) return UDA_ERR_UNSUPPORTED;
  WHParams p{};
  p.H = H; p.W = W; p.B = B; p.tiles_w = W / 128; p.Cin = Cin; p.Cout = Cout; p.dw_out = dw;
  // (64 -> 64 channels: only R = 2 rows fit, so twice as many halo rows as image rows at N = 192 — the old kernel wins)
  if (halo3_enabled() && Cout <= 64 && (Cin == 16 || Cin == 32 || Cin == 64) && !(Cin == 64 && Cout > 32)) {
    // one dY channel atom (16 / 32 / 64 channels; 24 travel as a zero-filled 32-channel box): three kernel rows per MMA
    const int ab = Cout <= 16 ? 16 : (Cout <= 32 ? 32 : 64);
    int rc = UDA_ERR_UNSUPPORTED;
#define UDA_W3(AAv, ABv) if (Cin == AAv && ab == ABv) rc = pick_r_and_launch3<AAv, ABv>(x, dy, p, st);
    UDA_W3(16, 16) UDA_W3(16, 32) UDA_W3(16, 64) UDA_W3(32, 16) UDA_W3(32, 32) UDA_W3(32, 64)
    UDA_W3(64, 16) UDA_W3(64, 32) UDA_W3(64, 64)
#undef UDA_W3
    if (rc != UDA_ERR_UNSUPPORTED) return rc;
  }
  const int atomB = Cout % 64 == 0 ? 64 : (Cout % 32 == 0 ? 32 : 16);
  const int BN = (Cout + atomB - 1) / atomB * atomB;
#define UDA_W(AAv, CAv, ABv, BNv) \
  if (atomB == ABv && BN == BNv) return pick_r_and_launch<AAv, CAv, ABv, BNv>(x, dy, p, st);
  if (Cin == 16) { UDA_W(16, 1, 16, 16) UDA_W(16, 1, 16, 32) UDA_W(16, 1, 32, 32) UDA_W(16, 1, 64, 64) UDA_W(16, 1, 64, 128) }
  if (Cin == 32) { UDA_W(32, 1, 16, 16) UDA_W(32, 1, 16, 32) UDA_W(32, 1, 32, 32) UDA_W(32, 1, 64, 64) UDA_W(32, 1, 64, 128) }
  if (Cin == 64) { UDA_W(64, 1, 16, 16) UDA_W(64, 1, 16, 32) UDA_W(64, 1, 32, 32) UDA_W(64, 1, 64, 64) }
  if (Cin == 128) { UDA_W(64, 2, 16, 16) UDA_W(64, 2, 16, 32) UDA_W(64, 2, 32, 32) }
#undef UDA_W
  return UDA_ERR_UNSUPPORTED;
}

}  // namespace tcconv
}  // namespace uda
