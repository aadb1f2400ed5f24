// Halo-tile tcgen05 kernel for the 4x4 stride-2 pad-1 convolution of a wide 16-channel tensor:
//   dx[B, h, w, C1] = conv4x4_s2_p1(dz[B, 2h, 2w, 16], W4[C1][4][4][16])
// = the backward of the transposed half of decoder conv1 (engine.upconv_bn_act) in decoder block 4 of the U-Net.
//
// On the persistent kernel each of the 16 taps is its own TMA box of 32-byte rows through the space-to-depth view:
// 16 x 256 operand rows per 256 output pixels, and the launch is bound by the TMA row rate (134 us against a 31 us HBM
// floor).  Here the space-to-depth halo of an R-row x 128-pixel output tile — rows i0-1 .. i0+R, both row parities,
// columns j0-1 .. j0+128, both column parities: a box {32 = (dj, c), 130, 2, R+2} of 64-byte rows — is loaded ONCE
// (3 rows per output pixel instead of 16), and the sixteen taps are descriptors at shifted addresses inside it:
//   kh -> (row shift, di) = (-1, 1), (0, 0), (0, 1), (+1, 0);   kw -> (column shift, dj) likewise, dj selecting the
//   first or second 32 bytes (16 channels) of a 64-byte row.
// All sixteen weight tiles stay resident; two TMEM accumulator sets overlap the epilogue with the next tile.
#include "conv_tc_internal.cuh"
#include <stdlib.h>

namespace uda {
namespace tcconv {
namespace {

using namespace tc;

constexpr int kThreads = kConvThreads;
constexpr int kHaloW = 130;
constexpr int kSmemBudget = 222 * 1024;
constexpr int kO = 16;           // channels of dz (the reduction of one tap)
constexpr int kBN = 32;          // output channels (C1 <= 32)

struct DHParams {
  int h, w, B, tiles_w, tiles_h, total_tiles;     // output (low-resolution) image; tiles per image
  int C1, stages;
  bf16* out;                                      // [B, h, w, C1]
};

template <int R>
__global__ void __launch_bounds__(kThreads, 1)
conv_tc_downhalo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                        const DHParams p) {
  constexpr int BN = kBN;
  constexpr int kRowA = 2 * kO * 2;                 // 64-byte rows: (dj, c)
  constexpr int kRowB = kO * 2;                     // 32-byte weight rows: one tap
  constexpr int kHaloBytes = (R + 2) * 2 * kHaloW * kRowA;
  constexpr int kHaloStride = (kHaloBytes + 1023) / 1024 * 1024;
  constexpr int kBBytes = BN * kRowB;               // 1 KB per tap
  constexpr int kWsBytes = 16 * kBBytes;
  constexpr uint32_t kAccCols = R * BN;
  constexpr uint32_t kTmemCols = 2 * kAccCols < 32 ? 32 : 2 * kAccCols;
  static_assert(2 * R * BN <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM columns");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWsBytes + S * kHaloStride);
  // bars: full[4], empty[4], tmem_full[2], tmem_empty[2], ws_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  const uint32_t ws_base = smem_u32(smem);
  const uint32_t ring_base = ws_base + kWsBytes;
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
    // ===================== TMA producer: sixteen weight tiles once, then one halo box per tile ==========
    if (elect_one()) {
      mbar_expect_tx(ws_bar, kWsBytes);
      for (int t = 0; t < 16; ++t) tma_load_2d(ws_base + t * kBBytes, &map_b, ws_bar, t * kO, 0);
      int it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
        const int b = t / tiles_per_img, tin = t % tiles_per_img;
        const int i0 = (tin / p.tiles_w) * R, j0 = (tin % p.tiles_w) * 128;
        const int s = it % S;
        mbar_wait(empty_bar(s), ((it / S) & 1) ^ 1);
        mbar_expect_tx(full_bar(s), kHaloBytes);
        tma_load_5d(ring_base + s * kHaloStride, &map_a, full_bar(s), 0, j0 - 1, 0, i0 - 1, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: 16 instructions (K = 16) per output row of 128 pixels =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN);
      mbar_wait(ws_bar, 0);
      tc_fence_after();
      int it = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
        const int q = it & 1, s = it % S;
        mbar_wait(tempty_bar(q), ((it >> 1) & 1) ^ 1);
        mbar_wait(full_bar(s), (it / S) & 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)q * kAccCols;
        const uint32_t halo = ring_base + s * kHaloStride;
#pragma unroll 1
        for (int sub = 0; sub < R; ++sub) {
#pragma unroll
          for (int kh = 0; kh < 4; ++kh) {
            const int ri = sub + 1 + (kh == 0 ? -1 : (kh == 3 ? 1 : 0)), di = (kh == 0 || kh == 2) ? 1 : 0;
            const uint32_t rowbase = halo + (uint32_t)((ri * 2 + di) * kHaloW) * kRowA;
#pragma unroll
            for (int kw = 0; kw < 4; ++kw) {
              const int cj = kw == 0 ? 0 : (kw == 3 ? 2 : 1), dj = (kw == 0 || kw == 2) ? 1 : 0;
              const uint64_t adesc = make_kmajor_desc(rowbase + cj * kRowA, kRowA) + (uint64_t)(2 * dj);
              const uint64_t bdesc = make_kmajor_desc(ws_base + (kh * 4 + kw) * kBBytes, kRowB);
              umma_bf16(acc + (uint32_t)sub * BN, adesc, bdesc, idesc, (kh > 0 || kw > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(empty_bar(s));
        umma_commit(tfull_bar(q));
      }
    }
  } else {
    // ===================== epilogue: one output row of 128 pixels per sub-tile =====================
    const int qw = warp & 3;
    const int eh = (warp - 2) >> 2;
    int j = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++j) {
      const int b = t / tiles_per_img, tin = t % tiles_per_img;
      const int i0 = (tin / p.tiles_w) * R, j0 = (tin % p.tiles_w) * 128;
      const int q = j & 1;
      mbar_wait(tfull_bar(q), (j >> 1) & 1);
      tc_fence_after();
      const int m = qw * 32 + lane;
#pragma unroll 1
      for (int sub = 0; sub < R; ++sub) {
        if ((sub % kEpiSplit) != eh) continue;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(qw * 32) << 16) + (uint32_t)q * kAccCols + (uint32_t)sub * BN, v);
        tmem_ld_wait();
        bf16* dst = p.out + (((long long)b * p.h + i0 + sub) * p.w + j0 + m) * p.C1;
#pragma unroll
        for (int k = 0; k < 32; k += 8) {
          if (k < p.C1) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[k + e]);
            st_vec<8>(dst + k, o);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(q));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int R>
int launch_downhalo(const CUtensorMap& ma, const CUtensorMap& mb, DHParams& p, cudaStream_t st) {
  constexpr int kHaloBytes = (R + 2) * 2 * kHaloW * 2 * kO * 2;
  constexpr int kHaloStride = (kHaloBytes + 1023) / 1024 * 1024;
  constexpr int kWsBytes = 16 * kBN * kO * 2;
  int S = (kSmemBudget - kWsBytes) / kHaloStride;
  if (S > 4) S = 4;
  if (S < 2) return UDA_ERR_UNSUPPORTED;
  p.stages = S;
  p.tiles_h = p.h / R;
  p.total_tiles = p.B * p.tiles_w * p.tiles_h;
  const int smem = kWsBytes + S * kHaloStride + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    UDA_CUDA_OK(cudaFuncSetAttribute(conv_tc_downhalo_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  UDA_CUDA_OK(launch_pdl(conv_tc_downhalo_kernel<R>, dim3(grid), dim3(kThreads), smem, st, ma, mb, p));
  UDA_LAUNCH_OK("conv_tc_downhalo_kernel");
  return UDA_OK;
}


// ------------------------------------------------------------------------------------------------
// Weight gradient of the same convolution:  dW[co][kh][kw][c] += sum_pixels dy[i][j][co] * x[2i+kh-1][2j+kw-1][c]
// (x = the wide 16-channel tensor, dy = the low-resolution tensor; decoder block 4: dW4 of the transposed half).
// The generic wgrad kernel loads one 32-byte-row TMA box per tap and pixel tile (148 us).  Here the space-to-depth halo
// of x is loaded once per tile and is an MN-major operand as it stands: a 64-byte row is 32 M rows (dj, c), four pixel
// shifts (LBO = one pixel) make the 128 rows of one MMA = (column shift cj, dj, c), of which (0,1), (1,0), (1,1), (2,0)
// are the taps kw = 0..3; the four kernel rows kh are four accumulators (row shift, di).  N = the dy channels.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mn_desc64(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((8u * 64u) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= 4ull << 61;            // SWIZZLE_64B
  return d;
}

struct DWParams {
  int h, w, B, tiles_w, tiles_h, total_tiles, tiles_per_cta;
  int Cout, stages;
  float* dw_out;   // [Cout][16][16] fp32
};

template <int R>
__global__ void __launch_bounds__(192, 1)
conv_tc_wgrad_downhalo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                              const DWParams p) {
  constexpr int kRow = 64;                                     // bytes per row of either operand
  constexpr int kHaloBytes = (R + 2) * 2 * kHaloW * kRow;
  constexpr int kHaloStride = (kHaloBytes + 8 * kRow + 1023) / 1024 * 1024;   // slack: the junk shift over-reads
  constexpr int kDyBytes = R * 128 * kRow;
  constexpr int kStageBytes = kHaloStride + kDyBytes;
  constexpr uint32_t kTmemCols = 128;                          // four accumulators (kh) of 32 columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * kStageBytes);
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
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int b = t / tiles_per_img, tin = t % tiles_per_img;
        const int i0 = (tin / p.tiles_w) * R, j0 = (tin % p.tiles_w) * 128;
        const int s = it % S;
        mbar_wait(empty_bar(s), ((it / S) & 1) ^ 1);
        const uint32_t st = ring_base + s * kStageBytes;
        mbar_expect_tx(full_bar(s), kHaloBytes + kDyBytes);
        tma_load_5d(st, &map_x, full_bar(s), 0, j0 - 1, 0, i0 - 1, b);
        tma_load_4d(st + kHaloStride, &map_dy, full_bar(s), 0, j0, i0, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, 32) | (1u << 15) | (1u << 16);   // both operands MN-major
      int it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int s = it % S;
        mbar_wait(full_bar(s), (it / S) & 1);
        tc_fence_after();
        const uint32_t st = ring_base + s * kStageBytes;
#pragma unroll 1
        for (int sub = 0; sub < R; ++sub) {
#pragma unroll
          for (int kh = 0; kh < 4; ++kh) {
            const int ri = sub + 1 + (kh == 0 ? -1 : (kh == 3 ? 1 : 0)), di = (kh == 0 || kh == 2) ? 1 : 0;
            const uint32_t a0 = st + (uint32_t)((ri * 2 + di) * kHaloW) * kRow;
            const uint32_t b0 = st + kHaloStride + (uint32_t)(sub * 128) * kRow;
#pragma unroll
            for (int k = 0; k < 8; ++k)     // 16 pixels per MMA
              umma_bf16(tmem_base + (uint32_t)kh * 32, mn_desc64(a0 + k * 16 * kRow, kRow), mn_desc64(b0 + k * 16 * kRow, kRow),
                        idesc, (it > 0 || sub > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(done_bar);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int cj = r >> 5, dj = (r >> 4) & 1, c = r & 15;
    const int kw = cj == 0 ? (dj ? 0 : -1) : (cj == 1 ? 1 + dj : (cj == 2 ? (dj ? -1 : 3) : -1));
    mbar_wait(done_bar, 0);
    tc_fence_after();
    if (t_end > t_begin) {
#pragma unroll 1
      for (int kh = 0; kh < 4; ++kh) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)kh * 32, v);
        tmem_ld_wait();
        if (kw >= 0) {
#pragma unroll
          for (int k = 0; k < 32; ++k)
            if (k < p.Cout) atomicAdd(p.dw_out + ((long long)k * 16 + kh * 4 + kw) * kO + c, __uint_as_float(v[k]));
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

// out[B, H/2, W/2, Cout] = conv4x4_s2_p1(x[B, H, W, 16], wmat[Cout][4][4][16]), Cout <= 32, (W/2) % 128 == 0.
// UDA_ERR_UNSUPPORTED (no message) otherwise; UDA_B200_DOWNHALO=0 switches the path off (read on every call).
int run_downconv_halo(const void* x, const void* wmat, void* y, int B, int H, int W, int Cin, int Cout, cudaStream_t st) {
  const char* e = getenv("UDA_B200_DOWNHALO");
  if (e && e[0] == '0') return UDA_ERR_UNSUPPORTED;
  if (Cin != kO || Cout > kBN || Cout % 8 || Cout < 8 || H % 2 || W % 2) return UDA_ERR_UNSUPPORTED;
  const int h = H / 2, w = W / 2;
  if (w % 128 || h % 2) return UDA_ERR_UNSUPPORTED;
  if (!(aligned<bf16>(x, 16) && aligned<bf16>(wmat, 16) && aligned<bf16>(y, 16))) return UDA_ERR_UNSUPPORTED;
  DHParams p{};
  p.h = h; p.w = w; p.B = B; p.tiles_w = w / 128; p.C1 = Cout; p.out = (bf16*)y;
  const int R = h % 4 == 0 ? 4 : 2;
  CUtensorMap ma, mb;
  {
    // space-to-depth view of x: {(dj, c), j, di, i, b}
    const uint64_t C = (uint64_t)Cin;
    uint64_t dims[5] = {2 * C, (uint64_t)w, 2, (uint64_t)h, (uint64_t)B};
    uint64_t str[4] = {2 * C * 2, (uint64_t)W * C * 2, 2 * (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[5] = {(uint32_t)(2 * Cin), (uint32_t)kHaloW, 2, (uint32_t)(R + 2), 1};
    if (int rc = make_tmap_bf16(&ma, x, 5, dims, str, box, 2 * Cin * 2)) return rc;
  }
  {
    const uint64_t Kt = (uint64_t)16 * Cin;
    uint64_t dims[2] = {Kt, (uint64_t)Cout};
    uint64_t str[1] = {Kt * 2};
    uint32_t box[2] = {(uint32_t)kO, (uint32_t)kBN};
    if (int rc = make_tmap_bf16(&mb, wmat, 2, dims, str, box, kO * 2)) return rc;
  }
  if (R == 4) return launch_downhalo<4>(ma, mb, p, st);
  return launch_downhalo<2>(ma, mb, p, st);
}

}  // namespace tcconv
}  // namespace uda

namespace uda {
namespace tcconv {

// dw[Cout][4][4][16] += wgrad of conv4x4_s2_p1(x[B, H, W, 16]) with dy[B, H/2, W/2, Cout], Cout <= 32, (W/2) % 128 == 0.
// UDA_ERR_UNSUPPORTED (no message) otherwise; UDA_B200_DOWNHALO=0 switches the path off.
int run_wgrad_downhalo(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin, int Cout, cudaStream_t st) {
  const char* e = getenv("UDA_B200_DOWNHALO");
  if (e && e[0] == '0') return UDA_ERR_UNSUPPORTED;
  if (Cin != kO || Cout > kBN || Cout % 8 || Cout < 8 || H % 2 || W % 2) return UDA_ERR_UNSUPPORTED;
  const int h = H / 2, w = W / 2;
  constexpr int R = 2;
  if (w % 128 || h % R) return UDA_ERR_UNSUPPORTED;
  if (!(aligned<bf16>(x, 16) && aligned<bf16>(dy, 16))) return UDA_ERR_UNSUPPORTED;
  constexpr int kHaloBytes = (R + 2) * 2 * kHaloW * 64;
  constexpr int kHaloStride = (kHaloBytes + 8 * 64 + 1023) / 1024 * 1024;
  constexpr int kStageBytes = kHaloStride + R * 128 * 64;
  int S = kSmemBudget / kStageBytes;
  if (S > 4) S = 4;
  if (S < 2) return UDA_ERR_UNSUPPORTED;
  DWParams p{};
  p.h = h; p.w = w; p.B = B; p.tiles_w = w / 128; p.tiles_h = h / R; p.total_tiles = B * p.tiles_w * p.tiles_h;
  p.Cout = Cout; p.stages = S; p.dw_out = dw;
  int ctas = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  p.tiles_per_cta = (p.total_tiles + ctas - 1) / ctas;
  ctas = (p.total_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  CUtensorMap mx, mdy;
  {
    const uint64_t C = (uint64_t)Cin;
    uint64_t dims[5] = {2 * C, (uint64_t)w, 2, (uint64_t)h, (uint64_t)B};
    uint64_t str[4] = {2 * C * 2, (uint64_t)W * C * 2, 2 * (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[5] = {(uint32_t)(2 * Cin), (uint32_t)kHaloW, 2, (uint32_t)(R + 2), 1};
    if (int rc = make_tmap_bf16(&mx, x, 5, dims, str, box, 64)) return rc;
  }
  {
    const uint64_t Co = (uint64_t)Cout;     // may be smaller than the 32-channel box: the rest reads as zero
    uint64_t dims[4] = {Co, (uint64_t)w, (uint64_t)h, (uint64_t)B};
    uint64_t str[3] = {Co * 2, (uint64_t)w * Co * 2, (uint64_t)h * w * Co * 2};
    uint32_t box[4] = {32, 128, (uint32_t)R, 1};
    if (int rc = make_tmap_bf16(&mdy, dy, 4, dims, str, box, 64)) return rc;
  }
  const int smem = S * kStageBytes + 1024 + 256;
  static bool configured = false;
  if (!configured) {
    UDA_CUDA_OK(cudaFuncSetAttribute(conv_tc_wgrad_downhalo_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  UDA_CUDA_OK(launch_pdl(conv_tc_wgrad_downhalo_kernel<R>, dim3(ctas), dim3(192), smem, st, mx, mdy, p));
  UDA_LAUNCH_OK("conv_tc_wgrad_downhalo_kernel");
  return UDA_OK;
}

}  // namespace tcconv
}  // namespace uda
