// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 NHWC activations, fp32 accumulate).
//
//   D[128 pixels x BN channels] = sum over (tap, channel chunk) A_tap[128 x KC] * W_tap[BN x KC]^T
//
//   * A operand: the NHWC activation tensor itself — a TMA tiled box {KC ch, TW, TH, NB} placed at the
//     output tile origin shifted by the tap; out-of-bounds pixels (padding) are zero-filled by TMA, so
//     there is no im2col buffer and no halo code.  Stride-2 convolutions read the same tensor through a
//     5-D "space-to-depth" view {2C, W/2, 2, H/2, B} in which every tap is again a plain box.
//   * B operand: OHWI weights viewed as a 2-D K-major matrix [Cout][taps*Cin].
//   * Both land in shared memory in the canonical K-major swizzled layout (128/64/32-byte rows) that
//     tcgen05.mma consumes directly; accumulators live in TMEM; one elected thread issues the MMAs;
//     a 4-warp epilogue reads TMEM with tcgen05.ld and writes bf16 NHWC (+bias, +addend) and/or the
//     fp32 NCHW logits edge.
//   * Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..5 = epilogue;
//     multi-stage smem ring with full/empty mbarriers.
//
// Replaces aten::_convolution (cuDNN) for the U-Net / discriminator convolutions the reference reaches
// through smp.Unet and DomainDiscriminator.  dgrad of stride-1 convolutions runs
// through the same kernel on flipped/transposed weights (uda_conv2d_weight_flip_transpose).
#include "conv_tc_internal.cuh"
#include <stdlib.h>

namespace uda {
namespace {

using namespace tc;
using namespace tcconv;

constexpr int kTcThreads = 192;  // 6 warps

struct FwdParams {
  int TW, TH, NB;              // tile = NB images x TH rows x TW cols = 128 output pixels
  int tiles_w, tiles_h;        // tiles per image row / column
  int MH, MW;                  // spatial size of the GEMM-M pixel grid the tiles cover
  int OH, OW, os, oh, ow;      // output tensor spatial size; M-grid pixel (i,j) -> output (i*os+oh, j*os+ow)
  int Cout, Cred;              // GEMM N, reduction channels per tap
  int ntaps, kchunks;          // taps, ceil(Cred / KC)
  int rank5;                   // 0: stride-1 4-D map {C,W,H,B}; 1: stride-2 5-D map {2C,W/2,2,H/2,B}
  signed char dh[kMaxTaps], dw[kMaxTaps];   // per-tap source offset (rows / cols; pair units for rank5)
  signed char ph[kMaxTaps], pw[kMaxTaps];   // rank5 only: row / column parity
  unsigned char wtap[kMaxTaps];             // per-tap index into the weight matrix' tap dimension
  bf16* out;                   // NHWC bf16 [B,MH,MW,Cout] or null
  float* out_nchw;             // fp32 [B,Cout,MH,MW] or null
  const float* bias;           // [Cout] or null
  const bf16* addend;          // same shape as out, may alias it, or null
};

template <int KC, int BN>
struct SmemLayout {
  static constexpr int kABytes = 128 * KC * 2;
  static constexpr int kBBytes = BN * KC * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;  // multiples of 1024 for every (KC,BN) used
  static constexpr int kMaxStages = 8;
  static constexpr int stages() {
    int s = (200 * 1024) / kStageBytes;
    return s > kMaxStages ? kMaxStages : s;
  }
  static constexpr int bytes() { return stages() * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/; }
};

template <int KC, int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   const FwdParams p) {
  using L = SmemLayout<KC, BN>;
  constexpr int S = L::stages();
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * L::kStageBytes);
  // bars[0..S) full, bars[S..2S) empty, bars[2S] tmem_full; then the TMEM base address
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * S);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // tile coordinates
  const int tiles_per_group = p.tiles_w * p.tiles_h;
  const int tile = blockIdx.x;
  const int grp = tile / tiles_per_group, tin = tile % tiles_per_group;
  const int b0 = grp * p.NB;
  const int h0 = (tin / p.tiles_w) * p.TH, w0 = (tin % p.tiles_w) * p.TW;
  const int n0 = blockIdx.y * BN;
  const int n_iters = p.ntaps * p.kchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
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

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        const int tap = it / p.kchunks, c0 = (it % p.kchunks) * KC;
        const uint32_t a_dst = smem_base + s * L::kStageBytes;
        const uint32_t b_dst = a_dst + L::kABytes;
        mbar_expect_tx(full_bar(s), L::kStageBytes);
        if (p.rank5)
          tma_load_5d(a_dst, &map_a, full_bar(s), p.pw[tap] * p.Cred + c0, w0 + p.dw[tap], p.ph[tap],
                      h0 + p.dh[tap], b0);
        else
          tma_load_4d(a_dst, &map_a, full_bar(s), c0, w0 + p.dw[tap], h0 + p.dh[tap], b0);
        tma_load_2d(b_dst, &map_b, full_bar(s), p.wtap[tap] * p.Cred + c0, n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, BN < 16 ? 16 : BN);
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % S;
        const uint32_t ph = (it / S) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * L::kStageBytes;
        const uint32_t b_addr = a_addr + L::kABytes;
        const uint64_t adesc = make_kmajor_desc(a_addr, KC * 2);
        const uint64_t bdesc = make_kmajor_desc(b_addr, KC * 2);
#pragma unroll
        for (int k = 0; k < KC / 16; ++k) {
          // advance 16 elements (32 bytes) along K inside the swizzle atom: +2 in the 16-byte address field
          umma_bf16(tmem_base, adesc + 2ull * k, bdesc + 2ull * k, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));  // frees the smem stage when these MMAs retire
      }
      umma_commit(tmem_full_bar);   // accumulator complete
    }
  } else {
    // ===== epilogue: 4 warps, warp w owns TMEM lanes 32*(w%4).. =====
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row of the tile = output pixel
    const int nb = r / (p.TH * p.TW);
    const int th = (r / p.TW) % p.TH, tw = r % p.TW;
    const int b = b0 + nb, h = (h0 + th) * p.os + p.oh, w = (w0 + tw) * p.os + p.ow;
    const long long pix = ((long long)b * p.OH + h) * p.OW + w;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      tmem_ld_wait();
      const int nbase = n0 + c;
      if (nbase >= p.Cout) break;
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (nbase + j < p.Cout) f[j] += __ldg(p.bias + nbase + j);
      }
      if (p.out) {
        bf16* dst = p.out + pix * p.Cout + nbase;
        const bf16* add = p.addend ? p.addend + pix * p.Cout + nbase : nullptr;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          if (nbase + j < p.Cout) {  // Cout % 8 == 0: whole 8-channel groups
            float o[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] = f[j + k];
            if (add) {
              float a8[8];
              ld_vec<8>(add + j, a8);
#pragma unroll
              for (int k = 0; k < 8; ++k) o[k] += a8[k];
            }
            st_vec<8>(dst + j, o);
          }
        }
      }
      if (p.out_nchw) {
        const long long hw = (long long)p.OH * p.OW;
        float* dst = p.out_nchw + ((long long)b * p.Cout + nbase) * hw + (long long)h * p.OW + w;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (nbase + j < p.Cout) dst[(long long)j * hw] = f[j];
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

// OHWI [O][KH][KW][I] -> flipped + transposed [I][KH][KW][O] (weights of the equivalent forward conv of dgrad)
__global__ void weight_flip_transpose_kernel(const bf16* __restrict__ w, bf16* __restrict__ wt, int O, int I, int KH,
                                             int KW) {
  const long long n = (long long)O * I * KH * KW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    // i indexes the destination [ci][kh][kw][co]
    const int co = (int)(i % O);
    long long t = i / O;
    const int kw = (int)(t % KW); t /= KW;
    const int kh = (int)(t % KH);
    const int ci = (int)(t / KH);
    wt[i] = w[(((long long)co * KH + (KH - 1 - kh)) * KW + (KW - 1 - kw)) * I + ci];
  }
}

// all conv weights of a network in one launch: table[i] = {element offset, O, I, KH, KW}.
// Per tap the job is a [O][I] -> [I][O] matrix transpose: 64 x 64 tiles through shared memory, so both the
// reads (64 consecutive ci of one co) and the writes (64 consecutive co of one ci) are full 128-byte segments.
__global__ void __launch_bounds__(256) weight_flip_transpose_batch_kernel(const bf16* __restrict__ base,
                                                                          bf16* __restrict__ out,
                                                                          const int* __restrict__ table) {
  __shared__ unsigned short tile[64][66];
  const int* e = table + 5 * blockIdx.y;
  const long long off = e[0];
  const int O = e[1], I = e[2], KH = e[3], KW = e[4];
  const unsigned short* w = reinterpret_cast<const unsigned short*>(base + off);
  unsigned short* wt = reinterpret_cast<unsigned short*>(out + off);
  const int taps = KH * KW, ot = (O + 63) / 64, it = (I + 63) / 64;
  const int ntiles = taps * ot * it;
  const int col = threadIdx.x & 63, row0 = threadIdx.x >> 6;   // 4 rows of 64 elements per pass
  for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int tap = t % taps, o0 = ((t / taps) % ot) * 64, i0 = (t / taps / ot) * 64;
    const int kh = tap / KW, kw = tap % KW;
    const int src_tap = (KH - 1 - kh) * KW + (KW - 1 - kw);
#pragma unroll 4
    for (int r = row0; r < 64; r += 4) {
      const int co = o0 + r, ci = i0 + col;
      tile[r][col] = (co < O && ci < I) ? w[((long long)co * taps + src_tap) * I + ci] : (unsigned short)0;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = row0; r < 64; r += 4) {
      const int ci = i0 + r, co = o0 + col;
      if (ci < I && co < O) wt[((long long)ci * taps + tap) * O + co] = tile[col][r];
    }
    __syncthreads();
  }
}

int floor_div2(int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); }

template <int KC, int BN>
int launch_fwd(const CUtensorMap& ma, const CUtensorMap& mb, const FwdParams& p, int n_tiles, cudaStream_t st) {
  using L = SmemLayout<KC, BN>;
  static bool configured = false;
  if (!configured) {
    UDA_CUDA_OK(cudaFuncSetAttribute(conv_tc_fwd_kernel<KC, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     L::bytes()));
    configured = true;
  }
  dim3 grid((unsigned)n_tiles, (unsigned)((p.Cout + BN - 1) / BN));
  conv_tc_fwd_kernel<KC, BN><<<grid, kTcThreads, L::bytes(), st>>>(ma, mb, p);
  UDA_LAUNCH_OK("conv_tc_fwd_kernel");
  return UDA_OK;
}

bool use_persistent() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UDA_B200_TC_PERSIST");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// Non-persistent path: one launch per tap class (conv_tc_fwd_kernel, one CTA per 128-pixel tile).
int run_gemm_conv_class(const GemmConv& g, int ci, cudaStream_t st) {
  const TapClass& c = g.cls[ci];
  const int MH = g.src_s2 ? g.SH / 2 : g.SH, MW = g.src_s2 ? g.SW / 2 : g.SW;
  const TilePlan tp = plan_tiles(g.B, MH, MW);
  const int KC = pick_kc(g.Cred), BN = pick_bn(g.Cout);
  UDA_REQUIRE(tp.ok && KC > 0 && c.ntaps >= 1 && c.ntaps <= kMaxTaps && g.Cout % 8 == 0, UDA_ERR_UNSUPPORTED,
              "conv_tc: shape not covered by the tensor-core kernel (B=%d grid=%dx%d Cred=%d Cout=%d taps=%d)", g.B,
              MH, MW, g.Cred, g.Cout, c.ntaps);
  UDA_REQUIRE(aligned<bf16>(g.src, 16) && aligned<bf16>(g.wmat, 16) && (!g.out || aligned<bf16>(g.out, 16)) &&
                  (!g.addend || aligned<bf16>(g.addend, 16)),
              UDA_ERR_BAD_ARG, "conv_tc: pointers must be 16-byte aligned");
  FwdParams p{};
  p.TW = tp.TW; p.TH = tp.TH; p.NB = tp.NB;
  p.tiles_w = MW / tp.TW; p.tiles_h = MH / tp.TH;
  p.MH = MH; p.MW = MW; p.OH = g.OH; p.OW = g.OW; p.os = g.os; p.oh = c.oh; p.ow = c.ow;
  p.Cout = g.Cout; p.Cred = g.Cred; p.ntaps = c.ntaps; p.kchunks = (g.Cred + KC - 1) / KC; p.rank5 = g.src_s2;
  for (int t = 0; t < c.ntaps; ++t) {
    p.dh[t] = (signed char)c.dh[t]; p.dw[t] = (signed char)c.dw[t];
    p.ph[t] = (signed char)c.ph[t]; p.pw[t] = (signed char)c.pw[t];
    p.wtap[t] = (unsigned char)c.wtap[t];
  }
  p.out = (bf16*)g.out; p.out_nchw = g.out_nchw; p.bias = g.bias; p.addend = (const bf16*)g.addend;

  CUtensorMap ma, mb;
  const uint64_t C = (uint64_t)g.Cred, H = (uint64_t)g.SH, W = (uint64_t)g.SW;
  if (!g.src_s2) {
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
  const int n_tiles = (g.B / tp.NB) * p.tiles_w * p.tiles_h;
#define UDA_TC(KCv, BNv) \
  if (KC == KCv && BN == BNv) return launch_fwd<KCv, BNv>(ma, mb, p, n_tiles, st);
  UDA_TC(64, 128) UDA_TC(64, 64) UDA_TC(64, 32)
  UDA_TC(32, 128) UDA_TC(32, 64) UDA_TC(32, 32)
  UDA_TC(16, 128) UDA_TC(16, 64) UDA_TC(16, 32)
#undef UDA_TC
  return set_error(UDA_ERR_UNSUPPORTED, "conv_tc: no kernel instance for KC=%d BN=%d", KC, BN);
}

bool use_halo() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UDA_B200_TC_HALO");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

int run_gemm_conv(const GemmConv& g, cudaStream_t st) {
  if (use_persistent()) {
    if (use_halo()) {
      int rc = run_gemm_conv_halo(g, st);
      if (rc != UDA_ERR_UNSUPPORTED) return rc;
      rc = run_gemm_conv_phalo(g, st);
      if (rc != UDA_ERR_UNSUPPORTED) return rc;
    }
    return run_gemm_conv_persistent(g, st);
  }
  if (g.fuse) return UDA_ERR_UNSUPPORTED;
  for (int ci = 0; ci < g.ncls; ++ci)
    if (int rc = run_gemm_conv_class(g, ci, st)) return rc;
  return UDA_OK;
}

bool fwd_shape_ok(int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad) {
  if (pick_kc(Cin) == 0 || Cout % 8 || Cout < 8) return false;
  if (KH * KW > kMaxTaps || KH != KW) return false;
  if (stride != 1 && stride != 2) return false;
  if (stride == 2 && (H % 2 || W % 2)) return false;
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  if (Ho <= 0 || Wo <= 0) return false;
  if (stride == 1 && (Ho != H || Wo != W)) return false;  // "same" convolutions only
  if (stride == 2 && (Ho != H / 2 || Wo != W / 2)) return false;
  return plan_tiles(B, Ho, Wo).ok;
}

// x: [B,H,W,Cin] bf16; w: [Cout][KH*KW][Cin] bf16; "same" (stride 1) or halving (stride 2) forward convolution
int run_fwd(const void* x, const void* w, const float* bias, const void* addend, void* y, float* y_nchw,
            double* bn_sums, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad,
            cudaStream_t st, int act = 0, float act_slope = 0.f, const BnFuse* fuse = nullptr) {
  UDA_REQUIRE(fwd_shape_ok(B, H, W, Cin, Cout, KH, KW, stride, pad), UDA_ERR_UNSUPPORTED,
              "conv_tc: shape not covered by the tensor-core kernel (B=%d H=%d W=%d Cin=%d Cout=%d k=%d s=%d p=%d)",
              B, H, W, Cin, Cout, KH, stride, pad);
  GemmConv g{};
  g.src = x; g.B = B; g.SH = H; g.SW = W; g.Cred = Cin; g.src_s2 = stride == 2;
  g.wmat = w; g.Cout = Cout; g.wtaps = KH * KW; g.ncls = 1;
  TapClass& c = g.cls[0];
  c.ntaps = KH * KW; c.oh = 0; c.ow = 0;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      const int t = kh * KW + kw, oh = kh - pad, ow = kw - pad;
      c.wtap[t] = t;
      if (stride == 1) {
        c.dh[t] = oh; c.dw[t] = ow; c.ph[t] = 0; c.pw[t] = 0;
      } else {
        c.dh[t] = floor_div2(oh); c.dw[t] = floor_div2(ow);
        c.ph[t] = oh - 2 * c.dh[t]; c.pw[t] = ow - 2 * c.dw[t];
      }
    }
  g.OH = stride == 1 ? H : H / 2; g.OW = stride == 1 ? W : W / 2; g.os = 1;
  g.bias = bias; g.addend = addend; g.out = y; g.out_nchw = y_nchw; g.bn_sums = bn_sums;
  g.act = act; g.act_slope = act_slope;
  g.fuse = fuse;
  return run_gemm_conv(g, st);
}

bool dgrad_shape_ok(int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad) {
  if (KH != KW || KH * KW > kMaxTaps) return false;
  if (pick_kc(Cout) == 0 || Cin % 8 || Cin < 8) return false;
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  if (stride == 1) return Ho == H && Wo == W && plan_tiles(B, H, W).ok;
  if (stride == 2) return H % 2 == 0 && W % 2 == 0 && Ho == H / 2 && Wo == W / 2 && plan_tiles(B, Ho, Wo).ok;
  return false;
}

// dX[B,H,W,Cin] (+= addend) from dY[B,Ho,Wo,Cout] with w_ft = [Cin][KH][KW][Cout] (flipped + transposed).
// stride 1: one "same" forward convolution of dY.  stride 2: one launch per output parity (ph,pw): the taps
// with kh = ph+pad (mod 2) read dY[i + (ph+pad-kh)/2, ...] and write dX[2i+ph, 2j+pw] (every dX pixel is
// produced by exactly one launch; parities without any tap are zero).
int run_dgrad(const void* dy, const void* w_ft, const void* addend, void* dx, int B, int H, int W, int Cin, int Cout,
              int KH, int KW, int stride, int pad, cudaStream_t st, const void* st_a = nullptr,
              const void* st_z = nullptr, float st_slope = 0.f, double* st_sums = nullptr, const float* bias = nullptr,
              double* bn_sums = nullptr, int act = 0, float act_slope = 0.f) {
  UDA_REQUIRE(dgrad_shape_ok(B, H, W, Cin, Cout, KH, KW, stride, pad), UDA_ERR_UNSUPPORTED,
              "conv_tc_dgrad: shape not covered (B=%d H=%d W=%d Cin=%d Cout=%d k=%d s=%d p=%d)", B, H, W, Cin, Cout,
              KH, stride, pad);
  UDA_REQUIRE(!st_sums || (use_persistent() && st_a && !(stride == 2 && KH < 2)), UDA_ERR_UNSUPPORTED,
              "conv_tc_dgrad: BatchNorm-backward statistics need the persistent kernels, `a`, and a dgrad that "
              "writes every input pixel (not 1x1 stride 2)");
  GemmConv g{};
  g.st_a = st_a; g.st_z = st_z; g.st_slope = st_slope; g.st_sums = st_sums;
  g.src = dy; g.B = B; g.Cred = Cout; g.src_s2 = 0;
  g.wmat = w_ft; g.Cout = Cin; g.wtaps = KH * KW;
  g.OH = H; g.OW = W; g.bias = bias; g.addend = addend; g.out = dx; g.out_nchw = nullptr;
  g.bn_sums = bn_sums; g.act = act; g.act_slope = act_slope;
  if (stride == 1) {
    g.SH = H; g.SW = W; g.os = 1; g.ncls = 1;
    TapClass& c = g.cls[0];
    c.ntaps = KH * KW; c.oh = 0; c.ow = 0;
    const int padp = KH - 1 - pad;
    for (int kh = 0; kh < KH; ++kh)
      for (int kw = 0; kw < KW; ++kw) {
        const int t = kh * KW + kw;
        c.dh[t] = kh - padp; c.dw[t] = kw - padp; c.ph[t] = c.pw[t] = 0; c.wtap[t] = t;  // w_ft is already flipped
      }
    return run_gemm_conv(g, st);
  }
  g.SH = H / 2; g.SW = W / 2; g.os = 2; g.ncls = 0;
  if (KH < 2 && !addend)  // some output parities receive no tap at all: they are zero
    UDA_CUDA_OK(cudaMemsetAsync(dx, 0, (size_t)B * H * W * Cin * 2, st));
  // classes with the most taps first (static round-robin of the persistent kernel balances better)
  for (int ph = 1; ph >= 0; --ph)
    for (int pw = 1; pw >= 0; --pw) {
        TapClass c{};
        int n = 0;
        for (int kh = 0; kh < KH; ++kh) {
          if (((ph + pad - kh) % 2 + 2) % 2) continue;
          for (int kw = 0; kw < KW; ++kw) {
            if (((pw + pad - kw) % 2 + 2) % 2) continue;
            c.dh[n] = floor_div2(ph + pad - kh); c.dw[n] = floor_div2(pw + pad - kw);
            c.ph[n] = c.pw[n] = 0;
            c.wtap[n] = (KH - 1 - kh) * KW + (KW - 1 - kw);  // position of tap (kh,kw) inside the flipped w_ft
            ++n;
          }
        }
        if (n == 0) continue;   // parities without any tap: zero gradient (memset above)
        c.ntaps = n; c.oh = ph; c.ow = pw;
        g.cls[g.ncls++] = c;
    }
  return run_gemm_conv(g, st);
}

// ================================================================================================
// wgrad:  dW[co][tap][ci] += sum_pixels dY[pix][co] * X[pix + tap][ci]
//
// GEMM with the pixel index as the reduction dimension.  Both operands are read straight from the NHWC
// tensors with the SAME TMA boxes as the forward kernel (128 pixels x channel atom), i.e. they sit in
// shared memory as [pixel rows][channels contiguous] — MN-major operands for tcgen05.mma (a_major =
// b_major = 1).  The 128 rows of D are (tap, ci) pairs: 128/ATOM_A channel atoms, each loaded with its own
// tap shift; the columns are output channels (BN/ATOM_B atoms).  A CTA accumulates a range of pixel tiles
// in TMEM and adds its partial result to the fp32 gradient with one atomic per element.
// ================================================================================================
struct WgradParams {
  int TW, TH, NB, tiles_w, tiles_h;     // pixel tiles over the OUTPUT grid (Ho x Wo)
  int Cin, Cout, ntaps, cchunks;        // cchunks = Cin / ATOM_A
  int rank5;                            // input read through the stride-2 space-to-depth view
  signed char dh[kMaxTaps], dw[kMaxTaps], ph[kMaxTaps], pw[kMaxTaps];
  int n_pixel_tiles, tiles_per_split;
  float* dw_out;                        // [Cout][ntaps][Cin] fp32
};

__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t smem_addr, uint32_t row_bytes, uint32_t atom_bytes) {
  // MN-major canonical layout (units of 16 B): ((8,n),(8,k)) : ((1,LBO),(8,SBO)) for 128-byte rows —
  // rows (k = pixels) are row_bytes apart, 8-row groups SBO = 8*row_bytes apart, channel atoms LBO apart.
  const uint64_t layout = row_bytes == 128 ? 2ull : (row_bytes == 64 ? 4ull : 6ull);
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((atom_bytes >> 4) & 0x3FFFu) << 16;   // LBO
  d |= (uint64_t)((8u * row_bytes) >> 4) << 32;          // SBO
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

template <int ATOM_A, int ATOM_B, int BN>
struct WgradSmem {
  static constexpr int kAAtoms = 128 / ATOM_A;
  static constexpr int kBAtoms = BN / ATOM_B;
  static constexpr int kAAtomBytes = 128 * ATOM_A * 2;
  static constexpr int kBAtomBytes = 128 * ATOM_B * 2;
  static constexpr int kABytes = kAAtoms * kAAtomBytes;   // 32 KB
  static constexpr int kBBytes = kBAtoms * kBAtomBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int stages() {
    int s = (196 * 1024) / kStageBytes;
    return s > 6 ? 6 : s;
  }
  static constexpr int bytes() { return stages() * kStageBytes + 1024 + 256; }
};

template <int ATOM_A, int ATOM_B, int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                     const WgradParams p) {
  using L = WgradSmem<ATOM_A, ATOM_B, BN>;
  constexpr int S = L::stages();
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * L::kStageBytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * S);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int mtile = blockIdx.x;
  const int n0 = blockIdx.y * BN;
  const int total_atoms = p.ntaps * p.cchunks;
  int valid_atoms = total_atoms - mtile * L::kAAtoms;
  if (valid_atoms > L::kAAtoms) valid_atoms = L::kAAtoms;
  const int pt_begin = blockIdx.z * p.tiles_per_split;
  int pt_end = pt_begin + p.tiles_per_split;
  if (pt_end > p.n_pixel_tiles) pt_end = p.n_pixel_tiles;
  const int n_iters = pt_end - pt_begin;   // >= 1 by construction of the grid
  const int tiles_per_group = p.tiles_w * p.tiles_h;

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

  if (warp == 0) {
    if (elect_one()) {
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % S;
        const uint32_t phs = (it / S) & 1;
        mbar_wait(empty_bar(s), phs ^ 1);
        const int tile = pt_begin + it;
        const int grp = tile / tiles_per_group, tin = tile % tiles_per_group;
        const int b0 = grp * p.NB, h0 = (tin / p.tiles_w) * p.TH, w0 = (tin % p.tiles_w) * p.TW;
        const uint32_t a_dst = smem_base + s * L::kStageBytes;
        const uint32_t b_dst = a_dst + L::kABytes;
        mbar_expect_tx(full_bar(s), valid_atoms * L::kAAtomBytes + L::kBBytes);
        for (int a = 0; a < valid_atoms; ++a) {
          const int gidx = mtile * L::kAAtoms + a;
          const int tap = gidx / p.cchunks, c0 = (gidx % p.cchunks) * ATOM_A;
          if (p.rank5)
            tma_load_5d(a_dst + a * L::kAAtomBytes, &map_x, full_bar(s), p.pw[tap] * p.Cin + c0, w0 + p.dw[tap],
                        p.ph[tap], h0 + p.dh[tap], b0);
          else
            tma_load_4d(a_dst + a * L::kAAtomBytes, &map_x, full_bar(s), c0, w0 + p.dw[tap], h0 + p.dh[tap], b0);
        }
#pragma unroll
        for (int j = 0; j < L::kBAtoms; ++j)
          tma_load_4d(b_dst + j * L::kBAtomBytes, &map_dy, full_bar(s), n0 + j * ATOM_B, w0, h0, b0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // M = 128 rows (tap,ci), N = BN output channels, both operands MN-major (bits 15 and 16)
      constexpr uint32_t idesc = make_idesc_bf16(128, BN < 16 ? 16 : BN) | (1u << 15) | (1u << 16);
      for (int it = 0; it < n_iters; ++it) {
        const int s = it % S;
        const uint32_t phs = (it / S) & 1;
        mbar_wait(full_bar(s), phs);
        tc_fence_after();
        const uint32_t a_addr = smem_base + s * L::kStageBytes;
        const uint32_t b_addr = a_addr + L::kABytes;
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // 16 pixels per MMA
          const uint64_t adesc = make_mnmajor_desc(a_addr + k * 16 * (ATOM_A * 2), ATOM_A * 2, L::kAAtomBytes);
          const uint64_t bdesc = make_mnmajor_desc(b_addr + k * 16 * (ATOM_B * 2), ATOM_B * 2, L::kBAtomBytes);
          umma_bf16(tmem_base, adesc, bdesc, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int a = r / ATOM_A, j = r % ATOM_A;
    const int gidx = mtile * L::kAAtoms + a;
    const bool row_ok = a < valid_atoms;
    const int tap = row_ok ? gidx / p.cchunks : 0;
    const int ci = row_ok ? (gidx % p.cchunks) * ATOM_A + j : 0;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t v[32];
      if (BN >= 32) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      } else {
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16), v);  // 32 columns are allocated; first BN are valid
      }
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const int co = n0 + c + k;
          if (c + k < BN && co < p.Cout)
            atomicAdd(p.dw_out + ((long long)co * p.ntaps + tap) * p.Cin + ci, __uint_as_float(v[k]));
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

int pick_atom(int c) { return c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : (c % 16 == 0 ? 16 : 0)); }

bool wgrad_shape_ok(int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad) {
  if (KH != KW || KH * KW > kMaxTaps) return false;
  if (pick_atom(Cin) == 0 || Cout % 8 || Cout < 8) return false;
  if (stride != 1 && stride != 2) return false;
  if (stride == 2 && (H % 2 || W % 2)) return false;
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  if (stride == 1 && (Ho != H || Wo != W)) return false;
  if (stride == 2 && (Ho != H / 2 || Wo != W / 2)) return false;
  return plan_tiles(B, Ho, Wo).ok;
}

template <int ATOM_A, int ATOM_B, int BN>
int launch_wgrad(const CUtensorMap& mx, const CUtensorMap& mdy, WgradParams& p, cudaStream_t st) {
  using L = WgradSmem<ATOM_A, ATOM_B, BN>;
  static bool configured = false;
  if (!configured) {
    UDA_CUDA_OK(cudaFuncSetAttribute(conv_tc_wgrad_kernel<ATOM_A, ATOM_B, BN>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, L::bytes()));
    configured = true;
  }
  const int m_tiles = (p.ntaps * p.cchunks + L::kAAtoms - 1) / L::kAAtoms;
  const int n_tiles = (p.Cout + BN - 1) / BN;
  int splits = (num_sms() + m_tiles * n_tiles - 1) / (m_tiles * n_tiles);
  if (splits > p.n_pixel_tiles) splits = p.n_pixel_tiles;
  if (splits < 1) splits = 1;
  p.tiles_per_split = (p.n_pixel_tiles + splits - 1) / splits;
  splits = (p.n_pixel_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  dim3 grid((unsigned)m_tiles, (unsigned)n_tiles, (unsigned)splits);
  conv_tc_wgrad_kernel<ATOM_A, ATOM_B, BN><<<grid, kTcThreads, L::bytes(), st>>>(mx, mdy, p);
  UDA_LAUNCH_OK("conv_tc_wgrad_kernel");
  return UDA_OK;
}

int run_wgrad(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin, int Cout, int KH, int KW,
              int stride, int pad, cudaStream_t st) {
  UDA_REQUIRE(wgrad_shape_ok(B, H, W, Cin, Cout, KH, KW, stride, pad), UDA_ERR_UNSUPPORTED,
              "conv_tc_wgrad: shape not covered (B=%d H=%d W=%d Cin=%d Cout=%d k=%d s=%d p=%d)", B, H, W, Cin, Cout,
              KH, stride, pad);
  UDA_REQUIRE(aligned<bf16>(dy, 16) && aligned<bf16>(x, 16) && aligned<float>(dw, 4), UDA_ERR_BAD_ARG,
              "conv_tc_wgrad: pointers must be 16-byte aligned");
  if (KH == 3 && KW == 3 && stride == 1 && pad == 1 && use_persistent() && use_halo()) {
    const int rc = run_wgrad_halo(dy, x, dw, B, H, W, Cin, Cout, st);
    if (rc != UDA_ERR_UNSUPPORTED) return rc;
  }
  if (KH == 4 && KW == 4 && stride == 2 && pad == 1 && use_persistent() && use_halo()) {   // wide 16-channel inputs
    const int rc = run_wgrad_downhalo(dy, x, dw, B, H, W, Cin, Cout, st);
    if (rc != UDA_ERR_UNSUPPORTED) return rc;
  }
  if (use_persistent()) {
    const int rc = run_wgrad_big(dy, x, dw, B, H, W, Cin, Cout, KH, KW, stride, pad, st);
    if (rc != UDA_ERR_UNSUPPORTED) return rc;
  }
  const int Ho = stride == 1 ? H : H / 2, Wo = stride == 1 ? W : W / 2;
  const TilePlan tp = plan_tiles(B, Ho, Wo);
  const int atomA = pick_atom(Cin);
  // output-channel atoms: 64-wide when possible, else 32 / 16 (the 24-class head: two 16-wide atoms, the
  // upper half of the second one is zero-filled by TMA)
  const int atomB = Cout % 64 == 0 ? 64 : (Cout % 32 == 0 ? 32 : 16);
  int BN = Cout >= 128 ? 128 : (Cout + atomB - 1) / atomB * atomB;
  if (BN > 128) BN = 128;
  WgradParams p{};
  p.TW = tp.TW; p.TH = tp.TH; p.NB = tp.NB; p.tiles_w = Wo / tp.TW; p.tiles_h = Ho / tp.TH;
  p.Cin = Cin; p.Cout = Cout; p.ntaps = KH * KW; p.cchunks = Cin / atomA; p.rank5 = stride == 2;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      const int t = kh * KW + kw, oh = kh - pad, ow = kw - pad;
      if (stride == 1) {
        p.dh[t] = (signed char)oh; p.dw[t] = (signed char)ow; p.ph[t] = p.pw[t] = 0;
      } else {
        const int ah = floor_div2(oh), aw = floor_div2(ow);
        p.dh[t] = (signed char)ah; p.dw[t] = (signed char)aw;
        p.ph[t] = (signed char)(oh - 2 * ah); p.pw[t] = (signed char)(ow - 2 * aw);
      }
    }
  p.n_pixel_tiles = (B / tp.NB) * p.tiles_w * p.tiles_h;
  p.dw_out = dw;
  CUtensorMap mx, mdy;
  const uint64_t C = (uint64_t)Cin;
  if (stride == 1) {
    uint64_t dims[4] = {C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {(uint32_t)atomA, (uint32_t)tp.TW, (uint32_t)tp.TH, (uint32_t)tp.NB};
    if (int rc = make_tmap_bf16(&mx, x, 4, dims, str, box, atomA * 2)) return rc;
  } else {
    uint64_t dims[5] = {2 * C, (uint64_t)W / 2, 2, (uint64_t)H / 2, (uint64_t)B};
    uint64_t str[4] = {2 * C * 2, (uint64_t)W * C * 2, 2 * (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[5] = {(uint32_t)atomA, (uint32_t)tp.TW, 1, (uint32_t)tp.TH, (uint32_t)tp.NB};
    if (int rc = make_tmap_bf16(&mx, x, 5, dims, str, box, atomA * 2)) return rc;
  }
  {
    const uint64_t Co = (uint64_t)Cout;
    uint64_t dims[4] = {Co, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
    uint64_t str[3] = {Co * 2, (uint64_t)Wo * Co * 2, (uint64_t)Ho * Wo * Co * 2};
    uint32_t box[4] = {(uint32_t)atomB, (uint32_t)tp.TW, (uint32_t)tp.TH, (uint32_t)tp.NB};
    if (int rc = make_tmap_bf16(&mdy, dy, 4, dims, str, box, atomB * 2)) return rc;
  }
#define UDA_WG(A, Bv, N) \
  if (atomA == A && atomB == Bv && BN == N) return launch_wgrad<A, Bv, N>(mx, mdy, p, st);
  UDA_WG(64, 64, 128) UDA_WG(64, 64, 64) UDA_WG(64, 32, 32) UDA_WG(64, 16, 16) UDA_WG(64, 16, 32)
  UDA_WG(32, 64, 128) UDA_WG(32, 64, 64) UDA_WG(32, 32, 32) UDA_WG(32, 16, 16) UDA_WG(32, 16, 32)
  UDA_WG(16, 64, 128) UDA_WG(16, 64, 64) UDA_WG(16, 32, 32) UDA_WG(16, 16, 16) UDA_WG(16, 16, 32)
#undef UDA_WG
  return set_error(UDA_ERR_UNSUPPORTED, "conv_tc_wgrad: no kernel instance for atoms %d/%d BN=%d", atomA, atomB, BN);
}

// ================================================================================================
// Cin = 3 stems (U-Net 7x7 s2 p3, discriminator 4x4 s2 p1) on the tensor cores.
//
// The image is repacked once per forward into a zero-padded 4-channel bf16 buffer Xs[B][H+8][W+8][4]
// (top/left border = pad).  For output pixel (ho,wo) and kernel row kh, the KS = 8 (k=7) or 4 (k=4)
// consecutive padded pixels starting at column 2*wo of padded row 2*ho+kh are KS*4 contiguous elements:
// a K chunk of 32 / 16 channels.  Consecutive output pixels start 2 pixels = 16 bytes apart, so the A
// operand is an OVERLAPPING-stride TMA view {KS*4, Wo (16 B), 2 (row parity), (H+8)/2 (row pairs), B}: each
// kernel row is one tap of a K=KS*4 implicit GEMM against weights repacked to [Cout][k][KS*4] (zero in the
// padded slots).  Same kernels as every other convolution; only the tensor map differs.
// ================================================================================================
constexpr int kStemPad = 8;   // Xs rows = H + 8, cols = W + 8

bool stem_shape_ok(int B, int H, int W, int Cin, int Cout, int K, int stride, int pad) {
  if (Cin != 3 || stride != 2 || H % 2 || W % 2 || Cout % 8 || Cout < 8) return false;
  if (!((K == 7 && pad == 3) || (K == 4 && pad == 1))) return false;
  if ((H + 2 * pad - K) / 2 + 1 != H / 2 || (W + 2 * pad - K) / 2 + 1 != W / 2) return false;
  return plan_tiles(B, H / 2, W / 2).ok;
}
inline int stem_slots(int K) { return K == 7 ? 8 : 4; }

__global__ void stem_pack_input_kernel(const float* __restrict__ x, bf16* __restrict__ xs, int B, int H, int W,
                                       int pad) {
  const int Hp = H + kStemPad, Wp = W + kStemPad;
  const long long total = (long long)B * Hp * Wp;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int wp = (int)(i % Wp);
    long long t = i / Wp;
    const int hp = (int)(t % Hp);
    const int b = (int)(t / Hp);
    const int h = hp - pad, w = wp - pad;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (h >= 0 && h < H && w >= 0 && w < W) {
      const float* s = x + ((long long)b * 3 * H + h) * W + w;
      v[0] = __ldg(s); v[1] = __ldg(s + (long long)H * W); v[2] = __ldg(s + 2LL * H * W);
    }
    st_vec<4>(xs + i * 4, v);
  }
}
// w [Cout][K][K][3] bf16 -> ws [Cout][K][KS*4] bf16 (slot kw*4+c; zero elsewhere)
__global__ void stem_pack_weight_kernel(const bf16* __restrict__ w, bf16* __restrict__ ws, int Cout, int K, int KS) {
  const int n = Cout * K * KS * 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = i % 4, kw = (i / 4) % KS, kh = (i / (4 * KS)) % K, co = i / (4 * KS * K);
    ws[i] = (c < 3 && kw < K) ? w[((co * K + kh) * K + kw) * 3 + c] : __float2bfloat16_rn(0.f);
  }
}
// dw [Cout][K][K][3] fp32 += dws [Cout][K][KS*4]
__global__ void stem_unpack_wgrad_kernel(const float* __restrict__ dws, float* __restrict__ dw, int Cout, int K,
                                         int KS) {
  const int n = Cout * K * K * 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int c = i % 3, kw = (i / 3) % K, kh = (i / (3 * K)) % K, co = i / (3 * K * K);
    dw[i] += dws[((co * K + kh) * KS + kw) * 4 + c];
  }
}

int make_stem_map(CUtensorMap* m, const void* xs, int B, int H, int W, int K, const TilePlan& tp) {
  const int KS = stem_slots(K);
  const uint64_t Hp = H + kStemPad, Wp = W + kStemPad;
  uint64_t dims[5] = {(uint64_t)KS * 4, (uint64_t)W / 2, 2, Hp / 2, (uint64_t)B};
  uint64_t str[4] = {16, Wp * 8, 2 * Wp * 8, Hp * Wp * 8};
  uint32_t box[5] = {(uint32_t)KS * 4, (uint32_t)tp.TW, 1, (uint32_t)tp.TH, (uint32_t)tp.NB};
  return make_tmap_bf16(m, xs, 5, dims, str, box, KS * 8);
}

int run_stem_fwd(const void* xs, const void* ws, const float* bias, void* y, double* bn_sums, int act, float act_slope,
                 int B, int H, int W,
                 int Cout, int K, cudaStream_t st) {
  const TilePlan tp = plan_tiles(B, H / 2, W / 2);
  CUtensorMap ma;
  if (int rc = make_stem_map(&ma, xs, B, H, W, K, tp)) return rc;
  GemmConv g{};
  g.src = xs; g.B = B; g.SH = H; g.SW = W; g.Cred = stem_slots(K) * 4; g.src_s2 = 1;
  g.wmat = ws; g.Cout = Cout; g.wtaps = K; g.ncls = 1;
  TapClass& c = g.cls[0];
  c.ntaps = K; c.oh = 0; c.ow = 0;
  for (int kh = 0; kh < K; ++kh) { c.dh[kh] = kh >> 1; c.ph[kh] = kh & 1; c.dw[kh] = 0; c.pw[kh] = 0; c.wtap[kh] = kh; }
  g.OH = H / 2; g.OW = W / 2; g.os = 1;
  g.bias = bias; g.addend = nullptr; g.out = y; g.out_nchw = nullptr; g.bn_sums = bn_sums;
  g.act = act; g.act_slope = act_slope;
  g.a_map = &ma; g.a_MH = H / 2; g.a_MW = W / 2; g.a_kc = stem_slots(K) * 4;
  return run_gemm_conv_persistent(g, st);
}

int run_stem_wgrad(const void* dy, const void* xs, float* dws, int B, int H, int W, int Cout, int K,
                   cudaStream_t st) {
  const int Ho = H / 2, Wo = W / 2, KS = stem_slots(K);
  const TilePlan tp = plan_tiles(B, Ho, Wo);
  const int atomA = KS * 4;   // 32 or 16 "channels" per kernel row
  const int atomB = Cout % 64 == 0 ? 64 : (Cout % 32 == 0 ? 32 : 16);
  int BN = Cout >= 128 ? 128 : (Cout + atomB - 1) / atomB * atomB;
  WgradParams p{};
  p.TW = tp.TW; p.TH = tp.TH; p.NB = tp.NB; p.tiles_w = Wo / tp.TW; p.tiles_h = Ho / tp.TH;
  p.Cin = atomA; p.Cout = Cout; p.ntaps = K; p.cchunks = 1; p.rank5 = 1;
  for (int kh = 0; kh < K; ++kh) { p.dh[kh] = (signed char)(kh >> 1); p.ph[kh] = (signed char)(kh & 1); p.dw[kh] = 0; p.pw[kh] = 0; }
  p.n_pixel_tiles = (B / tp.NB) * p.tiles_w * p.tiles_h;
  p.dw_out = dws;
  CUtensorMap mx, mdy;
  if (int rc = make_stem_map(&mx, xs, B, H, W, K, tp)) return rc;
  {
    const uint64_t Co = (uint64_t)Cout;
    uint64_t dims[4] = {Co, (uint64_t)Wo, (uint64_t)Ho, (uint64_t)B};
    uint64_t str[3] = {Co * 2, (uint64_t)Wo * Co * 2, (uint64_t)Ho * Wo * Co * 2};
    uint32_t box[4] = {(uint32_t)atomB, (uint32_t)tp.TW, (uint32_t)tp.TH, (uint32_t)tp.NB};
    if (int rc = make_tmap_bf16(&mdy, dy, 4, dims, str, box, atomB * 2)) return rc;
  }
#define UDA_WG(A, Bv, N) \
  if (atomA == A && atomB == Bv && BN == N) return launch_wgrad<A, Bv, N>(mx, mdy, p, st);
  UDA_WG(32, 64, 128) UDA_WG(32, 64, 64) UDA_WG(32, 32, 32) UDA_WG(32, 16, 16) UDA_WG(32, 16, 32)
  UDA_WG(16, 64, 128) UDA_WG(16, 64, 64) UDA_WG(16, 32, 32) UDA_WG(16, 16, 16) UDA_WG(16, 16, 32)
#undef UDA_WG
  return set_error(UDA_ERR_UNSUPPORTED, "stem_wgrad: no kernel instance for atoms %d/%d BN=%d", atomA, atomB, BN);
}

}  // namespace
}  // namespace uda

using namespace uda;

// ---- Cin = 3 stem entry points (xs: packed input of uda_stem_pack_input; ws: uda_stem_pack_weight) ----
#ifdef UDA_B200_EXPERIMENTS
namespace uda { namespace tcconv { long long* g_trace_buf = nullptr; int g_trace_series_left = 0; } }
// experiment builds only: device buffer of 16 int64 per CTA (see conv_tc_internal.cuh), or null to switch tracing off
extern "C" int uda_exp_set_trace(void* buf) {
  uda::tcconv::g_trace_buf = (long long*)buf; uda::tcconv::g_trace_series_left = 0; return 0;
}
// ... or `n` slices of 148 x 16 int64: every traced launch takes the next one (returns the slices still unused when
// called with buf == null)
extern "C" int uda_exp_set_trace_series(void* buf, int n) {
  const int left = uda::tcconv::g_trace_series_left;
  uda::tcconv::g_trace_buf = (long long*)buf; uda::tcconv::g_trace_series_left = buf ? n : 0;
  return left;
}
#endif

extern "C" int uda_stem_tc_supported(int B, int H, int W, int Cin, int Cout, int K, int stride, int pad) {
  if (!uda_device_supported() || !use_persistent()) return 0;
  return stem_shape_ok(B, H, W, Cin, Cout, K, stride, pad) ? 1 : 0;
}
extern "C" size_t uda_stem_packed_input_elems(int B, int H, int W) {
  return (size_t)B * (H + kStemPad) * (W + kStemPad) * 4;
}
extern "C" int uda_stem_pack_input(const float* x_nchw, void* xs, int B, int H, int W, int pad, void* stream) {
  UDA_REQUIRE(x_nchw && xs && B > 0 && H > 0 && W > 0 && pad >= 0 && pad <= 3, UDA_ERR_BAD_ARG, "stem_pack_input: bad argument");
  const long long total = (long long)B * (H + kStemPad) * (W + kStemPad);
  long long blocks = (total + 255) / 256;
  if (blocks > 8LL * num_sms()) blocks = 8LL * num_sms();
  stem_pack_input_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x_nchw, (bf16*)xs, B, H, W, pad);
  UDA_LAUNCH_OK("stem_pack_input_kernel");
  return UDA_OK;
}
extern "C" int uda_stem_pack_weight(const void* w, void* ws, int Cout, int K, void* stream) {
  UDA_REQUIRE(w && ws && Cout > 0 && (K == 7 || K == 4), UDA_ERR_BAD_ARG, "stem_pack_weight: bad argument");
  const int KS = stem_slots(K), n = Cout * K * KS * 4;
  stem_pack_weight_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const bf16*)w, (bf16*)ws, Cout, K, KS);
  UDA_LAUNCH_OK("stem_pack_weight_kernel");
  return UDA_OK;
}
extern "C" int uda_stem_tc_fwd(const void* xs, const void* ws, const float* bias, void* y, double* bn_sums, int B,
                               int H, int W, int Cout, int K, int pad, void* stream) {
  UDA_REQUIRE(xs && ws && y, UDA_ERR_BAD_ARG, "stem_tc_fwd: null pointer");
  UDA_REQUIRE(stem_shape_ok(B, H, W, 3, Cout, K, 2, pad), UDA_ERR_UNSUPPORTED, "stem_tc_fwd: shape not covered");
  return run_stem_fwd(xs, ws, bias, y, bn_sums, 0, 0.f, B, H, W, Cout, K, (cudaStream_t)stream);
}
// stem forward with the activation in the epilogue (eval mode: BatchNorm folded into ws / bias)
extern "C" int uda_stem_tc_fwd_act(const void* xs, const void* ws, const float* bias, void* y, float act_slope, int B,
                                   int H, int W, int Cout, int K, int pad, void* stream) {
  UDA_REQUIRE(xs && ws && y, UDA_ERR_BAD_ARG, "stem_tc_fwd_act: null pointer");
  UDA_REQUIRE(stem_shape_ok(B, H, W, 3, Cout, K, 2, pad), UDA_ERR_UNSUPPORTED, "stem_tc_fwd_act: shape not covered");
  return run_stem_fwd(xs, ws, bias, y, nullptr, 1, act_slope, B, H, W, Cout, K, (cudaStream_t)stream);
}
// dw [Cout][K][K][3] fp32 += wgrad; dws_scratch: fp32 [Cout][K][KS*4] scratch (KS = 8 for K = 7, 4 for K = 4)
extern "C" int uda_stem_tc_wgrad(const void* dy, const void* xs, float* dw, float* dws_scratch, int B, int H, int W,
                                 int Cout, int K, int pad, void* stream) {
  UDA_REQUIRE(dy && xs && dw && dws_scratch, UDA_ERR_BAD_ARG, "stem_tc_wgrad: null pointer");
  UDA_REQUIRE(stem_shape_ok(B, H, W, 3, Cout, K, 2, pad), UDA_ERR_UNSUPPORTED, "stem_tc_wgrad: shape not covered");
  cudaStream_t st = (cudaStream_t)stream;
  const int KS = stem_slots(K);
  UDA_CUDA_OK(cudaMemsetAsync(dws_scratch, 0, (size_t)Cout * K * KS * 4 * sizeof(float), st));
  if (int rc = run_stem_wgrad(dy, xs, dws_scratch, B, H, W, Cout, K, st)) return rc;
  const int n = Cout * K * K * 3;
  stem_unpack_wgrad_kernel<<<(n + 255) / 256, 256, 0, st>>>(dws_scratch, dw, Cout, K, KS);
  UDA_LAUNCH_OK("stem_unpack_wgrad_kernel");
  return UDA_OK;
}

extern "C" int uda_conv2d_tc_supported(int op, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride,
                                       int pad) {
  if (!uda_device_supported()) return 0;
  if (op == 0) return fwd_shape_ok(B, H, W, Cin, Cout, KH, KW, stride, pad) ? 1 : 0;
  if (op == 1) return dgrad_shape_ok(B, H, W, Cin, Cout, KH, KW, stride, pad) ? 1 : 0;
  if (op == 2) return wgrad_shape_ok(B, H, W, Cin, Cout, KH, KW, stride, pad) ? 1 : 0;
  return 0;
}

extern "C" int uda_conv2d_tc_fwd(const void* x, const void* w, const float* bias, void* y_nhwc, float* y_nchw_f32,
                                 double* bn_sums, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride,
                                 int pad, void* stream) {
  UDA_REQUIRE(x && w && (y_nhwc || y_nchw_f32), UDA_ERR_BAD_ARG, "conv_tc_fwd: null pointer");
  UDA_REQUIRE(bn_sums == nullptr || use_persistent(), UDA_ERR_UNSUPPORTED,
              "conv_tc_fwd: fused BN statistics need the persistent kernels");
  if (KH == 4 && KW == 4 && stride == 2 && pad == 1 && !bias && !bn_sums && y_nhwc && !y_nchw_f32 && use_persistent()) {
    const int rc = run_downconv_halo(x, w, y_nhwc, B, H, W, Cin, Cout, (cudaStream_t)stream);   // wide 16-channel inputs
    if (rc != UDA_ERR_UNSUPPORTED) return rc;
  }
  return run_fwd(x, w, bias, nullptr, y_nhwc, y_nchw_f32, bn_sums, B, H, W, Cin, Cout, KH, KW, stride, pad,
                 (cudaStream_t)stream);
}

// Inference form: y = act(conv(x, w) + bias (+ addend)) in ONE launch — with BatchNorm folded into w / bias
// (uda_bn_fold_conv) this is conv + BN (+ residual) + ReLU of an eval-mode network.  act_slope: 0 = ReLU,
// 0.2 = LeakyReLU, 1 = no activation.
extern "C" int uda_conv2d_tc_fwd_fused(const void* x, const void* w, const float* bias, const void* addend,
                                       float act_slope, void* y_nhwc, float* y_nchw_f32, int B, int H, int W, int Cin,
                                       int Cout, int KH, int KW, int stride, int pad, void* stream) {
  UDA_REQUIRE(x && w && (y_nhwc || y_nchw_f32), UDA_ERR_BAD_ARG, "conv_tc_fwd_fused: null pointer");
  UDA_REQUIRE(use_persistent(), UDA_ERR_UNSUPPORTED, "conv_tc_fwd_fused: the fused epilogue needs the persistent kernels");
  UDA_REQUIRE(!addend || aligned<bf16>(addend, 16), UDA_ERR_BAD_ARG, "conv_tc_fwd_fused: addend must be 16-byte aligned");
  return run_fwd(x, w, bias, addend, y_nhwc, y_nchw_f32, nullptr, B, H, W, Cin, Cout, KH, KW, stride, pad,
                 (cudaStream_t)stream, act_slope != 1.f ? 1 : 0, act_slope);
}

// training form of the skip-channel half of the decoder convolution: y = conv(x, w) + addend with the BatchNorm
// statistics of the SUM accumulated in the epilogue
extern "C" int uda_conv2d_tc_fwd_add(const void* x, const void* w, const void* addend, void* y_nhwc, double* bn_sums, int B,
                                     int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, void* stream) {
  UDA_REQUIRE(x && w && y_nhwc && addend, UDA_ERR_BAD_ARG, "conv_tc_fwd_add: null pointer");
  UDA_REQUIRE(use_persistent(), UDA_ERR_UNSUPPORTED, "conv_tc_fwd_add: needs the persistent kernels");
  UDA_REQUIRE(aligned<bf16>(addend, 16), UDA_ERR_BAD_ARG, "conv_tc_fwd_add: addend must be 16-byte aligned");
  return run_fwd(x, w, nullptr, addend, y_nhwc, nullptr, bn_sums, B, H, W, Cin, Cout, KH, KW, stride, pad,
                 (cudaStream_t)stream);
}

// Training form of conv + BatchNorm + activation (+ residual) as ONE launch (BnFuse, conv_tc_internal.cuh): returns
// UDA_ERR_UNSUPPORTED — before anything is launched and without an error message the caller would surface — when the
// layer's output tiles do not fit the TMEM of one wave (the caller then runs uda_conv2d_tc_fwd + uda_bn_apply_fused).
extern "C" int uda_conv2d_tc_fwd_bn_act(const void* x, const void* w, const void* addend, const void* residual, void* z,
                                        void* a, double* bn_sums, unsigned int* counter, const float* gamma,
                                        const float* beta, float* running_mean, float* running_var, float* mean,
                                        float* rstd, float* scale, float* shift, int B, int H, int W, int Cin, int Cout,
                                        int KH, int KW, int stride, int pad, float eps, float momentum, float slope,
                                        void* stream) {
  UDA_REQUIRE(x && w && z && a && bn_sums && counter && mean && rstd && scale && shift, UDA_ERR_BAD_ARG,
              "conv_tc_fwd_bn_act: null pointer");
  if (!use_persistent() || !fwd_shape_ok(B, H, W, Cin, Cout, KH, KW, stride, pad)) return UDA_ERR_UNSUPPORTED;
  UDA_REQUIRE(aligned<bf16>(a, 16) && (!addend || aligned<bf16>(addend, 16)) && (!residual || aligned<bf16>(residual, 16)),
              UDA_ERR_BAD_ARG, "conv_tc_fwd_bn_act: pointers must be 16-byte aligned");
  const int Ho = stride == 1 ? H : H / 2, Wo = stride == 1 ? W : W / 2;
  BnFuse f{};
  f.a_out = a; f.residual = residual; f.gamma = gamma; f.beta = beta; f.running_mean = running_mean;
  f.running_var = running_var; f.mean = mean; f.rstd = rstd; f.scale = scale; f.shift = shift;
  f.M = (long long)B * Ho * Wo; f.inv_m = 1.0 / (double)f.M; f.eps = eps; f.momentum = momentum; f.slope = slope; f.counter = counter;
  return run_fwd(x, w, nullptr, addend, z, nullptr, bn_sums, B, H, W, Cin, Cout, KH, KW, stride, pad,
                 (cudaStream_t)stream, 0, 0.f, &f);
}

// w_ft: weights from uda_conv2d_weight_flip_transpose ([Cin][KH][KW][Cout] bf16)
extern "C" int uda_conv2d_tc_dgrad(const void* dy, const void* w_ft, const void* addend, void* dx, int B, int H, int W,
                                   int Cin, int Cout, int KH, int KW, int stride, int pad, void* stream) {
  UDA_REQUIRE(dy && w_ft && dx, UDA_ERR_BAD_ARG, "conv_tc_dgrad: null pointer");
  return run_dgrad(dy, w_ft, addend, dx, B, H, W, Cin, Cout, KH, KW, stride, pad, (cudaStream_t)stream);
}

// dgrad that also accumulates the BatchNorm-backward statistics of the tensor it produces (include/uda_b200.h)
extern "C" int uda_conv2d_tc_dgrad_bnstats(const void* dy, const void* w_ft, const void* addend, void* dx, int B, int H,
                                           int W, int Cin, int Cout, int KH, int KW, int stride, int pad, const void* a,
                                           const void* z, float slope, double* sums, void* stream) {
  UDA_REQUIRE(dy && w_ft && dx && a && sums, UDA_ERR_BAD_ARG, "conv_tc_dgrad_bnstats: null pointer");
  UDA_REQUIRE(aligned<bf16>(a, 16) && (!z || aligned<bf16>(z, 16)), UDA_ERR_BAD_ARG,
              "conv_tc_dgrad_bnstats: a / z must be 16-byte aligned");
  return run_dgrad(dy, w_ft, addend, dx, B, H, W, Cin, Cout, KH, KW, stride, pad, (cudaStream_t)stream, a, z, slope, sums);
}

// ---- decoder conv1 without the upsampled / concatenated tensor -------------------------------------------------------
// conv3x3(cat(upsample2x(x), skip), W)  =  conv_transpose4x4_s2_p1(x, W4)  +  conv3x3(skip, Ws):
// on the nearest-upsampled image the nine taps of an output pixel of parity (a, c) fall on only 2 x 2 source pixels of
// x, so the 3x3 kernel collapses per parity into a 2x2 kernel — together exactly a 4x4 stride-2 transposed
// convolution, i.e. the DGRAD of a 4x4 stride-2 pad-1 convolution, which the tensor-core path already has (4 output
// parity classes x 4 taps in one launch, 16/36 of the FLOPs).  Row / column groups of the flipped-transposed weights:
// r = 0 <- {kh 0}, 1 <- {0, 1}, 2 <- {1, 2}, 3 <- {2}.
//   w     : bf16 [O][3][3][C1 + C2]   (x channels first, as torch.cat([up(x), skip], 1) orders them)
//   wx_ft : bf16 [O][4][4][C1] = sum over the tap group (fp32 sum, one rounding) — the `w_ft` of run_dgrad
//   ws    : bf16 [O][3][3][C2] contiguous copy of the skip channels
// w4 (optional): bf16 [C1][4][4][O] = weight_flip_transpose(wx_ft) — the weights of the 4x4 stride-2 convolution whose
// forward is dx; ws_ft (optional): bf16 [C2][3][3][O] = weight_flip_transpose(ws) for the skip dgrad.  One launch
// prepares everything a decoder block needs for a training step.
__global__ void __launch_bounds__(256)
upconv_split_weights_kernel(const bf16* __restrict__ w, bf16* __restrict__ wx_ft, bf16* __restrict__ ws,
                            bf16* __restrict__ w4, bf16* __restrict__ ws_ft, int O, int C1, int C2) {
  const int C = C1 + C2;
  const long long nx = (long long)O * 16 * C1, ns = (long long)O * 9 * C2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nx) {
    const int c = (int)(i % C1);
    const int rs = (int)((i / C1) % 16), o = (int)(i / ((long long)C1 * 16));
    const int r = rs >> 2, q = rs & 3;
    const int kh0 = r == 0 ? 0 : r - 1, kh1 = r == 3 ? 2 : (r == 0 ? 0 : r);      // rows of group r: [kh0, kh1]
    const int kw0 = q == 0 ? 0 : q - 1, kw1 = q == 3 ? 2 : (q == 0 ? 0 : q);
    float acc = 0.f;
    for (int kh = kh0; kh <= kh1; ++kh)
      for (int kw = kw0; kw <= kw1; ++kw) acc += __bfloat162float(w[((long long)o * 9 + kh * 3 + kw) * C + c]);
    const bf16 v = __float2bfloat16_rn(acc);
    wx_ft[i] = v;
    if (w4) w4[(((long long)c * 4 + (3 - r)) * 4 + (3 - q)) * O + o] = v;
  } else if (i < nx + ns) {
    const long long j = i - nx;
    const int c = (int)(j % C2);
    const long long ot = j / C2;      // o * 9 + tap
    const bf16 v = w[ot * C + C1 + c];
    ws[j] = v;
    if (ws_ft) {
      const int o = (int)(ot / 9), tap = (int)(ot % 9);
      ws_ft[((long long)c * 9 + (8 - tap)) * O + o] = v;
    }
  }
}
// backward of the split: dW[o][kh][kw][c] += sum over the groups (r, q) that contain (kh, kw) of dW4[c][3-r][3-q][o]
// (dW4 = fp32 wgrad of the 4x4 stride-2 convolution, [C1][4][4][O]) for the x channels; += dWs for the skip channels
__global__ void __launch_bounds__(256)
upconv_merge_wgrad_kernel(const float* __restrict__ dw4, const float* __restrict__ dws, float* __restrict__ dw, int O,
                          int C1, int C2) {
  const int C = C1 + C2;
  const long long n = (long long)O * 9 * C;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = (int)(i % C);
  const int tap = (int)((i / C) % 9), o = (int)(i / ((long long)C * 9));
  if (c >= C1) { dw[i] += dws[((long long)o * 9 + tap) * C2 + (c - C1)]; return; }
  const int kh = tap / 3, kw = tap % 3;
  float acc = 0.f;
  for (int r = kh; r <= kh + 1; ++r)          // kh belongs to groups r = kh and r = kh + 1
    for (int q = kw; q <= kw + 1; ++q)
      acc += dw4[(((long long)c * 4 + (3 - r)) * 4 + (3 - q)) * O + o];
  dw[i] += acc;
}
extern "C" int uda_upconv_split_weights(const void* w, void* wx_ft, void* ws, void* w4, void* ws_ft, int Cout, int C1,
                                        int C2, void* stream) {
  UDA_REQUIRE(w && wx_ft && (ws || C2 == 0) && Cout > 0 && C1 > 0 && C2 >= 0, UDA_ERR_BAD_ARG, "upconv_split_weights: bad argument");
  const long long n = (long long)Cout * (16 * C1 + 9 * C2);
  upconv_split_weights_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)w, (bf16*)wx_ft, (bf16*)ws, (bf16*)w4, (bf16*)ws_ft, Cout, C1, C2);
  UDA_LAUNCH_OK("upconv_split_weights_kernel");
  return UDA_OK;
}
extern "C" int uda_upconv_merge_wgrad(const float* dw4, const float* dws, float* dw, int Cout, int C1, int C2, void* stream) {
  UDA_REQUIRE(dw4 && dw && (dws || C2 == 0) && Cout > 0 && C1 > 0 && C2 >= 0, UDA_ERR_BAD_ARG, "upconv_merge_wgrad: bad argument");
  const long long n = (long long)Cout * 9 * (C1 + C2);
  upconv_merge_wgrad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dw4, dws, dw, Cout, C1, C2);
  UDA_LAUNCH_OK("upconv_merge_wgrad_kernel");
  return UDA_OK;
}
// y[B,H,W,Cout] = act(conv_transpose4x4_s2_p1(x[B,H/2,W/2,C1], wx_ft) + bias (+ addend)), optional BatchNorm statistics:
// the x-channel half of the decoder convolution (see above).  H, W are the OUTPUT (full-resolution) sizes.
extern "C" int uda_upconv_tc_fwd(const void* x, const void* wx_ft, const float* bias, const void* addend, float act_slope,
                                 void* y, double* bn_sums, int B, int H, int W, int C1, int Cout, void* stream) {
  UDA_REQUIRE(x && wx_ft && y, UDA_ERR_BAD_ARG, "upconv_tc_fwd: null pointer");
  UDA_REQUIRE(use_persistent(), UDA_ERR_UNSUPPORTED, "upconv_tc_fwd: needs the persistent kernels");
  if (!bias && !addend && act_slope == 1.f && H % 2 == 0 && W % 2 == 0) {   // wide, few-channel blocks: halo kernel
    const int rc = run_upconv_halo(x, wx_ft, y, bn_sums, B, H / 2, W / 2, C1, Cout, (cudaStream_t)stream);
    if (rc != UDA_ERR_UNSUPPORTED) return rc;
  }
  return run_dgrad(x, wx_ft, addend, y, B, H, W, Cout, C1, 4, 4, 2, 1, (cudaStream_t)stream, nullptr, nullptr, 0.f, nullptr,
                   bias, bn_sums, act_slope != 1.f ? 1 : 0, act_slope);
}

// ---- eval-mode BatchNorm folding:  w'[o] = w[o] * gamma[o] / sqrt(var[o] + eps),  b' = beta - mean * gamma / sqrt(...) (+ b * ...)
__global__ void bn_fold_conv_kernel(const float* __restrict__ w, const float* __restrict__ conv_bias,
                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                    bf16* __restrict__ w_out, float* __restrict__ bias_out, int Cout, int per_out) {
  const int o = blockIdx.x;
  const float s = gamma[o] * rsqrtf(var[o] + eps);
  for (int i = threadIdx.x; i < per_out; i += blockDim.x)
    w_out[(size_t)o * per_out + i] = __float2bfloat16_rn(w[(size_t)o * per_out + i] * s);
  if (threadIdx.x == 0) bias_out[o] = beta[o] + ((conv_bias ? conv_bias[o] : 0.f) - mean[o]) * s;
}
extern "C" int uda_bn_fold_conv(const float* w, const float* conv_bias, const float* gamma, const float* beta,
                                const float* running_mean, const float* running_var, float eps, void* w_folded,
                                float* bias_folded, int Cout, int per_out, void* stream) {
  UDA_REQUIRE(w && gamma && beta && running_mean && running_var && w_folded && bias_folded && Cout > 0 && per_out > 0,
              UDA_ERR_BAD_ARG, "bn_fold_conv: bad argument");
  bn_fold_conv_kernel<<<Cout, 128, 0, (cudaStream_t)stream>>>(w, conv_bias, gamma, beta, running_mean, running_var, eps,
                                                               (bf16*)w_folded, bias_folded, Cout, per_out);
  UDA_LAUNCH_OK("bn_fold_conv_kernel");
  return UDA_OK;
}

extern "C" int uda_conv2d_weight_flip_transpose(const void* w, void* w_ft, int Cout, int Cin, int KH, int KW,
                                                void* stream) {
  UDA_REQUIRE(w && w_ft && Cout > 0 && Cin > 0 && KH > 0 && KW > 0, UDA_ERR_BAD_ARG, "weight_flip_transpose: bad argument");
  const long long n = (long long)Cout * Cin * KH * KW;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  weight_flip_transpose_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)w, (bf16*)w_ft, Cout, Cin, KH, KW);
  UDA_LAUNCH_OK("weight_flip_transpose_kernel");
  return UDA_OK;
}

extern "C" int uda_conv2d_weight_flip_transpose_batch(const void* w_base, void* w_ft_base, const int* table_dev,
                                                      int n_weights, void* stream) {
  UDA_REQUIRE(w_base && w_ft_base && table_dev && n_weights > 0 && n_weights <= 65535, UDA_ERR_BAD_ARG,
              "weight_flip_transpose_batch: bad argument");
  weight_flip_transpose_batch_kernel<<<dim3(48, (unsigned)n_weights), 256, 0, (cudaStream_t)stream>>>(
      (const bf16*)w_base, (bf16*)w_ft_base, table_dev);
  UDA_LAUNCH_OK("weight_flip_transpose_batch_kernel");
  return UDA_OK;
}

extern "C" int uda_conv2d_tc_wgrad(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin, int Cout,
                                   int KH, int KW, int stride, int pad, void* stream) {
  UDA_REQUIRE(dy && x && dw, UDA_ERR_BAD_ARG, "conv_tc_wgrad: null pointer");
  return run_wgrad(dy, x, dw, B, H, W, Cin, Cout, KH, KW, stride, pad, (cudaStream_t)stream);
}
