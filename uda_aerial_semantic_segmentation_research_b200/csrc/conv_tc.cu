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
#include "tc_common.cuh"

namespace uda {
namespace {

using namespace tc;

constexpr int kTcThreads = 192;  // 6 warps
constexpr int kMaxTaps = 16;

struct FwdParams {
  int TW, TH, NB;              // tile = NB images x TH rows x TW cols = 128 output pixels
  int tiles_w, tiles_h;        // tiles per image row / column
  int MH, MW;                  // output spatial size
  int Cout, Cred;              // GEMM N, reduction channels per tap
  int ntaps, kchunks;          // taps, Cred / KC
  int rank5;                   // 0: stride-1 4-D map {C,W,H,B}; 1: stride-2 5-D map {2C,W/2,2,H/2,B}
  signed char dh[kMaxTaps], dw[kMaxTaps];   // per-tap source offset (rows / cols; pair units for rank5)
  signed char ph[kMaxTaps], pw[kMaxTaps];   // rank5 only: row / column parity
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
        tma_load_2d(b_dst, &map_b, full_bar(s), tap * p.Cred + c0, n0);
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
    const int b = b0 + nb, h = h0 + th, w = w0 + tw;
    const long long pix = ((long long)b * p.MH + h) * p.MW + w;
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
        const long long hw = (long long)p.MH * p.MW;
        float* dst = p.out_nchw + ((long long)b * p.Cout + nbase) * hw + (long long)h * p.MW + w;
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

struct TilePlan { int TW, TH, NB; bool ok; };

TilePlan plan_tiles(int B, int MH, int MW) {
  TilePlan t{0, 0, 0, false};
  if (MW <= 0 || MH <= 0) return t;
  t.TW = MW < 128 ? MW : 128;
  if (128 % t.TW || MW % t.TW) return t;
  int rows = 128 / t.TW;
  t.TH = rows < MH ? rows : MH;
  if (rows % t.TH || MH % t.TH) return t;
  t.NB = rows / t.TH;
  if (B % t.NB) return t;
  if (t.TW > 256 || t.TH > 256 || t.NB > 256) return t;
  t.ok = true;
  return t;
}

int pick_kc(int c) { return c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : (c % 16 == 0 ? 16 : 0)); }
int pick_bn(int cout) { return cout > 64 ? 128 : (cout > 32 ? 64 : 32); }

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

// x: [B,H,W,Cin] bf16; w: [Cout][KH*KW][Cin] bf16; generic "same"/stride-2 forward convolution
int run_fwd(const void* x, const void* w, const float* bias, const void* addend, void* y, float* y_nchw, int B,
            int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad, cudaStream_t st) {
  UDA_REQUIRE(fwd_shape_ok(B, H, W, Cin, Cout, KH, KW, stride, pad), UDA_ERR_UNSUPPORTED,
              "conv_tc: shape not covered by the tensor-core kernel (B=%d H=%d W=%d Cin=%d Cout=%d k=%d s=%d p=%d)",
              B, H, W, Cin, Cout, KH, stride, pad);
  UDA_REQUIRE(aligned<bf16>(x, 16) && aligned<bf16>(w, 16) && (!y || aligned<bf16>(y, 16)) &&
                  (!addend || aligned<bf16>(addend, 16)),
              UDA_ERR_BAD_ARG, "conv_tc: pointers must be 16-byte aligned");
  const int Ho = stride == 1 ? H : H / 2, Wo = stride == 1 ? W : W / 2;
  const TilePlan tp = plan_tiles(B, Ho, Wo);
  const int KC = pick_kc(Cin), BN = pick_bn(Cout);
  FwdParams p{};
  p.TW = tp.TW; p.TH = tp.TH; p.NB = tp.NB;
  p.tiles_w = Wo / tp.TW; p.tiles_h = Ho / tp.TH;
  p.MH = Ho; p.MW = Wo; p.Cout = Cout; p.Cred = Cin;
  p.ntaps = KH * KW; p.kchunks = Cin / KC; p.rank5 = stride == 2;
  for (int kh = 0; kh < KH; ++kh)
    for (int kw = 0; kw < KW; ++kw) {
      const int t = kh * KW + kw;
      const int oh = kh - pad, ow = kw - pad;
      if (stride == 1) {
        p.dh[t] = (signed char)oh; p.dw[t] = (signed char)ow; p.ph[t] = 0; p.pw[t] = 0;
      } else {
        const int ah = oh >= 0 ? oh / 2 : -((-oh + 1) / 2), aw = ow >= 0 ? ow / 2 : -((-ow + 1) / 2);  // floor
        p.dh[t] = (signed char)ah; p.dw[t] = (signed char)aw;
        p.ph[t] = (signed char)(oh - 2 * ah); p.pw[t] = (signed char)(ow - 2 * aw);
      }
    }
  p.out = (bf16*)y; p.out_nchw = y_nchw; p.bias = bias; p.addend = (const bf16*)addend;

  CUtensorMap ma, mb;
  const uint64_t C = (uint64_t)Cin;
  if (stride == 1) {
    uint64_t dims[4] = {C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[4] = {(uint32_t)KC, (uint32_t)tp.TW, (uint32_t)tp.TH, (uint32_t)tp.NB};
    if (int rc = make_tmap_bf16(&ma, x, 4, dims, str, box, KC * 2)) return rc;
  } else {
    uint64_t dims[5] = {2 * C, (uint64_t)W / 2, 2, (uint64_t)H / 2, (uint64_t)B};
    uint64_t str[4] = {2 * C * 2, (uint64_t)W * C * 2, 2 * (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
    uint32_t box[5] = {(uint32_t)KC, (uint32_t)tp.TW, 1, (uint32_t)tp.TH, (uint32_t)tp.NB};
    if (int rc = make_tmap_bf16(&ma, x, 5, dims, str, box, KC * 2)) return rc;
  }
  {
    const uint64_t Kt = (uint64_t)KH * KW * Cin;
    uint64_t dims[2] = {Kt, (uint64_t)Cout};
    uint64_t str[1] = {Kt * 2};
    uint32_t box[2] = {(uint32_t)KC, (uint32_t)BN};
    if (int rc = make_tmap_bf16(&mb, w, 2, dims, str, box, KC * 2)) return rc;
  }
  const int n_tiles = (B / tp.NB) * p.tiles_w * p.tiles_h;
#define UDA_TC(KCv, BNv) \
  if (KC == KCv && BN == BNv) return launch_fwd<KCv, BNv>(ma, mb, p, n_tiles, st);
  UDA_TC(64, 128) UDA_TC(64, 64) UDA_TC(64, 32)
  UDA_TC(32, 128) UDA_TC(32, 64) UDA_TC(32, 32)
  UDA_TC(16, 128) UDA_TC(16, 64) UDA_TC(16, 32)
#undef UDA_TC
  return set_error(UDA_ERR_UNSUPPORTED, "conv_tc: no kernel instance for KC=%d BN=%d", KC, BN);
}

}  // namespace
}  // namespace uda

using namespace uda;

extern "C" int uda_conv2d_tc_supported(int op, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride,
                                       int pad) {
  if (!uda_device_supported()) return 0;
  if (op == 0) return fwd_shape_ok(B, H, W, Cin, Cout, KH, KW, stride, pad) ? 1 : 0;
  if (op == 1) {
    // dgrad of a stride-1 "same" convolution == forward "same" convolution of dy with flipped/transposed weights
    if (stride != 1) return 0;
    const int Ho = (H + 2 * pad - KH) + 1, Wo = (W + 2 * pad - KW) + 1;
    if (Ho != H || Wo != W) return 0;
    return fwd_shape_ok(B, H, W, Cout, Cin, KH, KW, 1, KH - 1 - pad) ? 1 : 0;
  }
  return 0;  // wgrad: see conv_tc_wgrad.cu
}

extern "C" int uda_conv2d_tc_fwd(const void* x, const void* w, const float* bias, void* y_nhwc, float* y_nchw_f32,
                                 double* bn_sums, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride,
                                 int pad, void* stream) {
  UDA_REQUIRE(x && w && (y_nhwc || y_nchw_f32), UDA_ERR_BAD_ARG, "conv_tc_fwd: null pointer");
  UDA_REQUIRE(bn_sums == nullptr, UDA_ERR_UNSUPPORTED, "conv_tc_fwd: fused BN statistics are not implemented yet");
  return run_fwd(x, w, bias, nullptr, y_nhwc, y_nchw_f32, B, H, W, Cin, Cout, KH, KW, stride, pad,
                 (cudaStream_t)stream);
}

// w_ft: weights from uda_conv2d_weight_flip_transpose ([Cin][KH][KW][Cout] bf16)
extern "C" int uda_conv2d_tc_dgrad(const void* dy, const void* w_ft, const void* addend, void* dx, int B, int H, int W,
                                   int Cin, int Cout, int KH, int KW, int stride, int pad, void* stream) {
  UDA_REQUIRE(dy && w_ft && dx, UDA_ERR_BAD_ARG, "conv_tc_dgrad: null pointer");
  UDA_REQUIRE(stride == 1, UDA_ERR_UNSUPPORTED, "conv_tc_dgrad: stride %d not covered", stride);
  return run_fwd(dy, w_ft, nullptr, addend, dx, nullptr, B, H, W, Cout, Cin, KH, KW, 1, KH - 1 - pad,
                 (cudaStream_t)stream);
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

extern "C" int uda_conv2d_tc_wgrad(const void* dy, const void* x, float* dw, int B, int H, int W, int Cin, int Cout,
                                   int KH, int KW, int stride, int pad, void* stream) {
  return set_error(UDA_ERR_UNSUPPORTED, "conv_tc_wgrad: not implemented yet (use the direct kernel)");
}
