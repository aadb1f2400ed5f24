// Generic implicit-GEMM convolution on the FP32 pipe (any kernel size / stride / padding / channel
// count; fp32 or bf16 NHWC activations, fp32 accumulate).
//
// Role in the design (DESIGN.md "Convolutions"): this is (1) the fp32 "1e-4 parity mode" of the
// U-Net, (2) the on-GPU cross-check for the tcgen05 kernels in conv_tc.cu and (3) the path for the
// few shapes the tensor-core kernels do not cover yet (Cin=3 stems).  The hot bf16 shapes go
// through conv_tc.cu.
//
//   fwd   : Y[b,ho,wo,n]  = sum_{kh,kw,c} X[b,ho*s-p+kh,wo*s-p+kw,c] * W[n,kh,kw,c]   (+ bias[n])
//   dgrad : dX[b,hi,wi,n] = sum_{kh,kw,c} dY[b,(hi+p-kh)/s,(wi+p-kw)/s,c] * W[c,kh,kw,n]
//   wgrad : dW[n,kh,kw,c] += sum_{b,ho,wo} dY[b,ho,wo,n] * X[b,ho*s-p+kh,wo*s-p+kw,c]
// Weights are "OHWI" ([Cout][KH][KW][Cin], the physical layout of a channels_last torch parameter).
// Mirrors aten::_convolution as reached from smp.Unet / DomainDiscriminator.
#include "common.cuh"

namespace uda {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, CT = 256;

struct ConvGeom {
  int B, H, W, Cin;      // input  [B,H,W,Cin]
  int Ho, Wo, Cout;      // output [B,Ho,Wo,Cout]
  int KH, KW, stride, pad;
};

// MODE 0: forward, MODE 1: dgrad (roles of input/output swapped by the host: "A" is dY, result is dX)
template <typename T, typename WT, int MODE>
__global__ void __launch_bounds__(CT)
conv_igemm_kernel(const T* __restrict__ A, const WT* __restrict__ Wt, const float* __restrict__ bias,
                  const T* addend, T* out_nhwc, float* __restrict__ out_nchw, const ConvGeom g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  // GEMM view
  //  fwd  : M = B*Ho*Wo, N = Cout, reduction (tap, c<Cin),  source spatial = (H,W)
  //  dgrad: M = B*H*W,   N = Cin,  reduction (tap, c<Cout), source spatial = (Ho,Wo)
  const int taps = g.KH * g.KW;
  const int Cred = (MODE == 0) ? g.Cin : g.Cout;
  const int N = (MODE == 0) ? g.Cout : g.Cin;
  const int mh = (MODE == 0) ? g.Ho : g.H, mw = (MODE == 0) ? g.Wo : g.W;
  const int sh = (MODE == 0) ? g.H : g.Ho, sw = (MODE == 0) ? g.W : g.Wo;
  const long long M = (long long)g.B * mh * mw;
  const int K = taps * Cred;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;

  // loader assignment: one row (pixel / out-channel) and 4 consecutive k per thread
  const int lrow = tid / 4, lk = (tid % 4) * 4;
  const long long am = m0 + lrow;
  int ab = 0, ah = 0, aw = 0;
  const bool arow_ok = am < M;
  if (arow_ok) {
    long long t = am;
    aw = (int)(t % mw); t /= mw;
    ah = (int)(t % mh);
    ab = (int)(t / mh);
  }
  const int bn = n0 + lrow;
  const bool brow_ok = bn < N;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    // ---- gather A (activations) ----
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + lk + j;
      float v = 0.f;
      if (arow_ok && k < K) {
        const int tap = k / Cred, c = k - tap * Cred;
        const int kh = tap / g.KW, kw = tap - kh * g.KW;
        int hs, ws;
        bool ok;
        if (MODE == 0) {
          hs = ah * g.stride - g.pad + kh;
          ws = aw * g.stride - g.pad + kw;
          ok = hs >= 0 && hs < sh && ws >= 0 && ws < sw;
        } else {
          const int th = ah + g.pad - kh, tw = aw + g.pad - kw;
          ok = th >= 0 && tw >= 0 && (th % g.stride == 0) && (tw % g.stride == 0);
          hs = th / g.stride; ws = tw / g.stride;
          ok = ok && hs < sh && ws < sw;
        }
        if (ok) v = to_f(A[(((long long)ab * sh + hs) * sw + ws) * Cred + c]);
      }
      As[lk + j][lrow] = v;
    }
    // ---- gather B (weights, OHWI) ----
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + lk + j;
      float v = 0.f;
      if (brow_ok && k < K) {
        const int tap = k / Cred, c = k - tap * Cred;
        long long off = (MODE == 0) ? ((long long)bn * taps + tap) * g.Cin + c
                                    : ((long long)c * taps + tap) * g.Cin + bn;
        v = to_f(Wt[off]);
      }
      Bs[lk + j][lrow] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (addend) v += to_f(addend[m * N + n]);  // may alias out_nhwc (same element, same thread)
      if (out_nhwc) out_nhwc[m * N + n] = from_f<T>(v);
      if (out_nchw) {
        const long long hw = (long long)mh * mw;
        const long long b = m / hw, p = m - b * hw;
        out_nchw[(b * N + n) * hw + p] = v;
      }
    }
  }
}

// wgrad: rows = out-channel n, cols = kcol=(tap,c), reduction over pixels, split across blockIdx.z
template <typename T>
__global__ void __launch_bounds__(CT)
conv_wgrad_kernel(const T* __restrict__ dY, const T* __restrict__ X, float* __restrict__ dW, const ConvGeom g,
                  long long pix_per_split) {
  __shared__ __align__(16) float As[BK][BM + 4];  // [pix][n]
  __shared__ __align__(16) float Bs[BK][BN + 4];  // [pix][kcol]
  const int taps = g.KH * g.KW;
  const int Kw = taps * g.Cin;
  const long long Mpix = (long long)g.B * g.Ho * g.Wo;
  const int n0 = blockIdx.x * BM, c0 = blockIdx.y * BN;
  long long p_begin = (long long)blockIdx.z * pix_per_split;
  long long p_end = p_begin + pix_per_split;
  if (p_end > Mpix) p_end = Mpix;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  const int lp = tid / 16, lq = (tid % 16) * 4;  // loader: pixel slot, 4 consecutive columns

  // decompose this thread's 4 B-columns once
  int bt_kh[4], bt_kw[4], bt_c[4];
  bool bt_ok[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int kc = c0 + lq + j;
    bt_ok[j] = kc < Kw;
    const int tap = bt_ok[j] ? kc / g.Cin : 0;
    bt_c[j] = bt_ok[j] ? kc - tap * g.Cin : 0;
    bt_kh[j] = tap / g.KW;
    bt_kw[j] = tap - bt_kh[j] * g.KW;
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long p0 = p_begin; p0 < p_end; p0 += BK) {
    const long long pix = p0 + lp;
    const bool pok = pix < p_end;
    int b = 0, ho = 0, wo = 0;
    if (pok) {
      long long t = pix;
      wo = (int)(t % g.Wo); t /= g.Wo;
      ho = (int)(t % g.Ho);
      b = (int)(t / g.Ho);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + lq + j;
      As[lp][lq + j] = (pok && n < g.Cout) ? to_f(dY[pix * g.Cout + n]) : 0.f;
      float v = 0.f;
      if (pok && bt_ok[j]) {
        const int h = ho * g.stride - g.pad + bt_kh[j], w = wo * g.stride - g.pad + bt_kw[j];
        if (h >= 0 && h < g.H && w >= 0 && w < g.W)
          v = to_f(X[(((long long)b * g.H + h) * g.W + w) * g.Cin + bt_c[j]]);
      }
      Bs[lp][lq + j] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= g.Cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int kc = c0 + tx * 4 + j;
      if (kc >= Kw) continue;
      atomicAdd(dW + (long long)n * Kw + kc, acc[i][j]);
    }
  }
}

int check_geom(const ConvGeom& g, const char* who) {
  UDA_REQUIRE(g.B > 0 && g.H > 0 && g.W > 0 && g.Cin > 0 && g.Cout > 0, UDA_ERR_BAD_ARG, "%s: empty shape", who);
  UDA_REQUIRE(g.KH > 0 && g.KW > 0 && g.stride > 0 && g.pad >= 0, UDA_ERR_BAD_ARG, "%s: bad kernel geometry", who);
  UDA_REQUIRE(g.Ho == (g.H + 2 * g.pad - g.KH) / g.stride + 1 && g.Wo == (g.W + 2 * g.pad - g.KW) / g.stride + 1,
              UDA_ERR_BAD_ARG, "%s: output size %dx%d inconsistent with input %dx%d k=%d s=%d p=%d", who, g.Ho, g.Wo,
              g.H, g.W, g.KH, g.stride, g.pad);
  return UDA_OK;
}

}  // namespace
}  // namespace uda

using namespace uda;

static ConvGeom make_geom(int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad) {
  ConvGeom g;
  g.B = B; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout; g.KH = KH; g.KW = KW; g.stride = stride; g.pad = pad;
  g.Ho = (H + 2 * pad - KH) / stride + 1;
  g.Wo = (W + 2 * pad - KW) / stride + 1;
  return g;
}

// w_dtype: dtype of the weight buffer (fp32 master weights or the bf16 shadow copy)
extern "C" int uda_conv2d_direct_fwd(const void* x, int dtype, const void* w, int w_dtype, const float* bias,
                                     void* y_nhwc, float* y_nchw_f32, int B, int H, int W, int Cin, int Cout,
                                     int KH, int KW, int stride, int pad, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(x && w && (y_nhwc || y_nchw_f32), UDA_ERR_BAD_ARG, "conv_direct_fwd: null pointer");
  ConvGeom g = make_geom(B, H, W, Cin, Cout, KH, KW, stride, pad);
  if (int rc = check_geom(g, "conv_direct_fwd")) return rc;
  const long long M = (long long)B * g.Ho * g.Wo;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((Cout + BN - 1) / BN));
  if (dtype == UDA_BF16 && w_dtype == UDA_BF16)
    conv_igemm_kernel<bf16, bf16, 0><<<grid, CT, 0, st>>>((const bf16*)x, (const bf16*)w, bias, nullptr, (bf16*)y_nhwc, y_nchw_f32, g);
  else if (dtype == UDA_BF16 && w_dtype == UDA_F32)
    conv_igemm_kernel<bf16, float, 0><<<grid, CT, 0, st>>>((const bf16*)x, (const float*)w, bias, nullptr, (bf16*)y_nhwc, y_nchw_f32, g);
  else if (dtype == UDA_F32 && w_dtype == UDA_F32)
    conv_igemm_kernel<float, float, 0><<<grid, CT, 0, st>>>((const float*)x, (const float*)w, bias, nullptr, (float*)y_nhwc, y_nchw_f32, g);
  else
    return set_error(UDA_ERR_BAD_ARG, "conv_direct_fwd: dtype combination %d/%d", dtype, w_dtype);
  UDA_LAUNCH_OK("conv_igemm_kernel<fwd>");
  return UDA_OK;
}

// dX[B,H,W,Cin] = dgrad(dY[B,Ho,Wo,Cout]) (+ addend, same shape as dX; may alias dX)
extern "C" int uda_conv2d_direct_dgrad(const void* dy, int dtype, const void* w, int w_dtype, const void* addend,
                                       void* dx, int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride,
                                       int pad, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(dy && w && dx, UDA_ERR_BAD_ARG, "conv_direct_dgrad: null pointer");
  ConvGeom g = make_geom(B, H, W, Cin, Cout, KH, KW, stride, pad);
  if (int rc = check_geom(g, "conv_direct_dgrad")) return rc;
  const long long M = (long long)B * H * W;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((Cin + BN - 1) / BN));
  if (dtype == UDA_BF16 && w_dtype == UDA_BF16)
    conv_igemm_kernel<bf16, bf16, 1><<<grid, CT, 0, st>>>((const bf16*)dy, (const bf16*)w, nullptr, (const bf16*)addend, (bf16*)dx, nullptr, g);
  else if (dtype == UDA_BF16 && w_dtype == UDA_F32)
    conv_igemm_kernel<bf16, float, 1><<<grid, CT, 0, st>>>((const bf16*)dy, (const float*)w, nullptr, (const bf16*)addend, (bf16*)dx, nullptr, g);
  else if (dtype == UDA_F32 && w_dtype == UDA_F32)
    conv_igemm_kernel<float, float, 1><<<grid, CT, 0, st>>>((const float*)dy, (const float*)w, nullptr, (const float*)addend, (float*)dx, nullptr, g);
  else
    return set_error(UDA_ERR_BAD_ARG, "conv_direct_dgrad: dtype combination %d/%d", dtype, w_dtype);
  UDA_LAUNCH_OK("conv_igemm_kernel<dgrad>");
  return UDA_OK;
}

// dW (fp32, OHWI) += ...   (accumulates: zero it first for a fresh gradient)
extern "C" int uda_conv2d_direct_wgrad(const void* dy, const void* x, int dtype, float* dw, int B, int H, int W,
                                       int Cin, int Cout, int KH, int KW, int stride, int pad, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(dy && x && dw, UDA_ERR_BAD_ARG, "conv_direct_wgrad: null pointer");
  ConvGeom g = make_geom(B, H, W, Cin, Cout, KH, KW, stride, pad);
  if (int rc = check_geom(g, "conv_direct_wgrad")) return rc;
  const long long Mpix = (long long)B * g.Ho * g.Wo;
  const int Kw = KH * KW * Cin;
  const unsigned gm = (Cout + BM - 1) / BM, gn = (Kw + BN - 1) / BN;
  long long splits = (2LL * num_sms() + (long long)gm * gn - 1) / ((long long)gm * gn);
  long long max_splits = (Mpix + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  long long per = (Mpix + splits - 1) / splits;
  per = (per + BK - 1) / BK * BK;
  splits = (Mpix + per - 1) / per;
  dim3 grid(gm, gn, (unsigned)splits);
  if (dtype == UDA_BF16)
    conv_wgrad_kernel<bf16><<<grid, CT, 0, st>>>((const bf16*)dy, (const bf16*)x, dw, g, per);
  else if (dtype == UDA_F32)
    conv_wgrad_kernel<float><<<grid, CT, 0, st>>>((const float*)dy, (const float*)x, dw, g, per);
  else
    return set_error(UDA_ERR_BAD_ARG, "conv_direct_wgrad: dtype %d", dtype);
  UDA_LAUNCH_OK("conv_wgrad_kernel");
  return UDA_OK;
}
