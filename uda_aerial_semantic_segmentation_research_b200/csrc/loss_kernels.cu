// Fused loss + logit-gradient kernels (HBM-bound, warp-shuffle reduced, vectorised).
//
// Reference semantics (paths relative to the reference tree):
//   cross entropy            src/models/train.py:208,342          (nn.CrossEntropyLoss defaults)
//   DiceLoss                 src/models/losses.py:118-152
//   WeightedSegmentationLoss src/models/losses.py:154-215         (focal(w-CE) + Dice)
//   ConsistencyLoss          src/models/losses.py:62-90           (symmetric KL, T, batchmean)
//   AdversarialLoss (BCE)    src/models/losses.py:7-51
//   entropy minimisation     north-star extension
//
// Layout: logits / grads are contiguous NCHW ([B,C,H,W], fp32 or bf16) exactly as the reference
// passes them; targets are int64 [B,H,W].  One thread owns VEC consecutive pixels of one image and
// all C classes of them (C <= CPAD kept in registers), so every global access is a 16-byte
// (fp32) / 8-byte (bf16) coalesced vector per class plane and each logit is read once per pass.
#include "common.cuh"
#include "loss_stream.cuh"

namespace uda {
namespace {

constexpr int kLossThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;

struct SegLossParams {
  const void* logits;
  const long long* target;   // [B,HW] class indices (nullable when soft_target is given)
  const float* soft_target;  // [B,C,HW] float "one-hot"/soft targets for Dice (nullable)
  void* grad;                // [B,C,HW], same dtype as logits
  const float* class_w;      // [C] or null
  double* acc;               // [4]: sum pixel loss, sum denominator weight, #invalid targets, -
  double* dice_sums;         // [B*C*3]: I, sum p, sum t
  const float* dice_coef;    // [B*C*2]: coefA, coefB  (pass 2)
  const float* dev_scale;    // [1] exact CE gradient scale for pass 2
  int B, C;
  long long HW;
  int has_ce, focal, has_dice;
  float alpha, gamma;
  long long ignore_index;
  float ce_grad_scale;  // static scale used by the single-pass kernel
};

// PASS 0: single pass  (CE/focal loss sums + gradient)           -> reads P*C, writes P*C
// PASS 1: sums only    (CE/focal sums + per-(b,c) Dice sums)     -> reads P*C
// PASS 2: gradient     (CE/focal part + Dice part)               -> reads P*C, writes P*C
template <typename T, int CPAD, int VEC, int PASS>
__global__ void __launch_bounds__(kLossThreads, 1) seg_loss_kernel(const SegLossParams p) {
  extern __shared__ float smem[];
  const int b = blockIdx.y;
  const int C = p.C;
  const long long HW = p.HW;
  const long long nvec = HW / VEC;
  const T* zbase = reinterpret_cast<const T*>(p.logits) + (long long)b * C * HW;
  T* gbase = reinterpret_cast<T*>(p.grad) + (long long)b * C * HW;
  const long long* tbase = p.target ? p.target + (long long)b * HW : nullptr;
  const float* sbase = p.soft_target ? p.soft_target + (long long)b * C * HW : nullptr;

  float* coefA = smem;          // [CPAD]
  float* coefB = smem + CPAD;   // [CPAD]
  float* wsm = smem + 2 * CPAD; // [CPAD] class weights
  float* red = smem + 3 * CPAD; // reduction scratch
  if (threadIdx.x < CPAD) {
    int c = threadIdx.x;
    float a = 0.f, bb = 0.f, w = 1.f;
    if (c < C) {
      if (PASS == 2 && p.has_dice) {
        a = p.dice_coef[((long long)b * C + c) * 2 + 0];
        bb = p.dice_coef[((long long)b * C + c) * 2 + 1];
      }
      if (p.class_w) w = p.class_w[c];
    }
    coefA[c] = a; coefB[c] = bb; wsm[c] = w;
  }
  __syncthreads();
  const float ce_scale = (PASS == 2) ? (p.has_ce ? *p.dev_scale : 0.f) : p.ce_grad_scale;

  float loss_sum = 0.f, denom_sum = 0.f, invalid = 0.f;
  float psum[CPAD], inter[CPAD], tsum[CPAD];
  if constexpr (PASS == 1) {
#pragma unroll
    for (int c = 0; c < CPAD; ++c) { psum[c] = 0.f; inter[c] = 0.f; tsum[c] = 0.f; }
  }

  for (long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec;
       iv += (long long)gridDim.x * blockDim.x) {
    const long long px = iv * VEC;
    float z[CPAD][VEC];
#pragma unroll
    for (int c = 0; c < CPAD; ++c) {
      if (c < C) {
        ld_vec<VEC>(zbase + (long long)c * HW + px, z[c]);
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) z[c][j] = -INFINITY;
      }
    }
    int y[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) y[j] = -1;
    bool valid[VEC];
    if (tbase) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        long long t = __ldg(tbase + px + j);
        bool ign = (t == p.ignore_index);
        bool ok = (t >= 0 && t < C);
        valid[j] = ok && !ign;
        y[j] = valid[j] ? (int)t : -1;
        if (!ok && !ign) invalid += 1.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) valid[j] = false;
    }
    // softmax per pixel (in place: z -> p), keep log-prob of the target class
    float coef[VEC];  // CE/focal gradient coefficient per pixel
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float m = z[0][j];
#pragma unroll
      for (int c = 1; c < CPAD; ++c) m = fmaxf(m, z[c][j]);
      float s = 0.f, zy = 0.f;
#pragma unroll
      for (int c = 0; c < CPAD; ++c) {
        zy = (c == y[j]) ? z[c][j] : zy;
        float e = exp2f((z[c][j] - m) * kLog2e);
        z[c][j] = e;
        s += e;
      }
      float inv = 1.f / s;
#pragma unroll
      for (int c = 0; c < CPAD; ++c) z[c][j] *= inv;
      coef[j] = 0.f;
      if (p.has_ce && valid[j]) {
        float wy = wsm[y[j]];
        float nll = -(zy - m - logf(s));  // -log p_y
        float ce = wy * nll;
        if (p.focal) {
          float pt = expf(-ce);
          float omp = 1.f - pt;
          float pw = powf(omp, p.gamma);
          float dpw = p.gamma * powf(omp, p.gamma - 1.f);
          if (PASS != 2) { loss_sum += p.alpha * pw * ce; }
          coef[j] = p.alpha * (pw + dpw * pt * ce) * wy * ce_scale;
        } else {
          if (PASS != 2) { loss_sum += ce; denom_sum += wy; }
          coef[j] = wy * ce_scale;
        }
      }
    }
    if constexpr (PASS == 1) {
      if (p.has_dice) {
#pragma unroll
        for (int c = 0; c < CPAD; ++c) {
          if (c < C) {
            float tv[VEC];
            if (sbase) {
              ld_vec<VEC>(sbase + (long long)c * HW + px, tv);
            } else {
#pragma unroll
              for (int j = 0; j < VEC; ++j) tv[j] = (y[j] == c) ? 1.f : 0.f;
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
              psum[c] += z[c][j];
              inter[c] += z[c][j] * tv[j];
              tsum[c] += tv[j];
            }
          }
        }
      }
    } else {
      // gradient
      float dot[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) dot[j] = 0.f;
      const bool dice = (PASS == 2) && p.has_dice;
      if (dice) {
        // g_c = coefA_c * t_c - coefB_c ; dot = sum_c g_c p_c ; stash g in a second sweep
#pragma unroll
        for (int c = 0; c < CPAD; ++c) {
          if (c < C) {
            float tv[VEC];
            if (sbase) {
              ld_vec<VEC>(sbase + (long long)c * HW + px, tv);
            } else {
#pragma unroll
              for (int j = 0; j < VEC; ++j) tv[j] = (y[j] == c) ? 1.f : 0.f;
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j) dot[j] += (coefA[c] * tv[j] - coefB[c]) * z[c][j];
          }
        }
      }
#pragma unroll
      for (int c = 0; c < CPAD; ++c) {
        if (c < C) {
          float g[VEC];
          float tv[VEC];
          if (dice && sbase) ld_vec<VEC>(sbase + (long long)c * HW + px, tv);
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            float onehot = (y[j] == c) ? 1.f : 0.f;
            float v = coef[j] * (z[c][j] - onehot);
            if (dice) {
              float t = sbase ? tv[j] : onehot;
              v += z[c][j] * ((coefA[c] * t - coefB[c]) - dot[j]);
            }
            g[j] = v;
          }
          st_vec<VEC>(gbase + (long long)c * HW + px, g);
        }
      }
    }
  }

  if constexpr (PASS != 2) {
    float v[3] = {loss_sum, denom_sum, invalid}, o[3];
    block_sum<3>(v, red, o);
    if (threadIdx.x == 0) {
      if (o[0] != 0.f) atomicAdd(p.acc + 0, (double)o[0]);
      if (o[1] != 0.f) atomicAdd(p.acc + 1, (double)o[1]);
      if (o[2] != 0.f) atomicAdd(p.acc + 2, (double)o[2]);
    }
  }
  if constexpr (PASS == 1) {
    if (p.has_dice) {
      const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
      // red layout: [3][CPAD][nw]
#pragma unroll
      for (int c = 0; c < CPAD; ++c) {
        float a = warp_sum(inter[c]), bsum = warp_sum(psum[c]), t = warp_sum(tsum[c]);
        if (lane == 0) {
          red[(0 * CPAD + c) * nw + wid] = a;
          red[(1 * CPAD + c) * nw + wid] = bsum;
          red[(2 * CPAD + c) * nw + wid] = t;
        }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < 3 * CPAD; i += blockDim.x) {
        int k = i / CPAD, c = i % CPAD;
        if (c < C) {
          float s = 0.f;
          for (int w = 0; w < nw; ++w) s += red[i * nw + w];
          atomicAdd(p.dice_sums + ((long long)b * C + c) * 3 + k, (double)s);
        }
      }
    }
  }
}

// Finalise: loss scalars + Dice gradient coefficients + exact CE scale, all on device.
//   out[0] = CE/focal loss term, out[1] = Dice loss, out[2] = w_ce*out[0] + w_dice*out[1] (x out_scale)
//   out[3] = number of out-of-range targets (diagnostic)
struct SegFinalizeParams {
  const double* acc;
  const double* dice_sums;
  float* dice_coef;
  float* dev_scale;  // [2]: exact CE grad scale, rescale factor for the single-pass kernel
  float* out;
  int B, C;
  long long P;  // B*HW
  int has_ce, focal, has_dice, mean;
  float smooth, w_ce, w_dice, out_scale;
  float static_scale;  // what the single-pass kernel used
};

__global__ void seg_loss_finalize_kernel(const SegFinalizeParams p) {
  __shared__ double sh[32];
  double dsum = 0.0;
  const int n = p.B * p.C;
  if (p.has_dice) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      double I = p.dice_sums[i * 3 + 0], U = p.dice_sums[i * 3 + 1] + p.dice_sums[i * 3 + 2];
      double den = U + (double)p.smooth;
      double d = (2.0 * I + (double)p.smooth) / den;
      dsum += d;
      double k = -(double)p.w_dice * (double)p.out_scale / (double)n;  // d(total)/d(dice_bc)
      p.dice_coef[i * 2 + 0] = (float)(k * 2.0 / den);
      p.dice_coef[i * 2 + 1] = (float)(k * (2.0 * I + (double)p.smooth) / (den * den));
    }
  }
  dsum = warp_sum(dsum);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = dsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += sh[w];
    double dice = p.has_dice ? 1.0 - tot / (double)n : 0.0;
    double ce = 0.0, denom = 1.0;
    if (p.has_ce) {
      if (p.focal) denom = p.mean ? (double)p.P : 1.0;        // focal_loss.mean()/.sum() (losses.py:185-187)
      else denom = p.mean ? p.acc[1] : 1.0;                   // CE mean over non-ignored (weighted) pixels
      ce = p.acc[0] / denom;
    }
    double exact = (double)p.w_ce * (double)p.out_scale / denom;
    p.dev_scale[0] = (float)exact;
    p.dev_scale[1] = (p.static_scale != 0.f) ? (float)(exact / (double)p.static_scale) : 1.f;
    p.out[0] = (float)ce;
    p.out[1] = (float)dice;
    p.out[2] = (float)(((double)p.w_ce * ce + (double)p.w_dice * dice) * (double)p.out_scale);
    p.out[3] = (float)p.acc[2];
  }
}

// y[i] *= *s unless *s == 1 (then every block exits immediately).
template <typename T, int VEC>
__global__ void scale_by_dev_scalar_kernel(T* __restrict__ x, long long n, const float* __restrict__ s) {
  const float f = *s;
  if (f == 1.f) return;
  long long nv = n / VEC;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv;
       i += (long long)gridDim.x * blockDim.x) {
    float v[VEC];
    ld_vec<VEC>(x + i * VEC, v);
#pragma unroll
    for (int j = 0; j < VEC; ++j) v[j] *= f;
    st_vec<VEC>(x + i * VEC, v);
  }
}

template <typename T, int CPAD, int VEC, int PASS>
int launch_seg(const SegLossParams& p, cudaStream_t st) {
  const long long nvec = p.HW / VEC;
  long long bx = (nvec + kLossThreads - 1) / kLossThreads;
  long long want = (2LL * num_sms() + p.B - 1) / p.B;  // ~2 CTAs per SM in total (1 resident each)
  if (PASS == 1) want = (num_sms() + p.B - 1) / p.B;   // fewer, longer CTAs: amortise the block reductions
  if (bx > want) bx = want;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)p.B);
  size_t smem = (3 * CPAD + 3 * CPAD * (kLossThreads / 32) + 3 * (kLossThreads / 32)) * sizeof(float);
  seg_loss_kernel<T, CPAD, VEC, PASS><<<grid, kLossThreads, smem, st>>>(p);
  UDA_LAUNCH_OK("seg_loss_kernel");
  return UDA_OK;
}


// ------------------------------------------------------------------------------------------------
// Streaming (cp.async.bulk) version of seg_loss_kernel: same passes, same arithmetic, hard targets only.
// A CTA owns a contiguous range of tiles, so it crosses an image boundary at most a few times; the per-image
// state (Dice coefficients in, Dice sums out) is reloaded / flushed at those crossings.
// ------------------------------------------------------------------------------------------------
// ex2.approx on a non-positive argument: one MUFU, no denormal scaling (tiny results flush to zero)
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <typename T> __device__ __forceinline__ float ld_one(const uint8_t* row, int px);
template <> __device__ __forceinline__ float ld_one<float>(const uint8_t* r, int px) { return reinterpret_cast<const float*>(r)[px]; }
template <> __device__ __forceinline__ float ld_one<__nv_bfloat16>(const uint8_t* r, int px) {
  return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(r)[px]);
}
template <typename T> __device__ __forceinline__ void st_one(uint8_t* row, int px, float v);
template <> __device__ __forceinline__ void st_one<float>(uint8_t* r, int px, float v) { reinterpret_cast<float*>(r)[px] = v; }
template <> __device__ __forceinline__ void st_one<__nv_bfloat16>(uint8_t* r, int px, float v) {
  reinterpret_cast<__nv_bfloat16*>(r)[px] = __float2bfloat16_rn(v);
}

// The arithmetic is arranged for the instruction budget of an HBM-bound kernel (ncu on the first version: 29
// instructions per logit, issue-bound with 8 compute warps): per logit one FMNMX, one FFMA + MUFU.EX2, one FADD and
// one FMUL (+ one FFMA for the Dice sums / two for the Dice gradient); everything that depends on the target class
// (picked logit, one-hot terms of the gradient, Dice intersection / target counts) is done once per PIXEL through a
// shared-memory access at the target's row instead of a compare-select per class.
template <typename T, int CPAD, int PPT, int PASS, bool FULL>
__global__ void __launch_bounds__(pxstream::px_threads<PPT>(), 1)
seg_loss_stream_kernel(const SegLossParams p, const pxstream::PxIO io, const __grid_constant__ pxstream::PxMaps maps) {
  using namespace pxstream;
  constexpr int NT = px_compute_threads<PPT>();
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* tail = smem + (size_t)io.stages * io.stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  float* coefA = reinterpret_cast<float*>(tail + 64);   // [CPAD]
  float* coefB = coefA + CPAD;                           // [CPAD]
  float* wsm = coefB + CPAD;                             // [CPAD]
  // Dice intersection / target pixel counts of the current image, PRIVATE PER WARP ([nw][CPAD] each): with one copy per
  // CTA the two shared-memory atomics per pixel of up to 512 pixels land on <= 24 addresses and serialise (~20-way) —
  // that, not the loads, bounded pass 1 of CE + Dice (2.9 TB/s)
  constexpr int kNw = NT / 32;
  float* inter_s = wsm + CPAD;                           // [kNw][CPAD]
  float* tsum_s = inter_s + kNw * CPAD;                  // [kNw][CPAD]
  float* red = tsum_s + kNw * CPAD;                      // reduction scratch
  const int C = p.C;
  if (threadIdx.x < CPAD) wsm[threadIdx.x] = (threadIdx.x < C && p.class_w) ? p.class_w[threadIdx.x] : 1.f;
  for (int i = threadIdx.x; i < kNw * CPAD; i += blockDim.x) { inter_s[i] = 0.f; tsum_s[i] = 0.f; }
  PixelPipe<PPT, NT> pipe(io, smem, bars, &maps);    // (its constructor synchronises the CTA)
  const float ce_scale = (PASS == 2) ? (p.has_ce ? *p.dev_scale : 0.f) : p.ce_grad_scale;
  const bool dice = (PASS == 2) && p.has_dice;
  const bool sums = (PASS == 1) && p.has_dice;

  float loss_sum = 0.f, denom_sum = 0.f, invalid = 0.f;
  float psum[CPAD];
  if constexpr (PASS == 1) {
#pragma unroll
    for (int c = 0; c < CPAD; ++c) psum[c] = 0.f;
  }
  auto flush_dice = [&](int b) {
    if constexpr (PASS == 1) {
      const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = NT >> 5;
#pragma unroll
      for (int c = 0; c < CPAD; ++c) {
        const float bsum = warp_sum(psum[c]);
        if (lane == 0) red[c * nw + wid] = bsum;
        psum[c] = 0.f;
      }
      named_sync(1, NT);   // also orders every thread's shared-memory atomics of this image
      for (int c = threadIdx.x; c < C; c += NT) {
        float sp = 0.f, si = 0.f, stt = 0.f;
        for (int w = 0; w < nw; ++w) {
          sp += red[c * nw + w];
          si += inter_s[w * CPAD + c]; stt += tsum_s[w * CPAD + c];
          inter_s[w * CPAD + c] = 0.f; tsum_s[w * CPAD + c] = 0.f;
        }
        double* dst = p.dice_sums + ((long long)b * C + c) * 3;
        atomicAdd(dst + 0, (double)si);
        atomicAdd(dst + 1, (double)sp);
        atomicAdd(dst + 2, (double)stt);
      }
      named_sync(1, NT);
    }
  };

  int cur_b = -1;
  if (pipe.is_io()) {
    pipe.io_loop();
  } else {
    for (int k = 0; k < pipe.n_my; ++k) {
      const int b = pipe.image_of(k);
      if (b != cur_b) {
        if (sums && cur_b >= 0) flush_dice(cur_b);
        if (dice) {
          named_sync(1, NT);
          if (threadIdx.x < CPAD) {
            const int c = threadIdx.x;
            coefA[c] = c < C ? p.dice_coef[((long long)b * C + c) * 2 + 0] : 0.f;
            coefB[c] = c < C ? p.dice_coef[((long long)b * C + c) * 2 + 1] : 0.f;
          }
          named_sync(1, NT);
        }
        cur_b = b;
      }
      pipe.wait(k);
      const int px = threadIdx.x * PPT;
      if (px < pipe.npix_of(k)) {
        // row c of the tile, at this thread's pixels: compile-time offsets from one base address
        constexpr int RS = kPxTile * (int)sizeof(T);
        uint8_t* const rows = pipe.stage(k);
        uint8_t* const col = rows + (size_t)px * sizeof(T);
        float z[CPAD][PPT], m[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) m[j] = -INFINITY;
#pragma unroll
        for (int c = 0; c < CPAD; ++c) {
          if (FULL || c < C) {
            ld_px<T, PPT>(col + c * RS, 0, z[c]);
#pragma unroll
            for (int j = 0; j < PPT; ++j) m[j] = fmaxf(m[j], z[c][j]);
          }
        }
        // per-pixel target work: validity, picked logit (read from the target's row before it is overwritten)
        int y[PPT];
        float zy[PPT];
        const long long* trow = pipe.target_row(k);
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          const long long t = trow[px + j];
          const bool ign = (t == p.ignore_index);
          const bool ok = (t >= 0 && t < C);
          y[j] = (ok && !ign) ? (int)t : -1;
          if (!ok && !ign) invalid += 1.f;
          zy[j] = y[j] >= 0 ? ld_one<T>(col + y[j] * RS, j) : 0.f;
        }
        float s[PPT], nm2[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) { s[j] = 0.f; nm2[j] = -m[j] * kLog2e; }
#pragma unroll
        for (int c = 0; c < CPAD; ++c) {
          if (FULL || c < C) {
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
              z[c][j] = ex2_fast(fmaf(z[c][j], kLog2e, nm2[j]));
              s[j] += z[c][j];
            }
          }
        }
        float inv[PPT], coef[PPT], ey[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          inv[j] = 1.f / s[j];
          ey[j] = ex2_fast(fmaf(zy[j], kLog2e, nm2[j]));   // bit-identical to the loop's value for class y
          coef[j] = 0.f;
          if (p.has_ce && y[j] >= 0) {
            const float wy = wsm[y[j]];
            const float nll = (m[j] - zy[j]) + logf(s[j]);
            const float ce = wy * nll;
            if (p.focal) {
              const float pt = expf(-ce);
              const float omp = 1.f - pt;
              const float pw = powf(omp, p.gamma);
              const float dpw = p.gamma * powf(omp, p.gamma - 1.f);
              if (PASS != 2) { loss_sum += p.alpha * pw * ce; }
              coef[j] = p.alpha * (pw + dpw * pt * ce) * wy * ce_scale;
            } else {
              if (PASS != 2) { loss_sum += ce; denom_sum += wy; }
              coef[j] = wy * ce_scale;
            }
          }
        }
        if constexpr (PASS == 1) {
          if (sums) {
#pragma unroll
            for (int c = 0; c < CPAD; ++c) {
              if (FULL || c < C) {
#pragma unroll
                for (int j = 0; j < PPT; ++j) psum[c] = fmaf(z[c][j], inv[j], psum[c]);
              }
            }
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
              if (y[j] >= 0) {
                atomicAdd(inter_s + (threadIdx.x >> 5) * CPAD + y[j], ey[j] * inv[j]);
                atomicAdd(tsum_s + (threadIdx.x >> 5) * CPAD + y[j], 1.f);
              }
            }
          }
        } else {
          if (dice) {
            // dL/dz_c = p_c * (coef + u_c - sum_k u_k p_k) - coef*onehot_c,   u_c = coefA_c*onehot_c - coefB_c
            float dotB[PPT];
#pragma unroll
            for (int j = 0; j < PPT; ++j) dotB[j] = 0.f;
#pragma unroll
            for (int c = 0; c < CPAD; ++c) {
              if (FULL || c < C) {
                const float cb = coefB[c];
#pragma unroll
                for (int j = 0; j < PPT; ++j) dotB[j] = fmaf(cb, z[c][j], dotB[j]);
              }
            }
            float k1[PPT], cay[PPT];
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
              cay[j] = y[j] >= 0 ? coefA[y[j]] : 0.f;
              const float dot = (cay[j] * ey[j] - dotB[j]) * inv[j];
              k1[j] = inv[j] * (coef[j] - dot);
            }
#pragma unroll
            for (int c = 0; c < CPAD; ++c) {
              if (FULL || c < C) {
                const float cb = coefB[c];
                float g[PPT];
#pragma unroll
                for (int j = 0; j < PPT; ++j) g[j] = z[c][j] * fmaf(-cb, inv[j], k1[j]);
                st_px<T, PPT>(col + c * RS, 0, g);
              }
            }
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
              if (y[j] >= 0) {
                const float gy = ey[j] * (fmaf(-coefB[y[j]], inv[j], k1[j]) + cay[j] * inv[j]) - coef[j];
                st_one<T>(col + y[j] * RS, j, gy);
              }
            }
          } else {
            float kk[PPT];
#pragma unroll
            for (int j = 0; j < PPT; ++j) kk[j] = inv[j] * coef[j];
#pragma unroll
            for (int c = 0; c < CPAD; ++c) {
              if (FULL || c < C) {
                float g[PPT];
#pragma unroll
                for (int j = 0; j < PPT; ++j) g[j] = z[c][j] * kk[j];
                st_px<T, PPT>(col + c * RS, 0, g);
              }
            }
#pragma unroll
            for (int j = 0; j < PPT; ++j)
              if (y[j] >= 0) st_one<T>(col + y[j] * RS, j, fmaf(ey[j], kk[j], -coef[j]));
          }
        }
      }
      pipe.release(k);
    }
    if (sums && cur_b >= 0) flush_dice(cur_b);
  }
  if constexpr (PASS != 2) {
    float v[3] = {loss_sum, denom_sum, invalid}, o[3];
    // separate scratch: the IO warp gets here while the compute warps may still be inside flush_dice (red)
    block_sum<3>(v, red + CPAD * 17, o);
    if (threadIdx.x == 0) {
      if (o[0] != 0.f) atomicAdd(p.acc + 0, (double)o[0]);
      if (o[1] != 0.f) atomicAdd(p.acc + 1, (double)o[1]);
      if (o[2] != 0.f) atomicAdd(p.acc + 2, (double)o[2]);
    }
  }
}

inline bool loss_stream_enabled() {
  static const bool on = [] { const char* e = getenv("UDA_B200_LOSS_STREAM"); return !(e && e[0] == '0'); }();
  return on;
}

template <typename T, int CPAD, int PPT, int PASS, bool FULL>
int launch_seg_stream_t(const SegLossParams& p, const pxstream::PxIO& io, const pxstream::PxMaps& maps, cudaStream_t st) {
  constexpr int kNw = pxstream::px_compute_threads<PPT>() / 32;
  const size_t smem = 128 + (size_t)io.stages * io.stage_bytes + 64 + (3 + 2 * kNw) * CPAD * 4 +
                      (CPAD * 17 + 64) * 4;
  auto kfn = seg_loss_stream_kernel<T, CPAD, PPT, PASS, FULL>;
  static bool attr = false;
  if (!attr) {
    UDA_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  kfn<<<pxstream::px_grid(io), pxstream::px_threads<PPT>(), smem, st>>>(p, io, maps);
  UDA_LAUNCH_OK("seg_loss_stream_kernel");
  return UDA_OK;
}

// returns 1 when the streaming kernel was launched, 0 when the register kernel must be used, <0 on error
template <typename T, int PASS>
int try_seg_stream(const SegLossParams& p, cudaStream_t st) {
  if (!loss_stream_enabled() || !p.target || p.soft_target) return 0;
  pxstream::PxIO io{};
  io.in[0] = (const uint8_t*)p.logits;
  io.out[0] = (uint8_t*)p.grad;
  io.target = p.target;
  io.nten = 1; io.nout = (PASS == 1) ? 0 : 1;
  io.B = p.B; io.C = p.C; io.HW = p.HW; io.esize = (int)sizeof(T);
  int ppt = 0;
  // CE + Dice (two passes): 16 compute warps with one pixel each measured faster than 8 with a pixel pair (280 vs 310 us)
  pxstream::PxMaps maps;
  if (!pxstream::plan_px(io, ppt, 0, &maps, pxstream::kPxTile, p.has_dice ? 1 : pxstream::kDefaultPpt)) return 0;
  int rc;
#define UDA_SEG_STREAM(CP)                                                                         \
  rc = (p.C == CP) ? ((ppt == 2) ? launch_seg_stream_t<T, CP, 2, PASS, true>(p, io, maps, st)            \
                                 : launch_seg_stream_t<T, CP, 1, PASS, true>(p, io, maps, st))           \
                   : ((ppt == 2) ? launch_seg_stream_t<T, CP, 2, PASS, false>(p, io, maps, st)           \
                                 : launch_seg_stream_t<T, CP, 1, PASS, false>(p, io, maps, st))
  if (p.C <= 8) UDA_SEG_STREAM(8);
  else if (p.C <= 24) UDA_SEG_STREAM(24);
  else UDA_SEG_STREAM(32);
#undef UDA_SEG_STREAM
  return rc == UDA_OK ? 1 : rc;
}

template <typename T, int PASS>
int dispatch_seg(const SegLossParams& p, bool vec4, cudaStream_t st) {
  const int C = p.C;
  {
    const int rs = try_seg_stream<T, PASS>(p, st);
    if (rs != 0) return rs == 1 ? UDA_OK : rs;
  }
  if (vec4) {
    if (C <= 8) return launch_seg<T, 8, 4, PASS>(p, st);
    if (C <= 16) return launch_seg<T, 16, 4, PASS>(p, st);
    if (C <= 24) return launch_seg<T, 24, 4, PASS>(p, st);
    if (C <= 32) return launch_seg<T, 32, 4, PASS>(p, st);
    if (C <= 64) return launch_seg<T, 64, 1, PASS>(p, st);
  } else {
    if (C <= 8) return launch_seg<T, 8, 1, PASS>(p, st);
    if (C <= 16) return launch_seg<T, 16, 1, PASS>(p, st);
    if (C <= 24) return launch_seg<T, 24, 1, PASS>(p, st);
    if (C <= 32) return launch_seg<T, 32, 1, PASS>(p, st);
    if (C <= 64) return launch_seg<T, 64, 1, PASS>(p, st);
  }
  return set_error(UDA_ERR_UNSUPPORTED, "seg loss: C=%d > 64 classes is not supported", C);
}

template <typename T>
bool vec4_ok(const void* logits, const void* grad, const void* soft, long long HW) {
  return (HW % 4 == 0) && aligned<T>(logits, 4 * sizeof(T)) && (!grad || aligned<T>(grad, 4 * sizeof(T))) &&
         (!soft || aligned<float>(soft, 16));
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Consistency (symmetric KL with temperature), entropy minimisation, BCE-with-logits
// ------------------------------------------------------------------------------------------------
namespace {

// L = scale * sum_px [KL(p1||p2) + KL(p2||p1)],  p_i = softmax(z_i / T)      (losses.py:74-90)
// With a=z1/T, b=z2/T, d=b-a, e1=exp(a-m1), e2=exp(b-m2), s_i=sum e_i, w_i=sum e_i d,
// Delta = lse2-lse1:   KL12 = Delta - w1/s1,  KL21 = w2/s2 - Delta,  D_k = d_k - Delta,
//   dL/dz1_k = [-(p2_k-p1_k) - p1_k (D_k + KL12)] / T,   dL/dz2_k = [(p2_k-p1_k) + p2_k (D_k - KL21)] / T.
// One exp per logit; both gradients and the loss come out of the same single pass.
template <typename T, int CPAD, int VEC>
__global__ void __launch_bounds__(kLossThreads, 1)
consistency_kernel(const T* __restrict__ z1p, const T* __restrict__ z2p, T* __restrict__ g1p,
                   T* __restrict__ g2p, double* __restrict__ acc, int C, long long HW, float inv_T,
                   float scale) {
  extern __shared__ float smem[];
  const int b = blockIdx.y;
  const long long nvec = HW / VEC;
  const long long off = (long long)b * C * HW;
  float lsum = 0.f;
  for (long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec;
       iv += (long long)gridDim.x * blockDim.x) {
    const long long px = iv * VEC;
    float e1[CPAD][VEC], e2[CPAD][VEC], d[CPAD][VEC];
#pragma unroll
    for (int c = 0; c < CPAD; ++c) {
      if (c < C) {
        ld_vec<VEC>(z1p + off + (long long)c * HW + px, e1[c]);
        ld_vec<VEC>(z2p + off + (long long)c * HW + px, e2[c]);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          e1[c][j] *= inv_T; e2[c][j] *= inv_T;
          d[c][j] = e2[c][j] - e1[c][j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) { e1[c][j] = -INFINITY; e2[c][j] = -INFINITY; d[c][j] = 0.f; }
      }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float m1 = e1[0][j], m2 = e2[0][j];
#pragma unroll
      for (int c = 1; c < CPAD; ++c) { m1 = fmaxf(m1, e1[c][j]); m2 = fmaxf(m2, e2[c][j]); }
      float s1 = 0.f, s2 = 0.f, w1 = 0.f, w2 = 0.f;
#pragma unroll
      for (int c = 0; c < CPAD; ++c) {
        float x1 = exp2f((e1[c][j] - m1) * kLog2e), x2 = exp2f((e2[c][j] - m2) * kLog2e);
        e1[c][j] = x1; e2[c][j] = x2;
        s1 += x1; s2 += x2;
        w1 += x1 * d[c][j]; w2 += x2 * d[c][j];
      }
      const float r1 = 1.f / s1, r2 = 1.f / s2;
      const float delta = (m2 + logf(s2)) - (m1 + logf(s1));
      const float kl12 = delta - w1 * r1, kl21 = w2 * r2 - delta;
      lsum += kl12 + kl21;
      const float gs = inv_T * scale;
#pragma unroll
      for (int c = 0; c < CPAD; ++c) {
        float p1 = e1[c][j] * r1, p2 = e2[c][j] * r2;
        float D = d[c][j] - delta, q = p2 - p1;
        e1[c][j] = (-q - p1 * (D + kl12)) * gs;
        e2[c][j] = (q + p2 * (D - kl21)) * gs;
      }
    }
#pragma unroll
    for (int c = 0; c < CPAD; ++c) {
      if (c < C) {
        st_vec<VEC>(g1p + off + (long long)c * HW + px, e1[c]);
        st_vec<VEC>(g2p + off + (long long)c * HW + px, e2[c]);
      }
    }
  }
  float v[1] = {lsum}, o[1];
  block_sum<1>(v, smem, o);
  if (threadIdx.x == 0 && o[0] != 0.f) atomicAdd(acc, (double)o[0] * (double)scale);
}

// H = -sum_c p log p per pixel; L = scale * sum_px H; dL/dz_k = -scale * p_k (log p_k + H).
template <typename T, int CPAD, int VEC>
__global__ void __launch_bounds__(kLossThreads, 1)
entropy_kernel(const T* __restrict__ zp, T* __restrict__ gp, double* __restrict__ acc, int C,
               long long HW, float scale) {
  extern __shared__ float smem[];
  const int b = blockIdx.y;
  const long long nvec = HW / VEC;
  const long long off = (long long)b * C * HW;
  float lsum = 0.f;
  for (long long iv = (long long)blockIdx.x * blockDim.x + threadIdx.x; iv < nvec;
       iv += (long long)gridDim.x * blockDim.x) {
    const long long px = iv * VEC;
    float e[CPAD][VEC], x[CPAD][VEC];
#pragma unroll
    for (int c = 0; c < CPAD; ++c) {
      if (c < C) {
        ld_vec<VEC>(zp + off + (long long)c * HW + px, x[c]);
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) x[c][j] = -INFINITY;
      }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float m = x[0][j];
#pragma unroll
      for (int c = 1; c < CPAD; ++c) m = fmaxf(m, x[c][j]);
      float s = 0.f, w = 0.f;
#pragma unroll
      for (int c = 0; c < CPAD; ++c) {
        float xm = (c < C) ? x[c][j] - m : 0.f;
        float ex = (c < C) ? exp2f(xm * kLog2e) : 0.f;
        x[c][j] = xm; e[c][j] = ex;
        s += ex; w += ex * xm;
      }
      const float r = 1.f / s, ls = logf(s);
      const float H = ls - w * r;  // -(sum p (xm - ls))
      lsum += H;
#pragma unroll
      for (int c = 0; c < CPAD; ++c) e[c][j] = -scale * (e[c][j] * r) * ((x[c][j] - ls) + H);
    }
#pragma unroll
    for (int c = 0; c < CPAD; ++c)
      if (c < C) st_vec<VEC>(gp + off + (long long)c * HW + px, e[c]);
  }
  float v[1] = {lsum}, o[1];
  block_sum<1>(v, smem, o);
  if (threadIdx.x == 0 && o[0] != 0.f) atomicAdd(acc, (double)o[0] * (double)scale);
}

__global__ void acc_to_float_kernel(const double* __restrict__ acc, float* __restrict__ out, int n) {
  if ((int)threadIdx.x < n) out[threadIdx.x] = (float)acc[threadIdx.x];
}

// mean BCE-with-logits over n elements against a constant label y (losses.py:29-36,49-51);
// out[0] += scale * mean, grad = scale * (sigmoid(x) - y) / n.  n is tiny ([B,1]): one CTA.
__global__ void bce_logits_kernel(const float* __restrict__ x, float* __restrict__ grad, float* __restrict__ out,
                                  long long n, float y, float scale, int accumulate) {
  __shared__ float red[32];
  float s = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    float v = x[i];
    s += fmaxf(v, 0.f) - v * y + log1pf(expf(-fabsf(v)));
    if (grad) grad[i] = scale * (1.f / (1.f + expf(-v)) - y) / (float)n;
  }
  float v1[1] = {s}, o[1];
  block_sum<1>(v1, red, o);
  if (threadIdx.x == 0) {
    float r = scale * o[0] / (float)n;
    out[0] = accumulate ? out[0] + r : r;
  }
}


// ---- streaming versions (see loss_stream.cuh): tiles through shared memory, arithmetic arranged like
// seg_loss_stream_kernel's (one FMNMX, one FFMA + MUFU.EX2 and a handful of FADD/FFMA per logit) ----
constexpr float kLn2 = 0.6931471805599453f;

// (TP = 256-pixel tiles: two logit tensors in, two gradients out — 512-pixel tiles leave room for only TWO stages of
// 98 KB and the kernel sat at 3.2 TB/s whatever the warp count; four 49 KB stages keep loads, compute and stores apart)
template <typename T, int CPAD, int PPT, int TP>
__global__ void __launch_bounds__(pxstream::px_threads<PPT, TP>(), 1)
consistency_stream_kernel(const pxstream::PxIO io, const __grid_constant__ pxstream::PxMaps maps, double* __restrict__ acc,
                          float inv_T, float scale) {
  using namespace pxstream;
  constexpr int NT = px_compute_threads<PPT, TP>();
  constexpr int RS = TP * (int)sizeof(T);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* tail = smem + (size_t)io.stages * io.stage_bytes;
  PixelPipe<PPT, NT> pipe(io, smem, reinterpret_cast<uint64_t*>(tail), &maps);
  float* red = reinterpret_cast<float*>(tail + 64);
  const int C = io.C;
  const float k2 = inv_T * kLog2e;        // logits -> log2 units of the tempered softmax
  const float gs = inv_T * scale;
  float lsum = 0.f;
  if (pipe.is_io()) {
    pipe.io_loop();
  } else {
    for (int k = 0; k < pipe.n_my; ++k) {
      pipe.wait(k);
      const int px = threadIdx.x * PPT;
      if (px < pipe.npix_of(k)) {
        uint8_t* const col1 = pipe.stage(k) + (size_t)px * sizeof(T);
        uint8_t* const col2 = col1 + (size_t)C * RS;
        float e1[CPAD][PPT], e2[CPAD][PPT], dz[CPAD][PPT], m1[PPT], m2[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) { m1[j] = -INFINITY; m2[j] = -INFINITY; }
#pragma unroll
        for (int c = 0; c < CPAD; ++c) {
          if (c < C) {
            ld_px<T, PPT>(col1 + c * RS, 0, e1[c]);
            ld_px<T, PPT>(col2 + c * RS, 0, e2[c]);
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
              m1[j] = fmaxf(m1[j], e1[c][j]); m2[j] = fmaxf(m2[j], e2[c][j]);
              dz[c][j] = e2[c][j] - e1[c][j];
            }
          }
        }
        float s1[PPT], s2[PPT], w1[PPT], w2[PPT], n1[PPT], n2[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          s1[j] = 0.f; s2[j] = 0.f; w1[j] = 0.f; w2[j] = 0.f;
          n1[j] = -m1[j] * k2; n2[j] = -m2[j] * k2;
        }
#pragma unroll
        for (int c = 0; c < CPAD; ++c) {
          if (c < C) {
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
              const float x1 = ex2_fast(fmaf(e1[c][j], k2, n1[j])), x2 = ex2_fast(fmaf(e2[c][j], k2, n2[j]));
              e1[c][j] = x1; e2[c][j] = x2;
              s1[j] += x1; s2[j] += x2;
              w1[j] = fmaf(x1, dz[c][j], w1[j]); w2[j] = fmaf(x2, dz[c][j], w2[j]);
            }
          }
        }
        float u1[PPT], u2[PPT], A1[PPT], A2[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          const float r1 = 1.f / s1[j], r2 = 1.f / s2[j];
          // delta = logsumexp(z2/T) - logsumexp(z1/T);  d_c = (z2_c - z1_c)/T
          const float delta = (m2[j] - m1[j]) * inv_T + (logf(s2[j]) - logf(s1[j]));
          const float kl12 = delta - w1[j] * r1 * inv_T, kl21 = w2[j] * r2 * inv_T - delta;
          lsum += kl12 + kl21;
          u1[j] = gs * r1; u2[j] = gs * r2;
          A1[j] = 1.f + delta - kl12; A2[j] = 1.f - delta - kl21;
        }
#pragma unroll
        for (int c = 0; c < CPAD; ++c) {
          if (c < C) {
            float g1[PPT], g2[PPT];
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
              // g1 = gs*(p1*(A1 - d) - p2),  g2 = gs*(p2*(A2 + d) - p1)
              const float x1 = e1[c][j] * u1[j], x2 = e2[c][j] * u2[j];
              g1[j] = fmaf(x1, fmaf(dz[c][j], -inv_T, A1[j]), -x2);
              g2[j] = fmaf(x2, fmaf(dz[c][j], inv_T, A2[j]), -x1);
            }
            st_px<T, PPT>(col1 + c * RS, 0, g1);
            st_px<T, PPT>(col2 + c * RS, 0, g2);
          }
        }
      }
      pipe.release(k);
    }
  }
  float v[1] = {lsum}, o[1];
  block_sum<1>(v, red, o);
  if (threadIdx.x == 0 && o[0] != 0.f) atomicAdd(acc, (double)o[0] * (double)scale);
}

template <typename T, int CPAD, int PPT>
__global__ void __launch_bounds__(pxstream::px_threads<PPT>(), 1)
entropy_stream_kernel(const pxstream::PxIO io, const __grid_constant__ pxstream::PxMaps maps, double* __restrict__ acc,
                      float scale) {
  using namespace pxstream;
  constexpr int NT = px_compute_threads<PPT>();
  constexpr int RS = kPxTile * (int)sizeof(T);
  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* tail = smem + (size_t)io.stages * io.stage_bytes;
  PixelPipe<PPT, NT> pipe(io, smem, reinterpret_cast<uint64_t*>(tail), &maps);
  float* red = reinterpret_cast<float*>(tail + 64);
  const int C = io.C;
  float lsum = 0.f;
  if (pipe.is_io()) {
    pipe.io_loop();
  } else {
    for (int k = 0; k < pipe.n_my; ++k) {
      pipe.wait(k);
      const int px = threadIdx.x * PPT;
      if (px < pipe.npix_of(k)) {
        uint8_t* const col = pipe.stage(k) + (size_t)px * sizeof(T);
        float e[CPAD][PPT], t[CPAD][PPT], m[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) m[j] = -INFINITY;
#pragma unroll
        for (int c = 0; c < CPAD; ++c) {
          if (c < C) {
            ld_px<T, PPT>(col + c * RS, 0, t[c]);
#pragma unroll
            for (int j = 0; j < PPT; ++j) m[j] = fmaxf(m[j], t[c][j]);
          }
        }
        float s[PPT], w[PPT], nm[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) { s[j] = 0.f; w[j] = 0.f; nm[j] = -m[j] * kLog2e; }
#pragma unroll
        for (int c = 0; c < CPAD; ++c) {
          if (c < C) {
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
              t[c][j] = fmaf(t[c][j], kLog2e, nm[j]);        // (x - max) in log2 units
              e[c][j] = ex2_fast(t[c][j]);
              s[j] += e[c][j];
              w[j] = fmaf(e[c][j], t[c][j], w[j]);
            }
          }
        }
        float K[PPT], q0[PPT];
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
          const float r = 1.f / s[j], ls = logf(s[j]);
          const float H = ls - w[j] * r * kLn2;
          lsum += H;
          // dH/dx_c = -p_c * ((x_c - max) - ls + H)
          K[j] = -scale * r;
          q0[j] = H - ls;
        }
#pragma unroll
        for (int c = 0; c < CPAD; ++c) {
          if (c < C) {
            float g[PPT];
#pragma unroll
            for (int j = 0; j < PPT; ++j) g[j] = e[c][j] * K[j] * fmaf(t[c][j], kLn2, q0[j]);
            st_px<T, PPT>(col + c * RS, 0, g);
          }
        }
      }
      pipe.release(k);
    }
  }
  float v[1] = {lsum}, o[1];
  block_sum<1>(v, red, o);
  if (threadIdx.x == 0 && o[0] != 0.f) atomicAdd(acc, (double)o[0] * (double)scale);
}

template <int PPT, int TP = pxstream::kPxTile, typename K, typename... Args>
int launch_px_stream(K kfn, const pxstream::PxIO& io, const pxstream::PxMaps& maps, cudaStream_t st, const char* what,
                     Args... args) {
  const size_t smem = 128 + (size_t)io.stages * io.stage_bytes + 64 + 64 * 4;
  UDA_CUDA_OK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  kfn<<<pxstream::px_grid(io), pxstream::px_threads<PPT, TP>(), smem, st>>>(io, maps, args...);
  UDA_LAUNCH_OK(what);
  return UDA_OK;
}

#define UDA_PX_DISPATCH(KERN, T, ...)                                                               \
  do {                                                                                              \
    if (ppt == 2) {                                                                                 \
      if (io.C <= 8) return launch_px_stream<2>(KERN<T, 8, 2>, io, maps, st, #KERN, __VA_ARGS__);            \
      if (io.C <= 16) return launch_px_stream<2>(KERN<T, 16, 2>, io, maps, st, #KERN, __VA_ARGS__);          \
      if (io.C <= 24) return launch_px_stream<2>(KERN<T, 24, 2>, io, maps, st, #KERN, __VA_ARGS__);          \
      return launch_px_stream<2>(KERN<T, 32, 2>, io, maps, st, #KERN, __VA_ARGS__);                          \
    }                                                                                               \
    if (io.C <= 8) return launch_px_stream<1>(KERN<T, 8, 1>, io, maps, st, #KERN, __VA_ARGS__);              \
    if (io.C <= 16) return launch_px_stream<1>(KERN<T, 16, 1>, io, maps, st, #KERN, __VA_ARGS__);            \
    if (io.C <= 24) return launch_px_stream<1>(KERN<T, 24, 1>, io, maps, st, #KERN, __VA_ARGS__);            \
    return launch_px_stream<1>(KERN<T, 32, 1>, io, maps, st, #KERN, __VA_ARGS__);                            \
  } while (0)

template <typename T>
int consistency_stream(const pxstream::PxIO& io, const pxstream::PxMaps& maps, int ppt, double* acc, float inv_T, float scale,
                       cudaStream_t st) {
  constexpr int TP = pxstream::kPxTile / 2;
#define UDA_CS(CP)                                                                                                  \
  return ppt == 2 ? launch_px_stream<2, TP>(consistency_stream_kernel<T, CP, 2, TP>, io, maps, st, "consistency_stream_kernel", acc, inv_T, scale) \
                  : launch_px_stream<1, TP>(consistency_stream_kernel<T, CP, 1, TP>, io, maps, st, "consistency_stream_kernel", acc, inv_T, scale)
  if (io.C <= 8) UDA_CS(8);
  if (io.C <= 16) UDA_CS(16);
  if (io.C <= 24) UDA_CS(24);
  UDA_CS(32);
#undef UDA_CS
}
template <typename T>
int entropy_stream(const pxstream::PxIO& io, const pxstream::PxMaps& maps, int ppt, double* acc, float scale, cudaStream_t st) {
  UDA_PX_DISPATCH(entropy_stream_kernel, T, acc, scale);
}

template <typename T, int CPAD, int VEC>
int launch_consistency(const void* z1, const void* z2, void* g1, void* g2, double* acc, int B, int C,
                       long long HW, float inv_T, float scale, cudaStream_t st) {
  const long long nvec = HW / VEC;
  long long bx = (nvec + kLossThreads - 1) / kLossThreads;
  long long want = (2LL * num_sms() + B - 1) / B;
  if (bx > want) bx = want;
  if (bx < 1) bx = 1;
  consistency_kernel<T, CPAD, VEC><<<dim3((unsigned)bx, (unsigned)B), kLossThreads, 32 * sizeof(float), st>>>(
      (const T*)z1, (const T*)z2, (T*)g1, (T*)g2, acc, C, HW, inv_T, scale);
  UDA_LAUNCH_OK("consistency_kernel");
  return UDA_OK;
}
template <typename T, int CPAD, int VEC>
int launch_entropy(const void* z, void* g, double* acc, int B, int C, long long HW, float scale,
                   cudaStream_t st) {
  const long long nvec = HW / VEC;
  long long bx = (nvec + kLossThreads - 1) / kLossThreads;
  long long want = (2LL * num_sms() + B - 1) / B;
  if (bx > want) bx = want;
  if (bx < 1) bx = 1;
  entropy_kernel<T, CPAD, VEC><<<dim3((unsigned)bx, (unsigned)B), kLossThreads, 32 * sizeof(float), st>>>(
      (const T*)z, (T*)g, acc, C, HW, scale);
  UDA_LAUNCH_OK("entropy_kernel");
  return UDA_OK;
}

#define UDA_DISPATCH_CV(FN, T, C, vec2, ...)                                         \
  do {                                                                               \
    if (vec2) {                                                                      \
      if (C <= 8) return FN<T, 8, 2>(__VA_ARGS__);                                   \
      if (C <= 16) return FN<T, 16, 2>(__VA_ARGS__);                                 \
      if (C <= 24) return FN<T, 24, 2>(__VA_ARGS__);                                 \
      if (C <= 32) return FN<T, 32, 2>(__VA_ARGS__);                                 \
    } else {                                                                         \
      if (C <= 8) return FN<T, 8, 1>(__VA_ARGS__);                                   \
      if (C <= 16) return FN<T, 16, 1>(__VA_ARGS__);                                 \
      if (C <= 24) return FN<T, 24, 1>(__VA_ARGS__);                                 \
      if (C <= 32) return FN<T, 32, 1>(__VA_ARGS__);                                 \
    }                                                                                \
    if (C <= 64) return FN<T, 64, 1>(__VA_ARGS__);                                   \
    return set_error(UDA_ERR_UNSUPPORTED, "C=%d > 64 classes is not supported", C);  \
  } while (0)

template <typename T>
int consistency_dispatch(const void* z1, const void* z2, void* g1, void* g2, double* acc, int B, int C,
                         long long HW, float inv_T, float scale, cudaStream_t st) {
  if (loss_stream_enabled()) {
    pxstream::PxIO io{};
    io.in[0] = (const uint8_t*)z1; io.in[1] = (const uint8_t*)z2;
    io.out[0] = (uint8_t*)g1; io.out[1] = (uint8_t*)g2;
    io.nten = 2; io.nout = 2; io.B = B; io.C = C; io.HW = HW; io.esize = (int)sizeof(T);
    int ppt = 0;
    pxstream::PxMaps maps;
    if (pxstream::plan_px(io, ppt, 0, &maps, pxstream::kPxTile / 2, 1)) return consistency_stream<T>(io, maps, ppt, acc, inv_T, scale, st);
  }
  bool v2 = (HW % 2 == 0) && aligned<T>(z1, 2 * sizeof(T)) && aligned<T>(z2, 2 * sizeof(T)) &&
            aligned<T>(g1, 2 * sizeof(T)) && aligned<T>(g2, 2 * sizeof(T));
  UDA_DISPATCH_CV(launch_consistency, T, C, v2, z1, z2, g1, g2, acc, B, C, HW, inv_T, scale, st);
}
template <typename T>
int entropy_dispatch(const void* z, void* g, double* acc, int B, int C, long long HW, float scale,
                     cudaStream_t st) {
  if (loss_stream_enabled()) {
    pxstream::PxIO io{};
    io.in[0] = (const uint8_t*)z; io.out[0] = (uint8_t*)g;
    io.nten = 1; io.nout = 1; io.B = B; io.C = C; io.HW = HW; io.esize = (int)sizeof(T);
    int ppt = 0;
    pxstream::PxMaps maps;
    if (pxstream::plan_px(io, ppt, 0, &maps)) return entropy_stream<T>(io, maps, ppt, acc, scale, st);
  }
  bool v2 = (HW % 2 == 0) && aligned<T>(z, 2 * sizeof(T)) && aligned<T>(g, 2 * sizeof(T));
  UDA_DISPATCH_CV(launch_entropy, T, C, v2, z, g, acc, B, C, HW, scale, st);
}

}  // namespace
}  // namespace uda

// ================================================================================================
// C ABI (declared in include/uda_b200.h)
// ================================================================================================
using namespace uda;

extern "C" size_t uda_seg_loss_workspace_bytes(int B, int C) {
  // acc[4] doubles + dice_sums[B*C*3] doubles + dice_coef[B*C*2] floats + dev_scale[2] floats (+pad)
  size_t n = 4 * sizeof(double) + (size_t)B * C * 3 * sizeof(double) + (size_t)B * C * 2 * sizeof(float) +
             4 * sizeof(float);
  return (n + 255) / 256 * 256;
}

extern "C" int uda_seg_loss_fwd_bwd(const void* logits, int dtype, const long long* target,
                                    const float* soft_target, const float* class_weights, void* grad,
                                    float* out4, void* workspace, int B, int C, long long HW, int ce_mode,
                                    int use_dice, float alpha, float gamma, int mean_reduction,
                                    long long ignore_index, float smooth, float w_ce, float w_dice,
                                    float out_scale, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(logits && grad && out4 && workspace, UDA_ERR_BAD_ARG, "seg_loss: null pointer");
  UDA_REQUIRE(dtype == UDA_F32 || dtype == UDA_BF16, UDA_ERR_BAD_ARG, "seg_loss: dtype %d", dtype);
  UDA_REQUIRE(B > 0 && C > 0 && HW > 0, UDA_ERR_BAD_ARG, "seg_loss: empty shape B=%d C=%d HW=%lld", B, C, HW);
  UDA_REQUIRE(B <= 65535, UDA_ERR_UNSUPPORTED, "seg_loss: B=%d > 65535", B);
  UDA_REQUIRE(ce_mode >= 0 && ce_mode <= 2, UDA_ERR_BAD_ARG, "seg_loss: ce_mode %d", ce_mode);
  UDA_REQUIRE(ce_mode != 0 || use_dice, UDA_ERR_BAD_ARG, "seg_loss: nothing to compute");
  UDA_REQUIRE(target || (soft_target && ce_mode == 0), UDA_ERR_BAD_ARG,
              "seg_loss: index targets are required for the CE/focal term");
  UDA_REQUIRE(aligned<double>(workspace, 8), UDA_ERR_BAD_ARG, "seg_loss: workspace must be 8-byte aligned");

  double* acc = reinterpret_cast<double*>(workspace);
  double* dice_sums = acc + 4;
  float* dice_coef = reinterpret_cast<float*>(dice_sums + (size_t)B * C * 3);
  float* dev_scale = dice_coef + (size_t)B * C * 2;
  UDA_CUDA_OK(cudaMemsetAsync(workspace, 0, 4 * sizeof(double) + (size_t)B * C * 3 * sizeof(double), st));

  SegLossParams p{};
  p.logits = logits; p.target = target; p.soft_target = soft_target; p.grad = grad;
  p.class_w = class_weights; p.acc = acc; p.dice_sums = dice_sums; p.dice_coef = dice_coef;
  p.dev_scale = dev_scale; p.B = B; p.C = C; p.HW = HW;
  p.has_ce = ce_mode != 0; p.focal = ce_mode == 2; p.has_dice = use_dice != 0;
  p.alpha = alpha; p.gamma = gamma; p.ignore_index = ignore_index;
  const double P = (double)B * (double)HW;
  const float static_scale = (float)((double)w_ce * (double)out_scale / (mean_reduction ? P : 1.0));
  p.ce_grad_scale = static_scale;

  SegFinalizeParams f{};
  f.acc = acc; f.dice_sums = dice_sums; f.dice_coef = dice_coef; f.dev_scale = dev_scale; f.out = out4;
  f.B = B; f.C = C; f.P = (long long)P; f.has_ce = p.has_ce; f.focal = p.focal; f.has_dice = p.has_dice;
  f.mean = mean_reduction; f.smooth = smooth; f.w_ce = w_ce; f.w_dice = w_dice; f.out_scale = out_scale;
  f.static_scale = use_dice ? 0.f : static_scale;

  const bool bf = dtype == UDA_BF16;
  const bool v4 = bf ? vec4_ok<bf16>(logits, grad, soft_target, HW) : vec4_ok<float>(logits, grad, soft_target, HW);
  int rc;
  if (!use_dice) {
    rc = bf ? dispatch_seg<bf16, 0>(p, v4, st) : dispatch_seg<float, 0>(p, v4, st);
    if (rc) return rc;
    seg_loss_finalize_kernel<<<1, 256, 0, st>>>(f);
    UDA_LAUNCH_OK("seg_loss_finalize_kernel");
    // ignore_index / class-weighted mean: the denominator differs from B*HW only then; the kernel
    // exits immediately when the factor is exactly 1.
    const long long n = (long long)B * C * HW;
    const int blocks = 4 * num_sms();
    if (bf) {
      if (v4) scale_by_dev_scalar_kernel<bf16, 4><<<blocks, 256, 0, st>>>((bf16*)grad, n, dev_scale + 1);
      else scale_by_dev_scalar_kernel<bf16, 1><<<blocks, 256, 0, st>>>((bf16*)grad, n, dev_scale + 1);
    } else {
      if (v4) scale_by_dev_scalar_kernel<float, 4><<<blocks, 256, 0, st>>>((float*)grad, n, dev_scale + 1);
      else scale_by_dev_scalar_kernel<float, 1><<<blocks, 256, 0, st>>>((float*)grad, n, dev_scale + 1);
    }
    UDA_LAUNCH_OK("scale_by_dev_scalar_kernel");
    return UDA_OK;
  }
  rc = bf ? dispatch_seg<bf16, 1>(p, v4, st) : dispatch_seg<float, 1>(p, v4, st);
  if (rc) return rc;
  seg_loss_finalize_kernel<<<1, 256, 0, st>>>(f);
  UDA_LAUNCH_OK("seg_loss_finalize_kernel");
  rc = bf ? dispatch_seg<bf16, 2>(p, v4, st) : dispatch_seg<float, 2>(p, v4, st);
  return rc;
}

extern "C" int uda_scale_by_device_scalar(void* x, int dtype, long long n, const float* dev_scalar, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(x && dev_scalar && n >= 0, UDA_ERR_BAD_ARG, "scale: bad argument");
  if (n == 0) return UDA_OK;
  const int blocks = 4 * num_sms();
  if (dtype == UDA_BF16) {
    if (n % 4 == 0 && aligned<bf16>(x, 8)) scale_by_dev_scalar_kernel<bf16, 4><<<blocks, 256, 0, st>>>((bf16*)x, n, dev_scalar);
    else scale_by_dev_scalar_kernel<bf16, 1><<<blocks, 256, 0, st>>>((bf16*)x, n, dev_scalar);
  } else if (dtype == UDA_F32) {
    if (n % 4 == 0 && aligned<float>(x, 16)) scale_by_dev_scalar_kernel<float, 4><<<blocks, 256, 0, st>>>((float*)x, n, dev_scalar);
    else scale_by_dev_scalar_kernel<float, 1><<<blocks, 256, 0, st>>>((float*)x, n, dev_scalar);
  } else {
    return set_error(UDA_ERR_BAD_ARG, "scale: dtype %d", dtype);
  }
  UDA_LAUNCH_OK("scale_by_dev_scalar_kernel");
  return UDA_OK;
}

extern "C" int uda_consistency_fwd_bwd(const void* z1, const void* z2, int dtype, void* grad1, void* grad2,
                                       float* out1, void* workspace, int B, int C, long long HW,
                                       float temperature, float out_scale, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(z1 && z2 && grad1 && grad2 && out1 && workspace, UDA_ERR_BAD_ARG, "consistency: null pointer");
  UDA_REQUIRE(B > 0 && C > 0 && HW > 0 && B <= 65535, UDA_ERR_BAD_ARG, "consistency: bad shape");
  UDA_REQUIRE(temperature > 0.f, UDA_ERR_BAD_ARG, "consistency: temperature must be > 0");
  double* acc = reinterpret_cast<double*>(workspace);
  UDA_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(double), st));
  const float scale = out_scale / (2.f * (float)B);  // (kl1+kl2)/2 with 'batchmean' = sum/B
  int rc;
  if (dtype == UDA_BF16) rc = consistency_dispatch<bf16>(z1, z2, grad1, grad2, acc, B, C, HW, 1.f / temperature, scale, st);
  else if (dtype == UDA_F32) rc = consistency_dispatch<float>(z1, z2, grad1, grad2, acc, B, C, HW, 1.f / temperature, scale, st);
  else return set_error(UDA_ERR_BAD_ARG, "consistency: dtype %d", dtype);
  if (rc) return rc;
  acc_to_float_kernel<<<1, 32, 0, st>>>(acc, out1, 1);
  UDA_LAUNCH_OK("acc_to_float_kernel");
  return UDA_OK;
}

extern "C" int uda_entropy_fwd_bwd(const void* z, int dtype, void* grad, float* out1, void* workspace, int B,
                                   int C, long long HW, float out_scale, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(z && grad && out1 && workspace, UDA_ERR_BAD_ARG, "entropy: null pointer");
  UDA_REQUIRE(B > 0 && C > 0 && HW > 0 && B <= 65535, UDA_ERR_BAD_ARG, "entropy: bad shape");
  double* acc = reinterpret_cast<double*>(workspace);
  UDA_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(double), st));
  const float scale = out_scale / (float)((double)B * (double)HW);
  int rc;
  if (dtype == UDA_BF16) rc = entropy_dispatch<bf16>(z, grad, acc, B, C, HW, scale, st);
  else if (dtype == UDA_F32) rc = entropy_dispatch<float>(z, grad, acc, B, C, HW, scale, st);
  else return set_error(UDA_ERR_BAD_ARG, "entropy: dtype %d", dtype);
  if (rc) return rc;
  acc_to_float_kernel<<<1, 32, 0, st>>>(acc, out1, 1);
  UDA_LAUNCH_OK("acc_to_float_kernel");
  return UDA_OK;
}

extern "C" int uda_bce_logits_fwd_bwd(const float* x, float* grad, float* out1, long long n, float label,
                                      float scale, int accumulate, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  UDA_REQUIRE(x && out1 && n > 0, UDA_ERR_BAD_ARG, "bce: bad argument");
  bce_logits_kernel<<<1, 256, 0, st>>>(x, grad, out1, n, label, scale, accumulate);
  UDA_LAUNCH_OK("bce_logits_kernel");
  return UDA_OK;
}
