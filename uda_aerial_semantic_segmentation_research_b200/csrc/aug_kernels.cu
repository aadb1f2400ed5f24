// Device-side strong augmentation for the unsupervised fine-tuning step, sm_100a.
//
// The reference augments every target batch on the HOST, image by image, twice per step: tensor -> numpy ->
// albumentations pipeline -> tensor -> device (src/models/unsupervised_trainer.py:99-114, pipeline
// src/models/augmentation.py:40-80).  Here the random DECISIONS stay on the host (a few numbers per image) and the
// pixel work is ONE gather pass on the device per view: the geometric members of the pipeline (RandomRotate90, Flip,
// Transpose = an element of the dihedral group D4; ShiftScaleRotate = a similarity transform) compose into one 2x3
// inverse map per image, sampled bilinearly with OpenCV's BORDER_REFLECT_101 (albumentations' default border mode);
// RandomBrightnessContrast (x * alpha + beta) and GaussNoise (additive N(0, sigma^2), counter-based generator) are
// applied to the sampled value.  HBM-bound: 12 bytes read (4 taps mostly from L1/L2) + 12 bytes written per pixel.
// Not covered (left to the host pipeline if wanted): the blur family, optical / grid / elastic distortion, CLAHE /
// Sharpen / Emboss, HueSaturationValue.
#include "common.cuh"

namespace uda {
namespace {

struct AugRow {            // one row of the parameter table (12 floats per image)
  float m00, m01, m02;     // source x = m00 * x + m01 * y + m02      (output pixel centre -> source coordinates)
  float m10, m11, m12;     // source y = m10 * x + m11 * y + m12
  float alpha, beta;       // value' = value * alpha + beta
  float sigma;             // noise standard deviation (0 = none)
  float seed;              // per-image noise stream (integer value)
  float pad0, pad1;
};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * (n - 1);
  i = i % period;
  if (i < 0) i += period;
  return i < n ? i : period - i;
}
__device__ __forceinline__ uint32_t mix32(uint32_t x) {      // lowbias32 integer hash
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float gauss(uint32_t a, uint32_t b) {   // Box-Muller on two hashed uniforms
  const float u1 = ((mix32(a) >> 8) + 1) * (1.f / 16777217.f);
  const float u2 = (mix32(b) >> 8) * (1.f / 16777216.f);
  return sqrtf(-2.f * __logf(u1)) * __cosf(6.2831853f * u2);
}

__global__ void __launch_bounds__(256)
strong_augment_kernel(const float* __restrict__ in, float* __restrict__ out, const AugRow* __restrict__ rows, int C,
                      int H, int W) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * W) return;
  const int y = i / W, x = i - y * W;
  const AugRow r = rows[b];
  const float sx = r.m00 * x + r.m01 * y + r.m02;
  const float sy = r.m10 * x + r.m11 * y + r.m12;
  const float fx = floorf(sx), fy = floorf(sy);
  const float ax = sx - fx, ay = sy - fy;
  const int x0 = reflect101((int)fx, W), x1 = reflect101((int)fx + 1, W);
  const int y0 = reflect101((int)fy, H), y1 = reflect101((int)fy + 1, H);
  const long long plane = (long long)H * W;
  const float* src = in + (long long)b * C * plane;
  float* dst = out + (long long)b * C * plane + i;
  const uint32_t s = (uint32_t)r.seed;
  for (int c = 0; c < C; ++c) {
    const float* p = src + c * plane;
    const float v00 = __ldg(p + (long long)y0 * W + x0), v01 = __ldg(p + (long long)y0 * W + x1);
    const float v10 = __ldg(p + (long long)y1 * W + x0), v11 = __ldg(p + (long long)y1 * W + x1);
    float v = (v00 * (1.f - ax) + v01 * ax) * (1.f - ay) + (v10 * (1.f - ax) + v11 * ax) * ay;
    v = v * r.alpha + r.beta;
    if (r.sigma > 0.f) {
      const uint32_t k = (uint32_t)(c * plane + i);
      v += r.sigma * gauss(k * 2654435761u + s, (k ^ 0x9e3779b9u) * 40503u + s * 7919u + 1u);
    }
    dst[c * plane] = v;
  }
}

}  // namespace
}  // namespace uda

// images / out: fp32 NCHW [B,C,H,W]; table: device array of B rows x 12 floats
// {m00, m01, m02, m10, m11, m12, alpha, beta, sigma, seed, 0, 0} (see AugRow)
extern "C" int uda_strong_augment(const float* images, float* out, const float* table, int B, int C, int H, int W,
                                  void* stream) {
  UDA_REQUIRE(images && out && table && B > 0 && C > 0 && H > 0 && W > 0, UDA_ERR_BAD_ARG, "strong_augment: bad argument");
  UDA_REQUIRE(images != out, UDA_ERR_BAD_ARG, "strong_augment: in-place operation is not supported (gather)");
  UDA_REQUIRE(B <= 65535, UDA_ERR_UNSUPPORTED, "strong_augment: at most 65535 images per call");
  dim3 grid((unsigned)((H * W + 255) / 256), (unsigned)B);
  uda::strong_augment_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(images, out, (const uda::AugRow*)table, C, H, W);
  UDA_LAUNCH_OK("strong_augment_kernel");
  return UDA_OK;
}
