/* C restatement of the integer evaluation path — TEST INFRASTRUCTURE (see oracle/__init__.py).
 *
 *   oracle_argmax   : outputs.argmax(dim=1)                 reference src/models/predict.py:129
 *   oracle_fast_hist: SegmentationMetrics._fast_hist        reference src/analysis/metrics.py:17-27
 *
 * Scalar loops, bit-exact by construction; checked against the reference's own modules through
 * tests/golden/losses_*.npz (keys "argmax", "hist", "hist_ignore0").
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* logits [B][C][HW] float32 -> mask [B][HW] int64; first maximal index, NaN counts as maximal */
void oracle_argmax(const float* logits, int64_t* mask, int B, int C, int64_t HW) {
  for (int b = 0; b < B; ++b)
    for (int64_t p = 0; p < HW; ++p) {
      const float* z = logits + (int64_t)b * C * HW + p;
      float best = z[0];
      int64_t bi = 0;
      for (int c = 1; c < C; ++c) {
        float v = z[(int64_t)c * HW];
        if (v > best || (isnan(v) && !isnan(best))) { best = v; bi = c; }
      }
      mask[(int64_t)b * HW + p] = bi;
    }
}

/* hist[C][C] (rows = true, cols = pred); pixels with true outside [0,C) or == ignore are skipped */
void oracle_fast_hist(const int64_t* pred, const int64_t* truth, int64_t* hist, int64_t n, int C,
                      int64_t ignore_index, int has_ignore) {
  memset(hist, 0, sizeof(int64_t) * (size_t)C * (size_t)C);
  for (int64_t i = 0; i < n; ++i) {
    int64_t t = truth[i];
    if (t < 0 || t >= C) continue;
    if (has_ignore && t == ignore_index) continue;
    hist[t * C + pred[i]] += 1;
  }
}
