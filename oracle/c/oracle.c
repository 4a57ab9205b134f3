/*
 * oracle.c — plain-C scalar restatement of the reference's per-pixel algorithm.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): compiled by oracle/c/Makefile
 * into oracle/_build/liboracle.so, loaded by tests/ and by bench.py's cpu_baseline
 * leg through ctypes.  Never linked into libmdseg_b200.so.
 *
 * Each function cites the reference lines it follows (paths relative to the
 * reference root).  fp32 arithmetic mirrors ATen's operation order; compile with
 * -ffp-contract=off so that no FMA contraction changes it.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* lib/base_dataset.py:81-82 — label = self.lb_map[label] */
void orc_lut_remap_u8(const uint8_t* in, uint8_t* out, const uint8_t* lut, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = lut[in[i]];
}

/* lib/class_remap.py:39-48,55-64 on int64 maps: values outside [0,255] -> oob */
void orc_lut_remap_i64(const int64_t* in, int64_t* out, const uint8_t* lut, int oob, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = (in[i] >= 0 && in[i] < 256) ? (int64_t)lut[in[i]] : (int64_t)oob;
}

/* evaluate.py:89-93 — hist += bincount(label[keep]*C + pred[keep]).  returns #bad labels */
int64_t orc_confusion_i64(const int64_t* label, const int64_t* pred, const uint8_t* lut, int64_t* hist, int Ca,
                          int Cb, int ignore, int64_t n) {
  int64_t bad = 0;
  for (int64_t i = 0; i < n; ++i) {
    int64_t l = label[i];
    if (lut) l = (l >= 0 && l < 256) ? lut[l] : -1;
    if (l == ignore) continue;
    if (l < 0 || l >= Ca || pred[i] < 0 || pred[i] >= Cb) { ++bad; continue; }
    hist[l * Cb + pred[i]] += 1;
  }
  return bad;
}

/* evaluate.py:156-157 — legacy nearest: src = min(floor(dst * (in/out)), in-1), fp32 scale */
void orc_nearest_i64(const int64_t* in, int Hin, int Win, int64_t* out, int Ho, int Wo) {
  const float sy = (float)Hin / (float)Ho, sx = (float)Win / (float)Wo;
  for (int y = 0; y < Ho; ++y) {
    int ys = (int)floorf((float)y * sy);
    if (ys > Hin - 1) ys = Hin - 1;
    for (int x = 0; x < Wo; ++x) {
      int xs = (int)floorf((float)x * sx);
      if (xs > Win - 1) xs = Win - 1;
      out[(int64_t)y * Wo + x] = in[(int64_t)ys * Win + xs];
    }
  }
}

static void axis(int n_in, int n_out, int dst, int* i0, int* i1, float* l0, float* l1) {
  /* ATen area_pixel_compute_source_index(scale, dst, align_corners=true) */
  const float scale = n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.0f;
  const float s = scale * (float)dst;
  *i0 = (int)s;
  if (*i0 > n_in - 1) *i0 = n_in - 1;
  *i1 = *i0 + ((*i0 < n_in - 1) ? 1 : 0);
  *l1 = s - (float)*i0;
  *l0 = 1.0f - *l1;
}

/* lib/loss/loss_cross_datasets.py:1007 + lib/loss/ohem_ce_loss.py:27 for ONE image:
 * z = bilinear(src [C,h,w] -> (Y,X)), loss = logsumexp(z) - z[label]  (0 when ignored);
 * log_softmax order as ATen: max, Σ exp(z - max), (z_l - max) - log(Σ). */
void orc_up_ce_image_f32(const float* src, int C, int h, int w, const int64_t* label, int H, int W, int ignore,
                         float* loss, float* zbuf /* C floats scratch */) {
  for (int Y = 0; Y < H; ++Y) {
    int y0, y1;
    float l0h, l1h;
    axis(h, H, Y, &y0, &y1, &l0h, &l1h);
    for (int X = 0; X < W; ++X) {
      int x0, x1;
      float l0w, l1w;
      axis(w, W, X, &x0, &x1, &l0w, &l1w);
      float m = -INFINITY;
      for (int c = 0; c < C; ++c) {
        const float* p = src + (int64_t)c * h * w;
        const float z = l0h * (l0w * p[y0 * w + x0] + l1w * p[y0 * w + x1]) +
                        l1h * (l0w * p[y1 * w + x0] + l1w * p[y1 * w + x1]);
        zbuf[c] = z;
        if (z > m) m = z;
      }
      float s = 0.0f;
      for (int c = 0; c < C; ++c) s += expf(zbuf[c] - m);
      const int64_t l = label[(int64_t)Y * W + X];
      loss[(int64_t)Y * W + X] = (l == ignore || l < 0 || l >= C) ? 0.0f : -((zbuf[l] - m) - logf(s));
    }
  }
}

/* lib/loss/ohem_ce_loss.py:25-30 on full-resolution NCHW logits of one image */
void orc_ce_image_f32(const float* logits, int C, int64_t HW, const int64_t* label, int ignore, float* loss) {
  for (int64_t p = 0; p < HW; ++p) {
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) { const float z = logits[c * HW + p]; if (z > m) m = z; }
    float s = 0.0f;
    for (int c = 0; c < C; ++c) s += expf(logits[c * HW + p] - m);
    const int64_t l = label[p];
    loss[p] = (l == ignore || l < 0 || l >= C) ? 0.0f : -((logits[l * HW + p] - m) - logf(s));
  }
}
