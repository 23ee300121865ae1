/* oracle/oracle_nn.c — fp32 restatement of the TFLite reference kernels used by
 * the five backscrub models.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Each function cites the reference file it follows (paths relative to
 * /root/reference; TF/ = tensorflow/tensorflow/).  Loop order is the reference
 * loop order; the multiply-accumulate is fmaf() and exp is or_expf() — the
 * numeric contract in oracle.h.
 */
#include "oracle.h"

#include <math.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* scalar helpers                                                            */
/* ------------------------------------------------------------------------- */

/* Fixed-sequence expf: round-to-nearest range reduction by ln2 (two-part
 * constant), degree-5 polynomial in Horner/fmaf form (Cephes expf
 * coefficients), exact power-of-two scaling in two steps.  Every operation is
 * a single IEEE-754 binary32 op, so the CUDA path can reproduce it bit for
 * bit.  Stands in for std::exp in TF/lite/kernels/internal/reference/
 * logistic.h:30-57 and expf in lib/libbackscrub.cc:350-351. */
float or_expf(float x) {
  if (x != x) return x;
  if (x > 88.7228394f) return INFINITY;
  if (x < -87.3365479f) return 0.0f; /* below the normal range: flushed */
  float n = rintf(x * 1.44269504088896341f);
  float r = fmaf(n, -0.693359375f, x);
  r = fmaf(n, 2.12194440e-4f, r);
  float p = 1.9875691500E-4f;
  p = fmaf(p, r, 1.3981999507E-3f);
  p = fmaf(p, r, 8.3334519073E-3f);
  p = fmaf(p, r, 4.1665795894E-2f);
  p = fmaf(p, r, 1.6666665459E-1f);
  p = fmaf(p, r, 5.0000001201E-1f);
  float r2 = r * r;
  float y = fmaf(p, r2, r);
  y = y + 1.0f;
  int ni = (int)n;
  int n1 = ni / 2, n2 = ni - n1;
  union { uint32_t u; float f; } s1, s2;
  s1.u = (uint32_t)(n1 + 127) << 23;
  s2.u = (uint32_t)(n2 + 127) << 23;
  return (y * s1.f) * s2.f;
}

/* IEEE binary16 -> binary32, exact (TF/lite/kernels/internal/reference/dequantize.h:31
 * widens fp16 weights the same way). */
float or_half_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1f;
  uint32_t man = h & 0x3ffu;
  union { uint32_t u; float f; } v;
  if (exp == 0) {
    if (man == 0) { v.u = sign; return v.f; }
    /* subnormal half: value = man * 2^-24, exact in binary32 */
    float f = (float)man * 5.9604644775390625e-08f;
    v.f = f; v.u |= sign; return v.f;
  }
  if (exp == 31) { v.u = sign | 0x7f800000u | (man << 13); return v.f; }
  v.u = sign | ((exp + 112u) << 23) | (man << 13);
  return v.f;
}

static inline float act_clamp(float x, int act) {
  /* TF/lite/kernels/internal/common.h:59-64 ActivationFunctionWithMinMax with the
   * ranges of TF/lite/kernels/kernel_util.h CalculateActivationRange */
  switch (act) {
    case OR_ACT_RELU: return x < 0.f ? 0.f : x;
    case OR_ACT_RELU6: { float y = x < 0.f ? 0.f : x; return y > 6.f ? 6.f : y; }
    case OR_ACT_RELU_N1_TO_1: { float y = x < -1.f ? -1.f : x; return y > 1.f ? 1.f : y; }
    default: return x;
  }
}

/* TF/lite/kernels/padding.h:23-82 (ComputeOutSize, ComputePaddingWithOffset) */
void or_conv_out_size(int in_size, int k, int stride, int dil, int padding, int* out_size, int* pad_before) {
  int eff = (k - 1) * dil + 1;
  int o = (padding == OR_PAD_SAME) ? (in_size + stride - 1) / stride
                                   : (in_size + stride - eff) / stride;
  int total = (o - 1) * stride + eff - in_size;
  if (total < 0) total = 0;
  *out_size = o;
  *pad_before = total / 2;
}

/* ------------------------------------------------------------------------- */
/* ops                                                                       */
/* ------------------------------------------------------------------------- */

/* TF/lite/kernels/internal/reference/conv.h:25-99 */
void or_conv2d(const float* in, int ih, int iw, int ic,
               const float* w, int oc, int kh, int kw, const float* bias,
               int stride_h, int stride_w, int dil_h, int dil_w, int padding, int act,
               float* out, int oh, int ow) {
  int o_h, o_w, pad_h, pad_w;
  or_conv_out_size(ih, kh, stride_h, dil_h, padding, &o_h, &pad_h);
  or_conv_out_size(iw, kw, stride_w, dil_w, padding, &o_w, &pad_w);
  (void)o_h; (void)o_w;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int oy = 0; oy < oh; ++oy) {
    const int iy0 = oy * stride_h - pad_h;
    for (int ox = 0; ox < ow; ++ox) {
      const int ix0 = ox * stride_w - pad_w;
      for (int o = 0; o < oc; ++o) {
        float total = 0.f;
        for (int fy = 0; fy < kh; ++fy) {
          const int iy = iy0 + dil_h * fy;
          for (int fx = 0; fx < kw; ++fx) {
            const int ix = ix0 + dil_w * fx;
            if (ix < 0 || ix >= iw || iy < 0 || iy >= ih) continue;
            const float* ip = in + ((size_t)iy * iw + ix) * ic;
            const float* wp = w + (((size_t)o * kh + fy) * kw + fx) * ic;
            for (int c = 0; c < ic; ++c) total = fmaf(ip[c], wp[c], total);
          }
        }
        float b = bias ? bias[o] : 0.f;
        out[((size_t)oy * ow + ox) * oc + o] = act_clamp(total + b, act);
      }
    }
  }
}

/* TF/lite/kernels/internal/reference/depthwiseconv_float.h:25-96 */
void or_depthwise_conv2d(const float* in, int ih, int iw, int ic,
                         const float* w, int kh, int kw, const float* bias,
                         int stride_h, int stride_w, int dil_h, int dil_w, int padding,
                         int depth_mult, int act, float* out, int oh, int ow) {
  int o_h, o_w, pad_h, pad_w;
  or_conv_out_size(ih, kh, stride_h, dil_h, padding, &o_h, &pad_h);
  or_conv_out_size(iw, kw, stride_w, dil_w, padding, &o_w, &pad_w);
  (void)o_h; (void)o_w;
  const int odepth = ic * depth_mult;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int oy = 0; oy < oh; ++oy) {
    for (int ox = 0; ox < ow; ++ox) {
      for (int c = 0; c < ic; ++c) {
        for (int m = 0; m < depth_mult; ++m) {
          const int o = m + c * depth_mult;
          const int ix0 = ox * stride_w - pad_w;
          const int iy0 = oy * stride_h - pad_h;
          float total = 0.f;
          for (int fy = 0; fy < kh; ++fy) {
            for (int fx = 0; fx < kw; ++fx) {
              const int ix = ix0 + dil_w * fx;
              const int iy = iy0 + dil_h * fy;
              if (ix >= 0 && ix < iw && iy >= 0 && iy < ih) {
                total = fmaf(in[((size_t)iy * iw + ix) * ic + c],
                             w[((size_t)fy * kw + fx) * odepth + o], total);
              }
            }
          }
          float b = bias ? bias[o] : 0.f;
          out[((size_t)oy * ow + ox) * odepth + o] = act_clamp(total + b, act);
        }
      }
    }
  }
}

/* TF/lite/kernels/internal/reference/pooling.h:26-79.  Summation order deviation
 * (oracle.h): each window row is summed left-to-right, then the row sums are
 * summed top-to-bottom. */
void or_average_pool(const float* in, int ih, int iw, int c, int fh, int fw,
                     int stride_h, int stride_w, int padding, int act,
                     float* out, int oh, int ow) {
  int o_h, o_w, pad_h, pad_w;
  or_conv_out_size(ih, fh, stride_h, 1, padding, &o_h, &pad_h);
  or_conv_out_size(iw, fw, stride_w, 1, padding, &o_w, &pad_w);
  (void)o_h; (void)o_w;
  for (int oy = 0; oy < oh; ++oy) {
    for (int ox = 0; ox < ow; ++ox) {
      const int ix0 = ox * stride_w - pad_w, iy0 = oy * stride_h - pad_h;
      const int fx0 = ix0 < 0 ? -ix0 : 0, fx1 = (fw < iw - ix0) ? fw : iw - ix0;
      const int fy0 = iy0 < 0 ? -iy0 : 0, fy1 = (fh < ih - iy0) ? fh : ih - iy0;
      for (int ch = 0; ch < c; ++ch) {
        float total = 0.f;
        float count = 0.f;
        for (int fy = fy0; fy < fy1; ++fy) {
          float row = 0.f;
          for (int fx = fx0; fx < fx1; ++fx) {
            row += in[((size_t)(iy0 + fy) * iw + (ix0 + fx)) * c + ch];
            count += 1.f;
          }
          total += row;
        }
        float avg = total / count;
        out[((size_t)oy * ow + ox) * c + ch] = act_clamp(avg, act);
      }
    }
  }
}

/* TF/lite/kernels/internal/reference/fully_connected.h:27-61 */
void or_fully_connected(const float* in, int batches, int in_depth, const float* w,
                        int out_depth, const float* bias, int act, float* out) {
  for (int b = 0; b < batches; ++b) {
    for (int o = 0; o < out_depth; ++o) {
      float total = 0.f;
      for (int d = 0; d < in_depth; ++d)
        total = fmaf(in[(size_t)b * in_depth + d], w[(size_t)o * in_depth + d], total);
      float bv = bias ? bias[o] : 0.f;
      out[(size_t)b * out_depth + o] = act_clamp(total + bv, act);
    }
  }
}

/* TF/lite/kernels/internal/reference/resize_bilinear.h:29-117 (float path) */
static void interp_values(float value, float scale, int half_pixel, int in_size,
                          float* scaled, int* lo, int* hi) {
  if (half_pixel) *scaled = (value + 0.5f) * scale - 0.5f;
  else *scaled = value * scale;
  float fl = floorf(*scaled);
  int l = (int)fl; if (l < 0) l = 0;
  int h = (int)ceilf(*scaled); if (h > in_size - 1) h = in_size - 1;
  *lo = l; *hi = h;
}

void or_resize_bilinear(const float* in, int ih, int iw, int c, float* out, int oh, int ow,
                        int align_corners, int half_pixel) {
  float hs = (float)ih / oh, ws = (float)iw / ow;
  if (align_corners && oh > 1) hs = (float)(ih - 1) / (oh - 1);
  if (align_corners && ow > 1) ws = (float)(iw - 1) / (ow - 1);
  for (int y = 0; y < oh; ++y) {
    float fy; int y0, y1;
    interp_values((float)y, hs, half_pixel, ih, &fy, &y0, &y1);
    for (int x = 0; x < ow; ++x) {
      float fx; int x0, x1;
      interp_values((float)x, ws, half_pixel, iw, &fx, &x0, &x1);
      const float dy = fy - y0, dx = fx - x0;
      const float wy0 = 1 - dy, wx0 = 1 - dx;
      for (int ch = 0; ch < c; ++ch) {
        float a = in[((size_t)y0 * iw + x0) * c + ch] * wy0 * wx0;
        float b = in[((size_t)y1 * iw + x0) * c + ch] * dy * wx0;
        float d = in[((size_t)y0 * iw + x1) * c + ch] * wy0 * dx;
        float e = in[((size_t)y1 * iw + x1) * c + ch] * dy * dx;
        out[((size_t)y * ow + x) * c + ch] = ((a + b) + d) + e;
      }
    }
  }
}

/* TF/lite/kernels/internal/reference/hard_swish.h:45-56 */
void or_hard_swish(const float* in, float* out, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    float x = in[i];
    float t = x + 3.f;
    t = t < 0.f ? 0.f : t;
    t = t > 6.f ? 6.f : t;
    out[i] = (x * t) / 6.f;
  }
}

/* TF/lite/kernels/internal/reference/logistic.h:30-57 */
void or_logistic(const float* in, float* out, size_t n) {
  const float cutoff_upper = 16.619047164916992188f;
  const float cutoff_lower = -9.f;
  for (size_t i = 0; i < n; ++i) {
    float v = in[i], r;
    if (v > cutoff_upper) r = 1.0f;
    else if (v < cutoff_lower) r = or_expf(v);
    else r = 1.f / (1.f + or_expf(-v));
    out[i] = r;
  }
}

/* RELU / RELU6 as stand-alone ops (TF/lite/kernels/activations.cc; same clamps) */
void or_relu(const float* in, float* out, size_t n, int act) {
  for (size_t i = 0; i < n; ++i) out[i] = act_clamp(in[i], act);
}

/* TF/lite/kernels/internal/reference/add.h:28 */
void or_add(const float* a, const float* b, float* out, size_t n, int act) {
  for (size_t i = 0; i < n; ++i) out[i] = act_clamp(a[i] + b[i], act);
}

/* TF/lite/kernels/internal/reference/mul.h:45-110 (same-shape and channel broadcast) */
void or_mul(const float* a, const float* b, float* out, size_t n_outer, int c, int bcast, int act) {
  for (size_t i = 0; i < n_outer; ++i)
    for (int ch = 0; ch < c; ++ch) {
      float bv = bcast ? b[ch] : b[i * c + ch];
      out[i * c + ch] = act_clamp(a[i * c + ch] * bv, act);
    }
}

/* lib/transpose_conv_bias.cc:37-114 (scatter form) restated as the equivalent
 * gather: for every output element the contributions arrive in the reference's
 * (in_y, in_x, in_channel, filter_y, filter_x) order; padding per :203-228. */
void or_tconv_bias(const float* in, int ih, int iw, int ic, const float* w, int oc, int kh, int kw,
                   const float* bias, int stride_h, int stride_w, int padding_same,
                   float* out, int oh, int ow) {
  int pad_h_total = 0, pad_w_total = 0;
  if (padding_same) {
    pad_h_total = kh - (ih - 1) % stride_h - 1; if (pad_h_total < 0) pad_h_total = 0;
    pad_w_total = kw - (iw - 1) % stride_w - 1; if (pad_w_total < 0) pad_w_total = 0;
  }
  const int pad_h = pad_h_total / 2, pad_w = pad_w_total / 2;
  for (size_t i = 0; i < (size_t)oh * ow; ++i)
    for (int o = 0; o < oc; ++o) out[i * oc + o] = bias[o];
  for (int iy = 0; iy < ih; ++iy)
    for (int ix = 0; ix < iw; ++ix)
      for (int c = 0; c < ic; ++c) {
        const int ox0 = ix * stride_w - pad_w, oy0 = iy * stride_h - pad_h;
        const float v = in[((size_t)iy * iw + ix) * ic + c];
        for (int fy = 0; fy < kh; ++fy)
          for (int fx = 0; fx < kw; ++fx)
            for (int o = 0; o < oc; ++o) {
              const int ox = ox0 + fx, oy = oy0 + fy;
              if (ox >= 0 && ox < ow && oy >= 0 && oy < oh) {
                float* d = &out[((size_t)oy * ow + ox) * oc + o];
                *d = fmaf(v, w[(((size_t)o * kh + fy) * kw + fx) * ic + c], *d);
              }
            }
      }
}
