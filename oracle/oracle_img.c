/* oracle/oracle_img.c — scalar restatement of the OpenCV 8-bit primitives on the
 * backscrub hot path.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * OpenCV is a system dependency of the reference (CMakeLists.txt:31) and its
 * source is not under /root/reference, so these follow OpenCV's published
 * fixed-point algorithms and are pinned bit-for-bit against the in-container
 * cv2 4.13.0 by tests/test_oracle_img.py (call sites: lib/libbackscrub.cc:
 * 285-302,366-371; app/deepseg.cc:87-134; app/background.cc:178-194).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }
/* BORDER_REFLECT_101: -1 -> 1, -2 -> 2, n -> n-2, n+1 -> n-3 (single reflection; n >= 3 here) */
static inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; }
  return p;
}

/* cv::resize INTER_LINEAR, CV_8U: 11-bit fixed-point coefficients (INTER_RESIZE_COEF_BITS),
 * horizontal pass into int32 (scale 2^11), vertical pass with the >>4, >>16, +2, >>2
 * descale; when both scale factors are exactly 2 OpenCV switches to INTER_AREA
 * (2x2 mean, (sum+2)>>2). */
void or_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride,
                         uint8_t* dst, int dw, int dh, size_t dstride, int cn) {
  if (sw == dw * 2 && sh == dh * 2) {
    for (int y = 0; y < dh; ++y)
      for (int x = 0; x < dw; ++x)
        for (int c = 0; c < cn; ++c) {
          const uint8_t* s0 = src + (size_t)(2 * y) * sstride + (size_t)(2 * x) * cn + c;
          const uint8_t* s1 = s0 + sstride;
          dst[(size_t)y * dstride + (size_t)x * cn + c] = (uint8_t)((s0[0] + s0[cn] + s1[0] + s1[cn] + 2) >> 2);
        }
    return;
  }
  const double inv_scale_x = (double)dw / sw, inv_scale_y = (double)dh / sh;
  const double scale_x = 1.0 / inv_scale_x, scale_y = 1.0 / inv_scale_y;
  int* xofs = (int*)malloc(sizeof(int) * (size_t)dw);
  short* ialpha = (short*)malloc(sizeof(short) * 2 * (size_t)dw);
  for (int dx = 0; dx < dw; ++dx) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx -= (float)sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    xofs[dx] = sx;
    ialpha[2 * dx] = (short)lrintf((1.f - fx) * 2048.f);
    ialpha[2 * dx + 1] = (short)lrintf(fx * 2048.f);
  }
  int* row0 = (int*)malloc(sizeof(int) * (size_t)dw * cn);
  int* row1 = (int*)malloc(sizeof(int) * (size_t)dw * cn);
  for (int dy = 0; dy < dh; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)floorf(fy);
    fy -= (float)sy;
    const short b0 = (short)lrintf((1.f - fy) * 2048.f), b1 = (short)lrintf(fy * 2048.f);
    const int r0 = clampi(sy, 0, sh - 1), r1 = clampi(sy + 1, 0, sh - 1);
    const uint8_t* s0 = src + (size_t)r0 * sstride;
    const uint8_t* s1 = src + (size_t)r1 * sstride;
    for (int dx = 0; dx < dw; ++dx) {
      const int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
      const int a0 = ialpha[2 * dx], a1 = ialpha[2 * dx + 1];
      for (int c = 0; c < cn; ++c) {
        row0[dx * cn + c] = s0[sx * cn + c] * a0 + s0[sx1 * cn + c] * a1;
        row1[dx * cn + c] = s1[sx * cn + c] * a0 + s1[sx1 * cn + c] * a1;
      }
    }
    uint8_t* d = dst + (size_t)dy * dstride;
    for (int i = 0; i < dw * cn; ++i)
      d[i] = sat_u8((((b0 * (row0[i] >> 4)) >> 16) + ((b1 * (row1[i] >> 4)) >> 16) + 2) >> 2);
  }
  free(xofs); free(ialpha); free(row0); free(row1);
}

/* cv::bilateralFilter(8UC3, d = 5): radius 2, taps with sqrt(i^2+j^2) <= 2 in (i, j)
 * raster order, float LUT weights, BORDER_REFLECT_101, result cvRound(sum * (1/wsum)).
 * Per-tap accumulation is `sum = fmaf(v, w, sum)` in tap order (matches OpenCV's
 * v_muladd SIMD body on FMA hardware; see tests for the measured agreement). */
void or_bilateral_d5_u8c3(const uint8_t* src, uint8_t* dst, int w, int h, double sigma_color, double sigma_space) {
  const int radius = 2;
  const double gc = -0.5 / (sigma_color * sigma_color), gs = -0.5 / (sigma_space * sigma_space);
  float* color_w = (float*)malloc(sizeof(float) * 256 * 3);
  for (int i = 0; i < 256 * 3; ++i) color_w[i] = (float)exp((double)i * i * gc);
  float space_w[25]; int oi[25], oj[25]; int maxk = 0;
  for (int i = -radius; i <= radius; ++i)
    for (int j = -radius; j <= radius; ++j) {
      double r = sqrt((double)i * i + (double)j * j);
      if (r > radius) continue;
      space_w[maxk] = (float)exp(r * r * gs);
      oi[maxk] = i; oj[maxk] = j; ++maxk;
    }
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const uint8_t* p0 = src + ((size_t)y * w + x) * 3;
      const int b0 = p0[0], g0 = p0[1], r0 = p0[2];
      float wsum = 0.f, sb = 0.f, sg = 0.f, sr = 0.f;
      for (int k = 0; k < maxk; ++k) {
        const int yy = reflect101(y + oi[k], h), xx = reflect101(x + oj[k], w);
        const uint8_t* p = src + ((size_t)yy * w + xx) * 3;
        const int b = p[0], g = p[1], r = p[2];
        const float wk = space_w[k] * color_w[abs(b - b0) + abs(g - g0) + abs(r - r0)];
        wsum += wk;
        sb = fmaf((float)b, wk, sb);
        sg = fmaf((float)g, wk, sg);
        sr = fmaf((float)r, wk, sr);
      }
      const float inv = 1.f / wsum;
      uint8_t* d = dst + ((size_t)y * w + x) * 3;
      d[0] = sat_u8((int)lrintf(sb * inv));
      d[1] = sat_u8((int)lrintf(sg * inv));
      d[2] = sat_u8((int)lrintf(sr * inv));
    }
  free(color_w);
}

/* Mat::convertTo(CV_32F, alpha, beta) from CV_8U: (float)v * (float)alpha + (float)beta as
 * one fused multiply-add (OpenCV's v_fma body on FMA hardware). */
void or_convert_u8_f32(const uint8_t* src, float* dst, size_t n, float alpha, float beta) {
  for (size_t i = 0; i < n; ++i) dst[i] = fmaf((float)src[i], alpha, beta);
}

/* cv::blur 5x5 normalised box, CV_8UC1, BORDER_REFLECT_101: integer window sum S,
 * result cvRound(S * (1.0/25)) == (S + 12) / 25 (no exact .5 ties since 25 is odd). */
void or_box_blur5_u8(const uint8_t* src, size_t sstride, uint8_t* dst, size_t dstride, int w, int h) {
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int s = 0;
      for (int dy = -2; dy <= 2; ++dy) {
        const uint8_t* row = src + (size_t)reflect101(y + dy, h) * sstride;
        for (int dx = -2; dx <= 2; ++dx) s += row[reflect101(x + dx, w)];
      }
      dst[(size_t)y * dstride + x] = (uint8_t)lrint((double)s * (1.0 / 25));
    }
}

/* cv::cvtColor(COLOR_RGB2YUV), CV_8U: 14-bit fixed point (yuv_shift), channel 0 = "R". */
void or_rgb2yuv_u8(const uint8_t* src, uint8_t* dst, size_t npix) {
  for (size_t i = 0; i < npix; ++i) {
    const int R = src[3 * i], G = src[3 * i + 1], B = src[3 * i + 2];
    const int Y = (4899 * R + 9617 * G + 1868 * B + 8192) >> 14;
    const int U = ((B - Y) * 8061 + (128 << 14) + 8192) >> 14;
    const int V = ((R - Y) * 14369 + (128 << 14) + 8192) >> 14;
    dst[3 * i] = sat_u8(Y); dst[3 * i + 1] = sat_u8(U); dst[3 * i + 2] = sat_u8(V);
  }
}

/* app/deepseg.cc:87-106 convert_rgb_to_yuyv: RGB2YUV, then per pixel pair of the
 * flattened image [Y0, (V0+V1)/2, Y1, (U0+U1)/2] (V before U, truncating average). */
void or_convert_rgb_to_yuyv(const uint8_t* src, uint8_t* dst_yuyv, int w, int h) {
  const size_t npix = (size_t)w * h;
  uint8_t* yuv = (uint8_t*)malloc(npix * 3);
  or_rgb2yuv_u8(src, yuv, npix);
  for (size_t i = 0; i + 1 < npix; i += 2) {
    const int u = ((int)yuv[3 * i + 1] + (int)yuv[3 * (i + 1) + 1]) / 2;
    const int v = ((int)yuv[3 * i + 2] + (int)yuv[3 * (i + 1) + 2]) / 2;
    dst_yuyv[2 * i + 0] = yuv[3 * i];
    dst_yuyv[2 * i + 1] = (uint8_t)v;
    dst_yuyv[2 * i + 2] = yuv[3 * (i + 1)];
    dst_yuyv[2 * i + 3] = (uint8_t)u;
  }
  free(yuv);
}

/* app/deepseg.cc:108-134 alpha_blend: out = (a*m + b*(255-m)) / 255, C int division. */
void or_alpha_blend(const uint8_t* srca, const uint8_t* srcb, const uint8_t* mask, uint8_t* out, size_t npix) {
  for (size_t p = 0; p < npix; ++p) {
    const int aw = mask[p], bw = 255 - aw;
    for (int c = 0; c < 3; ++c)
      out[3 * p + c] = (uint8_t)(((int)srca[3 * p + c] * aw + (int)srcb[3 * p + c] * bw) / 255);
  }
}

/* cv::cvtColor(COLOR_YUV2BGR_YUYV), CV_8UC2 -> CV_8UC3: ITU-R BT.601 limited range, 20-bit fixed point.
 * This is the conversion OpenCV's V4L2 backend applies to YUYV camera frames when CAP_PROP_CONVERT_RGB is
 * set (app/deepseg.cc:553) and that app/deepseg.cc:725 applies to the output frame; pinned bit-exact
 * against cv2 in tests/test_oracle_img.py. */
void or_yuyv_to_bgr(const uint8_t* yuyv, uint8_t* bgr, int w, int h) {
  const int CY = 1220542, CUB = 2116026, CUG = -409993, CVG = -852492, CVR = 1673527, SH = 20;
  for (size_t i = 0; i + 1 < (size_t)w * h; i += 2) {
    const int y0 = yuyv[2 * i], u = (int)yuyv[2 * i + 1] - 128, y1 = yuyv[2 * i + 2], v = (int)yuyv[2 * i + 3] - 128;
    const int ruv = (1 << (SH - 1)) + CVR * v, guv = (1 << (SH - 1)) + CVG * v + CUG * u, buv = (1 << (SH - 1)) + CUB * u;
    const int ys[2] = {y0, y1};
    for (int k = 0; k < 2; ++k) {
      const int yy = (ys[k] - 16 < 0 ? 0 : ys[k] - 16) * CY;
      bgr[3 * (i + k) + 0] = sat_u8((yy + buv) >> SH);
      bgr[3 * (i + k) + 1] = sat_u8((yy + guv) >> SH);
      bgr[3 * (i + k) + 2] = sat_u8((yy + ruv) >> SH);
    }
  }
}


/* ---------------------------------------------------------------------------
 * cv::GaussianBlur(src, dst, Size(k,k), 0) for CV_8UC3, BORDER_DEFAULT (app/deepseg.cc:657-658, `-p bgblur:k`).
 *
 * OpenCV is a system dependency of the reference (not under /root/reference); this restates the
 * published bit-exact 8-bit path of OpenCV 4.x (modules/imgproc/src/smooth.dispatch.cpp:
 * getGaussianKernelBitExact + getGaussianKernelFixedPoint_ED + GaussianBlurFixedPoint) and is pinned
 * bit-exact against the in-container cv2 4.13 for k = 1..255 kernels and k in {3,5,7,9,11,25,51} images:
 *   sigma = 0.15 k + 0.35;   g[i] = exp(-(i-(k-1)/2)^2 / (2 sigma^2)) / sum   (double)
 *   q[i]  = round-half-even(g[i]*256 + carried rounding error) for the first half (mirrored),
 *   q[centre] = 256 - 2*sum(first half)        => sum q == 256 exactly (8.8 fixed point)
 *   rows:  h = sum_j q[j]*src[x+j]   (16 bit, exact)      cols: v = sum_j q[j]*h[y+j] (32 bit, exact)
 *   dst = (v + 32768) >> 16
 * k <= 9 with sigma 0 use OpenCV's tabulated small kernels (1; 1 2 1; 1 4 6 4 1; ...), which the formula
 * above does not produce.
 * --------------------------------------------------------------------------- */
int or_gaussian_kernel_q8(int k, int* q) {
  if (k < 1 || !(k & 1) || k > 255) return -1;
  static const double small[5][9] = {
      {1.0}, {0.25, 0.5, 0.25}, {0.0625, 0.25, 0.375, 0.25, 0.0625},
      {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125},
      {4.0 / 256, 13.0 / 256, 30.0 / 256, 51.0 / 256, 60.0 / 256, 51.0 / 256, 30.0 / 256, 13.0 / 256, 4.0 / 256}};
  double g[255];
  const int n2 = k / 2;
  if (k <= 9) {
    for (int i = 0; i < k; ++i) g[i] = small[n2][i];
  } else {
    const double sigma = fma((double)k, 0.15, 0.35);
    const double scale2x = -0.125 / (sigma * sigma);
    double sum = 0.0;
    for (int i = 0, x = 1 - k; i < n2; ++i, x += 2) { g[i] = exp((double)(x * x) * scale2x); sum += g[i]; }
    sum = sum * 2.0 + 1.0;
    const double mul = 1.0 / sum;
    for (int i = 0; i < n2; ++i) g[i] *= mul;
    g[n2] = mul;
  }
  double err = 0.0;
  int acc = 0;
  for (int i = 0; i < n2; ++i) {
    const double adj = g[i] * 256.0 + err;
    const int v = (int)nearbyint(adj);
    err = adj - (double)v;
    q[i] = q[k - 1 - i] = v;
    acc += v;
  }
  q[n2] = 256 - 2 * acc;
  return 0;
}

int or_gaussian_blur_u8c3(const uint8_t* src, int w, int h, size_t sstride, uint8_t* dst, size_t dstride, int k) {
  int q[255];
  if (or_gaussian_kernel_q8(k, q)) return -1;
  const int r = k / 2;
  uint16_t* tmp = (uint16_t*)malloc((size_t)w * h * 3 * sizeof(uint16_t));
  if (!tmp) return -1;
  for (int y = 0; y < h; ++y) {
    const uint8_t* s = src + (size_t)y * sstride;
    for (int x = 0; x < w; ++x)
      for (int c = 0; c < 3; ++c) {
        uint32_t a = 0;
        for (int j = 0; j < k; ++j) a += (uint32_t)q[j] * s[(size_t)reflect101(x + j - r, w) * 3 + c];
        tmp[((size_t)y * w + x) * 3 + c] = (uint16_t)a;
      }
  }
  for (int y = 0; y < h; ++y)
    for (size_t i = 0; i < (size_t)w * 3; ++i) {
      uint32_t a = 0;
      for (int j = 0; j < k; ++j) a += (uint32_t)q[j] * tmp[(size_t)reflect101(y + j - r, h) * w * 3 + i];
      dst[(size_t)y * dstride + i] = (uint8_t)((a + 32768u) >> 16);
    }
  free(tmp);
  return 0;
}

/* cv::flip(src, dst, code): code 1 = horizontal (around the y axis), 0 = vertical, -1 = both (app/deepseg.cc:667-673) */
void or_flip_u8c3(const uint8_t* src, uint8_t* dst, int w, int h, int flip_h, int flip_v) {
  for (int y = 0; y < h; ++y) {
    const uint8_t* s = src + (size_t)(flip_v ? h - 1 - y : y) * w * 3;
    uint8_t* d = dst + (size_t)y * w * 3;
    for (int x = 0; x < w; ++x) memcpy(d + (size_t)x * 3, s + (size_t)(flip_h ? w - 1 - x : x) * 3, 3);
  }
}
