/* oracle/oracle_pipeline.c — literal, deterministic restatement of the per-frame
 * pipeline: lib/libbackscrub.cc:161-259 (context/geometry), :279-376 (process),
 * app/background.cc:178-194 (grab_background resize), app/deepseg.cc:108-134
 * (alpha_blend), :87-106 (convert_rgb_to_yuyv).  TEST INFRASTRUCTURE ONLY.
 *
 * Deliberate, documented differences from the reference:
 *   - the IIR state `ofinal` starts at zero (the reference leaves it
 *     uninitialised, lib/libbackscrub.cc:257);
 *   - the mask is synchronous (frame t is blended with mask t; the reference's
 *     worker thread may hand back a stale mask, app/deepseg.cc:182-216);
 *   - body-pix: the reference slices the 33x33 `ofinal` with `in_roidim` given in
 *     257x257 input coordinates and throws (lib/libbackscrub.cc:241,257,368); the
 *     oracle scales that rectangle by out/in (integer floor).
 */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct or_maskgen {
  or_model* model;
  int type;
  float scaling, offset;
  int W, H;
  int mh, mw, mc, oh, ow, oc;
  int roidim[4], in_roidim[4], out_roidim[4];
  uint8_t* in_u8_bgr;   /* mh*mw*3, zero outside in_roidim */
  uint8_t* in_u8_rgb;
  uint8_t* filtered;
  float* input;
  uint8_t* ofinal;      /* oh*ow IIR state */
  uint8_t* mask;        /* H*W, 255 outside roidim */
  uint8_t* tmp_up;      /* roi_h*roi_w */
  uint8_t* scratch[3];  /* W*H*3 each: frame copy / background / work image of or_composite[_ex] (the reference allocates
                           cv::Mats per frame, app/deepseg.cc:110,649; kept here so the timed CPU arm scales with threads) */
};

static uint8_t* scratch_buf(or_maskgen* g, int i) {
  if (!g->scratch[i]) g->scratch[i] = (uint8_t*)malloc((size_t)g->W * g->H * 3);
  return g->scratch[i];
}

/* lib/libbackscrub.cc:116-130 */
int or_model_type_from_name(const char* path) {
  if (strstr(path, "body-pix")) return OR_MODEL_BODYPIX;
  if (strstr(path, "deeplab")) return OR_MODEL_DEEPLAB;
  if (strstr(path, "segm_")) return OR_MODEL_MEET;
  if (strstr(path, "selfie")) return OR_MODEL_MLKIT;
  return OR_MODEL_UNKNOWN;
}

void or_maskgen_delete(or_maskgen* g) {
  if (!g) return;
  or_model_free(g->model);
  free(g->in_u8_bgr); free(g->in_u8_rgb); free(g->filtered); free(g->input);
  free(g->ofinal); free(g->mask); free(g->tmp_up);
  free(g->scratch[0]); free(g->scratch[1]); free(g->scratch[2]);
  free(g);
}

or_maskgen* or_maskgen_new(const char* model_path, int width, int height, char* err, size_t errlen) {
  or_maskgen* g = (or_maskgen*)calloc(1, sizeof(*g));
  g->model = or_model_load(model_path, err, errlen);
  if (!g->model) { free(g); return NULL; }
  g->type = or_model_type_from_name(model_path);
  if (g->type == OR_MODEL_UNKNOWN) {
    if (err) snprintf(err, errlen, "unknown model type '%s'", model_path);
    or_maskgen_delete(g); return NULL;
  }
  /* lib/libbackscrub.cc:132-148 */
  if (g->type == OR_MODEL_DEEPLAB) { g->scaling = (float)(1 / 127.5); g->offset = -1.f; }
  else { g->scaling = (float)(1 / 255.0); g->offset = 0.f; }
  int s[4];
  or_model_tensor_shape(g->model, or_model_input(g->model), s);
  g->mh = s[1]; g->mw = s[2]; g->mc = s[3];
  or_model_tensor_shape(g->model, or_model_output(g->model), s);
  g->oh = s[1]; g->ow = s[2]; g->oc = s[3];
  g->W = width; g->H = height;
  /* lib/libbackscrub.cc:234-246 (float arithmetic, truncating conversion to int) */
  const float ratio = (float)g->mh / (float)g->mw;
  const float frameratio = (float)height / (float)width;
  if (frameratio < ratio) {
    g->roidim[0] = (int)(((float)width - (float)height / ratio) / 2);
    g->roidim[1] = 0;
    g->roidim[2] = (int)((float)height / ratio);
    g->roidim[3] = height;
    g->in_roidim[0] = 0; g->in_roidim[1] = 0; g->in_roidim[2] = g->mw; g->in_roidim[3] = g->mh;
  } else {
    g->roidim[0] = 0; g->roidim[1] = 0; g->roidim[2] = width; g->roidim[3] = height;
    g->in_roidim[0] = (int)(((float)g->mw - (float)g->mh / frameratio) / 2);
    g->in_roidim[1] = 0;
    g->in_roidim[2] = (int)((float)g->mh / frameratio);
    g->in_roidim[3] = g->mh;
  }
  if (g->oh == g->mh && g->ow == g->mw) memcpy(g->out_roidim, g->in_roidim, sizeof(g->in_roidim));
  else {
    g->out_roidim[0] = g->in_roidim[0] * g->ow / g->mw;
    g->out_roidim[1] = g->in_roidim[1] * g->oh / g->mh;
    g->out_roidim[2] = g->in_roidim[2] * g->ow / g->mw;
    g->out_roidim[3] = g->in_roidim[3] * g->oh / g->mh;
  }
  const size_t in_px = (size_t)g->mh * g->mw;
  g->in_u8_bgr = (uint8_t*)calloc(in_px * 3, 1);
  g->in_u8_rgb = (uint8_t*)calloc(in_px * 3, 1);
  g->filtered = (uint8_t*)calloc(in_px * 3, 1);
  g->input = (float*)calloc(in_px * 3, sizeof(float));
  g->ofinal = (uint8_t*)calloc((size_t)g->oh * g->ow, 1);
  g->mask = (uint8_t*)malloc((size_t)width * height);
  memset(g->mask, 255, (size_t)width * height);
  g->tmp_up = (uint8_t*)malloc((size_t)g->roidim[2] * g->roidim[3]);
  return g;
}

void or_maskgen_geometry(const or_maskgen* g, int roidim[4], int in_roidim[4], int out_roidim[4],
                         int in_hwc[3], int out_hwc[3]) {
  memcpy(roidim, g->roidim, 16); memcpy(in_roidim, g->in_roidim, 16); memcpy(out_roidim, g->out_roidim, 16);
  in_hwc[0] = g->mh; in_hwc[1] = g->mw; in_hwc[2] = g->mc;
  out_hwc[0] = g->oh; out_hwc[1] = g->ow; out_hwc[2] = g->oc;
}

/* lib/libbackscrub.cc:314-361 */
static void decide_and_smooth(or_maskgen* g, const float* tmp) {
  uint8_t* out = g->ofinal;
  const size_t total = (size_t)g->oh * g->ow;
  switch (g->type) {
    case OR_MODEL_DEEPLAB: {
      const int cnum = 21, pers = 15;
      for (size_t n = 0; n < total; ++n) {
        float maxval = -10000; int maxpos = 0;
        for (int i = 0; i < cnum; ++i)
          if (tmp[n * cnum + i] > maxval) { maxval = tmp[n * cnum + i]; maxpos = i; }
        uint8_t val = (maxpos == pers ? 0 : 255);
        out[n] = (uint8_t)((val & 0xE0) | (out[n] >> 3));
      }
      break;
    }
    case OR_MODEL_BODYPIX:
    case OR_MODEL_MLKIT:
      for (size_t n = 0; n < total; ++n) {
        uint8_t val = ((double)tmp[n] > 0.65 ? 0 : 255);
        out[n] = (uint8_t)((val & 0xE0) | (out[n] >> 3));
      }
      break;
    case OR_MODEL_MEET:
      for (size_t n = 0; n < total; ++n) {
        float exp0 = or_expf(tmp[2 * n]);
        float exp1 = or_expf(tmp[2 * n + 1]);
        float p0 = exp0 / (exp0 + exp1);
        float p1 = exp1 / (exp0 + exp1);
        uint8_t val = (p0 < p1 ? 0 : 255);
        out[n] = (uint8_t)((val & 0xE0) | (out[n] >> 3));
      }
      break;
    default: break;
  }
}

/* lib/libbackscrub.cc:366-374 */
static void upsample_and_blur(or_maskgen* g) {
  const int rw = g->roidim[2], rh = g->roidim[3];
  const uint8_t* sub = g->ofinal + (size_t)g->out_roidim[1] * g->ow + g->out_roidim[0];
  or_resize_linear_u8(sub, g->out_roidim[2], g->out_roidim[3], (size_t)g->ow, g->tmp_up, rw, rh, (size_t)rw, 1);
  uint8_t* mroi = g->mask + (size_t)g->roidim[1] * g->W + g->roidim[0];
  or_box_blur5_u8(g->tmp_up, (size_t)rw, mroi, (size_t)g->W, rw, rh);
}

int or_maskgen_post_from_output(or_maskgen* g, const float* model_output, uint8_t* mask_out) {
  decide_and_smooth(g, model_output);
  upsample_and_blur(g);
  if (mask_out) memcpy(mask_out, g->mask, (size_t)g->W * g->H);
  return 0;
}

int or_maskgen_process(or_maskgen* g, const uint8_t* frame_bgr, size_t stride, uint8_t* mask_out) {
  const size_t in_px = (size_t)g->mh * g->mw;
  /* :285-290 resize ROI into the (zero padded) model-sized BGR image, swap to RGB */
  const uint8_t* roi = frame_bgr + (size_t)g->roidim[1] * stride + (size_t)g->roidim[0] * 3;
  uint8_t* in_roi = g->in_u8_bgr + ((size_t)g->in_roidim[1] * g->mw + g->in_roidim[0]) * 3;
  or_resize_linear_u8(roi, g->roidim[2], g->roidim[3], stride, in_roi, g->in_roidim[2], g->in_roidim[3],
                      (size_t)g->mw * 3, 3);
  for (size_t p = 0; p < in_px; ++p) {
    g->in_u8_rgb[3 * p] = g->in_u8_bgr[3 * p + 2];
    g->in_u8_rgb[3 * p + 1] = g->in_u8_bgr[3 * p + 1];
    g->in_u8_rgb[3 * p + 2] = g->in_u8_bgr[3 * p];
  }
  /* :295-299 */
  or_bilateral_d5_u8c3(g->in_u8_rgb, g->filtered, g->mw, g->mh, 100.0, 100.0);
  /* :302 */
  or_convert_u8_f32(g->filtered, g->input, in_px * 3, g->scaling, g->offset);
  /* :307 */
  int rc = or_model_invoke(g->model, g->input);
  if (rc) return rc;
  return or_maskgen_post_from_output(g, or_model_tensor_data(g->model, or_model_output(g->model)), mask_out);
}

const uint8_t* or_maskgen_in_u8(const or_maskgen* g) { return g->in_u8_rgb; }
const uint8_t* or_maskgen_filtered_u8(const or_maskgen* g) { return g->filtered; }
const float* or_maskgen_input_f32(const or_maskgen* g) { return g->input; }
const float* or_maskgen_output_f32(const or_maskgen* g) { return or_model_tensor_data(g->model, or_model_output(g->model)); }
const uint8_t* or_maskgen_ofinal(const or_maskgen* g) { return g->ofinal; }
or_model* or_maskgen_model(or_maskgen* g) { return g->model; }

int or_composite(or_maskgen* g, const uint8_t* frame_bgr, size_t stride,
                 const uint8_t* bg_raw, int bw, int bh, size_t bstride,
                 uint8_t* out_rgb, uint8_t* out_yuyv, uint8_t* out_mask) {
  const size_t npix = (size_t)g->W * g->H;
  int rc = or_maskgen_process(g, frame_bgr, stride, out_mask);
  if (rc) return rc;
  /* app/background.cc:178-194: resize the raw background to the frame size, every frame */
  uint8_t* bg = scratch_buf(g, 1);
  or_resize_linear_u8(bg_raw, bw, bh, bstride, bg, g->W, g->H, (size_t)g->W * 3, 3);
  /* the camera frame as a packed buffer (alpha_blend walks data pointers linearly) */
  uint8_t* fr = scratch_buf(g, 0);
  for (int y = 0; y < g->H; ++y) memcpy(fr + (size_t)y * g->W * 3, frame_bgr + (size_t)y * stride, (size_t)g->W * 3);
  /* app/deepseg.cc:661: raw = alpha_blend(bg, raw, mask) */
  or_alpha_blend(bg, fr, g->mask, out_rgb, npix);
  /* app/deepseg.cc:681 */
  if (out_yuyv) or_convert_rgb_to_yuyv(out_rgb, out_yuyv, g->W, g->H);
  return 0;
}


int or_composite_ex(or_maskgen* g, const uint8_t* frame_bgr, size_t stride,
                    const uint8_t* bg_raw, int bw, int bh, size_t bstride, const or_frame_opts* o,
                    uint8_t* out_rgb, uint8_t* out_yuyv, uint8_t* out_mask) {
  const int W = g->W, H = g->H;
  const size_t npix = (size_t)W * H;
  const int ow = o->out_w > 0 ? o->out_w : W, oh = o->out_h > 0 ? o->out_h : H;
  int rc = or_maskgen_process(g, frame_bgr, stride, out_mask);
  if (rc) return rc;
  uint8_t* fr = scratch_buf(g, 0);
  uint8_t* bg = scratch_buf(g, 1);
  uint8_t* cur = scratch_buf(g, 2);
  for (int y = 0; y < H; ++y) memcpy(fr + (size_t)y * W * 3, frame_bgr + (size_t)y * stride, (size_t)W * 3);
  /* app/deepseg.cc:649-655: the grabbed background, else a copy of the camera frame */
  if (bg_raw) or_resize_linear_u8(bg_raw, bw, bh, bstride, bg, W, H, (size_t)W * 3, 3);
  else memcpy(bg, fr, npix * 3);
  /* :657-658 (in place in the reference; GaussianBlur clones the source when src == dst) */
  if (o->bgblur_k) {
    memcpy(cur, bg, npix * 3);
    if (or_gaussian_blur_u8c3(cur, W, H, (size_t)W * 3, bg, (size_t)W * 3, o->bgblur_k)) return -1;
  }
  or_alpha_blend(bg, fr, g->mask, cur, npix);                                   /* :661 */
  if (o->flip_h || o->flip_v) { or_flip_u8c3(cur, bg, W, H, o->flip_h, o->flip_v); memcpy(cur, bg, npix * 3); }   /* :667-673 */
  if (ow != W || oh != H) or_resize_linear_u8(cur, W, H, (size_t)W * 3, out_rgb, ow, oh, (size_t)ow * 3, 3);     /* :677-679 */
  else memcpy(out_rgb, cur, npix * 3);
  if (out_yuyv) or_convert_rgb_to_yuyv(out_rgb, out_yuyv, ow, oh);              /* :681 */
  return 0;
}
