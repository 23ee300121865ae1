// backscrub_b200/csrc/kernels_img.cu — integer image stages of the hot path (sm_100a).
//
// Replaces the OpenCV calls of lib/libbackscrub.cc:285-302 (ROI resize, BGR2RGB,
// bilateralFilter, convertTo), :314-361 (decision + IIR), :366-371 (mask upsample + 5x5
// blur) and app/deepseg.cc:108-134 (alpha_blend), :87-106 (convert_rgb_to_yuyv),
// app/background.cc:178-194 (background resize).  All integer arithmetic is OpenCV's
// fixed-point arithmetic restated (see oracle/oracle_img.c for the formulas and their
// cv2 pins); results are bit-exact against the oracle.
#include "kernels.h"
#include "post_math.h"

namespace bsb {

void count_launch();

// cv::resize INTER_LINEAR 8-bit, one output sample: horizontal taps at scale 2^11, vertical
// descale ((b0*(H0>>4))>>16) + ((b1*(H1>>4))>>16) + 2) >> 2.
BSB_D int lin_h(const uint8_t* row, int sx, int sx1, int cn, int c, int a0, int a1) {
  return (int)row[sx * cn + c] * a0 + (int)row[sx1 * cn + c] * a1;
}
BSB_D uint8_t lin_v(int h0, int h1, int b0, int b1) {
  return bsb_sat_u8((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
}

// background of batch frame b: one static image (bg_stride 0), one image per frame, or a ring of
// bg_count images entered at the device-side cursor (animated backgrounds, app/background.cc:126-176)
BSB_D size_t bg_frame_offset(const PostArgs& a, int b) {
  if (a.bg_cursor) b = (int)(((unsigned)__ldg(a.bg_cursor) + (unsigned)b * (unsigned)a.bg_advance) % (unsigned)a.bg_count);
  return (size_t)b * a.bg_stride;
}

// ---------------------------------------------------------------------------
// ROI crop -> resize -> BGR2RGB into the zero-padded model-sized image.
// One thread = one destination pixel (3 channels).
// ---------------------------------------------------------------------------
// IN_YUYV: `frames` are camera YUYV frames (W*2-byte rows); every tap is first converted the way
// cv::cvtColor(COLOR_YUV2BGR_YUYV) would have (cv::VideoCapture's conversion, app/deepseg.cc:553), so the result
// equals converting the whole frame first — without materialising the BGR frame.
template <bool IN_YUYV>
__global__ void __launch_bounds__(256) k_resize_roi_swap(int B, const uint8_t* frames, size_t frame_stride, size_t pitch,
                                                         int roi_x, int roi_y, int roi_w, int roi_h, ResizeTab t,
                                                         uint8_t* in_u8, int mw, int mh, int in_x, int in_y, int in_w, int in_h,
                                                         bool area2x2) {
  const long total = (long)B * in_w * in_h;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int dx = (int)(idx % in_w), dy = (int)((idx / in_w) % in_h), b = (int)(idx / ((long)in_w * in_h));
  uint8_t px[3];
  if (IN_YUYV) {
    const uint8_t* img = frames + (size_t)b * frame_stride;
    if (area2x2) {
      int p00[3], p01[3], p10[3], p11[3];
      const uint8_t* s0 = img + (size_t)(roi_y + 2 * dy) * pitch;
      const uint8_t* s1 = s0 + pitch;
      yuyv_px_to_bgr(s0, roi_x + 2 * dx, p00); yuyv_px_to_bgr(s0, roi_x + 2 * dx + 1, p01);
      yuyv_px_to_bgr(s1, roi_x + 2 * dx, p10); yuyv_px_to_bgr(s1, roi_x + 2 * dx + 1, p11);
#pragma unroll
      for (int c = 0; c < 3; ++c) px[c] = (uint8_t)((p00[c] + p01[c] + p10[c] + p11[c] + 2) >> 2);
    } else {
      const int sx = __ldg(t.xofs + dx), sx1 = min(sx + 1, roi_w - 1);
      const int a0 = __ldg(t.xw + 2 * dx), a1 = __ldg(t.xw + 2 * dx + 1);
      const int b0 = __ldg(t.yw + 2 * dy), b1 = __ldg(t.yw + 2 * dy + 1);
      const uint8_t* r0 = img + (size_t)(roi_y + __ldg(t.yofs0 + dy)) * pitch;
      const uint8_t* r1 = img + (size_t)(roi_y + __ldg(t.yofs1 + dy)) * pitch;
      int p00[3], p01[3], p10[3], p11[3];
      yuyv_px_to_bgr(r0, roi_x + sx, p00); yuyv_px_to_bgr(r0, roi_x + sx1, p01);
      yuyv_px_to_bgr(r1, roi_x + sx, p10); yuyv_px_to_bgr(r1, roi_x + sx1, p11);
#pragma unroll
      for (int c = 0; c < 3; ++c) px[c] = lin_v(p00[c] * a0 + p01[c] * a1, p10[c] * a0 + p11[c] * a1, b0, b1);
    }
  } else {
    const uint8_t* roi = frames + (size_t)b * frame_stride + (size_t)roi_y * pitch + (size_t)roi_x * 3;
    if (area2x2) {
      const uint8_t* s0 = roi + (size_t)(2 * dy) * pitch + (size_t)(2 * dx) * 3;
      const uint8_t* s1 = s0 + pitch;
#pragma unroll
      for (int c = 0; c < 3; ++c) px[c] = (uint8_t)((s0[c] + s0[3 + c] + s1[c] + s1[3 + c] + 2) >> 2);
    } else {
      const int sx = __ldg(t.xofs + dx), sx1 = min(sx + 1, roi_w - 1);
      const int a0 = __ldg(t.xw + 2 * dx), a1 = __ldg(t.xw + 2 * dx + 1);
      const int b0 = __ldg(t.yw + 2 * dy), b1 = __ldg(t.yw + 2 * dy + 1);
      const uint8_t* r0 = roi + (size_t)__ldg(t.yofs0 + dy) * pitch;
      const uint8_t* r1 = roi + (size_t)__ldg(t.yofs1 + dy) * pitch;
#pragma unroll
      for (int c = 0; c < 3; ++c) px[c] = lin_v(lin_h(r0, sx, sx1, 3, c, a0, a1), lin_h(r1, sx, sx1, 3, c, a0, a1), b0, b1);
    }
  }
  uint8_t* d = in_u8 + ((size_t)b * mh * mw + (size_t)(in_y + dy) * mw + (in_x + dx)) * 3;
  d[0] = px[2]; d[1] = px[1]; d[2] = px[0];   // BGR -> RGB
}

void launch_resize_roi_swap(cudaStream_t s, int B, const uint8_t* frames, size_t frame_stride, size_t frame_pitch,
                            int roi_x, int roi_y, int roi_w, int roi_h, ResizeTab tab,
                            uint8_t* in_u8, int mw, int mh, int in_x, int in_y, int in_w, int in_h, bool area2x2, bool in_yuyv) {
  const long total = (long)B * in_w * in_h;
  if (in_yuyv) BSB_LAUNCH(k_resize_roi_swap<true>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, B, frames, frame_stride, frame_pitch,
                          roi_x, roi_y, roi_w, roi_h, tab, in_u8, mw, mh, in_x, in_y, in_w, in_h, area2x2);
  else BSB_LAUNCH(k_resize_roi_swap<false>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, B, frames, frame_stride, frame_pitch,
                  roi_x, roi_y, roi_w, roi_h, tab, in_u8, mw, mh, in_x, in_y, in_w, in_h, area2x2);
  count_launch();
}

// ---------------------------------------------------------------------------
// cv::bilateralFilter(d=5, sigma 100/100) + convertTo(CV_32F, scale, offset).
// 13 taps (i^2+j^2 <= 4) in raster order, weights = space_w[k] * color_w[|db|+|dg|+|dr|],
// per-tap fmaf accumulation, cvRound(sum * (1/wsum)), BORDER_REFLECT_101.
// One thread = one pixel.
// ---------------------------------------------------------------------------
// One block = a 32 x 8 pixel tile: the tile plus its 2-pixel REFLECT_101 border is staged once in shared memory as
// packed words (b | g << 8 | r << 16), so each tap costs one LDS instead of three byte loads and two reflections.
constexpr int BF_TW = 32, BF_TH = 8, BF_SW = BF_TW + 4, BF_SH = BF_TH + 4;

__global__ void __launch_bounds__(256) k_bilateral_norm(int B, const uint8_t* in_u8, int mw, int mh,
                                                        const float* color_w, const float* space_w,
                                                        float scale, float offset, float* out_f32, uint8_t* out_u8) {
  __shared__ float cw[768];
  __shared__ unsigned tile[BF_SH * BF_SW];
  __shared__ float sw[16];
  (void)B;
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * BF_TW, y0 = blockIdx.y * BF_TH, b = blockIdx.z;
  const uint8_t* img = in_u8 + (size_t)b * mh * mw * 3;
  for (int i = tid; i < 768; i += 256) cw[i] = __ldg(color_w + i);
  if (tid < 13) sw[tid] = __ldg(space_w + tid);
  for (int i = tid; i < BF_SH * BF_SW; i += 256) {
    const int ty = i / BF_SW, tx = i - ty * BF_SW;
    const uint8_t* p = img + ((size_t)bsb_reflect101(y0 + ty - 2, mh) * mw + bsb_reflect101(x0 + tx - 2, mw)) * 3;
    tile[i] = (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16);
  }
  __syncthreads();
  const int lx = tid & (BF_TW - 1), ly = tid >> 5;
  const int x = x0 + lx, y = y0 + ly;
  if (x >= mw || y >= mh) return;
  const unsigned* t0 = tile + (ly + 2) * BF_SW + lx + 2;
  const unsigned cc = t0[0];
  const int c0 = (int)(cc & 255u), c1 = (int)((cc >> 8) & 255u), c2 = (int)(cc >> 16);
  float wsum = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f;
  int k = 0;
#pragma unroll
  for (int i = -2; i <= 2; ++i) {
#pragma unroll
    for (int j = -2; j <= 2; ++j) {
      if (i * i + j * j > 4) continue;
      const unsigned v = t0[i * BF_SW + j];
      const int v0 = (int)(v & 255u), v1 = (int)((v >> 8) & 255u), v2 = (int)(v >> 16);
      const float w = sw[k] * cw[abs(v0 - c0) + abs(v1 - c1) + abs(v2 - c2)];
      ++k;
      wsum = wsum + w;
      s0 = fmaf((float)v0, w, s0);
      s1 = fmaf((float)v1, w, s1);
      s2 = fmaf((float)v2, w, s2);
    }
  }
  const size_t idx = ((size_t)b * mh + y) * mw + x;
  const float inv = bsb_div(1.f, wsum);
  const int r0 = bsb_sat_u8(__float2int_rn(s0 * inv)), r1 = bsb_sat_u8(__float2int_rn(s1 * inv)), r2 = bsb_sat_u8(__float2int_rn(s2 * inv));
  if (out_f32) {
    float* o = out_f32 + idx * 3;
    o[0] = fmaf((float)r0, scale, offset);
    o[1] = fmaf((float)r1, scale, offset);
    o[2] = fmaf((float)r2, scale, offset);
  }
  if (out_u8) { uint8_t* u = out_u8 + idx * 3; u[0] = (uint8_t)r0; u[1] = (uint8_t)r1; u[2] = (uint8_t)r2; }
}

void launch_bilateral_norm(cudaStream_t s, int B, const uint8_t* in_u8, int mw, int mh,
                           const float* color_w, const float* space_w, float scale, float offset,
                           float* out_f32, uint8_t* out_u8_dbg) {
  BSB_LAUNCH(k_bilateral_norm, dim3((unsigned)ceil_div(mw, BF_TW), (unsigned)ceil_div(mh, BF_TH), (unsigned)B), dim3(256), 0, s, B, in_u8, mw, mh,
             color_w, space_w, scale, offset, out_f32, out_u8_dbg);
  count_launch();
}

// ---------------------------------------------------------------------------
// Decision + temporal IIR, lib/libbackscrub.cc:314-361.  One thread = one model-output
// pixel, looping over the B consecutive frames of the batch so the 3-tap state
// out = (val & 0xE0) | (out >> 3) advances in frame order.
// ---------------------------------------------------------------------------
BSB_D unsigned decide(int model_type, const float* t) {
  if (model_type == MODEL_DEEPLAB) {
    float maxval = -10000.f; int maxpos = 0;
    for (int i = 0; i < 21; ++i) { const float v = __ldg(t + i); if (v > maxval) { maxval = v; maxpos = i; } }
    return (maxpos == 15) ? 0u : 1u;
  }
  if (model_type == MODEL_MEET) {
    const float e0 = bsb_expf(__ldg(t)), e1 = bsb_expf(__ldg(t + 1));
    const float p0 = bsb_div(e0, e0 + e1), p1 = bsb_div(e1, e0 + e1);
    return (p0 < p1) ? 0u : 1u;
  }
  // `tmp[n] > 0.65` compares float with the double literal; equivalent to > 0.65f
  // (0.65 lies strictly between two adjacent floats) — SURVEY.md Appendix A
  return (__ldg(t) > 0.65f) ? 0u : 1u;
}

// The per-frame decisions of a pixel are independent (computed 8 at a time so their loads
// overlap); only the 3-tap state update is sequential, and it is pure ALU on a bit mask.
__global__ void __launch_bounds__(128) k_decision_iir(int model_type, int B, const float* out_f, int npix, int oc,
                                                      uint8_t* state, uint8_t* ofinal, int ow, int opitch, int oframe) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= npix) return;
  const int opos = (n / ow) * opitch + (n % ow);      // row-padded position inside one frame of ofinal
  unsigned st = state[n];
  for (int b0 = 0; b0 < B; b0 += 8) {
    unsigned bits = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (b0 + j < B) bits |= decide(model_type, out_f + ((size_t)(b0 + j) * npix + n) * oc) << j;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (b0 + j < B) {
        const unsigned val = ((bits >> j) & 1u) ? 255u : 0u;
        st = (val & 0xE0u) | (st >> 3);
        ofinal[(size_t)(b0 + j) * oframe + opos] = (uint8_t)st;
      }
  }
  state[n] = (uint8_t)st;
}

// Frame-parallel form of the same stage.  The 3-tap state is a finite window: unrolling
//   st_b = (val_b & 0xE0) | (st_{b-1} >> 3)
// gives st_b = E0*d_b | 1C*d_{b-1} | 03*d_{b-2} for b >= 2, st_1 = E0*d_1 | 1C*d_0 | (s >> 6), st_0 = E0*d_0 | (s >> 3)
// (s = the state the previous call left, any byte; d = 1 where val = 255), so nothing but the three latest decisions
// is carried from frame to frame.  k_decision_iir walks the B frames of a pixel in ONE thread — 36 864 threads for the
// Meet output however large the batch is, 12 % of the GPU's thread slots, 0.4 TB/s (profiles/r2_launch_shares_meet720_b256.txt:
// 195 us per 256-frame launch).  Here a block owns 32 pixels x all B frames: its eight warps take the frames round-robin,
// park the decisions in shared memory ([B][32] bytes) and, after one barrier, each thread assembles the state bytes of
// its frames from three of them.  The block that owns a pixel is the only one that touches its state byte: read before
// the barrier, written (by the owner of frame B-1) after it.
constexpr int DEC_PX = 32, DEC_WARPS = 8;

BSB_D void decision_window_store(const uint8_t* d, int tx, int ty, int B, unsigned s_in, uint8_t* state_px, uint8_t* ofinal_px, int oframe) {
  for (int b = ty; b < B; b += DEC_WARPS) {
    unsigned st = d[b * DEC_PX + tx] ? 0xE0u : 0u;
    if (b >= 1) st |= d[(b - 1) * DEC_PX + tx] ? 0x1Cu : 0u; else st |= s_in >> 3;
    if (b >= 2) st |= d[(b - 2) * DEC_PX + tx] ? 0x03u : 0u; else if (b == 1) st |= s_in >> 6;
    ofinal_px[(size_t)b * oframe] = (uint8_t)st;
    if (b == B - 1) *state_px = (uint8_t)st;
  }
}

__global__ void __launch_bounds__(DEC_PX * DEC_WARPS) k_decision_par(int model_type, int B, const float* out_f, int npix, int oc,
                                                                     uint8_t* state, uint8_t* ofinal, int ow, int opitch, int oframe) {
  BSB_DYN_SMEM(smem_raw);
  uint8_t* d = reinterpret_cast<uint8_t*>(smem_raw);               // [B][32] decisions
  const int tx = threadIdx.x % DEC_PX, ty = threadIdx.x / DEC_PX;
  const int n = blockIdx.x * DEC_PX + tx;
  const bool valid = n < npix;
  const unsigned s_in = valid ? state[n] : 0u;
  if (valid) {
#pragma unroll 4
    for (int b = ty; b < B; b += DEC_WARPS) d[b * DEC_PX + tx] = (uint8_t)decide(model_type, out_f + ((size_t)b * npix + n) * oc);
  }
  __syncthreads();
  if (!valid) return;
  decision_window_store(d, tx, ty, B, s_in, state + n, ofinal + (n / ow) * opitch + (n % ow), oframe);
}

void launch_decision_iir(cudaStream_t s, int model_type, int B, const float* model_out, int oh, int ow, int oc,
                         uint8_t* state, uint8_t* ofinal, int opitch) {
  const int npix = oh * ow;
  if (tuning().dec_par && (size_t)B * DEC_PX <= 48 * 1024) {
    BSB_LAUNCH(k_decision_par, dim3((unsigned)ceil_div(npix, DEC_PX)), dim3(DEC_PX * DEC_WARPS), (size_t)B * DEC_PX, s, model_type, B, model_out,
               npix, oc, state, ofinal, ow, opitch, oh * opitch);
    count_launch();
    return;
  }
  BSB_LAUNCH(k_decision_iir, dim3((unsigned)ceil_div(npix, 128)), dim3(128), 0, s, model_type, B, model_out, npix, oc, state, ofinal,
             ow, opitch, oh * opitch);
  count_launch();
}

// DeepLab: RESIZE_BILINEAR (33x33x21 -> 257x257x21) + argmax + IIR in one pass.  One thread = one output pixel; the four
// neighbours and weights are those of k_resize_bilinear, the value of class i is formed with the same expression
// (-fmad=false: every product and sum rounds separately), the argmax scans the classes in order with a strict >.
__global__ void __launch_bounds__(128) k_decision_up_iir(int B, const float* low, int ih, int iw, int ld, float hs, float ws, bool half_pixel,
                                                         int oh, int ow, uint8_t* state, uint8_t* ofinal, int opitch, int oframe) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= oh * ow) return;
  const int y = n / ow, x = n - y * ow;
  float fy, fx; int y0, y1, x0, x1;
  bsb_resize_interp((float)y, hs, half_pixel, ih, &fy, &y0, &y1);
  bsb_resize_interp((float)x, ws, half_pixel, iw, &fx, &x0, &x1);
  const float dy = fy - (float)y0, dx = fx - (float)x0;
  const float wy0 = 1.f - dy, wx0 = 1.f - dx;
  const int o00 = (y0 * iw + x0) * ld, o10 = (y1 * iw + x0) * ld, o01 = (y0 * iw + x1) * ld, o11 = (y1 * iw + x1) * ld;
  const int opos = y * opitch + x;
  unsigned st = state[n];
  for (int b = 0; b < B; ++b) {
    const float* f = low + (size_t)b * ih * iw * ld;
    float maxval = -10000.f; int maxpos = 0;
#pragma unroll 7
    for (int i = 0; i < 21; ++i) {
      const float a = __ldg(f + o00 + i) * wy0 * wx0;
      const float bb = __ldg(f + o10 + i) * dy * wx0;
      const float d = __ldg(f + o01 + i) * wy0 * dx;
      const float e = __ldg(f + o11 + i) * dy * dx;
      const float v = ((a + bb) + d) + e;
      if (v > maxval) { maxval = v; maxpos = i; }
    }
    const unsigned val = (maxpos == 15) ? 0u : 255u;
    st = (val & 0xE0u) | (st >> 3);
    ofinal[(size_t)b * oframe + opos] = (uint8_t)st;
  }
  state[n] = (uint8_t)st;
}

// frame-parallel form (see k_decision_par): a block = 32 output pixels x all B frames
__global__ void __launch_bounds__(DEC_PX * DEC_WARPS) k_decision_up_par(int B, const float* low, int ih, int iw, int ld, float hs, float ws, bool half_pixel,
                                                                        int oh, int ow, uint8_t* state, uint8_t* ofinal, int opitch, int oframe) {
  BSB_DYN_SMEM(smem_raw);
  uint8_t* d = reinterpret_cast<uint8_t*>(smem_raw);
  const int tx = threadIdx.x % DEC_PX, ty = threadIdx.x / DEC_PX;
  const int n = blockIdx.x * DEC_PX + tx;
  const bool valid = n < oh * ow;
  const unsigned s_in = valid ? state[n] : 0u;
  const int y = valid ? n / ow : 0, x = valid ? n - y * ow : 0;
  if (valid) {
    float fy, fx; int y0, y1, x0, x1;
    bsb_resize_interp((float)y, hs, half_pixel, ih, &fy, &y0, &y1);
    bsb_resize_interp((float)x, ws, half_pixel, iw, &fx, &x0, &x1);
    const float dy = fy - (float)y0, dx = fx - (float)x0;
    const float wy0 = 1.f - dy, wx0 = 1.f - dx;
    const int o00 = (y0 * iw + x0) * ld, o10 = (y1 * iw + x0) * ld, o01 = (y0 * iw + x1) * ld, o11 = (y1 * iw + x1) * ld;
    for (int b = ty; b < B; b += DEC_WARPS) {
      const float* f = low + (size_t)b * ih * iw * ld;
      float maxval = -10000.f; int maxpos = 0;
#pragma unroll 7
      for (int i = 0; i < 21; ++i) {
        const float a = __ldg(f + o00 + i) * wy0 * wx0;
        const float bb = __ldg(f + o10 + i) * dy * wx0;
        const float dd = __ldg(f + o01 + i) * wy0 * dx;
        const float e = __ldg(f + o11 + i) * dy * dx;
        const float v = ((a + bb) + dd) + e;
        if (v > maxval) { maxval = v; maxpos = i; }
      }
      d[b * DEC_PX + tx] = (maxpos == 15) ? 0 : 1;
    }
  }
  __syncthreads();
  if (!valid) return;
  decision_window_store(d, tx, ty, B, s_in, state + n, ofinal + y * opitch + x, oframe);
}

void launch_decision_up_iir(cudaStream_t s, int B, const float* low, int ih, int iw, int ld, bool align_corners, bool half_pixel,
                            int oh, int ow, uint8_t* state, uint8_t* ofinal, int opitch) {
  float hs = (float)ih / (float)oh, ws = (float)iw / (float)ow;
  if (align_corners && oh > 1) hs = (float)(ih - 1) / (float)(oh - 1);
  if (align_corners && ow > 1) ws = (float)(iw - 1) / (float)(ow - 1);
  if (tuning().dec_par && (size_t)B * DEC_PX <= 48 * 1024) {
    BSB_LAUNCH(k_decision_up_par, dim3((unsigned)ceil_div(oh * ow, DEC_PX)), dim3(DEC_PX * DEC_WARPS), (size_t)B * DEC_PX, s, B, low, ih, iw, ld, hs, ws,
               half_pixel, oh, ow, state, ofinal, opitch, oh * opitch);
    count_launch();
    return;
  }
  BSB_LAUNCH(k_decision_up_iir, dim3((unsigned)ceil_div(oh * ow, 128)), dim3(128), 0, s, B, low, ih, iw, ld, hs, ws, half_pixel, oh, ow, state,
             ofinal, opitch, oh * opitch);
  count_launch();
}

// ---------------------------------------------------------------------------
// Fused post stage: mask upsample (cv::resize 8UC1) + cv::blur 5x5 (REFLECT_101 inside the
// ROI) + alpha_blend + optional RGB->YUYV + optional mask store.
//
// Block = 256 threads = one 128 x 32 pixel tile; thread = 16 consecutive pixels of one row
// (48 B per 3-channel array: three 16-byte vector accesses).  Shared memory holds the
// upsampled tile with a 2-pixel halo (u8) and its vertical 5-sums (u16); box sums are
// integer, so the separable evaluation is exact.
// ---------------------------------------------------------------------------
constexpr int PT_W = 128, PT_H = 32, PT_PX = 16;
constexpr int PT_UW = PT_W + 4, PT_UH = PT_H + 4, PT_US = PT_UW + 4;   // smem row stride (bytes)

struct Px16 { uint4 v[3]; };

BSB_D void load48(const uint8_t* p, bool fast, Px16& r) {
  if (fast) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    r.v[0] = __ldg(q); r.v[1] = __ldg(q + 1); r.v[2] = __ldg(q + 2);
  } else {
    uint8_t* d = reinterpret_cast<uint8_t*>(&r);
    for (int i = 0; i < 48; ++i) d[i] = p[i];
  }
}

BSB_D void rgb2yuv(int R, int G, int Bc, int& Y, int& U, int& V) {
  // cv::cvtColor(COLOR_RGB2YUV) 8-bit, 14-bit fixed point (oracle_img.c:or_rgb2yuv_u8)
  Y = (4899 * R + 9617 * G + 1868 * Bc + 8192) >> 14;
  U = ((Bc - Y) * 8061 + (128 << 14) + 8192) >> 14;
  V = ((R - Y) * 14369 + (128 << 14) + 8192) >> 14;
  Y = Y > 255 ? 255 : Y;
  U = U < 0 ? 0 : (U > 255 ? 255 : U);
  V = V < 0 ? 0 : (V > 255 ? 255 : V);
}

__global__ void __launch_bounds__(256) k_post(PostArgs a) {
  __shared__ __align__(16) uint8_t Us[PT_UH * PT_US];
  __shared__ __align__(16) unsigned short Vs[PT_H * PT_UW];
  const int b = blockIdx.z;
  const int tx0 = blockIdx.x * PT_W, ty0 = blockIdx.y * PT_H;
  const int tid = threadIdx.x;
  const bool hits_roi = tx0 < a.roi_x + a.roi_w && tx0 + PT_W > a.roi_x && ty0 < a.roi_y + a.roi_h && ty0 + PT_H > a.roi_y;

  if (hits_roi) {
    // ---- A: upsampled mask tile with halo, reflect-101 at the ROI border ----
    const uint8_t* src = a.ofinal + (size_t)b * a.opitch * a.oh + (size_t)a.out_y * a.opitch + a.out_x;
    for (int i = tid; i < PT_UH * PT_UW; i += 256) {
      const int uy = i / PT_UW, ux = i % PT_UW;
      const int gy = bsb_reflect101(ty0 - a.roi_y - 2 + uy, a.roi_h);
      const int gx = bsb_reflect101(tx0 - a.roi_x - 2 + ux, a.roi_w);
      if (a.area2x2) {      // exact 2x down-scale: cv::resize takes its INTER_AREA 2x2 mean ((a + b + c + d + 2) >> 2)
        const uint8_t* s0 = src + (size_t)(2 * gy) * a.opitch + 2 * gx;
        const uint8_t* s1 = s0 + a.opitch;
        Us[uy * PT_US + ux] = (uint8_t)(((int)s0[0] + s0[1] + s1[0] + s1[1] + 2) >> 2);
        continue;
      }
      const int sx = __ldg(a.tab.xofs + gx), sx1 = min(sx + 1, a.out_w - 1);
      const int a0 = __ldg(a.tab.xw + 2 * gx), a1 = __ldg(a.tab.xw + 2 * gx + 1);
      const int b0 = __ldg(a.tab.yw + 2 * gy), b1 = __ldg(a.tab.yw + 2 * gy + 1);
      const uint8_t* r0 = src + (size_t)__ldg(a.tab.yofs0 + gy) * a.opitch;
      const uint8_t* r1 = src + (size_t)__ldg(a.tab.yofs1 + gy) * a.opitch;
      Us[uy * PT_US + ux] = lin_v(lin_h(r0, sx, sx1, 1, 0, a0, a1), lin_h(r1, sx, sx1, 1, 0, a0, a1), b0, b1);
    }
    __syncthreads();
    // ---- B: vertical 5-sums ----
    for (int i = tid; i < PT_H * PT_UW; i += 256) {
      const int y = i / PT_UW, x = i % PT_UW;
      const uint8_t* u = Us + y * PT_US + x;
      Vs[i] = (unsigned short)(u[0] + u[PT_US] + u[2 * PT_US] + u[3 * PT_US] + u[4 * PT_US]);
    }
    __syncthreads();
  }

  const int lx = (tid % (PT_W / PT_PX)) * PT_PX, ly = tid / (PT_W / PT_PX);
  const int x0 = tx0 + lx, y = ty0 + ly;
  if (y >= a.H || x0 >= a.W) return;
  const int npx = min(PT_PX, a.W - x0);

  // ---- C: horizontal 5-sums -> mask ----
  uint8_t m[PT_PX];
  const bool row_in = hits_roi && y >= a.roi_y && y < a.roi_y + a.roi_h;
#pragma unroll
  for (int i = 0; i < PT_PX; ++i) {
    const int x = x0 + i;
    unsigned mv = 255u;
    if (row_in && x >= a.roi_x && x < a.roi_x + a.roi_w) {
      const unsigned short* v = Vs + ly * PT_UW + lx + i;
      mv = bsb_box25((unsigned)v[0] + v[1] + v[2] + v[3] + v[4]);
    }
    m[i] = (uint8_t)mv;
  }

  // ---- D: blend (+ YUYV, + mask) ----
  const size_t fo = (size_t)b * a.frame_stride + (size_t)y * a.frame_pitch + (size_t)x0 * 3;
  const size_t bo = bg_frame_offset(a, b) + (size_t)y * a.bg_pitch + (size_t)x0 * 3;
  const bool full = npx == PT_PX;
  const bool fast_in = full && ((reinterpret_cast<uintptr_t>(a.frames + fo) | reinterpret_cast<uintptr_t>(a.bg + bo)) & 15) == 0;
  Px16 f, g, o;
  if (full) { load48(a.frames + fo, fast_in, f); load48(a.bg + bo, fast_in, g); }
  else {
    uint8_t* fd = reinterpret_cast<uint8_t*>(&f); uint8_t* gd = reinterpret_cast<uint8_t*>(&g);
    for (int i = 0; i < npx * 3; ++i) { fd[i] = a.frames[fo + i]; gd[i] = a.bg[bo + i]; }
  }
  const uint8_t* fb = reinterpret_cast<const uint8_t*>(&f);
  const uint8_t* gb = reinterpret_cast<const uint8_t*>(&g);
  uint8_t* ob = reinterpret_cast<uint8_t*>(&o);
#pragma unroll
  for (int i = 0; i < PT_PX; ++i) {
    const unsigned aw = m[i], bw = 255u - aw;
#pragma unroll
    for (int c = 0; c < 3; ++c) ob[3 * i + c] = (uint8_t)bsb_div255((unsigned)gb[3 * i + c] * aw + (unsigned)fb[3 * i + c] * bw);
  }
  if (a.out) {
    uint8_t* op = a.out + (size_t)b * a.out_stride + (size_t)y * a.out_pitch + (size_t)x0 * 3;
    if (full && (reinterpret_cast<uintptr_t>(op) & 15) == 0) {
      uint4* q = reinterpret_cast<uint4*>(op); q[0] = o.v[0]; q[1] = o.v[1]; q[2] = o.v[2];
    } else for (int i = 0; i < npx * 3; ++i) op[i] = ob[i];
  }
  if (a.mask) {
    uint8_t* mp = a.mask + (size_t)b * a.mask_stride + (size_t)y * a.W + x0;
    if (full && (reinterpret_cast<uintptr_t>(mp) & 15) == 0) *reinterpret_cast<uint4*>(mp) = *reinterpret_cast<const uint4*>(m);
    else for (int i = 0; i < npx; ++i) mp[i] = m[i];
  }
  if (a.yuyv) {
    // pairs are taken over the flattened image (app/deepseg.cc:97-104); W even => pairs never straddle rows
    __align__(16) uint8_t yy[PT_PX * 2];
#pragma unroll
    for (int i = 0; i < PT_PX; i += 2) {
      int Y0, U0, V0, Y1, U1, V1;
      rgb2yuv(ob[3 * i], ob[3 * i + 1], ob[3 * i + 2], Y0, U0, V0);
      rgb2yuv(ob[3 * i + 3], ob[3 * i + 4], ob[3 * i + 5], Y1, U1, V1);
      yy[2 * i] = (uint8_t)Y0; yy[2 * i + 1] = (uint8_t)((V0 + V1) / 2);
      yy[2 * i + 2] = (uint8_t)Y1; yy[2 * i + 3] = (uint8_t)((U0 + U1) / 2);
    }
    uint8_t* yp = a.yuyv + (size_t)b * a.yuyv_stride + ((size_t)y * a.W + x0) * 2;
    if (full && (reinterpret_cast<uintptr_t>(yp) & 15) == 0) {
      uint4* q = reinterpret_cast<uint4*>(yp); const uint4* sv = reinterpret_cast<const uint4*>(yy); q[0] = sv[0]; q[1] = sv[1];
    } else for (int i = 0; i < (npx & ~1) * 2; ++i) yp[i] = yy[i];
  }
}

// ---------------------------------------------------------------------------
// Fast path of the same fused stage (W % 16 == 0, tight 16-byte aligned rows).
// The kernel is bound by integer instruction issue, not by HBM, so it is organised to
// minimise instructions per pixel:
//   A0  column / row interpolation parameters of the tile -> shared memory (once)
//   A1  horizontal pass of cv::resize for the few source rows the tile touches (Hs, >>4)
//   A2  vertical pass -> upsampled mask tile with halo (Us, u16)
//   B   vertical 5-sums on packed u16x2 lanes with a sliding window (Vs)
//   C   horizontal 5-sums on packed lanes -> mask = (S + 12) / 25 as a multiply-shift
//   D   alpha blend with two channels per 32-bit IMAD (16-bit lanes share the pixel's
//       mask) and one via DP4A, division by 255 on packed lanes, RGB->YUV with DP2A,
//       16-byte streaming loads / stores.
// ---------------------------------------------------------------------------

template <bool OUT, bool YUYV>
__global__ void __launch_bounds__(256, 4) k_post_fast(PostArgs a) {
  __shared__ __align__(16) unsigned short Hs[PF_RMAX * PF_US];
  __shared__ __align__(16) unsigned short Us[PF_UH * PF_US];
  __shared__ __align__(16) unsigned short Vs[PF_H * PF_US];
  __shared__ __align__(16) uint4 rows[PF_UH];     // r0, r1, b0 << 16, b1 << 16
  __shared__ __align__(16) uint8_t Ps[PF_RMAX * PF_PS];   // source patch of the small mask
  const int b = blockIdx.z;
  const int tx0 = blockIdx.x * PF_W, ty0 = blockIdx.y * PF_H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool hits_roi = tx0 < a.roi_x + a.roi_w && tx0 + PF_W > a.roi_x && ty0 < a.roi_y + a.roi_h && ty0 + PF_H > a.roi_y;
  int tile_const = -1;                       // 255 / 0 when the whole mask tile is known to be constant
  {
    // pull this thread's 48 bytes of camera frame into L2 while the mask tile is being built
    const int px0 = tx0 + (tid & 7) * PF_PX, py = ty0 + (tid >> 3);
    if ((OUT || YUYV) && py < a.H && px0 < a.W) prefetch_l2(a.frames + (size_t)b * a.frame_stride + (size_t)py * a.frame_pitch + (size_t)px0 * 3);
  }

  if (hits_roi) {
    // yofs0/yofs1/xofs are monotonic, so the source patch this tile touches follows from the smallest /
    // largest (reflected) destination coordinates of the tile — no reduction needed.
    const int gy_lo = ty0 - a.roi_y - 2, gy_hi = gy_lo + PF_UH - 1;
    const int gy_min = gy_lo < 0 ? 0 : min(gy_lo, a.roi_h - 1);
    const int gy_max = gy_hi >= a.roi_h ? a.roi_h - 1 : max(gy_hi, 0);
    const int gx_lo = tx0 - a.roi_x - 2, gx_hi = gx_lo + PF_UW - 1;
    const int gx_min = gx_lo < 0 ? 0 : min(gx_lo, a.roi_w - 1);
    const int gx_max = gx_hi >= a.roi_w ? a.roi_w - 1 : max(gx_hi, 0);
    const int rmin = __ldg(a.tab.yofs0 + gy_min);
    const int nrows = __ldg(a.tab.yofs1 + gy_max) - rmin + 1;
    const int cmin = (int)(__ldg(&a.tab.xcol[gx_min].x) & 0xffffu);
    const int ncols = (int)(__ldg(&a.tab.xcol[gx_max].x) >> 16) - cmin + 1;
    // ---- P: source patch of the small mask -> shared memory (coalesced); row parameters for A2.
    //      If every source pixel of the patch is 255 (or 0) the upsampled + blurred tile is 255 (or 0)
    //      whatever the interpolation weights are (their sums are 2048 +- 1), so A1/A2/B are skipped. ----
    unsigned p_and = 255u, p_or = 0u;
    {
      const uint8_t* src = a.ofinal + (size_t)b * a.opitch * a.oh + (size_t)(a.out_y + rmin) * a.opitch + a.out_x + cmin;
      for (int r = warp; r < nrows; r += 8)
        for (int c = lane; c < ncols; c += 32) {
          const unsigned v = src[(size_t)r * a.opitch + c];
          Ps[r * PF_PS + c] = (uint8_t)v;
          p_and &= v; p_or |= v;
        }
      if (tid >= 192 && tid < 192 + PF_UH) {
        const int uy = tid - 192;
        int gy = gy_lo + uy;
        gy = gy < 0 ? -gy : gy; gy = gy >= a.roi_h ? 2 * a.roi_h - 2 - gy : gy;     // reflect-101 (single fold)
        gy = min(max(gy, 0), a.roi_h - 1);
        rows[uy] = make_uint4((unsigned)__ldg(a.tab.yofs0 + gy), (unsigned)__ldg(a.tab.yofs1 + gy),
                              (unsigned)(int)__ldg(a.tab.yw + 2 * gy) << 16, (unsigned)(int)__ldg(a.tab.yw + 2 * gy + 1) << 16);
      }
    }
    const int all_hi = __syncthreads_and(p_and == 255u);
    const int all_lo = all_hi ? 0 : __syncthreads_and(p_or == 0u);
    tile_const = all_hi ? 255 : (all_lo ? 0 : -1);
    if (tile_const < 0) {
    // ---- A1: horizontal pass of cv::resize on the patch rows: Hs = (s[sx]*a0 + s[sx1]*a1) >> 4.
    //      A thread owns a column (parameters in registers) and walks every second patch row. ----
    {
      const int phase = tid >> 7;
      for (int ux = tid & 127; ux < PF_UW; ux += 128) {
        int gx = gx_lo + ux;
        gx = gx < 0 ? -gx : gx; gx = gx >= a.roi_w ? 2 * a.roi_w - 2 - gx : gx;
        gx = min(max(gx, 0), a.roi_w - 1);
        const uint2 xc = __ldg(a.tab.xcol + gx);
        const int sx = (int)(xc.x & 0xffffu) - cmin, sx1 = (int)(xc.x >> 16) - cmin;
        const int a0 = (int)(short)(xc.y & 0xffffu), a1 = (int)(short)(xc.y >> 16);
        const uint8_t* pr = Ps + phase * PF_PS;
        unsigned short* hp = Hs + phase * PF_US + ux;
        for (int r = phase; r < nrows; r += 2) {
          *hp = (unsigned short)(((int)pr[sx] * a0 + (int)pr[sx1] * a1) >> 4);
          pr += 2 * PF_PS; hp += 2 * PF_US;
        }
      }
    }
    __syncthreads();
    // ---- A2: vertical pass -> upsampled tile.  ((b*h) >> 16) is one IMAD.HI with b pre-shifted by 16.
    //      A warp takes a row; each lane produces columns lane, lane+32, ... ----
    for (int uy = warp; uy < PF_UH; uy += 8) {
      const uint4 rp = rows[uy];
      const unsigned short* h0 = Hs + ((int)rp.x - rmin) * PF_US + lane;
      const unsigned short* h1 = Hs + ((int)rp.y - rmin) * PF_US + lane;
      unsigned short* up = Us + uy * PF_US + lane;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        up[32 * k] = (unsigned short)((__umulhi(rp.z, (unsigned)h0[32 * k]) + __umulhi(rp.w, (unsigned)h1[32 * k]) + 2u) >> 2);
      if (lane < PF_UW - 128)
        up[128] = (unsigned short)((__umulhi(rp.z, (unsigned)h0[128]) + __umulhi(rp.w, (unsigned)h1[128]) + 2u) >> 2);
    }
    __syncthreads();
    // ---- B: vertical 5-sums, two columns per 32-bit word, sliding window over 8 rows ----
    for (int it = tid; it < (PF_UW / 2) * 4; it += 256) {
      const int pair = it % (PF_UW / 2), seg = it / (PF_UW / 2);
      const unsigned* up = reinterpret_cast<const unsigned*>(Us + (seg * 8) * PF_US) + pair;
      unsigned* vp = reinterpret_cast<unsigned*>(Vs + (seg * 8) * PF_US) + pair;
      unsigned u[12];
#pragma unroll
      for (int k = 0; k < 12; ++k) u[k] = up[k * (PF_US / 2)];
      unsigned v = u[0] + u[1] + u[2] + u[3] + u[4];
      vp[0] = v;
#pragma unroll
      for (int k = 1; k < 8; ++k) { v = v + u[k + 4] - u[k - 1]; vp[k * (PF_US / 2)] = v; }
    }
    __syncthreads();
    }
  }

  const int lx = (tid & 7) * PF_PX, ly = tid >> 3;
  const int x0 = tx0 + lx, y = ty0 + ly;
  if (y >= a.H || x0 >= a.W) return;

  // ---- issue the streaming loads early ----
  unsigned f[12], g[12];
  if (OUT || YUYV) {
    const uint8_t* fp = a.frames + (size_t)b * a.frame_stride + (size_t)y * a.frame_pitch + (size_t)x0 * 3;
    const uint8_t* gp = a.bg + bg_frame_offset(a, b) + (size_t)y * a.bg_pitch + (size_t)x0 * 3;
    if (a.wide) {
      load48_wide(fp, (tid & 1) != 0, a.frame_l1 != 0, f);
      load48_wide(gp, (tid & 1) != 0, true, g);
    } else {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const uint4 t = a.frame_l1 ? __ldg(reinterpret_cast<const uint4*>(fp + 16 * k)) : ldg_stream(fp + 16 * k);
        f[4 * k] = t.x; f[4 * k + 1] = t.y; f[4 * k + 2] = t.z; f[4 * k + 3] = t.w;
        const uint4 s = __ldg(reinterpret_cast<const uint4*>(gp + 16 * k));
        g[4 * k] = s.x; g[4 * k + 1] = s.y; g[4 * k + 2] = s.z; g[4 * k + 3] = s.w;
      }
    }
  }

  // ---- C: horizontal 5-sums on packed lanes -> 16 mask values ----
  unsigned m[PF_PX];
  const bool row_in = hits_roi && y >= a.roi_y && y < a.roi_y + a.roi_h;
  post_mask16(a, row_in, tile_const, x0, Vs + ly * PF_US + lx, m);

  if (a.mask) {
    uint4 mv;
    mv.x = m[0] | (m[1] << 8) | (m[2] << 16) | (m[3] << 24);
    mv.y = m[4] | (m[5] << 8) | (m[6] << 16) | (m[7] << 24);
    mv.z = m[8] | (m[9] << 8) | (m[10] << 16) | (m[11] << 24);
    mv.w = m[12] | (m[13] << 8) | (m[14] << 16) | (m[15] << 24);
    stg_stream(a.mask + (size_t)b * a.mask_stride + (size_t)y * a.W + x0, mv);
  }
  if (!(OUT || YUYV)) return;

  // ---- D: blend (+ RGB->YUV) ----
  unsigned o[12], yy[8];
  post_blend16<OUT, YUYV>(f, g, m, o, yy);
  if (OUT) {
    uint8_t* op = a.out + (size_t)b * a.out_stride + (size_t)y * a.out_pitch + (size_t)x0 * 3;
    if (a.wide) store48_wide(op, (tid & 1) != 0, o);
    else {
#pragma unroll
      for (int k = 0; k < 3; ++k) stg_stream(op + 16 * k, make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]));
    }
  }
  if (YUYV) {
    uint8_t* yp = a.yuyv + (size_t)b * a.yuyv_stride + ((size_t)y * a.W + x0) * 2;
    if (a.wide) stg256(yp, yy);
    else {
      stg_stream(yp, make_uint4(yy[0], yy[1], yy[2], yy[3]));
      stg_stream(yp + 16, make_uint4(yy[4], yy[5], yy[6], yy[7]));
    }
  }
}

static bool post_fast_ok(const PostArgs& a) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (a.W % 16 != 0 || a.area2x2) return false;
  if (!al16(a.frames) || !al16(a.bg) || a.frame_pitch % 16 || a.frame_stride % 16 || a.bg_pitch % 16 || a.bg_stride % 16) return false;
  if (a.out && (!al16(a.out) || a.out_pitch % 16 || a.out_stride % 16)) return false;
  if (a.yuyv && (!al16(a.yuyv) || a.yuyv_stride % 16)) return false;
  if (a.mask && (!al16(a.mask) || a.mask_stride % 16)) return false;
  if (a.ow > 32000 || a.oh > 32000 || a.roi_w < 8 || a.roi_h < 8) return false;
  // source rows touched by one 36-row tile must fit the Hs buffer: ceil(36 * scale) + 2
  const double scale_y = (double)a.out_h / (double)a.roi_h;
  if ((int)(PF_UH * scale_y) + 3 > PF_RMAX) return false;
  const double scale_x = (double)a.out_w / (double)a.roi_w;
  if ((int)(PF_UW * scale_x) + 4 > PF_PS || a.tab.xcol == nullptr) return false;
  return true;
}

bool launch_post_tma(cudaStream_t s, const PostArgs& a);   // kernels_post.cu

void launch_post(cudaStream_t s, const PostArgs& a_in) {
  if (launch_post_tma(s, a_in)) return;
  PostArgs a = a_in;
  // measurement switches (profiles/r1_post_ab_run29.txt): both on is the fastest at 720p and at 4k
  const int wide_en = tuning().post_wide;
  a.frame_l1 = tuning().post_l1;
  auto al32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; };
  a.wide = wide_en && a.W % 32 == 0 && al32(a.frames) && al32(a.bg) && a.frame_pitch % 32 == 0 && a.frame_stride % 32 == 0 &&
           a.bg_pitch % 32 == 0 && a.bg_stride % 32 == 0 && (!a.out || (al32(a.out) && a.out_pitch % 32 == 0 && a.out_stride % 32 == 0)) &&
           (!a.yuyv || (al32(a.yuyv) && a.yuyv_stride % 32 == 0));
  if (post_fast_ok(a)) {
    dim3 grid((unsigned)ceil_div(a.W, PF_W), (unsigned)ceil_div(a.H, PF_H), (unsigned)a.B);
    if (a.yuyv && a.out) { auto k = k_post_fast<true, true>; BSB_LAUNCH(k, grid, dim3(256), 0, s, a); }
    else if (a.yuyv) { auto k = k_post_fast<false, true>; BSB_LAUNCH(k, grid, dim3(256), 0, s, a); }
    else if (a.out) { auto k = k_post_fast<true, false>; BSB_LAUNCH(k, grid, dim3(256), 0, s, a); }
    else { auto k = k_post_fast<false, false>; BSB_LAUNCH(k, grid, dim3(256), 0, s, a); }
    count_launch();
    return;
  }
  dim3 grid((unsigned)ceil_div(a.W, PT_W), (unsigned)ceil_div(a.H, PT_H), (unsigned)a.B);
  BSB_LAUNCH(k_post, grid, dim3(256), 0, s, a);
  count_launch();
}

// ---------------------------------------------------------------------------
// cv::resize(src -> dst) 8UC3 (background provider).  One thread = one dst pixel.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_resize_u8c3(const uint8_t* src, int sw, int sh, size_t spitch,
                                                     uint8_t* dst, int dw, int dh, size_t dpitch, ResizeTab t, bool area2x2,
                                                     size_t sstride, size_t dstride) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)dw * dh) return;
  src += (size_t)blockIdx.y * sstride; dst += (size_t)blockIdx.y * dstride;
  const int dx = (int)(idx % dw), dy = (int)(idx / dw);
  uint8_t* d = dst + (size_t)dy * dpitch + (size_t)dx * 3;
  if (area2x2) {
    const uint8_t* s0 = src + (size_t)(2 * dy) * spitch + (size_t)(2 * dx) * 3;
    const uint8_t* s1 = s0 + spitch;
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c] = (uint8_t)((s0[c] + s0[3 + c] + s1[c] + s1[3 + c] + 2) >> 2);
    return;
  }
  const int sx = __ldg(t.xofs + dx), sx1 = min(sx + 1, sw - 1);
  const int a0 = __ldg(t.xw + 2 * dx), a1 = __ldg(t.xw + 2 * dx + 1);
  const int b0 = __ldg(t.yw + 2 * dy), b1 = __ldg(t.yw + 2 * dy + 1);
  const uint8_t* r0 = src + (size_t)__ldg(t.yofs0 + dy) * spitch;
  const uint8_t* r1 = src + (size_t)__ldg(t.yofs1 + dy) * spitch;
#pragma unroll
  for (int c = 0; c < 3; ++c) d[c] = lin_v(lin_h(r0, sx, sx1, 3, c, a0, a1), lin_h(r1, sx, sx1, 3, c, a0, a1), b0, b1);
  (void)sh;
}

void launch_resize_u8c3(cudaStream_t s, const uint8_t* src, int sw, int sh, size_t spitch,
                        uint8_t* dst, int dw, int dh, size_t dpitch, ResizeTab tab, bool area2x2,
                        int n, size_t sstride, size_t dstride) {
  const long total = (long)dw * dh;
  BSB_LAUNCH(k_resize_u8c3, dim3((unsigned)((total + 255) / 256), (unsigned)n), dim3(256), 0, s, src, sw, sh, spitch, dst, dw, dh, dpitch, tab, area2x2,
             sstride, dstride);
  count_launch();
}

// ---------------------------------------------------------------------------
// YUYV -> BGR ingest (cv::cvtColor COLOR_YUV2BGR_YUYV, BT.601 limited range, 20-bit fixed point).
// One thread = 8 pixels: one 16-byte load, 24 bytes out.
// ---------------------------------------------------------------------------
BSB_D void yuv2bgr_px(int y, int ruv, int guv, int buv, uint8_t* o) {
  const int yy = max(y - 16, 0) * 1220542;
  o[0] = bsb_sat_u8((yy + buv) >> 20); o[1] = bsb_sat_u8((yy + guv) >> 20); o[2] = bsb_sat_u8((yy + ruv) >> 20);
}

__global__ void __launch_bounds__(256) k_yuyv_to_bgr(const uint8_t* yuyv, uint8_t* bgr, size_t ngroups, size_t npix, int vec_ok) {
  const size_t gidx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gidx >= ngroups) return;          // (only in the last, partial block: full blocks reach the barrier below together)
  const size_t p0 = gidx * 8;
  const int n = (npix - p0) < 8 ? (int)(npix - p0) : 8;
  unsigned w[4] = {0u, 0u, 0u, 0u};     // 8 pixels = 4 words of Y0 U Y1 V
  if (n == 8) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(yuyv + 2 * p0));
    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
  } else {
    for (int i = 0; i < 2 * n; ++i) w[i >> 2] |= (unsigned)yuyv[2 * p0 + i] << (8 * (i & 3));
  }
  unsigned o[6];                              // 24 output bytes, assembled in registers
  yuyv8_to_bgr24(w, o);
  // full blocks: stage the 24-byte groups in shared memory and write 16-byte vectors (a warp then writes 768
  // contiguous bytes in 1.5 passes instead of three 8-byte-strided passes that each touch every sector)
  __shared__ __align__(16) uint2 stage[256 * 3];
  const size_t blk_p0 = (size_t)blockIdx.x * blockDim.x * 8;
  const bool full_block = blk_p0 + (size_t)blockDim.x * 8 <= npix && blockDim.x == 256 && vec_ok;
  if (full_block) {
    uint2* q = stage + threadIdx.x * 3;
    q[0] = make_uint2(o[0], o[1]); q[1] = make_uint2(o[2], o[3]); q[2] = make_uint2(o[4], o[5]);
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(bgr + 3 * blk_p0);
    const uint4* src = reinterpret_cast<const uint4*>(stage);
    dst[threadIdx.x] = src[threadIdx.x];
    if (threadIdx.x < 128) dst[256 + threadIdx.x] = src[256 + threadIdx.x];
    return;
  }
  uint8_t* d = bgr + 3 * p0;
  if (n == 8 && (reinterpret_cast<uintptr_t>(d) & 7) == 0) {
    uint2* q = reinterpret_cast<uint2*>(d);
    q[0] = make_uint2(o[0], o[1]); q[1] = make_uint2(o[2], o[3]); q[2] = make_uint2(o[4], o[5]);
  } else {
    for (int i = 0; i < 3 * n; ++i) d[i] = (uint8_t)(o[i >> 2] >> (8 * (i & 3)));
  }
}

void launch_yuyv_to_bgr(cudaStream_t s, const uint8_t* yuyv, uint8_t* bgr, size_t npix_total) {
  const size_t ngroups = (npix_total + 7) / 8;
  const int vec_ok = (reinterpret_cast<uintptr_t>(bgr) & 15) == 0 ? 1 : 0;
  BSB_LAUNCH(k_yuyv_to_bgr, dim3((unsigned)((ngroups + 255) / 256)), dim3(256), 0, s, yuyv, bgr, ngroups, npix_total, vec_ok);
  count_launch();
}

// ---------------------------------------------------------------------------
// Stand-alone alpha_blend / convert_rgb_to_yuyv (stage-level parity through the C-ABI).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_alpha_blend(const uint8_t* a, const uint8_t* b, const uint8_t* mask, uint8_t* out, size_t npix) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const unsigned aw = mask[p], bw = 255u - aw;
#pragma unroll
  for (int c = 0; c < 3; ++c) out[3 * p + c] = (uint8_t)bsb_div255((unsigned)a[3 * p + c] * aw + (unsigned)b[3 * p + c] * bw);
}

void launch_alpha_blend(cudaStream_t s, const uint8_t* a, const uint8_t* b, const uint8_t* mask, uint8_t* out, size_t npix) {
  BSB_LAUNCH(k_alpha_blend, dim3((unsigned)((npix + 255) / 256)), dim3(256), 0, s, a, b, mask, out, npix);
  count_launch();
}

__global__ void __launch_bounds__(256) k_rgb_to_yuyv(const uint8_t* rgb, uint8_t* yuyv, size_t npairs, size_t rgb_stride, size_t yuyv_stride) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  rgb += (size_t)blockIdx.y * rgb_stride; yuyv += (size_t)blockIdx.y * yuyv_stride;
  const uint8_t* s = rgb + 6 * p;
  int Y0, U0, V0, Y1, U1, V1;
  rgb2yuv(s[0], s[1], s[2], Y0, U0, V0);
  rgb2yuv(s[3], s[4], s[5], Y1, U1, V1);
  uint8_t* d = yuyv + 4 * p;
  d[0] = (uint8_t)Y0; d[1] = (uint8_t)((V0 + V1) / 2); d[2] = (uint8_t)Y1; d[3] = (uint8_t)((U0 + U1) / 2);
}

void launch_rgb_to_yuyv(cudaStream_t s, const uint8_t* rgb, uint8_t* yuyv, size_t npix, int n, size_t rgb_stride, size_t yuyv_stride) {
  const size_t npairs = npix / 2;
  BSB_LAUNCH(k_rgb_to_yuyv, dim3((unsigned)((npairs + 255) / 256), (unsigned)n), dim3(256), 0, s, rgb, yuyv, npairs, rgb_stride, yuyv_stride);
  count_launch();
}

}  // namespace bsb
