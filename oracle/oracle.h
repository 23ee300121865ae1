/* oracle/oracle.h — CPU oracle for the backscrub per-frame hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or as
 * the timed CPU baseline.  The product library (backscrub_b200/csrc) neither
 * includes, links nor calls anything here and has no CPU fallback.
 *
 * What it restates (plain C, scalar, one thread unless OpenMP is enabled):
 *   - the .tflite loader + fp32 graph interpreter, following the TFLite
 *     *reference* kernels the survey names (see oracle_nn.c for file:line),
 *   - OpenCV's 8-bit image primitives used by lib/libbackscrub.cc and
 *     app/deepseg.cc (oracle_img.c), pinned against the in-container cv2,
 *   - the pipeline bs_maskgen_new/process + alpha_blend + convert_rgb_to_yuyv
 *     (oracle_pipeline.c).
 *
 * Parity status: every op is pinned by the TFLite single-op KATs transcribed
 * in tests/ and by oracle/_ref (the reference's own transpose_conv_bias.cc
 * compiled in place); the integer image ops are pinned bit-exact against
 * cv2 4.13.  WHOLE-MODEL outputs are parity-UNPINNED against the reference's
 * own TFLite+XNNPACK binary (it has no image->mask golden and cannot be built
 * offline).  What stands in: OpenCV 4.13's dnn module (a third-party TFLite
 * importer + kernels; OpenCV is the reference's system dependency) agrees with
 * this interpreter on the whole MLKit graph to 1e-5 and on every DeepLab /
 * BodyPix layer it imports correctly (tests/test_oracle_model.py, committed as
 * tests/golden/model_cv2dnn_golden.npz), and a torch-fp64 evaluation agrees on
 * every tensor of all five graphs.
 *
 * Numeric contract shared with the CUDA path (documented deviations from
 * the reference kernels; each is below the 1e-5/3e-6 tolerances the
 * reference's own tests use):
 *   - multiply-accumulate is a fused fmaf() in the reference loop order,
 *   - exp() is the fixed polynomial or_expf() below (|err| <= ~1 ulp),
 *   - global AVERAGE_POOL sums each row left-to-right, then the row sums
 *     top-to-bottom (pooling.h sums the whole window in one sequence),
 *   - denormals are flushed (the reference runs Invoke under FTZ/DAZ,
 *     tensorflow/lite/interpreter.cc:226).
 */
#ifndef BS_ORACLE_H
#define BS_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- op kinds (TFLite builtin codes, schema.fbs enum :229) ---- */
enum { OR_ADD = 0, OR_AVERAGE_POOL_2D = 1, OR_CONCATENATION = 2, OR_CONV_2D = 3,
       OR_DEPTHWISE_CONV_2D = 4, OR_DEQUANTIZE = 6, OR_FULLY_CONNECTED = 9,
       OR_LOGISTIC = 14, OR_MUL = 18, OR_RELU = 19, OR_RELU6 = 21,
       OR_RESIZE_BILINEAR = 23, OR_CUSTOM = 32, OR_HARD_SWISH = 117 };
enum { OR_ACT_NONE = 0, OR_ACT_RELU = 1, OR_ACT_RELU_N1_TO_1 = 2, OR_ACT_RELU6 = 3 };
enum { OR_PAD_SAME = 0, OR_PAD_VALID = 1 };

/* ---- scalar helpers (exported so tests can pin them) ---- */
float or_expf(float x);
float or_half_to_float(uint16_t h);

/* ---- single ops (batch = 1 unless stated; NHWC; weights OHWI) ---- */
void or_conv2d(const float* in, int ih, int iw, int ic,
               const float* w, int oc, int kh, int kw, const float* bias,
               int stride_h, int stride_w, int dil_h, int dil_w, int padding, int act,
               float* out, int oh, int ow);
void or_depthwise_conv2d(const float* in, int ih, int iw, int ic,
                         const float* w, int kh, int kw, const float* bias,
                         int stride_h, int stride_w, int dil_h, int dil_w, int padding,
                         int depth_mult, int act, float* out, int oh, int ow);
void or_conv_out_size(int in_size, int k, int stride, int dil, int padding, int* out_size, int* pad_before);
void or_average_pool(const float* in, int ih, int iw, int c, int fh, int fw,
                     int stride_h, int stride_w, int padding, int act,
                     float* out, int oh, int ow);
void or_fully_connected(const float* in, int batches, int in_depth, const float* w,
                        int out_depth, const float* bias, int act, float* out);
void or_resize_bilinear(const float* in, int ih, int iw, int c, float* out, int oh, int ow,
                        int align_corners, int half_pixel);
void or_hard_swish(const float* in, float* out, size_t n);
void or_logistic(const float* in, float* out, size_t n);
void or_relu(const float* in, float* out, size_t n, int act);
void or_add(const float* a, const float* b, float* out, size_t n, int act);
/* a: [n_outer, c]; b: either [n_outer, c] (bcast=0) or [c] (bcast=1) */
void or_mul(const float* a, const float* b, float* out, size_t n_outer, int c, int bcast, int act);
void or_tconv_bias(const float* in, int ih, int iw, int ic, const float* w, int oc, int kh, int kw,
                   const float* bias, int stride_h, int stride_w, int padding_same,
                   float* out, int oh, int ow);

/* ---- model loader / interpreter ---- */
typedef struct or_model or_model;
or_model* or_model_load(const char* path, char* err, size_t errlen);
void or_model_free(or_model* m);
int or_model_num_tensors(const or_model* m);
int or_model_num_ops(const or_model* m);
int or_model_input(const or_model* m);
int or_model_output(const or_model* m);
int or_model_tensor_shape(const or_model* m, int t, int shape[4]); /* returns rank */
int or_model_tensor_is_const(const or_model* m, int t);
const float* or_model_tensor_data(const or_model* m, int t);        /* f32 view (consts widened) */
int or_model_op(const or_model* m, int op, int* kind, int inputs[4], int* n_inputs, int* output);
int or_model_invoke(or_model* m, const float* input);               /* 0 on success */
double or_model_flops(const or_model* m);

/* ---- OpenCV-exact 8-bit image primitives ---- */
/* cv::resize(src, dst, dsize) default INTER_LINEAR, 8UC{1,3}; strides in bytes */
void or_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstride,
                         uint8_t* dst, int dw, int dh, size_t dstride, int cn);
/* cv::bilateralFilter(src8UC3, dst, 5, 100, 100), BORDER_DEFAULT */
void or_bilateral_d5_u8c3(const uint8_t* src, uint8_t* dst, int w, int h, double sigma_color, double sigma_space);
/* Mat::convertTo(CV_32F, alpha, beta) for 8U input */
void or_convert_u8_f32(const uint8_t* src, float* dst, size_t n, float alpha, float beta);
/* cv::blur(src, dst, Size(5,5)) 8UC1 BORDER_REFLECT_101 */
void or_box_blur5_u8(const uint8_t* src, size_t sstride, uint8_t* dst, size_t dstride, int w, int h);
/* cv::cvtColor(COLOR_RGB2YUV) 8UC3 (channel 0 treated as R) */
void or_rgb2yuv_u8(const uint8_t* src, uint8_t* dst, size_t npix);
/* app/deepseg.cc:87-106 */
void or_convert_rgb_to_yuyv(const uint8_t* src, uint8_t* dst_yuyv, int w, int h);
/* app/deepseg.cc:108-134 (srca = background, srcb = camera frame) */
void or_alpha_blend(const uint8_t* srca, const uint8_t* srcb, const uint8_t* mask, uint8_t* out, size_t npix);
/* cv::cvtColor(COLOR_YUV2BGR_YUYV): the YUYV camera frame -> BGR conversion (app/deepseg.cc:553,725) */
void or_yuyv_to_bgr(const uint8_t* yuyv, uint8_t* bgr, int w, int h);

/* cv::GaussianBlur(src, dst, Size(k,k), 0) 8UC3 BORDER_DEFAULT, OpenCV's bit-exact 8.8 fixed-point path
 * (app/deepseg.cc:657-658).  q receives the k fixed-point taps (sum 256).  Both return 0 on success. */
int or_gaussian_kernel_q8(int k, int* q);
int or_gaussian_blur_u8c3(const uint8_t* src, int w, int h, size_t sstride, uint8_t* dst, size_t dstride, int k);
/* cv::flip (app/deepseg.cc:667-673) */
void or_flip_u8c3(const uint8_t* src, uint8_t* dst, int w, int h, int flip_h, int flip_v);

/* ---- pipeline (lib/libbackscrub.cc:161-376 restated, deterministic) ---- */
enum { OR_MODEL_UNKNOWN = 0, OR_MODEL_BODYPIX, OR_MODEL_DEEPLAB, OR_MODEL_MEET, OR_MODEL_MLKIT };
typedef struct or_maskgen or_maskgen;
int or_model_type_from_name(const char* path);
or_maskgen* or_maskgen_new(const char* model_path, int width, int height, char* err, size_t errlen);
void or_maskgen_delete(or_maskgen* g);
/* geometry: roidim / in_roidim / out_roidim as x,y,w,h ; model in/out dims */
void or_maskgen_geometry(const or_maskgen* g, int roidim[4], int in_roidim[4], int out_roidim[4],
                         int in_hwc[3], int out_hwc[3]);
/* frame: BGR u8 W x H, stride in bytes.  mask_out: W x H tightly packed (255 = background). */
int or_maskgen_process(or_maskgen* g, const uint8_t* frame_bgr, size_t stride, uint8_t* mask_out);
/* intermediates of the last process() call, for stage-by-stage parity */
const uint8_t* or_maskgen_in_u8(const or_maskgen* g);      /* after resize+BGR2RGB, mh*mw*3 */
const uint8_t* or_maskgen_filtered_u8(const or_maskgen* g);/* after bilateral */
const float* or_maskgen_input_f32(const or_maskgen* g);
const float* or_maskgen_output_f32(const or_maskgen* g);
const uint8_t* or_maskgen_ofinal(const or_maskgen* g);     /* oh*ow IIR state */
or_model* or_maskgen_model(or_maskgen* g);
/* skip the CNN: run only decision+IIR+upsample+blur on a given output tensor (tests) */
int or_maskgen_post_from_output(or_maskgen* g, const float* model_output, uint8_t* mask_out);
/* whole frame: mask -> grab_background resize -> alpha_blend -> (optional) YUYV.
 * bg_raw: BGR u8 bw x bh.  out_rgb: W*H*3 ; out_yuyv may be NULL; out_mask may be NULL. */
int or_composite(or_maskgen* g, const uint8_t* frame_bgr, size_t stride,
                 const uint8_t* bg_raw, int bw, int bh, size_t bstride,
                 uint8_t* out_rgb, uint8_t* out_yuyv, uint8_t* out_mask);

/* the whole main-loop body of app/deepseg.cc:640-681 with its options:
 *   bg_raw NULL  -> the background is a copy of the camera frame (only meaningful with bgblur_k, :652-654)
 *   bgblur_k     -> cv::GaussianBlur(bg, k x k, 0) after grab_background (:657-658); 0 = off
 *   flip_h/flip_v-> cv::flip (:667-673);  out_w x out_h -> cv::resize to the virtual-camera size (:677-679)
 * out_rgb / out_yuyv are out_w x out_h; out_mask stays W x H. */
typedef struct { int bgblur_k, flip_h, flip_v, out_w, out_h; } or_frame_opts;
int or_composite_ex(or_maskgen* g, const uint8_t* frame_bgr, size_t stride,
                    const uint8_t* bg_raw, int bw, int bh, size_t bstride, const or_frame_opts* o,
                    uint8_t* out_rgb, uint8_t* out_yuyv, uint8_t* out_mask);

#ifdef __cplusplus
}
#endif
#endif
