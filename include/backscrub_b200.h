/* include/backscrub_b200.h — C ABI of the B200-native backscrub hot path.
 *
 * This is the drop-in boundary for the reference's per-frame path.  The reference exposes
 * a C++ API (std::string, cv::Mat&) — lib/libbackscrub.h:13-39 — plus two free functions
 * compiled into the app (app/deepseg.cc:87-134) and the background provider
 * (app/background.h:14-23).  Everything below is plain C (pointers + sizes, no C++ or
 * torch types); include/libbackscrub.h and include/background.h are the header-only
 * cv::Mat shims that map the reference's signatures onto it, and INTEGRATION.md shows
 * the binding a maintainer adds.
 *
 * Conventions (same as the reference):
 *   - frames are 8-bit 3-channel, in the channel order the camera delivers (BGR in
 *     backscrub), row pitch given in bytes; W x H fixed at bsb_maskgen_new;
 *   - masks are 8-bit 1-channel W x H, 255 = background, 0 = person;
 *   - errors: NULL / negative return + message via the ondebug callback (or stderr) and
 *     bsb_last_error(); no exceptions cross the boundary;
 *   - one context is used by one thread at a time; contexts are independent (one per
 *     stream / GPU); the library has no global mutable state except the error string.
 *   - There is NO CPU fallback: every entry point that computes requires a CUDA device
 *     and fails loudly without one.
 */
#ifndef BACKSCRUB_B200_H
#define BACKSCRUB_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BSB_API __attribute__((visibility("default")))

typedef struct bsb_ctx bsb_ctx;

/* Optional callbacks, as in lib/libbackscrub.h:21-33.  onprep / oninfer / onmask fire in
 * this order, once per bsb_maskgen_process() call, on the calling thread, after the GPU
 * work of that stage has been *enqueued and completed* (the call synchronises the stage
 * boundary only when the callback is non-NULL). */
typedef void (*bsb_debug_cb)(void* caller_ctx, const char* msg);
typedef void (*bsb_stage_cb)(void* caller_ctx);

/* flags for bsb_maskgen_new_ex */
enum {
  BSB_FLAG_KEEP_TENSORS = 1,  /* keep every intermediate activation (tests / debugging) */
  BSB_FLAG_NO_GRAPH = 2,      /* launch kernels eagerly instead of one CUDA graph per batch */
  BSB_FLAG_TENSOR_CORES = 4,  /* tcgen05 3xTF32 1x1 convs for ANY model (default only for the GEMM-dominated DeepLab / BodyPix graphs) */
  BSB_FLAG_FUSE_BLOCKS = 8,   /* round-1 experiment: one-kernel inverted-residual blocks (bit-exact; superseded by the chain kernel) */
  BSB_FLAG_EXACT = 16         /* fp32 FFMA 1x1 convs everywhere: every activation bit-identical to the CPU oracle (no tensor cores) */
};

/* replaces bs_tensorflow_version() (lib/libbackscrub.h:13, lib/libbackscrub.cc:150):
 * a static string naming the inference runtime. */
BSB_API const char* bsb_version(void);

/* Thread-local text of the last error raised by any entry point ("" if none). */
BSB_API const char* bsb_last_error(void);

/* Number of CUDA devices visible (0 => nothing below can run). */
BSB_API int bsb_device_count(void);

/* replaces bs_maskgen_new (lib/libbackscrub.h:16-34, lib/libbackscrub.cc:161-259).
 * Model family is chosen by file-name substring (body-pix / deeplab / segm_ / selfie),
 * lib/libbackscrub.cc:116-130.  `threads` is accepted and ignored (advisory in the
 * reference too).  Returns NULL after reporting through ondebug. */
BSB_API bsb_ctx* bsb_maskgen_new(const char* modelname, size_t threads, size_t width, size_t height,
                                 bsb_debug_cb ondebug, bsb_stage_cb onprep, bsb_stage_cb oninfer,
                                 bsb_stage_cb onmask, void* caller_ctx);

/* Same, with the GPU-side knobs: CUDA device ordinal, maximum frames per launch
 * (consecutive frames of ONE stream; the IIR advances in order), flags above. */
BSB_API bsb_ctx* bsb_maskgen_new_ex(const char* modelname, size_t width, size_t height, int device,
                                    int max_batch, unsigned flags, bsb_debug_cb ondebug, bsb_stage_cb onprep,
                                    bsb_stage_cb oninfer, bsb_stage_cb onmask, void* caller_ctx);

/* replaces bs_maskgen_delete (lib/libbackscrub.h:37, lib/libbackscrub.cc:261-277); NULL-safe. */
BSB_API void bsb_maskgen_delete(bsb_ctx* ctx);

/* replaces bs_maskgen_process (lib/libbackscrub.h:39, lib/libbackscrub.cc:279-376).
 * frame: host pointer, W x H x 3, `frame_pitch` bytes per row (never written).
 * *mask / *mask_pitch receive a pointer to context-owned HOST storage holding the W x H
 * mask — valid until the next process call on this context (the reference returns a
 * cv::Mat header aliasing ctx.mask the same way, lib/libbackscrub.cc:374).
 * Returns 1 on success, 0 on error (the reference returns bool). */
BSB_API int bsb_maskgen_process(bsb_ctx* ctx, const uint8_t* frame, size_t frame_pitch,
                                const uint8_t** mask, size_t* mask_pitch);

/* ---- background provider (app/background.h:14-23, app/background.cc:178-194) ----
 * Decoding stays on the host (cv::VideoCapture / imread in the shim); the library takes
 * the decoded raw frame and performs grab_background's cv::resize(raw -> W x H) on the
 * GPU.  Call again whenever the decoded frame changes (video backgrounds). */
BSB_API int bsb_set_background(bsb_ctx* ctx, const uint8_t* bg_raw, int bg_w, int bg_h, size_t bg_pitch);
/* the resized background (what grab_background hands back), copied to host.  Until a background is set it is
 * plain green, like `bg` in app/deepseg.cc:603. */
BSB_API int bsb_get_background(bsb_ctx* ctx, uint8_t* out, size_t out_pitch);
/* Animated background (the video branch of app/background.cc:126-176 decodes; :178-194 resizes): `count` decoded
 * images, `frame_stride` bytes apart, are resized into a device-resident ring.  Batch frame b of a composite call
 * blends ring image (cursor + b*advance) % count, and the cursor then moves on by n_frames*advance.  advance 0
 * leaves pacing to the caller (bsb_set_background_cursor), as the reference paces by wall clock. */
BSB_API int bsb_set_background_ring(bsb_ctx* ctx, const uint8_t* frames, int count, int bg_w, int bg_h, size_t bg_pitch,
                                    size_t frame_stride, int advance);
BSB_API int bsb_set_background_cursor(bsb_ctx* ctx, int index);

/* ---- per-frame options of the reference's main loop around alpha_blend (app/deepseg.cc:649-679) ----
 * `-p bgblur:k` (:415-434, :657-658): cv::GaussianBlur(bg, bg, Size(k,k), 0) on the background — the grabbed
 * one, or, when no background was set, a copy of the camera frame (:652-654).  k odd in 1..255 ("strength value
 * must be odd"), 0 = off.  OpenCV's bit-exact 8-bit path. */
BSB_API int bsb_set_bgblur(bsb_ctx* ctx, int ksize);
/* `-H`/`-V` cv::flip of the composited frame (:667-673) and cv::resize to the virtual-camera geometry when it
 * differs from the capture geometry (:677-679); both before the YUYV conversion.  out_w/out_h <= 0 = frame size.
 * After this call `out` buffers are out_w x out_h x 3 and `out_yuyv` out_w x out_h x 2; the mask stays W x H. */
BSB_API int bsb_set_output(bsb_ctx* ctx, int flip_h, int flip_v, int out_w, int out_h);
BSB_API int bsb_output_size(bsb_ctx* ctx, int* out_w, int* out_h);

/* ---- fused per-frame path: one CUDA-graph launch per call ---------------------------
 * mask generation (bs_maskgen_process) + alpha_blend(bg, frame, mask)
 * (app/deepseg.cc:108-134, :661) + convert_rgb_to_yuyv (app/deepseg.cc:87-106, :681).
 * HOST buffers; n_frames consecutive frames of the stream (1 <= n <= max_batch), laid out
 * `frame_stride` bytes apart.  Any of out / out_yuyv / out_mask may be NULL. */
BSB_API int bsb_composite(bsb_ctx* ctx, int n_frames, const uint8_t* frames, size_t frame_pitch, size_t frame_stride,
                          uint8_t* out, size_t out_pitch, size_t out_stride,
                          uint8_t* out_yuyv, size_t yuyv_stride, uint8_t* out_mask, size_t mask_stride);

/* Same with DEVICE pointers (frames already resident in HBM; tightly packed W*3 / W*2 / W
 * rows).  Asynchronous on the context's stream unless `sync` is non-zero. */
BSB_API int bsb_composite_device(bsb_ctx* ctx, int n_frames, const uint8_t* d_frames, size_t frame_stride,
                                 uint8_t* d_out, size_t out_stride, uint8_t* d_yuyv, size_t yuyv_stride,
                                 uint8_t* d_mask, size_t mask_stride, int sync);
/* Camera-format ingest: the same fused call fed with YUYV frames (W*2 bytes per row, tightly packed) — the
 * YUYV -> BGR conversion cv::VideoCapture performs for the reference (CAP_PROP_CONVERT_RGB,
 * app/deepseg.cc:553; cv::COLOR_YUV2BGR_YUYV) runs on the GPU ahead of the pipeline.  HOST buffers. */
BSB_API int bsb_composite_yuyv(bsb_ctx* ctx, int n_frames, const uint8_t* yuyv_frames, size_t in_stride,
                               uint8_t* out, size_t out_stride, uint8_t* out_yuyv, size_t yuyv_stride,
                               uint8_t* out_mask, size_t mask_stride);
/* MJPG camera ingest (app/deepseg.cc:548-553: the capture is opened with FOURCC MJPG and cv::VideoCapture decodes each
 * JPEG to BGR): `n_frames` JPEG images (pointer + size each, HOST memory) are decoded on the GPU by NVJPG (libnvjpeg,
 * loaded at run time; a library decoder for a wire format either side of the path, like FFmpeg is for the reference)
 * straight into the context's BGR frame buffer, then the fused call runs as in bsb_composite.  Decoded pixels may differ
 * from libjpeg's by a few LSB (different IDCT / chroma up-sampling), so bsb_decode_mjpg returns the frame the pipeline saw.
 * Returns 0 with an error message if libnvjpeg is not available. */
BSB_API int bsb_composite_mjpg(bsb_ctx* ctx, int n_frames, const uint8_t* const* jpegs, const size_t* jpeg_sizes,
                               uint8_t* out, size_t out_stride, uint8_t* out_yuyv, size_t yuyv_stride, uint8_t* out_mask, size_t mask_stride);
/* the decode step alone: one JPEG -> W x H x 3 BGR (tightly packed HOST buffer) */
BSB_API int bsb_decode_mjpg(bsb_ctx* ctx, const uint8_t* jpeg, size_t jpeg_size, uint8_t* bgr_out);
/* DEVICE pointers, asynchronous unless `sync`. */
BSB_API int bsb_composite_yuyv_device(bsb_ctx* ctx, int n_frames, const uint8_t* d_yuyv_frames,
                                      uint8_t* d_out, size_t out_stride, uint8_t* d_yuyv, size_t yuyv_stride,
                                      uint8_t* d_mask, size_t mask_stride, int sync);
/* ---- asynchronous mask front-end: the reference's CalcMask (app/deepseg.cc:159-286) ----
 * A worker thread owns the mask-generation context.  set_input_frame clones the frame and wakes the worker (a
 * frame that arrives while the worker is busy replaces the pending one: latest frame wins); get_output_mask copies
 * the newest finished mask ONCE (returns 1) and otherwise leaves `out` untouched and returns 0, so the caller keeps
 * blending with its previous mask exactly like `ai.get_output_mask(mask)` in app/deepseg.cc:641; -1 after a
 * processing error (the reference exits the process there).  timings: waitns, prepns, tfltns, maskns, loopns
 * (app/deepseg.cc:233-237, printed by `-d`). */
typedef struct bsb_calcmask bsb_calcmask;
BSB_API bsb_calcmask* bsb_calcmask_new(const char* modelname, size_t threads, size_t width, size_t height, int device);
BSB_API void bsb_calcmask_delete(bsb_calcmask* cm);
BSB_API int bsb_calcmask_set_input_frame(bsb_calcmask* cm, const uint8_t* frame, size_t frame_pitch);
BSB_API int bsb_calcmask_get_output_mask(bsb_calcmask* cm, uint8_t* out, size_t out_pitch);
BSB_API int bsb_calcmask_timings(bsb_calcmask* cm, long ns[5]);
/* frames the worker has finished; the 1-based index (count of set_input_frame calls) of the frame behind the newest mask */
BSB_API long bsb_calcmask_frames_done(bsb_calcmask* cm);
BSB_API long bsb_calcmask_mask_serial(bsb_calcmask* cm);

/* ---- background provider object: the reference's background_t (app/background.cc:13-202) ----
 * still image: the decoded image is kept; grab returns it resized and frame number 1.
 * video: a reader thread pulls decoded frames through `read` (1 = a frame, 0 = end of stream; the pixels must stay
 * valid until the next call), publishes the latest one under a mutex, counts frames, paces itself to `fps`
 * (all sources play in real time, app/background.cc:84-92) and at end of stream calls `rewind` (1 = repositioned
 * at frame 0) and starts over with frame number 0, or stops when the source cannot rewind (:93-102).
 * `first`: the frame load_background() decoded while probing (may be NULL); start_frame: 0 after a successful
 * rewind of the probe reads, else 2 (:146-149).  grab: cv::resize(latest -> width x height) on the GPU, returns the
 * frame number (1 for stills, may wrap to 0), -1 on error (app/background.cc:178-194).  grab_into: the same resize
 * straight into a context's device-resident background (no host round trip).  thumbnail: 160-pixel-wide copy of the
 * latest frame, refreshed by the reader when debug > 1 (:65-78); returns 0 and the size (0 x 0 if none yet). */
typedef struct bsb_background bsb_background;
typedef int (*bsb_bg_read_cb)(void* user, const uint8_t** data, int* width, int* height, size_t* pitch);
typedef int (*bsb_bg_rewind_cb)(void* user);
BSB_API bsb_background* bsb_background_new_still(int device, const uint8_t* raw, int width, int height, size_t pitch, int debug);
BSB_API bsb_background* bsb_background_new_video(int device, double fps, int start_frame, bsb_bg_read_cb read, bsb_bg_rewind_cb rewind,
                                                 void* user, const uint8_t* first, int width, int height, size_t pitch, int debug);
BSB_API void bsb_background_delete(bsb_background* bg);
BSB_API int bsb_background_grab(bsb_background* bg, int width, int height, uint8_t* out, size_t out_pitch);
BSB_API int bsb_background_grab_into(bsb_background* bg, bsb_ctx* ctx);
BSB_API int bsb_background_thumbnail(bsb_background* bg, uint8_t* out, size_t capacity, int* width, int* height);
BSB_API int bsb_background_frame(bsb_background* bg);
BSB_API int bsb_background_running(bsb_background* bg);

BSB_API int bsb_synchronize(bsb_ctx* ctx);
/* the context's cudaStream_t (for event timing on the launching stream) */
BSB_API void* bsb_stream(bsb_ctx* ctx);

/* ---- stand-alone stages (stage-level parity; HOST buffers, tightly packed) -----------
 * app/deepseg.cc:108-134 */
BSB_API int bsb_alpha_blend(int device, const uint8_t* srca, const uint8_t* srcb, const uint8_t* mask,
                            uint8_t* out, size_t npix);
/* app/deepseg.cc:87-106 */
BSB_API int bsb_convert_rgb_to_yuyv(int device, const uint8_t* rgb, uint8_t* yuyv, int width, int height);
/* cv::cvtColor(COLOR_YUV2BGR_YUYV) (camera ingest, app/deepseg.cc:553,725) */
BSB_API int bsb_convert_yuyv_to_bgr(int device, const uint8_t* yuyv, uint8_t* bgr, int width, int height);
/* cv::GaussianBlur(src, dst, Size(k,k), 0) 8UC3 (app/deepseg.cc:657-658) and its 8.8 fixed-point taps (host only) */
BSB_API int bsb_gaussian_blur(int device, const uint8_t* src, uint8_t* dst, int width, int height, int ksize);
BSB_API int bsb_gaussian_taps(int ksize, int* taps);
/* cv::flip(src, dst, code) 8UC3 (app/deepseg.cc:667-673) */
BSB_API int bsb_flip(int device, const uint8_t* src, uint8_t* dst, int width, int height, int flip_h, int flip_v);
/* cv::resize(src, dst, Size(dw, dh)) 8UC3 (app/background.cc:178-194) */
BSB_API int bsb_resize_u8c3(int device, const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);

/* 1x1 convolution stage on HOST buffers: out[M][N] = act(A[M][K] * W[N][K]^T + bias).  variant 0: exact FFMA
 * kernels, chosen by shape (k-ascending fmaf, bit-identical to the oracle); 1: tcgen05 3xTF32 tensor-core kernel
 * (needs K % 4 == 0); 2 / 3: force the classic / the register-tiled exact FFMA kernel (same bits as 0).
 * act is a TFLite fused-activation code (0 none, 1 relu, 3 relu6). */
BSB_API int bsb_pointwise(int device, int variant, int M, int K, int N, const float* A, const float* W,
                          const float* bias, int act, float* out);
/* ms per launch of the pointwise kernel `variant` (0, 2, 3, ...: exact FFMA kernels; 1: the tensor-core kernel selected by the
 * "tc_variant" tuning switch) on an M x K x N problem (CUDA events, zero data) */
BSB_API double bsb_time_pointwise(int device, int variant, int M, int K, int N, int iters);

/* ---- introspection (tests, bench) -------------------------------------------------- */
/* geometry: roidim / in_roidim / out_roidim as x,y,w,h (lib/libbackscrub.cc:234-246);
 * model input / output dims as h,w,c */
/* frame size the context was created for */
BSB_API int bsb_frame_size(bsb_ctx* ctx, int* width, int* height);
BSB_API int bsb_geometry(bsb_ctx* ctx, int roidim[4], int in_roidim[4], int out_roidim[4], int in_hwc[3], int out_hwc[3]);
/* run only the CNN on a caller-provided fp32 NHWC input batch (host), output to host */
BSB_API int bsb_infer(bsb_ctx* ctx, int n_frames, const float* input, float* output);
/* copy an intermediate activation (frame 0) to host; needs BSB_FLAG_KEEP_TENSORS.
 * Returns the element count, 0 if the tensor was folded away, -1 on error. */
BSB_API long bsb_get_tensor(bsb_ctx* ctx, int tensor_index, float* out, long capacity);
/* stage buffers of the last call, frame `frame`: which = 0 model-sized RGB u8 (after resize +
 * BGR2RGB), 1 bilateral-filtered u8, 2 ofinal (IIR state after that frame) */
BSB_API long bsb_get_stage_u8(bsb_ctx* ctx, int which, int frame, uint8_t* out, long capacity);
/* reset the temporal IIR state to zero (start of a new stream) */
BSB_API int bsb_reset_state(bsb_ctx* ctx);
/* kernel launches per n-frame call (graph nodes) and launches issued so far by this process */
BSB_API int bsb_launches_per_call(bsb_ctx* ctx, int n_frames);
BSB_API long bsb_total_launches(void);
/* Average device time (ms, CUDA events on the context's stream) of one stage of the per-call
 * kernel sequence, run `iters` times back to back on the context-owned device buffers:
 * stage 0 = pre-proc (resize+bilateral+normalise), 1 = CNN, 2 = decision+IIR,
 * 3 = post (mask upsample + blur + blend + YUYV), 4 = whole call.  Returns < 0 on error.
 * Measurement aid for bench.py's roofline block; it advances the IIR state. */
BSB_API double bsb_time_stage(bsb_ctx* ctx, int stage, int n_frames, int iters);
/* 1 if the last fused call read camera YUYV frames in place (pre-processing and post kernels convert per tap / per tile),
 * 0 if a BGR frame was materialised first (k_yuyv_to_bgr) or the input was BGR */
BSB_API int bsb_yuyv_native(bsb_ctx* ctx);
/* 1 if at least one 1x1 conv of this context runs on the tensor cores */
BSB_API int bsb_uses_tensor_cores(bsb_ctx* ctx);
/* algorithmic FLOPs of one CNN frame (2*MAC) */
BSB_API double bsb_model_flops(bsb_ctx* ctx);
/* Process-wide measurement switches of the kernel launchers (A/B runs in bench.py / tools/): they select between
 * bit-identical kernel variants / schedules and never change results.  Names (defaults in csrc/kernels.h, struct Tuning):
 * planner fusions "cnn_chain", "pool_merge", "up_pw", "up_staged", "head", "pw_dws2", "stem_pw", "stem_x2", "dec_up", "sub_batch_mb";
 * kernel variants "pw_variant", "dw_px", "dw_plane", "dw_plane_cs", "epi_static", "dec_par", "tc_variant", "tc_min_k", "tc_mask_hi";
 * post stage "post_tma", "post_tile" (0 = by frame size), "post_wide", "post_l1"; host-buffer calls "e2e_chunk" (frames per
 * overlapped chunk of bsb_composite_yuyv, 0 = serial).  Set them before creating contexts (plans and captured CUDA graphs
 * keep the variant they were built with).  Returns 1, or 0 for an unknown name. */
BSB_API int bsb_set_tuning(const char* name, int value);

#ifdef __cplusplus
}
#endif
#endif
