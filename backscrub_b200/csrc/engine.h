// backscrub_b200/csrc/engine.h — per-stream context: model plan, HBM layout, CUDA graph.
//
// Host-side equivalent of backscrub_ctx_t + bs_maskgen_new/process
// (lib/libbackscrub.cc:28-54,161-376), re-designed for one GPU per process:
//   * the .tflite graph is planned once into a short list of fused kernel launches
//     (unary ops, residual ADDs, SE channel scales folded into producers/consumers),
//   * activations live in one liveness-packed arena sized for `max_batch` frames so a
//     whole batch stays L2-resident between layers,
//   * a whole call (pre-proc -> CNN -> decision/IIR -> mask upsample/blur/blend/YUYV) is
//     captured once per batch size into a CUDA graph and replayed with one launch.
#pragma once
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "kernels.h"
#include "tflite_model.h"

namespace bsb {

struct TensorInfo {
  int h = 1, w = 1, c = 1, ld = 1;
  size_t frame_elems = 0;     // h*w*ld floats per frame
  size_t offset = 0;          // float offset inside the arena
  int alias_parent = -1;      // >= 0: lives inside that tensor's buffer (in-place CONCATENATION)
  int alias_off = 0;          // channel offset inside the parent
  bool materialized = false;
  int first_def = 1 << 30, last_use = -1;
};

struct Step {
  enum Kind { CONV, PW, DW, POOL, RESIZE, TCONV, ELT, COPY, BLOCK, CHAIN, HEAD } kind = ELT;
  int block = -1;             // BLOCK: index into Engine::blocks_ (the four fused sub-steps); CHAIN: index into Engine::chains_
  int op_index = -1;
  int in = -1, in2 = -1, out = -1, scale = -1, in_add = -1, residual = -1;
  size_t w_off = 0, b_off = 0; bool has_bias = false;
  int act1 = 0, act2 = 0, act3 = 0;
  int K = 0, N = 0, n4 = 0;
  int kh = 1, kw = 1, sh = 1, sw = 1, dh = 1, dw = 1, pt = 0, pl = 0;
  int elt_mode = 0;
  bool align_corners = false, half_pixel = false;
  int copy_off = 0;
  int up_from = -1;           // PW: >= 0 -> the input is RESIZE_BILINEAR(up_from), interpolated on the fly (k_upsample_pw)
  bool use_tc = false; size_t tc_hi_off = 0, tc_lo_off = 0; int kpad = 0, npad = 0;   // tensor-core pointwise
  // POOL: optional fused fully-connected chain (squeeze-excite) after the global average pool
  struct Fc { size_t w_off = 0, b_off = 0; bool has_bias = false; int K = 0, N = 0, n4 = 0, act1 = 0, act2 = 0; };
  int n_fc = 0;
  Fc fc[2];
  int launches() const { return 1; }
};

struct HostResizeTab {
  std::vector<int> xofs, yofs0, yofs1;
  std::vector<short> xw, yw;
  bool area2x2 = false;
  int sw = 0;
};
// OpenCV INTER_LINEAR 8-bit coefficient tables for sw x sh -> dw x dh
HostResizeTab build_resize_tab(int sw, int sh, int dw, int dh);

struct DevResizeTab { void* blob = nullptr; ResizeTab tab{}; bool area2x2 = false; };
// tables -> device memory of the current device (frees a previous blob)
bool upload_resize_tab(const HostResizeTab& h, DevResizeTab* d, std::string* err);

struct Callbacks {
  void (*ondebug)(void*, const char*) = nullptr;
  void (*onprep)(void*) = nullptr;
  void (*oninfer)(void*) = nullptr;
  void (*onmask)(void*) = nullptr;
  void* caller_ctx = nullptr;
};

class Engine {
 public:
  static Engine* create(const std::string& model_path, int width, int height, int device, int max_batch,
                        unsigned flags, const Callbacks& cb, std::string* err);
  ~Engine();

  // frames on device: n frames, row pitch / frame stride in bytes.  Any output may be null.
  bool run(int n, const uint8_t* d_frames, size_t pitch, size_t stride, uint8_t* d_out, size_t out_stride,
           uint8_t* d_yuyv, size_t yuyv_stride, uint8_t* d_mask, size_t mask_stride, bool use_callbacks, std::string* err,
           const uint8_t* d_yuyv_in = nullptr);
  // same, from camera YUYV frames (tightly packed W*2 rows): converted to BGR into the context staging first
  bool run_yuyv(int n, const uint8_t* d_yuyv_in, uint8_t* d_out, size_t out_stride, uint8_t* d_yuyv, size_t yuyv_stride,
                uint8_t* d_mask, size_t mask_stride, std::string* err);
  bool infer(int n, const float* h_in, float* h_out, std::string* err);
  bool set_background(const uint8_t* bg_raw, int bw, int bh, size_t pitch, std::string* err) {
    return set_background_ring(bg_raw, 1, bw, bh, pitch, 0, 0, err);
  }
  // animated background (app/background.cc:126-176 decodes, :178-194 resizes): `count` decoded images are
  // resized into a device ring; batch frame b of a call blends ring image (cursor + b*advance) % count and the
  // cursor then moves on by n*advance (advance 0: the caller moves it with set_background_cursor)
  bool set_background_ring(const uint8_t* frames, int count, int bw, int bh, size_t pitch, size_t frame_stride, int advance,
                           std::string* err);
  bool set_background_cursor(int index, std::string* err);
  // app/deepseg.cc:657-658 `-p bgblur:k`: Gaussian-blur the background (the grabbed one, else a copy of the
  // camera frame); k = 0 switches it off
  bool set_bgblur(int k, std::string* err);
  // app/deepseg.cc:667-679: cv::flip and cv::resize to the virtual-camera size before the YUYV conversion
  bool set_output(bool flip_h, bool flip_v, int out_w, int out_h, std::string* err);
  bool sync(std::string* err);
  bool reset_state(std::string* err);
  double time_stage(int stage, int n, int iters, std::string* err);

  // accessors
  int W() const { return W_; }
  int H() const { return H_; }
  int out_w() const { return out_w_; }
  int out_h() const { return out_h_; }
  int bgblur() const { return bgblur_k_; }
  int bg_count() const { return bg_count_; }
  int max_batch() const { return max_batch_; }
  int device() const { return device_; }
  cudaStream_t stream() const { return stream_; }
  const int* roidim() const { return roidim_; }
  const int* in_roidim() const { return in_roidim_; }
  const int* out_roidim() const { return out_roidim_; }
  void in_hwc(int* v) const { v[0] = mh_; v[1] = mw_; v[2] = 3; }
  void out_hwc(int* v) const { v[0] = oh_; v[1] = ow_; v[2] = oc_; }
  double flops() const { return flops_; }
  int launches_per_call() const {
    if (last_call_launches_ > 0) return last_call_launches_;      // counted while the last call was enqueued / captured
    int n = 4;
    for (const Step& s : steps_) n += s.launches();
    if (bgblur_k_ && !has_bg_) n += 2;
    if (bg_count_ > 1 && bg_advance_) n += 1;
    const bool resized = out_w_ != W_ || out_h_ != H_;
    if (flip_h_ || flip_v_ || resized) n += (flip_h_ || flip_v_ ? 1 : 0) + (resized ? 1 : 0) + 1;   // + stand-alone YUYV
    return n;
  }
  long get_tensor(int t, float* out, long cap, std::string* err);
  long get_stage_u8(int which, int frame, uint8_t* out, long cap, std::string* err);

  // device staging owned by the context (host-buffer API)
  uint8_t* d_frames() const { return d_frames_; }
  uint8_t* d_out() const { return d_out_; }
  uint8_t* d_yuyv() const { return d_yuyv_; }
  uint8_t* d_mask() const { return d_mask_; }
  uint8_t* d_bg() const { return d_bg_; }          // resized, unblurred (what grab_background returns)
  uint8_t* d_yuyv_in() const { return d_yuyv_in_; }
  uint8_t* h_mask() const { return h_mask_; }
  bool has_background() const { return has_bg_; }
  bool last_native() const { return last_native_; }
  bool uses_tensor_cores() const { return uses_tc_; }

 private:
  Engine() = default;
  bool plan(std::string* err);
  bool upload(std::string* err);
  void enqueue_pre(int n, const uint8_t* d_frames, size_t pitch, size_t stride, const uint8_t* d_yuyv_in = nullptr);
  PostArgs post_args(int n, const uint8_t* d_frames, size_t pitch, size_t stride, uint8_t* d_out, size_t out_stride,
                     uint8_t* d_yuyv, size_t yuyv_stride, uint8_t* d_mask, size_t mask_stride, const uint8_t* d_yuyv_in) const;
  bool yuyv_native(int n, const uint8_t* d_frames, size_t pitch, size_t stride, uint8_t* d_out, size_t out_stride,
                   uint8_t* d_yuyv, size_t yuyv_stride, uint8_t* d_mask, size_t mask_stride, const uint8_t* d_yuyv_in) const;
  void enqueue_all(int n, const uint8_t* d_frames, size_t pitch, size_t stride, uint8_t* d_out, size_t out_stride,
                   uint8_t* d_yuyv, size_t yuyv_stride, uint8_t* d_mask, size_t mask_stride, const uint8_t* d_yuyv_in,
                   bool native, bool sync_cbs);
  bool refresh_bg_yuyv(std::string* err);
  void enqueue_cnn(int n, bool from_u8);
  void enqueue_decision(int n);
  void enqueue_post(int n, const uint8_t* d_frames, size_t pitch, size_t stride, uint8_t* d_out, size_t out_stride,
                    uint8_t* d_yuyv, size_t yuyv_stride, uint8_t* d_mask, size_t mask_stride, const uint8_t* d_yuyv_in = nullptr);
  // base of tensor t for the frames [frame_off_, ...) of the batch (frame_off_ != 0 only inside a sub-batched segment)
  float* tptr(int t) const {
    const TensorInfo& I = tinfo_[t];
    if (I.alias_parent >= 0) {
      const TensorInfo& P = tinfo_[I.alias_parent];
      return arena_ + P.offset + (size_t)frame_off_ * P.frame_elems + I.alias_off;
    }
    return arena_ + I.offset + (size_t)frame_off_ * I.frame_elems;
  }
  void run_step(size_t si, int n, bool from_u8, bool* first, bool* skip_next);
  void find_segments();

  Graph g_;
  int model_type_ = 0;
  float scaling_ = 0.f, offset_ = 0.f;
  int W_ = 0, H_ = 0, device_ = 0, max_batch_ = 1;
  unsigned flags_ = 0;
  Callbacks cb_;
  int mh_ = 0, mw_ = 0, oh_ = 0, ow_ = 0, oc_ = 0;
  int roidim_[4] = {0, 0, 0, 0}, in_roidim_[4] = {0, 0, 0, 0}, out_roidim_[4] = {0, 0, 0, 0};
  double flops_ = 0;

  std::vector<TensorInfo> tinfo_;
  std::vector<Step> steps_;
  // Optional (tuning().sub_batch_mb, off by default): a run of per-frame steps whose tensors are too large for the L2 at
  // the full batch is executed a few frames at a time, so that each step finds its input where the previous one left it
  // instead of streaming it through HBM.  Measured on DeepLab 720p x 32 (run r2t): the smaller launches cost more than
  // the HBM traffic they save (15.3 k frames/s off, 14.4 k with 64 MB groups, 13.4 k with 16 MB groups) — kept as a
  // bit-exact, tested switch.  seg_len_[si] > 0: steps [si, si + seg_len_[si]) form a segment whose largest tensor has
  // seg_frame_bytes_[si] bytes per frame.
  std::vector<int> seg_len_;
  std::vector<size_t> seg_frame_bytes_;
  int frame_off_ = 0;
  struct FusedBlock { Step expand, dw, pool, project; };
  std::vector<FusedBlock> blocks_;
  // the low-resolution middle of the graph run by one kernel (kernels_chain.cu): the original steps in order, each
  // tagged with the chain op it became
  struct ChainPlan { int h = 0, w = 0; std::vector<Step> seq; std::vector<int> types; ChainOp* d_ops = nullptr; int n_ops = 0; };
  std::vector<ChainPlan> chains_;
  bool detect_chain();
  // decoder stage run by one kernel (k_head): 1x1 conv -> depthwise 3x3 + residual [-> transposed conv]
  struct HeadPlan { Step p, d, t; bool has_t = false; bool s2 = false; };   // s2: 1x1 + stride-2 depthwise (k_pw_dws2), else a decoder stage (k_head)
  std::vector<HeadPlan> heads_;
  void detect_heads();
  void detect_pw_dws2();
  std::vector<float> wblob_h_;
  size_t arena_elems_ = 0;
  bool tc_enabled_ = false, uses_tc_ = false;   // tensor-core 1x1 convs allowed / actually planned for at least one layer
  bool stem_u8_ok_ = false;          // step 0 is a 3->16 dense conv that is the only reader of the graph input
  int dec_up_step_ = -1;             // DeepLab: index of the final RESIZE_BILINEAR step the decision kernel performs itself (pipeline calls)
  bool stem_pw_ok_ = false;          // ... and step 1 is a plain 16 -> 16 1x1 conv of its output (runs inside the stem kernel)

  // device memory
  cudaStream_t stream_ = nullptr;
  float* wblob_ = nullptr;
  float* arena_ = nullptr;
  float* lut_ = nullptr;             // color_w[768] + space_w[16]
  float* rowsum_ = nullptr;          // scratch of the global-average-pool row sums
  size_t rowsum_elems_ = 0;
  unsigned* pool_counters_ = nullptr; // [B] arrival counters of the one-launch pool + SE kernel (zero between launches)
  uint8_t* in_u8_ = nullptr;         // [B][mh][mw][3] zero outside in_roidim
  uint8_t* filt_u8_ = nullptr;       // [B][mh][mw][3] (KEEP_TENSORS only)
  uint8_t* state_ = nullptr;         // [oh*ow] IIR state
  uint8_t* ofinal_ = nullptr;        // [B][oh][opitch_]
  int opitch_ = 0;                   // ofinal row pitch (ow rounded up to 16 bytes: TMA-addressable)
  uint8_t* d_frames_ = nullptr, *d_out_ = nullptr, *d_yuyv_ = nullptr, *d_mask_ = nullptr, *d_bg_ = nullptr, *d_bg_raw_ = nullptr, *d_yuyv_in_ = nullptr;
  size_t bg_raw_cap_ = 0;
  uint8_t* h_mask_ = nullptr;        // pinned host W*H
  bool has_bg_ = false;
  DevResizeTab tab_in_, tab_up_, tab_bg_, tab_out_;
  int bg_w_ = 0, bg_h_ = 0;

  // optional stages around the blend (app/deepseg.cc:649-679)
  bool ensure(void** p, size_t* cap, size_t need, std::string* err);   // grow-only device buffer; drops the graphs
  bool refresh_bg_blur(std::string* err);
  void drop_graphs();
  int bg_count_ = 1, bg_advance_ = 0;
  size_t bg_cap_ = 0;
  int* d_bg_cursor_ = nullptr;
  int bgblur_k_ = 0;
  GaussTaps taps_{};
  uint8_t* d_bg_eff_ = nullptr;      // blurred copy of the background ring
  uint8_t* d_bg_frames_ = nullptr;   // [B][H][W][3] blurred camera frames (bgblur without a background)
  uint16_t* d_gauss_tmp_ = nullptr;  // [B][H][W*3] row sums
  size_t bg_eff_cap_ = 0, bg_frames_cap_ = 0, gauss_tmp_cap_ = 0;
  bool flip_h_ = false, flip_v_ = false;
  int out_w_ = 0, out_h_ = 0;
  uint8_t* d_stage_a_ = nullptr, *d_stage_b_ = nullptr, *d_stage_c_ = nullptr;
  int2* d_tile_geo_ = nullptr;       // k_post_tma patch geometry per tile row | per 64-wide tile column | per 128-wide tile column
  int geo_nty_ = 0, geo_ntx64_ = 0;
  uint8_t* d_bg_yuyv_ = nullptr;     // YUYV of the effective background (ring): all-background tiles are copies
  size_t bg_yuyv_cap_ = 0;
  bool bg_yuyv_valid_ = false;
  bool last_native_ = false;         // the last call read camera YUYV in place (no BGR frame was materialised)
  int last_call_launches_ = 0;
  size_t stage_a_cap_ = 0, stage_b_cap_ = 0, stage_c_cap_ = 0, out_cap_ = 0, yuyv_cap_ = 0;

  // everything a captured launch sequence bakes into its kernel arguments
  struct GraphKey {
    int n; const void* f; size_t pitch, stride; const void* o; const void* y; const void* m; const void* yin;
    size_t ostride, ystride, mstride;
    bool operator<(const GraphKey& r) const {
      return std::tie(n, f, pitch, stride, o, y, m, yin, ostride, ystride, mstride) <
             std::tie(r.n, r.f, r.pitch, r.stride, r.o, r.y, r.m, r.yin, r.ostride, r.ystride, r.mstride);
    }
  };
#ifndef BSB_EMU
  std::map<GraphKey, cudaGraphExec_t> graphs_;
#endif
};

}  // namespace bsb
