// backscrub_b200/csrc/frontend.cu — host-side callers of the hot path, native C++ behind the C ABI:
//
//   * bsb_calcmask_*   — the reference's `CalcMask` (app/deepseg.cc:159-286): a worker thread that owns the
//     mask-generation context, double-buffered input frames, double-buffered masks, "latest frame wins" /
//     "a mask is handed out once" semantics, and the per-stage nanosecond counters `-d` prints
//     (app/deepseg.cc:702-718).
//   * bsb_background_* — the reference's background provider (app/background.cc:13-202): still image or a
//     paced, looping reader thread over a frame source, the latest decoded frame under a mutex, the frame
//     counter grab_background() returns, thumbnails; the per-grab cv::resize (:178-194) runs on the GPU.
//
// Decoding itself (cv::VideoCapture / cv::imread) is I/O and stays with the application: the video provider pulls
// frames through two callbacks, which backscrub_b200/shim/background_shim.cc implements over cv::VideoCapture.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/backscrub_b200.h"
#include "engine.h"

namespace {

using clk = std::chrono::steady_clock;
long ns_since(clk::time_point t0) { return (long)std::chrono::duration_cast<std::chrono::nanoseconds>(clk::now() - t0).count(); }

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// CalcMask
// ---------------------------------------------------------------------------------------------------------
struct bsb_calcmask {
  bsb_ctx* ctx = nullptr;
  int W = 0, H = 0;
  // buffers (app/deepseg.cc:166-174): the producer writes frame_next, the worker swaps and reads frame_current
  std::vector<uint8_t> frame[2], mask[2];
  int frame_next = 0, frame_current = 1, mask_current = 0, mask_out = 1;
  std::mutex lock_frame, lock_mask;
  std::condition_variable cond_new_frame;
  bool new_frame = false;
  std::atomic<bool> new_mask{false};
  std::atomic<bool> running{true};
  std::atomic<bool> failed{false};
  std::atomic<long> frames_done{0};
  std::atomic<long> frames_set{0};
  std::atomic<long> mask_serial{0};          // serial number (frames_set at swap time) of the frame behind mask_out
  long cur_serial = 0;
  std::thread thread;
  // timing (app/deepseg.cc:233-237, written by the worker and its callbacks)
  clk::time_point t0;
  std::atomic<long> waitns{0}, prepns{0}, tfltns{0}, maskns{0}, loopns{0};

  static void onprep(void* c) { auto* s = static_cast<bsb_calcmask*>(c); s->prepns = ns_since(s->t0); s->t0 = clk::now(); }
  static void oninfer(void* c) { auto* s = static_cast<bsb_calcmask*>(c); s->tfltns = ns_since(s->t0); s->t0 = clk::now(); }
  static void onmask(void* c) { auto* s = static_cast<bsb_calcmask*>(c); s->maskns = ns_since(s->t0); s->t0 = clk::now(); }

  void run() {
    while (running.load()) {
      const clk::time_point tloop = clk::now();
      t0 = tloop;
      {
        std::unique_lock<std::mutex> hold(lock_frame);
        cond_new_frame.wait(hold, [&] { return new_frame; });
        new_frame = false;
        std::swap(frame_next, frame_current);
        cur_serial = frames_set.load();
      }
      if (!running.load()) break;
      waitns = ns_since(t0);
      t0 = clk::now();
      const uint8_t* mptr = nullptr; size_t mpitch = 0;
      if (!bsb_maskgen_process(ctx, frame[frame_current].data(), (size_t)W * 3, &mptr, &mpitch)) {
        // the reference prints "failed to process video frame" and exits the process (app/deepseg.cc:204-206);
        // a library reports instead: every later get_output_mask returns -1
        failed = true;
        break;
      }
      std::vector<uint8_t>& m = mask[mask_current];
      for (int y = 0; y < H; ++y) std::memcpy(m.data() + (size_t)y * W, mptr + (size_t)y * mpitch, (size_t)W);
      {
        std::lock_guard<std::mutex> hold(lock_mask);
        std::swap(mask_out, mask_current);
        mask_serial = cur_serial;
        new_mask = true;
      }
      frames_done.fetch_add(1);
      loopns = ns_since(tloop);
    }
  }
};

extern "C" {

bsb_calcmask* bsb_calcmask_new(const char* modelname, size_t threads, size_t width, size_t height, int device) {
  bsb_calcmask* c = new bsb_calcmask();
  (void)threads;
  // callbacks with `this` as the caller context, exactly like the CalcMask constructor (app/deepseg.cc:246)
  c->ctx = bsb_maskgen_new_ex(modelname, width, height, device, 1, 0, nullptr, bsb_calcmask::onprep, bsb_calcmask::oninfer, bsb_calcmask::onmask, c);
  if (!c->ctx) { delete c; return nullptr; }       // the reference throws "Could not create mask context"
  c->W = (int)width; c->H = (int)height;
  for (int i = 0; i < 2; ++i) { c->frame[i].assign(width * height * 3, 0); c->mask[i].assign(width * height, 255); }
  c->thread = std::thread(&bsb_calcmask::run, c);
  return c;
}

void bsb_calcmask_delete(bsb_calcmask* c) {
  if (!c) return;
  c->running = false;
  {
    std::lock_guard<std::mutex> hold(c->lock_frame);
    c->new_frame = true;                              // wake the worker (app/deepseg.cc:263-266)
  }
  c->cond_new_frame.notify_all();
  if (c->thread.joinable()) c->thread.join();
  bsb_maskgen_delete(c->ctx);
  delete c;
}

int bsb_calcmask_set_input_frame(bsb_calcmask* c, const uint8_t* frame, size_t pitch) {
  if (!c || !frame || pitch < (size_t)c->W * 3) return 0;
  {
    std::lock_guard<std::mutex> hold(c->lock_frame);
    std::vector<uint8_t>& f = c->frame[c->frame_next];          // `*frame_next = frame.clone()` (app/deepseg.cc:274)
    for (int y = 0; y < c->H; ++y) std::memcpy(f.data() + (size_t)y * c->W * 3, frame + (size_t)y * pitch, (size_t)c->W * 3);
    c->frames_set.fetch_add(1);
    c->new_frame = true;
  }
  c->cond_new_frame.notify_all();
  return 1;
}

int bsb_calcmask_get_output_mask(bsb_calcmask* c, uint8_t* out, size_t pitch) {
  if (!c || !out || pitch < (size_t)c->W) return -1;
  if (c->failed.load()) return -1;
  if (!c->new_mask.load()) return 0;                              // `out` keeps the caller's previous mask (app/deepseg.cc:279)
  std::lock_guard<std::mutex> hold(c->lock_mask);
  const std::vector<uint8_t>& m = c->mask[c->mask_out];
  for (int y = 0; y < c->H; ++y) std::memcpy(out + (size_t)y * pitch, m.data() + (size_t)y * c->W, (size_t)c->W);
  c->new_mask = false;
  return 1;
}

int bsb_calcmask_timings(bsb_calcmask* c, long ns[5]) {
  if (!c || !ns) return 0;
  ns[0] = c->waitns; ns[1] = c->prepns; ns[2] = c->tfltns; ns[3] = c->maskns; ns[4] = c->loopns;
  return 1;
}

long bsb_calcmask_frames_done(bsb_calcmask* c) { return c ? c->frames_done.load() : -1; }
long bsb_calcmask_mask_serial(bsb_calcmask* c) { return c ? c->mask_serial.load() : -1; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// background provider
// ---------------------------------------------------------------------------------------------------------
struct bsb_background {
  int device = 0, debug = 0;
  bool video = false;
  std::atomic<bool> run{false};
  double fps = 0.0;
  int frame = 0;                                   // frames delivered since the last loop (guarded by rawmux)
  std::vector<uint8_t> raw; int rw = 0, rh = 0;    // latest decoded frame, tightly packed BGR
  std::mutex rawmux;
  std::vector<uint8_t> thumb; int tw = 0, th = 0;
  std::mutex thumbmux;
  bsb_bg_read_cb read = nullptr; bsb_bg_rewind_cb rewind = nullptr; void* user = nullptr;
  std::thread thread;
  // GPU resize state
  cudaStream_t stream = nullptr;
  uint8_t* d_raw = nullptr; size_t raw_cap = 0;
  uint8_t* d_out = nullptr; size_t out_cap = 0;
  bsb::DevResizeTab tab; int tab_sw = 0, tab_sh = 0, tab_dw = 0, tab_dh = 0;
  std::mutex gpumux;
  std::string err;

  bool store(const uint8_t* data, int w, int h, size_t pitch) {
    if (!data || w <= 0 || h <= 0 || pitch < (size_t)w * 3) return false;
    raw.resize((size_t)w * h * 3);
    for (int y = 0; y < h; ++y) std::memcpy(raw.data() + (size_t)y * w * 3, data + (size_t)y * pitch, (size_t)w * 3);
    rw = w; rh = h;
    return true;
  }

  // cv::resize(src, dst, Size(dw, dh)) on the GPU (bit-exact OpenCV INTER_LINEAR / the 2x2 INTER_AREA special case)
  bool gpu_resize(const uint8_t* src, int sw, int sh, uint8_t* out, size_t out_pitch, int dw, int dh) {
    std::lock_guard<std::mutex> hold(gpumux);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { err = "no CUDA device available (this library has no CPU path)"; return false; }
    if (cudaSetDevice(device) != cudaSuccess) { err = "invalid CUDA device"; return false; }
    if (!stream && cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) { err = "cudaStreamCreate failed"; return false; }
    const size_t sb = (size_t)sw * sh * 3, db = (size_t)dw * dh * 3;
    if (sb > raw_cap) { if (d_raw) cudaFree(d_raw); d_raw = nullptr; raw_cap = 0; if (cudaMalloc((void**)&d_raw, sb) != cudaSuccess) { err = "cudaMalloc failed"; return false; } raw_cap = sb; }
    if (db > out_cap) { if (d_out) cudaFree(d_out); d_out = nullptr; out_cap = 0; if (cudaMalloc((void**)&d_out, db) != cudaSuccess) { err = "cudaMalloc failed"; return false; } out_cap = db; }
    if (sw != tab_sw || sh != tab_sh || dw != tab_dw || dh != tab_dh) {
      cudaStreamSynchronize(stream);
      if (!bsb::upload_resize_tab(bsb::build_resize_tab(sw, sh, dw, dh), &tab, &err)) return false;
      tab_sw = sw; tab_sh = sh; tab_dw = dw; tab_dh = dh;
    }
    if (cudaMemcpyAsync(d_raw, src, sb, cudaMemcpyHostToDevice, stream) != cudaSuccess) { err = "H2D copy failed"; return false; }
    bsb::launch_resize_u8c3(stream, d_raw, sw, sh, (size_t)sw * 3, d_out, dw, dh, (size_t)dw * 3, tab.tab, tab.area2x2);
    if (cudaMemcpy2DAsync(out, out_pitch, d_out, (size_t)dw * 3, (size_t)dw * 3, (size_t)dh, cudaMemcpyDeviceToHost, stream) != cudaSuccess) { err = "D2H copy failed"; return false; }
    if (cudaStreamSynchronize(stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) { err = "CUDA error in background resize"; return false; }
    return true;
  }

  // reader thread (app/background.cc:29-104): read -> publish -> pace to `fps` -> loop at end of stream
  void reader() {
    if (debug) std::fprintf(stderr, "background: thread start\n");
    auto last = clk::now();
    auto next = last;
    while (run.load()) {
      const uint8_t* data = nullptr; int w = 0, h = 0; size_t pitch = 0;
      if (read(user, &data, &w, &h, &pitch) == 1) {
        {
          std::lock_guard<std::mutex> hold(rawmux);
          if (store(data, w, h, pitch)) frame += 1;
        }
        const auto now0 = clk::now();
        if (debug > 1) make_thumb();
        last = now0;
        // some sources are real-time, others are not: pacing makes all of them play in real time
        if (fps > 0.0) next += std::chrono::nanoseconds((long)(1e9 / fps));
        auto now = now0;
        while (now < next && run.load()) {
          std::this_thread::sleep_until(std::min(next, now + std::chrono::milliseconds(20)));   // stays responsive to delete
          now = clk::now();
        }
      } else {
        bool looped = false;
        {
          std::lock_guard<std::mutex> hold(rawmux);
          if (frame > 0 && rewind && rewind(user) == 1) { frame = 0; looped = true; }
        }
        if (!looped) {
          if (debug) std::fprintf(stderr, "background: thread stopping at end of stream and not resettable\n");
          break;
        }
      }
    }
    run = false;
    if (debug) std::fprintf(stderr, "background: thread stop\n");
  }

  // 160-pixel-wide thumbnail of the latest frame (app/background.cc:65-78; the text overlay is debug UI and not reproduced)
  void make_thumb() {
    std::vector<uint8_t> src; int w, h;
    { std::lock_guard<std::mutex> hold(rawmux); src = raw; w = rw; h = rh; }
    if (src.empty()) return;
    const int t_h = (h * 160) / w;
    if (t_h <= 0) return;
    std::vector<uint8_t> t((size_t)160 * t_h * 3);
    if (!gpu_resize(src.data(), w, h, t.data(), 160 * 3, 160, t_h)) return;
    std::lock_guard<std::mutex> hold(thumbmux);
    thumb.swap(t); tw = 160; th = t_h;
  }
};

extern "C" {

bsb_background* bsb_background_new_still(int device, const uint8_t* raw, int w, int h, size_t pitch, int debug) {
  bsb_background* b = new bsb_background();
  b->device = device; b->debug = debug;
  if (!b->store(raw, w, h, pitch)) { delete b; return nullptr; }
  return b;
}

bsb_background* bsb_background_new_video(int device, double fps, int start_frame, bsb_bg_read_cb read, bsb_bg_rewind_cb rewind,
                                         void* user, const uint8_t* first, int w, int h, size_t pitch, int debug) {
  if (!read) return nullptr;
  bsb_background* b = new bsb_background();
  b->device = device; b->debug = debug; b->video = true; b->fps = fps; b->frame = start_frame;
  b->read = read; b->rewind = rewind; b->user = user;
  if (first && !b->store(first, w, h, pitch)) { delete b; return nullptr; }   // the frame load_background() already decoded
  b->run = true;
  b->thread = std::thread(&bsb_background::reader, b);
  return b;
}

void bsb_background_delete(bsb_background* b) {
  if (!b) return;
  if (b->video) {
    b->run = false;
    if (b->thread.joinable()) b->thread.join();
  }
  if (b->stream) { cudaSetDevice(b->device); cudaStreamSynchronize(b->stream); }
  if (b->d_raw) cudaFree(b->d_raw);
  if (b->d_out) cudaFree(b->d_out);
  if (b->tab.blob) cudaFree(b->tab.blob);
  if (b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

int bsb_background_grab(bsb_background* b, int width, int height, uint8_t* out, size_t out_pitch) {
  if (!b || !out || width <= 0 || height <= 0 || out_pitch < (size_t)width * 3) return -1;
  // frame and frame number are taken under the mutex, as app/background.cc:184-188 does
  std::lock_guard<std::mutex> hold(b->rawmux);
  if (b->raw.empty()) return -1;
  if (!b->gpu_resize(b->raw.data(), b->rw, b->rh, out, out_pitch, width, height)) {
    if (b->debug) std::fprintf(stderr, "background: %s\n", b->err.c_str());
    return -1;
  }
  return b->video ? b->frame : 1;
}

int bsb_background_grab_into(bsb_background* b, bsb_ctx* ctx) {
  if (!b || !ctx) return -1;
  std::lock_guard<std::mutex> hold(b->rawmux);
  if (b->raw.empty()) return -1;
  if (!bsb_set_background(ctx, b->raw.data(), b->rw, b->rh, (size_t)b->rw * 3)) return -1;
  return b->video ? b->frame : 1;
}

int bsb_background_thumbnail(bsb_background* b, uint8_t* out, size_t capacity, int* w, int* h) {
  if (!b) return -1;
  std::lock_guard<std::mutex> hold(b->thumbmux);
  if (w) *w = b->tw;
  if (h) *h = b->th;
  if (out && !b->thumb.empty()) {
    if (capacity < b->thumb.size()) return -1;
    std::memcpy(out, b->thumb.data(), b->thumb.size());
  }
  return 0;
}

int bsb_background_frame(bsb_background* b) {
  if (!b) return -1;
  std::lock_guard<std::mutex> hold(b->rawmux);
  return b->video ? b->frame : 1;
}

int bsb_background_running(bsb_background* b) { return b && b->video && b->run.load() ? 1 : 0; }

}  // extern "C"
