// tests/cpp/shim_test.cpp — drives the reference-compatible C++ API (include/libbackscrub.h,
// include/background.h) exactly the way app/deepseg.cc's CalcMask does (app/deepseg.cc:203,246,269):
//   bs_maskgen_new(modelname, threads, w, h, ondebug, onprep, oninfer, onmask, ctx)
//   bs_maskgen_process(ctx, frame, mask)   -> mask aliases context storage
//   bs_maskgen_delete(ctx)
// Usage: shim_test <model.tflite> <W> <H> <frames.bgr (n*W*H*3 raw bytes)> <n> <masks.out>
// Writes the n masks (W*H bytes each) and prints the callback order; the pytest wrapper compares
// the masks with the oracle.  Exit code 3 = no CUDA device (context creation failed loudly).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "background.h"
#include "libbackscrub.h"

static std::string g_events;
static void on_debug(void*, const char* msg) { std::fprintf(stderr, "debug: %s", msg); }
static void on_prep(void*) { g_events += "P"; }
static void on_infer(void*) { g_events += "I"; }
static void on_mask(void*) { g_events += "M"; }

int main(int argc, char** argv) {
  if (argc < 7) { std::fprintf(stderr, "usage: %s model W H frames.bgr n masks.out\n", argv[0]); return 2; }
  const std::string model = argv[1];
  const int W = std::atoi(argv[2]), H = std::atoi(argv[3]), n = std::atoi(argv[5]);
  std::printf("runtime: %s\n", bs_tensorflow_version());
  void* ctx = bs_maskgen_new(model, 2, W, H, on_debug, on_prep, on_infer, on_mask, nullptr);
  if (!ctx) { std::printf("bs_maskgen_new returned nullptr\n"); return 3; }
  std::vector<uint8_t> frames((size_t)n * W * H * 3);
  FILE* f = std::fopen(argv[4], "rb");
  if (!f || std::fread(frames.data(), 1, frames.size(), f) != frames.size()) { std::fprintf(stderr, "cannot read frames\n"); return 2; }
  std::fclose(f);
  FILE* out = std::fopen(argv[6], "wb");
  for (int i = 0; i < n; ++i) {
    cv::Mat frame(H, W, CV_8UC3, frames.data() + (size_t)i * W * H * 3);
    cv::Mat mask;
    if (!bs_maskgen_process(ctx, frame, mask)) { std::fprintf(stderr, "bs_maskgen_process failed\n"); return 1; }
    if (mask.rows != H || mask.cols != W || mask.type() != CV_8UC1) { std::fprintf(stderr, "bad mask header\n"); return 1; }
    for (int y = 0; y < H; ++y) std::fwrite(mask.data + (size_t)y * mask.step, 1, W, out);
  }
  std::fclose(out);
  // background adapter: resize a synthetic 64x48 ramp to W x H through the GPU
  cv::Mat raw(48, 64, CV_8UC3), bg;
  for (int i = 0; i < 48 * 64 * 3; ++i) raw.data[i] = (uint8_t)(i * 7);
  const int rc = bsb_grab_background(static_cast<bsb_ctx*>(ctx), raw, W, H, bg);
  std::printf("grab_background rc=%d size=%dx%d\n", rc, bg.cols, bg.rows);
  bs_maskgen_delete(ctx);
  bs_maskgen_delete(nullptr);                     // nullptr-safe like the reference
  std::printf("callbacks: %s\n", g_events.c_str());
  return 0;
}
