// tests/cpp/dropin_test.cc — a translation unit that uses the reference-facing C++ API exactly the way
// app/deepseg.cc does: the same two include lines (app/deepseg.cc:24-25) and the same call shapes
//   :246  bs_maskgen_new(modelname.c_str(), threads, width, height, nullptr, onprep, oninfer, onmask, this)
//   :203  bs_maskgen_process(maskctx, *frame_current, *mask_current)
//   :269  bs_maskgen_delete(maskctx)
//   :351  bs_tensorflow_version()
//   :596  load_background(path, debug)         :649  grab_background(pbk, width, height, bg)
// It is linked with backscrub_b200/shim/{libbackscrub_shim,background_shim}.cc + libbackscrub_b200.so and built
// twice by the tests: with this repo's include/ first, and (when /root/reference exists) with the reference's own
// headers first, as the reference's CMakeLists.txt:72 would have it.
//
// usage: dropin_test <model> <W> <H> <frames.bgr> <n> <background file> <out prefix>
//   writes <prefix>.masks (n*W*H), <prefix>.bg (W*H*3 of the first grab), prints frame numbers / callback order.
//   exit code 3 = bs_maskgen_new returned nullptr (no CUDA device: the library has no CPU path).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "lib/libbackscrub.h"
#include "background.h"

static std::string g_events;

class CalcMaskLike {
 public:
  void* maskctx;
  static void onprep(void*) { g_events += "P"; }
  static void oninfer(void*) { g_events += "I"; }
  static void onmask(void*) { g_events += "M"; }
  CalcMaskLike(const std::string& modelname, size_t threads, size_t width, size_t height) {
    maskctx = bs_maskgen_new(modelname.c_str(), threads, width, height, nullptr, onprep, oninfer, onmask, this);
  }
  ~CalcMaskLike() { bs_maskgen_delete(maskctx); }
};

int main(int argc, char** argv) {
  if (argc < 8) { std::fprintf(stderr, "usage: %s model W H frames.bgr n background prefix\n", argv[0]); return 2; }
  const std::string model = argv[1], bgpath = argv[6], prefix = argv[7];
  const int W = std::atoi(argv[2]), H = std::atoi(argv[3]), n = std::atoi(argv[5]);
  std::printf("runtime: %s\n", bs_tensorflow_version());

  // ---- background provider first: it must behave without a mask context ----
  auto pbk(load_background(bgpath, 2));
  std::printf("load_background: %s\n", pbk ? "ok" : "nullptr");
  auto none(load_background(bgpath + ".does-not-exist", 0));
  std::printf("load_background(missing): %s\n", none ? "ok" : "nullptr");
  cv::Mat bg;
  std::printf("grab_background(nullptr) = %d\n", grab_background(nullptr, W, H, bg));

  CalcMaskLike ai(model, 2, W, H);
  if (!ai.maskctx) { std::printf("bs_maskgen_new returned nullptr\n"); return 3; }

  std::vector<uint8_t> frames((size_t)n * W * H * 3);
  FILE* f = std::fopen(argv[4], "rb");
  if (!f || std::fread(frames.data(), 1, frames.size(), f) != frames.size()) { std::fprintf(stderr, "cannot read frames\n"); return 2; }
  std::fclose(f);
  FILE* out = std::fopen((prefix + ".masks").c_str(), "wb");
  for (int i = 0; i < n; ++i) {
    cv::Mat frame(H, W, CV_8UC3, frames.data() + (size_t)i * W * H * 3);
    cv::Mat mask;
    if (!bs_maskgen_process(ai.maskctx, frame, mask)) { std::fprintf(stderr, "bs_maskgen_process failed\n"); return 1; }
    if (mask.rows != H || mask.cols != W || mask.type() != CV_8UC1) { std::fprintf(stderr, "bad mask header\n"); return 1; }
    cv::Mat copy = mask.clone();                        // the app clones under its lock (app/deepseg.cc:208-213)
    for (int y = 0; y < H; ++y) std::fwrite(copy.data + (size_t)y * copy.step, 1, W, out);
  }
  std::fclose(out);
  {
    // a frame smaller than the context size is an error (the reference's cv::Mat::operator() would throw), and a
    // frame with extra rows / columns is read through its top-left W x H window and still yields a W x H mask
    cv::Mat small(H / 2, W / 2, CV_8UC3), mask;
    std::memset(small.data, 0, (size_t)(H / 2) * small.step);
    std::printf("short frame -> %d\n", (int)bs_maskgen_process(ai.maskctx, small, mask));
    cv::Mat big(H + 8, W + 16, CV_8UC3);
    for (int y = 0; y < H + 8; ++y) std::memset(big.data + (size_t)y * big.step, 7, big.step);
    for (int y = 0; y < H; ++y) std::memcpy(big.data + (size_t)y * big.step, frames.data() + ((size_t)(n - 1) * H + y) * W * 3, (size_t)W * 3);
    const bool ok = bs_maskgen_process(ai.maskctx, big, mask);
    std::printf("oversize frame -> %d mask %dx%d\n", (int)ok, mask.cols, mask.rows);
  }

  if (pbk) {
    int frm = grab_background(pbk, W, H, bg);
    std::printf("grab_background frame=%d size=%dx%d\n", frm, bg.cols, bg.rows);
    FILE* fb = std::fopen((prefix + ".bg").c_str(), "wb");
    for (int y = 0; y < H; ++y) std::fwrite(bg.data + (size_t)y * bg.step, 1, (size_t)W * 3, fb);
    std::fclose(fb);
    // a video keeps playing in real time and loops: sample the frame counter for a while
    std::printf("frames:");
    for (int i = 0; i < 12; ++i) {
      std::this_thread::sleep_for(std::chrono::milliseconds(50));
      std::printf(" %d", grab_background(pbk, W, H, bg));
    }
    std::printf("\n");
    cv::Mat thumb;
    const int trc = grab_thumbnail(pbk, thumb);
    std::printf("grab_thumbnail rc=%d size=%dx%d\n", trc, thumb.cols, thumb.rows);
  }
  std::printf("callbacks: %s\n", g_events.c_str());
  return 0;
}
