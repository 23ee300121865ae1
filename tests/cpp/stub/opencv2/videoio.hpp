// tests/cpp/stub/opencv2/videoio.hpp — test-only stand-in for cv::VideoCapture: "decodes" a raw container
//   magic "RAWV" | int32 width | int32 height | int32 frames | int32 fps | int32 seekable | frames x (h*w*3) BGR bytes
// so background_shim.cc's probe / reader / rewind logic can be exercised without FFmpeg.
#pragma once
#include <cstdio>
#include <string>
#include <vector>

#include "core/mat.hpp"

namespace cv {
enum { CAP_ANY = 0, CAP_PROP_POS_FRAMES = 1, CAP_PROP_FPS = 5, CAP_PROP_FOURCC = 6, CAP_PROP_FRAME_COUNT = 7, CAP_PROP_CONVERT_RGB = 16 };
class VideoCapture {
 public:
  bool open(const std::string& path, int = CAP_ANY) {
    release();
    f_ = std::fopen(path.c_str(), "rb");
    if (!f_) return false;
    char magic[4]; int hdr[5];
    if (std::fread(magic, 1, 4, f_) != 4) { release(); return false; }
    if (std::memcmp(magic, "RAWI", 4) == 0 && std::fread(hdr, 4, 2, f_) == 2) {
      // like OpenCV's image-sequence backend: a still image opens as a one-frame "video"
      w_ = hdr[0]; h_ = hdr[1]; n_ = 1; fps_ = 0; seekable_ = false; pos_ = 0; base_ = 12;
      return true;
    }
    if (std::memcmp(magic, "RAWV", 4) != 0 || std::fread(hdr, 4, 5, f_) != 5) { release(); return false; }
    w_ = hdr[0]; h_ = hdr[1]; n_ = hdr[2]; fps_ = hdr[3]; seekable_ = hdr[4] != 0; pos_ = 0; base_ = 24;
    return true;
  }
  bool isOpened() const { return f_ != nullptr; }
  void release() { if (f_) std::fclose(f_); f_ = nullptr; }
  bool read(Mat& out) {
    if (!f_ || pos_ >= n_) return false;
    out.create(h_, w_, CV_8UC3);
    if (std::fread(out.data, 1, (size_t)w_ * h_ * 3, f_) != (size_t)w_ * h_ * 3) return false;
    ++pos_;
    return true;
  }
  bool set(int prop, double v) {
    if (prop == CAP_PROP_CONVERT_RGB) return true;
    if (prop == CAP_PROP_POS_FRAMES && f_ && seekable_) { pos_ = (int)v; std::fseek(f_, base_ + (long)pos_ * w_ * h_ * 3, SEEK_SET); return true; }
    return false;
  }
  double get(int prop) const { return prop == CAP_PROP_FPS ? (double)fps_ : (prop == CAP_PROP_FRAME_COUNT ? (double)n_ : 0.0); }
  ~VideoCapture() { release(); }
 private:
  FILE* f_ = nullptr; int w_ = 0, h_ = 0, n_ = 0, fps_ = 0, pos_ = 0; long base_ = 24; bool seekable_ = false;
};
}  // namespace cv
