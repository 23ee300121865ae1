// tests/cpp/stub/opencv2/core/mat.hpp — MINIMAL stand-in for OpenCV's cv::Mat, test-only.
// This image has no OpenCV C++ headers, so the reference-compatible C++ layer (include/lib/libbackscrub.h,
// include/background.h and the shims that define them) is compiled against this stub to prove it builds, links and
// behaves; with a real OpenCV the genuine <opencv2/core/mat.hpp> is found instead.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>

#define CV_8U 0
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC2 CV_MAKETYPE(CV_8U, 2)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)

namespace cv {
class Mat {
 public:
  int rows = 0, cols = 0;
  uint8_t* data = nullptr;
  size_t step = 0;
  Mat() = default;
  Mat(int r, int c, int type, void* ext, size_t st = 0) : rows(r), cols(c), data(static_cast<uint8_t*>(ext)), type_(type) {
    step = st ? st : (size_t)c * channels();
  }
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    if (own_ && r == rows && c == cols && type == type_) return;
    rows = r; cols = c; type_ = type; step = (size_t)c * channels();
    own_.reset(static_cast<uint8_t*>(std::malloc(step * (size_t)r + 1)), std::free);
    data = own_.get();
  }
  Mat clone() const {
    Mat m;
    if (empty()) return m;
    m.create(rows, cols, type_);
    for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * channels());
    return m;
  }
  void copyTo(Mat& dst) const { dst = clone(); }
  void release() { own_.reset(); data = nullptr; rows = cols = 0; step = 0; }
  int type() const { return type_; }
  int channels() const { return (type_ >> 3) + 1; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  size_t total() const { return (size_t)rows * cols; }
  size_t elemSize() const { return (size_t)channels(); }
 private:
  int type_ = 0;
  std::shared_ptr<uint8_t> own_;
};
}  // namespace cv
