// tests/cpp/stub/opencv2/imgcodecs.hpp — test-only stand-in for cv::imread: magic "RAWI" | int32 w | int32 h | BGR bytes
#pragma once
#include <cstdio>
#include <string>

#include "core/mat.hpp"

namespace cv {
inline Mat imread(const std::string& path) {
  Mat m;
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return m;
  char magic[4]; int hdr[2];
  if (std::fread(magic, 1, 4, f) == 4 && std::memcmp(magic, "RAWI", 4) == 0 && std::fread(hdr, 4, 2, f) == 2) {
    m.create(hdr[1], hdr[0], CV_8UC3);
    if (std::fread(m.data, 1, (size_t)hdr[0] * hdr[1] * 3, f) != (size_t)hdr[0] * hdr[1] * 3) m.release();
  }
  std::fclose(f);
  return m;
}
}  // namespace cv
