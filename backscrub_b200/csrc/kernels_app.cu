// backscrub_b200/csrc/kernels_app.cu — the optional per-frame stages of the reference's main loop
// that sit either side of alpha_blend (app/deepseg.cc:649-679):
//   * `-p bgblur:k`  cv::GaussianBlur(bg, bg, Size(k,k), 0)            (:657-658)
//   * `-H` / `-V`    cv::flip(raw, raw, 1 / 0 / -1)                     (:667-673)
//   * virtual-camera geometry  cv::resize(raw, raw, vidGeo)             (:677-679) — k_resize_u8c3 (kernels_img.cu)
// Integer arithmetic is OpenCV's bit-exact 8-bit path restated (oracle/oracle_img.c pins it on cv2).
#include "kernels.h"

namespace bsb {

void count_launch();

// ---------------------------------------------------------------------------
// Gaussian blur, 8UC3, BORDER_REFLECT_101, OpenCV fixed point: taps q (8.8, sum 256);
//   rows: h = sum_j q[j]*src[x+j-r]   (<= 255*256, 16 bit)     cols: (sum_j q[j]*h[y+j-r] + 32768) >> 16
// Both sums are exact integers, so tap order is free.
//
// Row pass: one block = 256 pixels of one row; the reflected source span is staged in shared
// memory once, every thread then produces three byte lanes.
// ---------------------------------------------------------------------------
constexpr int GAUSS_TW = 256;

__global__ void __launch_bounds__(256) k_gauss_rows(const uint8_t* src, size_t pitch, size_t frame_stride, uint16_t* tmp,
                                                    int W, int H, GaussTaps g) {
  __shared__ uint8_t s[(GAUSS_TW + 2 * 127) * 3 + 2];
  const int r = g.k >> 1;
  const int x0 = blockIdx.x * GAUSS_TW, y = blockIdx.y, b = blockIdx.z;
  const int tw = min(GAUSS_TW, W - x0);
  const uint8_t* row = src + (size_t)b * frame_stride + (size_t)y * pitch;
  const int span = tw + 2 * r;
  for (int i = threadIdx.x; i < span; i += blockDim.x) {
    const int sx = bsb_reflect101(x0 + i - r, W);
    s[3 * i] = row[3 * sx]; s[3 * i + 1] = row[3 * sx + 1]; s[3 * i + 2] = row[3 * sx + 2];
  }
  __syncthreads();
  uint16_t* d = tmp + ((size_t)b * H + y) * (size_t)W * 3 + (size_t)x0 * 3;
  for (int l = threadIdx.x; l < tw * 3; l += blockDim.x) {
    unsigned acc = 0;
    for (int j = 0; j < g.k; ++j) acc += (unsigned)g.q[j] * s[l + 3 * j];
    d[l] = (uint16_t)acc;
  }
}

// Column pass: one block = 128 byte lanes x 32 rows; the (32 + 2r) reflected rows of 16-bit row sums
// are staged in dynamic shared memory.
constexpr int GAUSS_CL = 128, GAUSS_CH = 32;

__global__ void __launch_bounds__(256) k_gauss_cols(const uint16_t* tmp, uint8_t* dst, size_t pitch, size_t frame_stride,
                                                    int W, int H, GaussTaps g) {
  BSB_DYN_SMEM(smem_raw);
  uint16_t* sm = reinterpret_cast<uint16_t*>(smem_raw);
  const int r = g.k >> 1, lanes = W * 3;
  const int l0 = blockIdx.x * GAUSS_CL, y0 = blockIdx.y * GAUSS_CH, b = blockIdx.z;
  const int nl = min(GAUSS_CL, lanes - l0), nr = min(GAUSS_CH, H - y0);
  const uint16_t* t = tmp + (size_t)b * H * lanes;
  const int rows = nr + 2 * r;
  for (int i = threadIdx.x; i < rows * GAUSS_CL; i += blockDim.x) {
    const int rr = i / GAUSS_CL, l = i % GAUSS_CL;
    if (l < nl) sm[i] = t[(size_t)bsb_reflect101(y0 + rr - r, H) * lanes + l0 + l];
  }
  __syncthreads();
  const int l = threadIdx.x % GAUSS_CL;
  if (l >= nl) return;
  for (int yy = threadIdx.x / GAUSS_CL; yy < nr; yy += 256 / GAUSS_CL) {
    unsigned acc = 32768u;
    for (int j = 0; j < g.k; ++j) acc += (unsigned)g.q[j] * sm[(yy + j) * GAUSS_CL + l];
    dst[(size_t)b * frame_stride + (size_t)(y0 + yy) * pitch + l0 + l] = (uint8_t)(acc >> 16);
  }
}

size_t gauss_cols_smem(int k) { return (size_t)(GAUSS_CH + 2 * (k >> 1)) * GAUSS_CL * sizeof(uint16_t); }

void launch_gauss_blur(cudaStream_t s, int n, const uint8_t* src, size_t spitch, size_t sstride, uint16_t* tmp,
                       uint8_t* dst, size_t dpitch, size_t dstride, int W, int H, const GaussTaps& g) {
  BSB_LAUNCH(k_gauss_rows, dim3((unsigned)ceil_div(W, GAUSS_TW), (unsigned)H, (unsigned)n), dim3(256), 0, s,
             src, spitch, sstride, tmp, W, H, g);
  count_launch();
  const size_t smem = gauss_cols_smem(g.k);
#ifndef BSB_EMU
  static size_t configured = 0;   // largest opt-in so far (per process; every context uses the same kernel)
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(k_gauss_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
#endif
  BSB_LAUNCH(k_gauss_cols, dim3((unsigned)ceil_div(W * 3, GAUSS_CL), (unsigned)ceil_div(H, GAUSS_CH), (unsigned)n), dim3(256), smem, s,
             tmp, dst, dpitch, dstride, W, H, g);
  count_launch();
}

// OpenCV getGaussianKernelBitExact + getGaussianKernelFixedPoint_ED for sigma = 0 (8 fractional bits).
bool gauss_taps(int k, GaussTaps* out) {
  if (k < 1 || !(k & 1) || k > 255) return false;
  static const double small[5][9] = {
      {1.0}, {0.25, 0.5, 0.25}, {0.0625, 0.25, 0.375, 0.25, 0.0625},
      {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125},
      {4.0 / 256, 13.0 / 256, 30.0 / 256, 51.0 / 256, 60.0 / 256, 51.0 / 256, 30.0 / 256, 13.0 / 256, 4.0 / 256}};
  double v[255];
  const int n2 = k / 2;
  if (k <= 9) {
    for (int i = 0; i < k; ++i) v[i] = small[n2][i];
  } else {
    const double sigma = std::fma((double)k, 0.15, 0.35), scale2x = -0.125 / (sigma * sigma);
    double sum = 0.0;
    for (int i = 0, x = 1 - k; i < n2; ++i, x += 2) { v[i] = std::exp((double)(x * x) * scale2x); sum += v[i]; }
    const double mul = 1.0 / (sum * 2.0 + 1.0);
    for (int i = 0; i < n2; ++i) v[i] *= mul;
    v[n2] = mul;
  }
  double err = 0.0;
  int acc = 0;
  for (int i = 0; i < n2; ++i) {
    const double adj = v[i] * 256.0 + err;
    const int q = (int)std::nearbyint(adj);
    err = adj - (double)q;
    out->q[i] = out->q[k - 1 - i] = (uint16_t)q;
    acc += q;
  }
  out->q[n2] = (uint16_t)(256 - 2 * acc);
  out->k = k;
  return true;
}

// ---------------------------------------------------------------------------
// cv::flip for n packed W x H x 3 frames.  One thread = one destination pixel.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_flip_u8c3(const uint8_t* src, size_t sstride, uint8_t* dst, size_t dstride, int W, int H,
                                                   int flip_h, int flip_v) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (size_t)W * H) return;
  const int x = (int)(p % W), y = (int)(p / W);
  const uint8_t* sp = src + (size_t)blockIdx.y * sstride + ((size_t)(flip_v ? H - 1 - y : y) * W + (flip_h ? W - 1 - x : x)) * 3;
  uint8_t* dp = dst + (size_t)blockIdx.y * dstride + 3 * p;
  dp[0] = sp[0]; dp[1] = sp[1]; dp[2] = sp[2];
}

void launch_flip_u8c3(cudaStream_t s, int n, const uint8_t* src, size_t sstride, uint8_t* dst, size_t dstride, int W, int H, bool flip_h, bool flip_v) {
  const size_t npix = (size_t)W * H;
  BSB_LAUNCH(k_flip_u8c3, dim3((unsigned)((npix + 255) / 256), (unsigned)n), dim3(256), 0, s, src, sstride, dst, dstride, W, H,
             flip_h ? 1 : 0, flip_v ? 1 : 0);
  count_launch();
}

__global__ void k_advance_cursor(int* cursor, int step, int count) { *cursor = (int)(((unsigned)*cursor + (unsigned)step) % (unsigned)count); }

void launch_advance_cursor(cudaStream_t s, int* cursor, int step, int count) {
  BSB_LAUNCH(k_advance_cursor, dim3(1), dim3(1), 0, s, cursor, step, count);
  count_launch();
}

}  // namespace bsb
