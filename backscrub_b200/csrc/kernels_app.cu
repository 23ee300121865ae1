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

// ---------------------------------------------------------------------------
// Fused tile kernel for 3 <= k <= 31 (the default strength is 25): one block = 64 x 32 output pixels.
//   A  the reflected (32+2r) x (64+2r) source patch is staged de-interleaved (one byte plane per channel),
//   B  row sums with DP4A: a thread owns 4 adjacent outputs of one plane row; it reads aligned 32-bit words
//      and the taps come pre-shifted per output (hq[i][g], byte b = q[4g+b-i]) so no byte realignment is needed;
//      the 16-bit sums are stored transposed (row index fastest) so that
//   C  column sums run on DP2A: a thread owns two vertically adjacent outputs and reads (row, row+1) pairs as
//      words; cw[g] packs the tap pairs of both outputs (low half: q[2g],q[2g+1]; high half: q[2g-1],q[2g]),
//   D  the re-interleaved tile goes out in 16-byte stores.
// All sums are exact integers (<= 255*256*256 < 2^24), so the regrouping is bit-neutral.
// ---------------------------------------------------------------------------
constexpr int GF_TW = 64, GF_TH = 32, GF_MAXR = 15, GF_RM = GF_TH + 2 * GF_MAXR, GF_PW = 100, GF_HR = 66;
constexpr int GF_NW = 9, GF_NV = 16;

struct GaussFast { int k, nw, nv; unsigned hq[4][GF_NW]; unsigned cw[GF_NV]; };

// NWC / NVC > 0: tap-word counts fixed at compile time (the default strength 25 -> 7 / 13), so the dot-product
// chains are straight-line code; 0: run-time loops for the other strengths.
template <int NWC, int NVC>
__global__ void __launch_bounds__(256) k_gauss_fused(const uint8_t* src, size_t spitch, size_t sstride, uint8_t* dst, size_t dpitch,
                                                     size_t dstride, int W, int H, GaussFast g, int vec_ok) {
  __shared__ __align__(16) uint8_t P[3 * GF_RM * GF_PW];      // planes; reused as the interleaved output tile in C/D
  __shared__ __align__(16) uint16_t Ht[3 * GF_TW * GF_HR];
  const int r = g.k >> 1, tid = threadIdx.x;
  const int x0 = blockIdx.x * GF_TW, y0 = blockIdx.y * GF_TH, b = blockIdx.z;
  const int tw = min(GF_TW, W - x0), th = min(GF_TH, H - y0);
  const int rows = GF_TH + 2 * r, cols = GF_TW + 2 * r;
  const uint8_t* frame = src + (size_t)b * sstride;
  const int warp = tid >> 5, lane = tid & 31;

  // ---- A: stage + de-interleave (loads of three column groups and two rows are issued before any store: the
  //      patch comes from L2/HBM and the loop is latency bound unless several requests per thread are in flight)
  const bool interior = x0 - r >= 0 && x0 + GF_TW + r <= W && y0 - r >= 0 && y0 + GF_TH + r <= H;
  if (interior) {
    const uint8_t* base = frame + (size_t)(y0 - r) * spitch + (size_t)(x0 - r) * 3;
#pragma unroll 2
    for (int row = warp; row < rows; row += 8) {
      const uint8_t* srow = base + (size_t)row * spitch;
      uint8_t v[3][3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int col = lane + 32 * j;
        if (col < cols) { const uint8_t* sp = srow + col * 3; v[j][0] = sp[0]; v[j][1] = sp[1]; v[j][2] = sp[2]; }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int col = lane + 32 * j;
        if (col < cols) { uint8_t* pp = P + row * GF_PW + col; pp[0] = v[j][0]; pp[GF_RM * GF_PW] = v[j][1]; pp[2 * GF_RM * GF_PW] = v[j][2]; }
      }
    }
  } else {
    for (int row = warp; row < rows; row += 8) {
      const uint8_t* srow = frame + (size_t)bsb_reflect101(y0 + row - r, H) * spitch;
      for (int col = lane; col < cols; col += 32) {
        const uint8_t* sp = srow + (size_t)bsb_reflect101(x0 + col - r, W) * 3;
        uint8_t* pp = P + row * GF_PW + col;
        pp[0] = sp[0]; pp[GF_RM * GF_PW] = sp[1]; pp[2 * GF_RM * GF_PW] = sp[2];
      }
    }
  }
  __syncthreads();

  // ---- B: row sums (DP4A), stored transposed ----
  const int nw = NWC > 0 ? NWC : g.nw, nv = NVC > 0 ? NVC : g.nv;
  for (int cm = warp; cm < 3 * (GF_TW / 4); cm += 8) {          // one (plane, 4-pixel group) per warp, lanes walk the rows
    const int c = cm / (GF_TW / 4), m = cm % (GF_TW / 4);
    for (int row = lane; row < rows; row += 32) {
      const unsigned* pw = reinterpret_cast<const unsigned*>(P + (c * GF_RM + row) * GF_PW + 4 * m);
      unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      if constexpr (NWC > 0) {
#pragma unroll
        for (int q = 0; q < NWC; ++q) {
          const unsigned w = pw[q];
          a0 = __dp4a(w, g.hq[0][q], a0); a1 = __dp4a(w, g.hq[1][q], a1);
          a2 = __dp4a(w, g.hq[2][q], a2); a3 = __dp4a(w, g.hq[3][q], a3);
        }
      } else {
#pragma unroll 1
        for (int q = 0; q < nw; ++q) {
          const unsigned w = pw[q];
          a0 = __dp4a(w, g.hq[0][q], a0); a1 = __dp4a(w, g.hq[1][q], a1);
          a2 = __dp4a(w, g.hq[2][q], a2); a3 = __dp4a(w, g.hq[3][q], a3);
        }
      }
      uint16_t* hp = Ht + (c * GF_TW + 4 * m) * GF_HR + row;
      hp[0] = (uint16_t)a0; hp[GF_HR] = (uint16_t)a1; hp[2 * GF_HR] = (uint16_t)a2; hp[3 * GF_HR] = (uint16_t)a3;
    }
  }
  __syncthreads();

  // ---- C: column sums (DP2A), re-interleaved into the output tile ----
  uint8_t* O = P;                                               // [GF_TH][GF_TW * 3]
  for (int it = tid; it < 3 * (GF_TH / 2) * GF_TW; it += 256) {
    const int x = it % GF_TW, yp = (it / GF_TW) % (GF_TH / 2), c = it / (GF_TW * (GF_TH / 2));
    const unsigned* hw = reinterpret_cast<const unsigned*>(Ht + (c * GF_TW + x) * GF_HR + 2 * yp);
    unsigned a0 = 32768u, a1 = 32768u;
    if constexpr (NVC > 0) {
#pragma unroll
      for (int q = 0; q < NVC; ++q) {
        const unsigned w = hw[q];
        a0 = __dp2a_lo(w, g.cw[q], a0); a1 = __dp2a_hi(w, g.cw[q], a1);
      }
    } else {
#pragma unroll 1
      for (int q = 0; q < nv; ++q) {
        const unsigned w = hw[q];
        a0 = __dp2a_lo(w, g.cw[q], a0); a1 = __dp2a_hi(w, g.cw[q], a1);
      }
    }
    O[(2 * yp) * (GF_TW * 3) + x * 3 + c] = (uint8_t)(a0 >> 16);
    O[(2 * yp + 1) * (GF_TW * 3) + x * 3 + c] = (uint8_t)(a1 >> 16);
  }
  __syncthreads();

  // ---- D: write out ----
  uint8_t* out = dst + (size_t)b * dstride + (size_t)y0 * dpitch + (size_t)x0 * 3;
  if (vec_ok && tw == GF_TW) {
    for (int i = tid; i < th * (GF_TW * 3 / 16); i += 256) {
      const int row = i / (GF_TW * 3 / 16), v = i % (GF_TW * 3 / 16);
      *reinterpret_cast<uint4*>(out + (size_t)row * dpitch + 16 * v) = *reinterpret_cast<const uint4*>(O + row * (GF_TW * 3) + 16 * v);
    }
  } else {
    for (int i = tid; i < th * tw * 3; i += 256) {
      const int row = i / (tw * 3), l = i - row * (tw * 3);
      out[(size_t)row * dpitch + l] = O[row * (GF_TW * 3) + l];
    }
  }
}

static bool gauss_fast_params(const GaussTaps& t, GaussFast* f) {
  if (t.k < 3 || t.k > 2 * GF_MAXR + 1) return false;
  f->k = t.k; f->nw = (t.k + 3 + 3) / 4; f->nv = (t.k + 1 + 1) / 2;
  auto q = [&](int j) -> unsigned { return (j >= 0 && j < t.k) ? (unsigned)t.q[j] : 0u; };
  for (int i = 0; i < 4; ++i)
    for (int w = 0; w < GF_NW; ++w) {
      unsigned v = 0;
      for (int b = 0; b < 4; ++b) v |= q(4 * w + b - i) << (8 * b);
      f->hq[i][w] = v;
    }
  for (int w = 0; w < GF_NV; ++w) f->cw[w] = q(2 * w) | (q(2 * w + 1) << 8) | (q(2 * w - 1) << 16) | (q(2 * w) << 24);
  return true;
}

size_t gauss_cols_smem(int k) { return (size_t)(GAUSS_CH + 2 * (k >> 1)) * GAUSS_CL * sizeof(uint16_t); }

void launch_gauss_blur(cudaStream_t s, int n, const uint8_t* src, size_t spitch, size_t sstride, uint16_t* tmp,
                       uint8_t* dst, size_t dpitch, size_t dstride, int W, int H, const GaussTaps& g) {
  GaussFast f;
  if (gauss_fast_params(g, &f)) {
    const int vec_ok = (reinterpret_cast<uintptr_t>(dst) % 16 == 0 && dpitch % 16 == 0 && dstride % 16 == 0) ? 1 : 0;
    const dim3 grid((unsigned)ceil_div(W, GF_TW), (unsigned)ceil_div(H, GF_TH), (unsigned)n);
    if (f.nw == 7 && f.nv == 13) { auto kern = k_gauss_fused<7, 13>; BSB_LAUNCH(kern, grid, dim3(256), 0, s, src, spitch, sstride, dst, dpitch, dstride, W, H, f, vec_ok); }
    else { auto kern = k_gauss_fused<0, 0>; BSB_LAUNCH(kern, grid, dim3(256), 0, s, src, spitch, sstride, dst, dpitch, dstride, W, H, f, vec_ok); }
    count_launch();
    return;
  }
  BSB_LAUNCH(k_gauss_rows, dim3((unsigned)ceil_div(W, GAUSS_TW), (unsigned)H, (unsigned)n), dim3(256), 0, s,
             src, spitch, sstride, tmp, W, H, g);
  count_launch();
  const size_t smem = gauss_cols_smem(g.k);
  ensure_dyn_smem(reinterpret_cast<const void*>(k_gauss_cols), smem);
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
