// backscrub_b200/csrc/kernels_nn.cu — fp32 CNN kernels for sm_100a (CUDA cores).
//
// These replace the XNNPACK/TFLite CPU kernels behind Interpreter::Invoke()
// (lib/libbackscrub.cc:307; op semantics: reference tensorflow/lite/kernels/internal/
// reference/{conv,depthwiseconv_float,pooling,resize_bilinear,hard_swish,logistic,
// fully_connected}.h and lib/transpose_conv_bias.cc).  Every output element accumulates
// its products with fmaf in the reference loop order (ky, kx, cin), so results are
// bit-identical to the CPU oracle; parallelism comes from pixels x channels x frames.
#include <atomic>

#include "kernels.h"

namespace bsb {

static std::atomic<long> g_launches{0};
long launch_count() { return g_launches.load(); }
void count_launch() { g_launches.fetch_add(1); }

struct EpiDev {
  const float* bias; const float* residual; int ld_res; int act1, act2, act3;
};
static inline EpiDev to_dev(const Epilogue& e) { return EpiDev{e.bias, e.residual, e.ld_res, e.act1, e.act2, e.act3}; }

BSB_D float epilogue(float total, int ch, size_t pix, const EpiDev& e) {
  float v = total + (e.bias ? __ldg(e.bias + ch) : 0.f);
  v = bsb_act(v, e.act1);
  v = bsb_act(v, e.act2);
  if (e.residual) v = bsb_act(v + __ldg(e.residual + pix * (size_t)e.ld_res + ch), e.act3);
  return v;
}

// ---------------------------------------------------------------------------
// Dense KxK conv, small Cin (stem).  One thread = one output pixel x 4 output channels.
// Weights [kh][kw][ic][oc4] staged in shared memory (broadcast reads).
// ---------------------------------------------------------------------------
struct ConvArgs {
  const float* in; const float* w; float* out;
  int B, ih, iw, ic, ld_in, oc, oc4, kh, kw, sh, sw, dh, dw, pt, pl, oh, ow, ld_out;
  EpiDev e;
};

__global__ void __launch_bounds__(256) k_conv_direct(ConvArgs a) {
  BSB_DYN_SMEM(smem_raw);
  float* ws = reinterpret_cast<float*>(smem_raw);
  const int wcount = a.kh * a.kw * a.ic * a.oc4;
  for (int i = threadIdx.x; i < wcount; i += blockDim.x) ws[i] = __ldg(a.w + i);
  __syncthreads();
  const int groups = a.oc4 / 4;
  const long total = (long)a.B * a.oh * a.ow * groups;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const long pix = idx / groups;
  const int ox = (int)(pix % a.ow);
  const int oy = (int)((pix / a.ow) % a.oh);
  const int b = (int)(pix / ((long)a.ow * a.oh));
  const float* inb = a.in + (size_t)b * a.ih * a.iw * a.ld_in;
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  const int iy0 = oy * a.sh - a.pt, ix0 = ox * a.sw - a.pl;
  for (int fy = 0; fy < a.kh; ++fy) {
    const int iy = iy0 + a.dh * fy;
    if (iy < 0 || iy >= a.ih) continue;
    for (int fx = 0; fx < a.kw; ++fx) {
      const int ix = ix0 + a.dw * fx;
      if (ix < 0 || ix >= a.iw) continue;
      const float* ip = inb + ((size_t)iy * a.iw + ix) * a.ld_in;
      const float* wp = ws + ((fy * a.kw + fx) * a.ic) * a.oc4 + g * 4;
      for (int c = 0; c < a.ic; ++c) {
        const float v = __ldg(ip + c);
        const float4 w4 = *reinterpret_cast<const float4*>(wp + c * a.oc4);
        acc0 = fmaf(v, w4.x, acc0); acc1 = fmaf(v, w4.y, acc1);
        acc2 = fmaf(v, w4.z, acc2); acc3 = fmaf(v, w4.w, acc3);
      }
    }
  }
  float* op = a.out + (size_t)pix * a.ld_out;
  const float accs[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ch = g * 4 + j;
    if (ch < a.oc) op[ch] = epilogue(accs[j], ch, (size_t)pix, a.e);
  }
}

void launch_conv_direct(cudaStream_t s, int B, const float* in, int ih, int iw, int ic, int ld_in,
                        const float* w_t, int oc, int kh, int kw, int stride_h, int stride_w,
                        int dil_h, int dil_w, int pad_t, int pad_l,
                        float* out, int oh, int ow, int ld_out, const Epilogue& e) {
  ConvArgs a{in, w_t, out, B, ih, iw, ic, ld_in, oc, (oc + 3) / 4 * 4, kh, kw, stride_h, stride_w, dil_h, dil_w,
             pad_t, pad_l, oh, ow, ld_out, to_dev(e)};
  const long total = (long)B * oh * ow * (a.oc4 / 4);
  const size_t smem = sizeof(float) * (size_t)kh * kw * ic * a.oc4;
  BSB_LAUNCH(k_conv_direct, dim3((unsigned)((total + 255) / 256)), dim3(256), smem, s, a);
  count_launch();
}

// ---------------------------------------------------------------------------
// Pointwise (1x1) conv / fully connected as a shared-memory tiled FFMA GEMM.
//   out[m][n] = epilogue( sum_k A'[m][k] * W[k][n] ),  A' = A (* scale[frame][k]) (+ add[m][k])
// 256 threads; thread tile TM rows x 4 cols; K consumed in ascending chunks of 16 so each
// accumulator sees its products in k order (bit-exact vs the oracle).
// ---------------------------------------------------------------------------
struct PWArgs {
  const float* A; const float* w; float* out; const float* in_scale; const float* in_add;
  int M, K, N, n4, ld_a, ld_out, rows_per_frame, ld_add;
  EpiDev e;
};

template <int BN, int TM>
__global__ void __launch_bounds__(256) k_pointwise(PWArgs a) {
  constexpr int BK = 16;
  constexpr int CT = BN / 4;          // column threads
  constexpr int RT = 256 / CT;        // row threads
  constexpr int BM = RT * TM;
  constexpr int LDS_A = BM + 2;       // 4*(BM+2) % 32 == 8: conflict-free transposed stores
  __shared__ float As[BK * LDS_A];
  __shared__ __align__(16) float Ws[BK * BN];
  const int tid = threadIdx.x;
  const int tx = tid % CT, ty = tid / CT;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  float acc[TM][4];
#pragma unroll
  for (int i = 0; i < TM; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }

  for (int k0 = 0; k0 < a.K; k0 += BK) {
    // ---- stage A tile (BM x BK) transposed into As[k][m] ----
    for (int q = tid; q < BM * (BK / 4); q += 256) {
      const int m = q / (BK / 4), kq = (q % (BK / 4)) * 4;
      const int gm = m0 + m, gk = k0 + kq;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gm < a.M) {
        const float* ap = a.A + (size_t)gm * a.ld_a + gk;
        if (gk + 3 < a.K && (a.ld_a & 3) == 0) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(ap));
          v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) if (gk + j < a.K) v[j] = __ldg(ap + j);
        }
        if (a.in_scale) {
          const float* sp = a.in_scale + (size_t)(gm / a.rows_per_frame) * a.K + gk;
#pragma unroll
          for (int j = 0; j < 4; ++j) if (gk + j < a.K) v[j] = v[j] * __ldg(sp + j);
        }
        if (a.in_add) {
          const float* dp = a.in_add + (size_t)gm * a.ld_add + gk;
#pragma unroll
          for (int j = 0; j < 4; ++j) if (gk + j < a.K) v[j] = v[j] + __ldg(dp + j);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) As[(kq + j) * LDS_A + m] = v[j];
    }
    // ---- stage W tile (BK x BN) ----
    for (int q = tid; q < BK * (BN / 4); q += 256) {
      const int k = q / (BN / 4), nq = (q % (BN / 4)) * 4;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + k < a.K && n0 + nq < a.n4) t = __ldg(reinterpret_cast<const float4*>(a.w + (size_t)(k0 + k) * a.n4 + n0 + nq));
      *reinterpret_cast<float4*>(&Ws[k * BN + nq]) = t;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[k * BN + tx * 4]);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const float av = As[k * LDS_A + ty + RT * i];
        acc[i][0] = fmaf(av, w4.x, acc[i][0]);
        acc[i][1] = fmaf(av, w4.y, acc[i][1]);
        acc[i][2] = fmaf(av, w4.z, acc[i][2]);
        acc[i][3] = fmaf(av, w4.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty + RT * i;
    if (gm >= a.M) continue;
    float* op = a.out + (size_t)gm * a.ld_out;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ch = n0 + tx * 4 + j;
      if (ch < a.N) op[ch] = epilogue(acc[i][j], ch, (size_t)gm, a.e);
    }
  }
}

void launch_pointwise(cudaStream_t s, int M, int K, int N, const float* A, int ld_a,
                      const float* w_kn, int n4, float* out, int ld_out, const Epilogue& e,
                      const float* in_scale, int rows_per_frame, const float* in_add, int ld_add) {
  PWArgs a{A, w_kn, out, in_scale, in_add, M, K, N, n4, ld_a, ld_out, rows_per_frame > 0 ? rows_per_frame : 1, ld_add, to_dev(e)};
  // choose the N tile that wastes the fewest columns; ties go to the wider tile
  const int pad16 = (N + 15) / 16 * 16, pad32 = (N + 31) / 32 * 32, pad64 = (N + 63) / 64 * 64;
  int bn = 64;
  if (pad32 < pad64) bn = 32;
  if (pad16 < (bn == 64 ? pad64 : pad32)) bn = 16;
  if (bn == 64) {
    BSB_LAUNCH((k_pointwise<64, 4>), dim3((unsigned)ceil_div(M, 64), (unsigned)ceil_div(N, 64)), dim3(256), 0, s, a);
  } else if (bn == 32) {
    BSB_LAUNCH((k_pointwise<32, 4>), dim3((unsigned)ceil_div(M, 128), (unsigned)ceil_div(N, 32)), dim3(256), 0, s, a);
  } else {
    BSB_LAUNCH((k_pointwise<16, 4>), dim3((unsigned)ceil_div(M, 256), (unsigned)ceil_div(N, 16)), dim3(256), 0, s, a);
  }
  count_launch();
}

// ---------------------------------------------------------------------------
// Depthwise KxK.  One thread = one output pixel x 4 channels (float4 over C).
// ---------------------------------------------------------------------------
struct DWArgs {
  const float* in; const float* w; float* out;
  int B, ih, iw, c, ld_in, kh, kw, sh, sw, dh, dw, pt, pl, oh, ow, ld_out;
  EpiDev e;
};

template <int VEC>
__global__ void __launch_bounds__(256) k_depthwise(DWArgs a) {
  const int groups = (a.c + VEC - 1) / VEC;
  const long total = (long)a.B * a.oh * a.ow * groups;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c0 = (int)(idx % groups) * VEC;
  const long pix = idx / groups;
  const int ox = (int)(pix % a.ow);
  const int oy = (int)((pix / a.ow) % a.oh);
  const int b = (int)(pix / ((long)a.ow * a.oh));
  const float* inb = a.in + (size_t)b * a.ih * a.iw * a.ld_in;
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  const int iy0 = oy * a.sh - a.pt, ix0 = ox * a.sw - a.pl;
  for (int fy = 0; fy < a.kh; ++fy) {
    const int iy = iy0 + a.dh * fy;
    if (iy < 0 || iy >= a.ih) continue;
    for (int fx = 0; fx < a.kw; ++fx) {
      const int ix = ix0 + a.dw * fx;
      if (ix < 0 || ix >= a.iw) continue;
      const float* ip = inb + ((size_t)iy * a.iw + ix) * a.ld_in + c0;
      const float* wp = a.w + (size_t)(fy * a.kw + fx) * a.c + c0;
      if (VEC == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(ip));
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(wp));
        acc[0] = fmaf(v.x, w4.x, acc[0]); acc[1 % VEC] = fmaf(v.y, w4.y, acc[1 % VEC]);
        acc[2 % VEC] = fmaf(v.z, w4.z, acc[2 % VEC]); acc[3 % VEC] = fmaf(v.w, w4.w, acc[3 % VEC]);
      } else {
        acc[0] = fmaf(__ldg(ip), __ldg(wp), acc[0]);
      }
    }
  }
  float* op = a.out + (size_t)pix * a.ld_out + c0;
#pragma unroll
  for (int j = 0; j < VEC; ++j) op[j] = epilogue(acc[j], c0 + j, (size_t)pix, a.e);
}

void launch_depthwise(cudaStream_t s, int B, const float* in, int ih, int iw, int c, int ld_in,
                      const float* w, int kh, int kw, int stride_h, int stride_w, int dil_h, int dil_w,
                      int pad_t, int pad_l, float* out, int oh, int ow, int ld_out, const Epilogue& e) {
  DWArgs a{in, w, out, B, ih, iw, c, ld_in, kh, kw, stride_h, stride_w, dil_h, dil_w, pad_t, pad_l, oh, ow, ld_out, to_dev(e)};
  const bool vec = (c % 4 == 0) && (ld_in % 4 == 0) && (ld_out % 4 == 0);
  const long total = (long)B * oh * ow * (vec ? c / 4 : c);
  if (vec) BSB_LAUNCH(k_depthwise<4>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, a);
  else BSB_LAUNCH(k_depthwise<1>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, a);
  count_launch();
}

// ---------------------------------------------------------------------------
// Global average pool.  Summation order (shared with the oracle): each row left to
// right, then the row sums top to bottom; total / (float)(H*W).
// One block per frame; row sums in shared memory.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_global_avgpool(const float* in, int h, int w, int c, int ld_in,
                                                        float* out, int ld_out, int act) {
  BSB_DYN_SMEM(smem_raw);
  float* rows = reinterpret_cast<float*>(smem_raw);   // [h][c]
  const int b = blockIdx.x;
  const float* inb = in + (size_t)b * h * w * ld_in;
  for (int i = threadIdx.x; i < h * c; i += blockDim.x) {
    const int y = i / c, ch = i % c;
    const float* p = inb + (size_t)y * w * ld_in + ch;
    float r = 0.f;
    for (int x = 0; x < w; ++x) r = r + __ldg(p + (size_t)x * ld_in);
    rows[i] = r;
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < c; ch += blockDim.x) {
    float t = 0.f;
    for (int y = 0; y < h; ++y) t = t + rows[y * c + ch];
    const float count = (float)(h * w);
    out[(size_t)b * ld_out + ch] = bsb_act(bsb_div(t, count), act);
  }
}

void launch_global_avgpool(cudaStream_t s, int B, const float* in, int h, int w, int c, int ld_in,
                           float* out, int ld_out, int act) {
  BSB_LAUNCH(k_global_avgpool, dim3((unsigned)B), dim3(256), sizeof(float) * (size_t)h * c, s, in, h, w, c, ld_in, out, ld_out, act);
  count_launch();
}

// ---------------------------------------------------------------------------
// RESIZE_BILINEAR, float, NHWC.  One thread = one output pixel x 4 channels.
// ---------------------------------------------------------------------------
BSB_D void interp(float value, float scale, bool half_pixel, int in_size, float* scaled, int* lo, int* hi) {
  *scaled = half_pixel ? (value + 0.5f) * scale - 0.5f : value * scale;
  const float fl = floorf(*scaled);
  int l = (int)fl; if (l < 0) l = 0;
  int h = (int)ceilf(*scaled); if (h > in_size - 1) h = in_size - 1;
  *lo = l; *hi = h;
}

template <int VEC>
__global__ void __launch_bounds__(256) k_resize_bilinear(const float* in, int B, int ih, int iw, int c, int ld_in,
                                                         float* out, int oh, int ow, int ld_out,
                                                         float hs, float ws, bool half_pixel) {
  const int groups = (c + VEC - 1) / VEC;
  const long total = (long)B * oh * ow * groups;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c0 = (int)(idx % groups) * VEC;
  const long pix = idx / groups;
  const int x = (int)(pix % ow), y = (int)((pix / ow) % oh), b = (int)(pix / ((long)ow * oh));
  float fy, fx; int y0, y1, x0, x1;
  interp((float)y, hs, half_pixel, ih, &fy, &y0, &y1);
  interp((float)x, ws, half_pixel, iw, &fx, &x0, &x1);
  const float dy = fy - (float)y0, dx = fx - (float)x0;
  const float wy0 = 1.f - dy, wx0 = 1.f - dx;
  const float* inb = in + (size_t)b * ih * iw * ld_in + c0;
  const float* p00 = inb + ((size_t)y0 * iw + x0) * ld_in;
  const float* p10 = inb + ((size_t)y1 * iw + x0) * ld_in;
  const float* p01 = inb + ((size_t)y0 * iw + x1) * ld_in;
  const float* p11 = inb + ((size_t)y1 * iw + x1) * ld_in;
  float* op = out + (size_t)pix * ld_out + c0;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    if (c0 + j >= c) break;
    const float a = __ldg(p00 + j) * wy0 * wx0;
    const float bb = __ldg(p10 + j) * dy * wx0;
    const float d = __ldg(p01 + j) * wy0 * dx;
    const float e = __ldg(p11 + j) * dy * dx;
    op[j] = ((a + bb) + d) + e;
  }
}

void launch_resize_bilinear(cudaStream_t s, int B, const float* in, int ih, int iw, int c, int ld_in,
                            float* out, int oh, int ow, int ld_out, bool align_corners, bool half_pixel) {
  float hs = (float)ih / (float)oh, ws = (float)iw / (float)ow;
  if (align_corners && oh > 1) hs = (float)(ih - 1) / (float)(oh - 1);
  if (align_corners && ow > 1) ws = (float)(iw - 1) / (float)(ow - 1);
  const long total = (long)B * oh * ow * ((c + 3) / 4);
  BSB_LAUNCH(k_resize_bilinear<4>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s,
             in, B, ih, iw, c, ld_in, out, oh, ow, ld_out, hs, ws, half_pixel);
  count_launch();
}

// ---------------------------------------------------------------------------
// Convolution2DTransposeBias, k = 2x2, stride 2, SAME (even output): no overlap, so
// out[2y+fy][2x+fx][o] = bias[o] + sum_ic in[y][x][ic] * w[o][fy][fx][ic], ic ascending
// (the order in which lib/transpose_conv_bias.cc:80-108 scatters into one output).
// One thread = one output pixel; weights in shared memory.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tconv2x2(const float* in, int B, int ih, int iw, int ic, int ld_in,
                                                  const float* w, const float* bias, int oc,
                                                  float* out, int oh, int ow, int ld_out, int act2) {
  BSB_DYN_SMEM(smem_raw);
  float* ws = reinterpret_cast<float*>(smem_raw);    // [oc][2][2][ic]
  for (int i = threadIdx.x; i < oc * 4 * ic; i += blockDim.x) ws[i] = __ldg(w + i);
  __syncthreads();
  const long total = (long)B * oh * ow;
  const long pix = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const int x = (int)(pix % ow), y = (int)((pix / ow) % oh), b = (int)(pix / ((long)ow * oh));
  const int iy = y >> 1, ix = x >> 1, fy = y & 1, fx = x & 1;
  if (iy >= ih || ix >= iw) return;
  const float* ip = in + ((size_t)b * ih * iw + (size_t)iy * iw + ix) * ld_in;
  float* op = out + (size_t)pix * ld_out;
  for (int o = 0; o < oc; ++o) {
    const float* wp = ws + ((o * 2 + fy) * 2 + fx) * ic;
    float acc = __ldg(bias + o);
    for (int c = 0; c < ic; ++c) acc = fmaf(__ldg(ip + c), wp[c], acc);
    op[o] = bsb_act(acc, act2);
  }
}

void launch_tconv2x2(cudaStream_t s, int B, const float* in, int ih, int iw, int ic, int ld_in,
                     const float* w, const float* bias, int oc, float* out, int oh, int ow, int ld_out, int act2) {
  const long total = (long)B * oh * ow;
  BSB_LAUNCH(k_tconv2x2, dim3((unsigned)((total + 255) / 256)), dim3(256), sizeof(float) * (size_t)oc * 4 * ic, s,
             in, B, ih, iw, ic, ld_in, w, bias, oc, out, oh, ow, ld_out, act2);
  count_launch();
}

// ---------------------------------------------------------------------------
// Element-wise (only what the planner could not fold into a producer/consumer).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_eltwise(int mode, long rows, int hw, int c, const float* a, int ld_a,
                                                 const float* b, int ld_b, const float* scale,
                                                 float* out, int ld_out, int act) {
  const long total = rows * c;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ch = (int)(idx % c);
  const long row = idx / c;
  const float av = __ldg(a + (size_t)row * ld_a + ch);
  float v;
  switch (mode) {
    case 0: v = av; break;
    case 1: v = av + __ldg(b + (size_t)row * ld_b + ch); break;
    case 2: v = av * __ldg(b + (size_t)row * ld_b + ch); break;
    case 3: v = av * __ldg(scale + (size_t)(row / hw) * c + ch); break;
    default: v = av * __ldg(scale + (size_t)(row / hw) * c + ch); v = v + __ldg(b + (size_t)row * ld_b + ch); break;
  }
  out[(size_t)row * ld_out + ch] = bsb_act(v, act);
}

void launch_eltwise(cudaStream_t s, int mode, int B, int hw, int c, const float* a, int ld_a,
                    const float* b, int ld_b, const float* scale, float* out, int ld_out, int act) {
  const long rows = (long)B * hw;
  const long total = rows * c;
  BSB_LAUNCH(k_eltwise, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, mode, rows, hw, c, a, ld_a, b, ld_b, scale, out, ld_out, act);
  count_launch();
}

__global__ void __launch_bounds__(256) k_copy_channels(long rows, int c, const float* in, int ld_in, float* out, int ld_out) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * c) return;
  const int ch = (int)(idx % c);
  const long row = idx / c;
  out[(size_t)row * ld_out + ch] = __ldg(in + (size_t)row * ld_in + ch);
}

void launch_copy_channels(cudaStream_t s, int rows, int c, const float* in, int ld_in, float* out, int ld_out) {
  const long total = (long)rows * c;
  BSB_LAUNCH(k_copy_channels, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, (long)rows, c, in, ld_in, out, ld_out);
  count_launch();
}

}  // namespace bsb
