// backscrub_b200/csrc/kernels_chain.cu — the low-resolution middle of the MobileNetV3-style graphs in ONE kernel.
//
// The Meet / MLKit graphs spend 20+ operators (and, unfused, 30 kernel launches) on tensors of 9x16 / 16x16 pixels:
// depthwise 5x5, squeeze-excite (global pool -> FC -> FC -> sigmoid -> channel scale), 1x1 project / expand
// (reference kernels: depthwiseconv_float.h:25-96, pooling.h:26-79, fully_connected.h:27-61, conv.h:25-99; delegated to
// XNNPACK by the reference, xnnpack_delegate.cc:1840-2497).  None of those launches can fill a B200, and the SE
// reductions serialise them.  Here one CTA owns one frame and keeps every activation of the chain in shared memory
// (see kernels.h: X / D / E); the batch x streams dimension provides the parallelism (one frame per SM at a time).
//
// Arithmetic contract: identical to the stand-alone kernels (kernels_nn.cu) and therefore to the oracle —
// every output accumulates fmaf in ascending k / (fy, fx) order, the pool sums each row left to right and then the
// rows top to bottom, the SE scale multiplies the operand (one rounding) before the fmaf.
#include <algorithm>

#include "kernels.h"

namespace bsb {

void count_launch();

namespace {

constexpr int CH_T = 512;            // threads per CTA
constexpr int CH_XLD = 32;           // X row stride (floats): block inputs / outputs have <= 32 channels
constexpr int CH_DLD = 128;          // D row stride: expanded tensors have <= 128 channels
constexpr int CH_ELD = 32;           // E slice width

constexpr int CH_PAD = 2;            // zero halo of the E slice (covers 3x3 and 5x5 stride-1 windows)
struct ChainLayout { int off_x, off_d, off_e, e_floats, off_w1, off_wd, off_vec, total; };   // float offsets
BSB_HD ChainLayout chain_layout(int h, int w) {
  ChainLayout L;
  const int P = h * w;
  L.off_x = 0;
  L.off_d = L.off_x + P * CH_XLD;
  L.off_e = L.off_d + P * CH_DLD;                 // zero-haloed E slice; PW / FC weight staging and the pool row sums alias it
  int e_floats = (h + 2 * CH_PAD) * (w + 2 * CH_PAD) * CH_ELD;
  if (e_floats < 4096) e_floats = 4096;
  if (e_floats < h * CH_DLD) e_floats = h * CH_DLD;
  L.e_floats = e_floats;
  L.off_w1 = L.off_e + e_floats;
  L.off_wd = L.off_w1 + 32 * 32;                  // expand weights of one slice [cin <= 32][32]
  L.off_vec = L.off_wd + 25 * 32;                 // depthwise taps of one slice [k*k <= 25][32]
  L.total = L.off_vec + 3 * 128;                  // v0 | v1 | sv
  return L;
}

BSB_D float chain_act2(float v, int a1, int a2) { return bsb_act(bsb_act(v, a1), a2); }

// dst[p][n] = act(sum_k A[p][k] (* sv[k]) * Ws[k][n] + bias) (+ residual), R rows x 4 columns per thread, k ascending.
// A: shared [P][lda]; Ws: shared [K][n4] (zero padded); dst: shared, row p at dst + drow(p) * ldd (may be the residual
// source: each element is read and then written by the same thread).  pw > 0: dst is the zero-haloed E slice
// (row p = pixel (p / pw, p % pw) of a (h + 4) x (pw + 4) grid).  ACT2 >= 0 fixes the second activation at compile time
// (the first one must then be ACT_NONE); ACT2 < 0 is the generic path.
template <int R, int ACT2>
BSB_D void chain_gemm_t(const float* A, int lda, int K, const float* sv, const float* Ws, int n4, int N, const float* bias,
                        int a1, int a2, bool residual, int a3, float* dst, int ldd, int P, int pw) {
  const int cg = n4 / 4, rg = (P + R - 1) / R;
  for (int t = threadIdx.x; t < cg * rg; t += blockDim.x) {
    const int tx = t % cg, ty = t / cg;
    const int n0 = tx * 4, p0 = ty * R;
    float acc[R][4];
#pragma unroll
    for (int i = 0; i < R; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    const float* ar[R];
#pragma unroll
    for (int i = 0; i < R; ++i) ar[i] = A + (size_t)min(p0 + i, P - 1) * lda;
    for (int k = 0; k < K; k += 4) {
      float4 av[R];
#pragma unroll
      for (int i = 0; i < R; ++i) av[i] = *reinterpret_cast<const float4*>(ar[i] + k);
      if (sv) {
        const float4 sc = *reinterpret_cast<const float4*>(sv + k);
#pragma unroll
        for (int i = 0; i < R; ++i) { av[i].x = av[i].x * sc.x; av[i].y = av[i].y * sc.y; av[i].z = av[i].z * sc.z; av[i].w = av[i].w * sc.w; }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 w4 = *reinterpret_cast<const float4*>(Ws + (size_t)(k + j) * n4 + n0);
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const float a = j == 0 ? av[i].x : (j == 1 ? av[i].y : (j == 2 ? av[i].z : av[i].w));
          acc[i][0] = fmaf(a, w4.x, acc[i][0]); acc[i][1] = fmaf(a, w4.y, acc[i][1]);
          acc[i][2] = fmaf(a, w4.z, acc[i][2]); acc[i][3] = fmaf(a, w4.w, acc[i][3]);
        }
      }
    }
    float bv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bv[j] = (bias && n0 + j < N) ? __ldg(bias + n0 + j) : 0.f;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int p = p0 + i;
      if (p >= P) break;
      const int drow = pw > 0 ? ((p / pw + CH_PAD) * (pw + 2 * CH_PAD) + (p % pw) + CH_PAD) : p;
      float* dp = dst + (size_t)drow * ldd + n0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n0 + j >= N) break;
        float v = acc[i][j] + bv[j];
        v = ACT2 >= 0 ? bsb_act(v, ACT2) : chain_act2(v, a1, a2);
        if (residual) v = bsb_act(v + dp[j], a3);
        dp[j] = v;
      }
    }
  }
}

template <int R>
BSB_D void chain_gemm(const float* A, int lda, int K, const float* sv, const float* Ws, int n4, int N, const float* bias,
                      int a1, int a2, bool residual, int a3, float* dst, int ldd, int P, int pw) {
  // the activation combinations of the MobileNetV3-style graphs get their own instantiation (no per-element switch)
  if (a1 == ACT_NONE && a2 == ACT_NONE) chain_gemm_t<R, ACT_NONE>(A, lda, K, sv, Ws, n4, N, bias, a1, a2, residual, a3, dst, ldd, P, pw);
  else if (a1 == ACT_NONE && a2 == ACT_HARD_SWISH) chain_gemm_t<R, ACT_HARD_SWISH>(A, lda, K, sv, Ws, n4, N, bias, a1, a2, residual, a3, dst, ldd, P, pw);
  else if (a1 == ACT_NONE && a2 == ACT_RELU6) chain_gemm_t<R, ACT_RELU6>(A, lda, K, sv, Ws, n4, N, bias, a1, a2, residual, a3, dst, ldd, P, pw);
  else if (a1 == ACT_NONE && a2 == ACT_RELU) chain_gemm_t<R, ACT_RELU>(A, lda, K, sv, Ws, n4, N, bias, a1, a2, residual, a3, dst, ldd, P, pw);
  else chain_gemm_t<R, -1>(A, lda, K, sv, Ws, n4, N, bias, a1, a2, residual, a3, dst, ldd, P, pw);
}

// global [K][n4] -> shared, all loads in flight
BSB_D void chain_stage(float* ws, const float* w, int count) {
  for (int i = threadIdx.x * 4; i < count; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(ws + i) = __ldg(reinterpret_cast<const float4*>(w + i));
}

// one fully-connected layer of the SE path: out[n] = act(sum_k in[k] * w[k][n] + bias), k ascending.  The weights are
// pulled into shared memory (`stage`, `cap` floats) by the whole CTA in chunks of rows — one coalesced L2 round trip per
// chunk instead of K dependent ones — and the running sums stay in registers across chunks.
BSB_D void chain_fc(const float* in, const FcLayer& f, float* out, float* stage, int cap) {
  const int rows_per_chunk = max(1, cap / f.n4);
  const int n = threadIdx.x;
  float acc = 0.f;
  for (int k0 = 0; k0 < f.K; k0 += rows_per_chunk) {
    const int rows = min(rows_per_chunk, f.K - k0);
    __syncthreads();                                              // the previous chunk (or the caller's use of `stage`) is consumed
    chain_stage(stage, f.w + (size_t)k0 * f.n4, rows * f.n4);
    __syncthreads();
    if (n < f.N) for (int k = 0; k < rows; ++k) acc = fmaf(in[k0 + k], stage[k * f.n4 + n], acc);
  }
  if (n < f.N) out[n] = chain_act2(acc + (f.bias ? __ldg(f.bias + n) : 0.f), f.act1, f.act2);
}

// depthwise KS x KS, stride 1, SAME, of one 32-channel slice: E (zero-haloed (h+4) x (w+4) grid of [32]) -> D[:, c0 : c0 + cw].
// A lane keeps its channel, so its KS*KS taps live in registers; a thread produces 4 consecutive pixels of a row from
// KS x (KS + 3) loads.  Out-of-image taps read the halo's zeros: fmaf(+0, w, acc) leaves acc unchanged bit for bit (an
// accumulator that starts at +0 can never become -0 by adding zeros), so skipping them as the oracle does gives the same bits.
template <int KS, int ACT2>
BSB_D void chain_dw_slice_t(const float* E, const float* wds, const ChainOp& op, int c0, int cw, int h, int w, float* D) {
  const int j = threadIdx.x & 31;
  if (j >= cw) return;
  float wr[KS * KS];
#pragma unroll
  for (int t = 0; t < KS * KS; ++t) wr[t] = wds[t * 32 + j];
  const float bias = op.bd ? __ldg(op.bd + c0 + j) : 0.f;
  const int ew = w + 2 * CH_PAD, strips = (w + 3) / 4;
  const bool full = (w & 3) == 0;                 // every strip has 4 pixels and its KS + 3 columns lie inside the haloed grid
  for (int sidx = threadIdx.x >> 5; sidx < h * strips; sidx += blockDim.x >> 5) {
    const int oy = sidx / strips, ox0 = (sidx - oy * strips) * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // window origin in the haloed grid: (oy - pt + CH_PAD, ox0 - pl + CH_PAD); pt, pl <= CH_PAD
    const float* e0 = E + ((size_t)(oy - op.pt + CH_PAD) * ew + (ox0 - op.pl + CH_PAD)) * CH_ELD + j;
#pragma unroll
    for (int fy = 0; fy < KS; ++fy) {
      float v[KS + 3];
      if (full) {
#pragma unroll
        for (int c = 0; c < KS + 3; ++c) v[c] = e0[((size_t)fy * ew + c) * CH_ELD];
      } else {
#pragma unroll
        for (int c = 0; c < KS + 3; ++c) v[c] = (ox0 - op.pl + CH_PAD + c < ew) ? e0[((size_t)fy * ew + c) * CH_ELD] : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int fx = 0; fx < KS; ++fx) acc[q] = fmaf(v[q + fx], wr[fy * KS + fx], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (ox0 + q < w) {
        const float r = acc[q] + bias;
        D[(size_t)(oy * w + ox0 + q) * CH_DLD + c0 + j] = ACT2 >= 0 ? bsb_act(r, ACT2) : chain_act2(r, op.dact1, op.dact2);
      }
  }
}

// depthwise KS x KS, stride 1, SAME, of one 32-channel slice: E (zero-haloed (h+4) x (w+4) grid of [32]) -> D[:, c0 : c0 + cw].
// A lane keeps its channel, so its KS*KS taps live in registers; a thread produces 4 consecutive pixels of a row from
// KS x (KS + 3) loads.  Out-of-image taps read the halo's zeros: fmaf(+0, w, acc) leaves acc unchanged bit for bit (an
// accumulator that starts at +0 can never become -0 by adding zeros), so skipping them as the oracle does gives the same bits.
template <int KS>
BSB_D void chain_dw_slice(const float* E, const float* wds, const ChainOp& op, int c0, int cw, int h, int w, float* D) {
  if (op.dact1 == ACT_NONE && op.dact2 == ACT_HARD_SWISH) chain_dw_slice_t<KS, ACT_HARD_SWISH>(E, wds, op, c0, cw, h, w, D);
  else if (op.dact1 == ACT_NONE && op.dact2 == ACT_RELU6) chain_dw_slice_t<KS, ACT_RELU6>(E, wds, op, c0, cw, h, w, D);
  else if (op.dact1 == ACT_NONE && op.dact2 == ACT_RELU) chain_dw_slice_t<KS, ACT_RELU>(E, wds, op, c0, cw, h, w, D);
  else chain_dw_slice_t<KS, -1>(E, wds, op, c0, cw, h, w, D);
}

}  // namespace

template <int R>
__global__ void __launch_bounds__(CH_T, 1) k_chain(const ChainOp* ops, int n_ops, int h, int w) {
  BSB_DYN_SMEM(smem_raw);
  float* sm = reinterpret_cast<float*>(smem_raw);
  const int P = h * w;
  const ChainLayout L = chain_layout(h, w);
  float* X = sm + L.off_x; float* D = sm + L.off_d; float* E = sm + L.off_e;
  float* w1s = sm + L.off_w1; float* wds = sm + L.off_wd;
  float* v0 = sm + L.off_vec; float* v1 = v0 + 128; float* sv = v1 + 128;
  const int tid = threadIdx.x, T = blockDim.x;
  const int b = blockIdx.x;

  for (int oi = 0; oi < n_ops; ++oi) {
    const ChainOp& op = ops[oi];
    __syncthreads();                                      // the previous op's results are complete
    switch (op.type) {
      case CH_DWG: {
        // depthwise k x k, stride s, SAME padding, straight from the global input tensor into D[p][c]
        const float* gin = op.gin + (size_t)b * op.gin_frame;
        const int C = op.cin, groups = C / 4;
        for (int i = tid; i < P * groups; i += T) {
          const int p = i / groups, c0 = (i - p * groups) * 4;
          const int oy = p / w, ox = p - oy * w;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int fy = 0; fy < op.k; ++fy) {
            const int iy = oy * op.s - op.pt + fy;
            if (iy < 0 || iy >= op.ih) continue;
            for (int fx = 0; fx < op.k; ++fx) {
              const int ix = ox * op.s - op.pl + fx;
              if (ix < 0 || ix >= op.iw) continue;
              const float4 v = __ldg(reinterpret_cast<const float4*>(gin + ((size_t)iy * op.iw + ix) * op.gin_ld + c0));
              const float4 wv = __ldg(reinterpret_cast<const float4*>(op.wd + (size_t)(fy * op.k + fx) * C + c0));
              acc.x = fmaf(v.x, wv.x, acc.x); acc.y = fmaf(v.y, wv.y, acc.y); acc.z = fmaf(v.z, wv.z, acc.z); acc.w = fmaf(v.w, wv.w, acc.w);
            }
          }
          float* d = D + (size_t)p * CH_DLD + c0;
          d[0] = chain_act2(acc.x + (op.bd ? __ldg(op.bd + c0) : 0.f), op.dact1, op.dact2);
          d[1] = chain_act2(acc.y + (op.bd ? __ldg(op.bd + c0 + 1) : 0.f), op.dact1, op.dact2);
          d[2] = chain_act2(acc.z + (op.bd ? __ldg(op.bd + c0 + 2) : 0.f), op.dact1, op.dact2);
          d[3] = chain_act2(acc.w + (op.bd ? __ldg(op.bd + c0 + 3) : 0.f), op.dact1, op.dact2);
        }
        break;
      }
      case CH_EXPAND_DW: {
        // X [P][cin] -> expand 1x1 (+act) -> depthwise k x k stride 1 (+act) -> D [P][cout], one 32-channel slice at a time
        // (E was used as scratch by the ops in between: its zero halo is rebuilt first)
        {
          const int ew = w + 2 * CH_PAD, eh = h + 2 * CH_PAD;
          for (int i = tid; i < eh * ew * (CH_ELD / 4); i += T) {
            const int cell = i / (CH_ELD / 4), ey = cell / ew, ex = cell - ey * ew;
            if (ey < CH_PAD || ey >= h + CH_PAD || ex < CH_PAD || ex >= w + CH_PAD)
              *reinterpret_cast<float4*>(E + (size_t)cell * CH_ELD + 4 * (i % (CH_ELD / 4))) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        for (int c0 = 0; c0 < op.cout; c0 += 32) {
          const int cw = min(32, op.cout - c0);
          if (c0) __syncthreads();                        // the previous slice's E / w1s / wds are fully consumed
          for (int i = tid; i < op.cin * 32; i += T) {
            const int kk = i >> 5, j = i & 31;
            w1s[i] = j < cw ? __ldg(op.w + (size_t)kk * op.n4 + c0 + j) : 0.f;
          }
          for (int i = tid; i < op.k * op.k * 32; i += T) {
            const int tp = i >> 5, j = i & 31;
            wds[i] = j < cw ? __ldg(op.wd + (size_t)tp * op.cout + c0 + j) : 0.f;
          }
          __syncthreads();
          chain_gemm<R>(X, CH_XLD, op.cin, nullptr, w1s, 32, cw, op.b ? op.b + c0 : nullptr, op.act1, op.act2, false, 0, E, CH_ELD, P, w);
          __syncthreads();
          if (op.k == 5) chain_dw_slice<5>(E, wds, op, c0, cw, h, w, D);
          else chain_dw_slice<3>(E, wds, op, c0, cw, h, w, D);
        }
        break;
      }
      case CH_SE: {
        // global average pool: row sums (x ascending) -> E (aliased as rs[h][C]), then rows (y ascending), / (h*w)
        const float* src = op.src ? D : X;
        const int ld = op.src ? CH_DLD : CH_XLD, C = op.cin;
        float* rs = E;
        for (int i = tid; i < h * C; i += T) {
          const int y = i / C, c = i - y * C;
          const float* sp = src + (size_t)(y * w) * ld + c;
          float r = 0.f;
          for (int x = 0; x < w; ++x) r = r + sp[(size_t)x * ld];
          rs[i] = r;
        }
        __syncthreads();
        for (int c = tid; c < C; c += T) {
          float t = 0.f;
          for (int y = 0; y < h; ++y) t = t + rs[y * C + c];
          v0[c] = bsb_act(bsb_div(t, (float)(h * w)), op.pool_act);
        }
        __syncthreads();
        if (op.n_fc == 1) chain_fc(v0, op.f0, sv, E, L.e_floats);
        else {
          chain_fc(v0, op.f0, v1, E, L.e_floats);
          chain_fc(v1, op.f1, sv, E, L.e_floats);
        }
        break;
      }
      case CH_PW: {
        const float* A = op.src ? D : X;
        const int lda = op.src ? CH_DLD : CH_XLD;
        float* dst = op.dst ? D : X;
        const int ldd = op.dst ? CH_DLD : CH_XLD;
        float* Ws = E;                                    // [K][n4] staging (E is dead outside EXPAND_DW)
        chain_stage(Ws, op.w, op.cin * op.n4);
        __syncthreads();
        chain_gemm<R>(A, lda, op.cin, op.use_scale ? sv : nullptr, Ws, op.n4, op.cout, op.b, op.act1, op.act2, op.residual != 0, op.act3, dst, ldd, P, 0);
        break;
      }
      case CH_SCALE_STORE: {
        float* go = op.gout + (size_t)b * op.gout_frame;
        const int C = op.cin, groups = C / 4;
        for (int i = tid; i < P * groups; i += T) {
          const int p = i / groups, c0 = (i - p * groups) * 4;
          const float* d = D + (size_t)p * CH_DLD + c0;
          *reinterpret_cast<float4*>(go + (size_t)p * op.gout_ld + c0) =
              make_float4(bsb_act(d[0] * sv[c0], op.act1), bsb_act(d[1] * sv[c0 + 1], op.act1), bsb_act(d[2] * sv[c0 + 2], op.act1),
                          bsb_act(d[3] * sv[c0 + 3], op.act1));
        }
        break;
      }
      default: break;
    }
  }
}

size_t chain_smem_bytes(int h, int w) {
  const int P = h * w;
  if (P < 1 || P > 256 || h > 64) return 0;
  const size_t bytes = (size_t)chain_layout(h, w).total * sizeof(float);
  return bytes <= 226 * 1024 ? bytes : 0;
}

void launch_chain(cudaStream_t s, int B, int h, int w, const ChainOp* d_ops, int n_ops) {
  const size_t smem = chain_smem_bytes(h, w);
  // threads and row tile so that the narrow GEMMs ((P / R) x 8 thread tiles) fill the CTA exactly once:
  // 9x16 (Meet): 48 x 8 = 384 tiles of 3 rows; 16x16 (MLKit): 64 x 8 = 512 tiles of 4 rows
  const int P = h * w;
  if (P % 3 == 0 && (P / 3) * 8 <= 384) {
    ensure_dyn_smem(reinterpret_cast<const void*>(k_chain<3>), smem);
    BSB_LAUNCH(k_chain<3>, dim3((unsigned)B), dim3(384), smem, s, d_ops, n_ops, h, w);
  } else {
    ensure_dyn_smem(reinterpret_cast<const void*>(k_chain<4>), smem);
    BSB_LAUNCH(k_chain<4>, dim3((unsigned)B), dim3(CH_T), smem, s, d_ops, n_ops, h, w);
  }
  count_launch();
}

}  // namespace bsb
