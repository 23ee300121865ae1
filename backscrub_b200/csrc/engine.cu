// backscrub_b200/csrc/engine.cu — planner + per-stream engine (see engine.h).
#include "engine.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

namespace bsb {

#define CUDA_OK(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      if (err) *err = std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr; \
      return false;                                                                        \
    }                                                                                      \
  } while (0)

// ---------------------------------------------------------------------------
// OpenCV INTER_LINEAR 8-bit tables (cv::resize; formulas pinned in oracle/oracle_img.c)
// ---------------------------------------------------------------------------
HostResizeTab build_resize_tab(int sw, int sh, int dw, int dh) {
  HostResizeTab t;
  t.sw = sw;
  t.area2x2 = (sw == dw * 2 && sh == dh * 2);
  const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
  t.xofs.resize(dw); t.xw.resize(2 * (size_t)dw);
  for (int dx = 0; dx < dw; ++dx) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)std::floor(fx);
    fx -= (float)sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    t.xofs[dx] = sx;
    t.xw[2 * dx] = (short)std::lrintf((1.f - fx) * 2048.f);
    t.xw[2 * dx + 1] = (short)std::lrintf(fx * 2048.f);
  }
  t.yofs0.resize(dh); t.yofs1.resize(dh); t.yw.resize(2 * (size_t)dh);
  for (int dy = 0; dy < dh; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)std::floor(fy);
    fy -= (float)sy;
    t.yofs0[dy] = std::min(std::max(sy, 0), sh - 1);
    t.yofs1[dy] = std::min(std::max(sy + 1, 0), sh - 1);
    t.yw[2 * dy] = (short)std::lrintf((1.f - fy) * 2048.f);
    t.yw[2 * dy + 1] = (short)std::lrintf(fy * 2048.f);
  }
  return t;
}

bool upload_resize_tab(const HostResizeTab& h, DevResizeTab* d, std::string* err) {
  const size_t nx = h.xofs.size(), ny = h.yofs0.size();
  auto al = [](size_t n) { return (n + 15) / 16 * 16; };
  const size_t o_xofs = 0, o_y0 = o_xofs + al(nx * 4), o_y1 = o_y0 + al(ny * 4), o_xw = o_y1 + al(ny * 4), o_yw = o_xw + al(nx * 4);
  const size_t o_xcol = o_yw + al(ny * 4);
  const size_t total = o_xcol + al(nx * 8);
  std::vector<uint8_t> blob(total, 0);
  {
    std::vector<uint32_t> xc(2 * nx);
    for (size_t i = 0; i < nx; ++i) {
      const int sx = h.xofs[i], sx1 = std::min(sx + 1, h.sw - 1);
      xc[2 * i] = (uint32_t)sx | ((uint32_t)sx1 << 16);
      xc[2 * i + 1] = (uint32_t)(uint16_t)h.xw[2 * i] | ((uint32_t)(uint16_t)h.xw[2 * i + 1] << 16);
    }
    std::memcpy(blob.data() + o_xcol, xc.data(), nx * 8);
  }
  std::memcpy(blob.data() + o_xofs, h.xofs.data(), nx * 4);
  std::memcpy(blob.data() + o_y0, h.yofs0.data(), ny * 4);
  std::memcpy(blob.data() + o_y1, h.yofs1.data(), ny * 4);
  std::memcpy(blob.data() + o_xw, h.xw.data(), nx * 4);
  std::memcpy(blob.data() + o_yw, h.yw.data(), ny * 4);
  if (d->blob) { cudaFree(d->blob); d->blob = nullptr; }
  CUDA_OK(cudaMalloc(&d->blob, blob.size()));
  CUDA_OK(cudaMemcpy(d->blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
  uint8_t* b = static_cast<uint8_t*>(d->blob);
  d->tab.xofs = reinterpret_cast<const int*>(b + o_xofs);
  d->tab.yofs0 = reinterpret_cast<const int*>(b + o_y0);
  d->tab.yofs1 = reinterpret_cast<const int*>(b + o_y1);
  d->tab.xw = reinterpret_cast<const short*>(b + o_xw);
  d->tab.yw = reinterpret_cast<const short*>(b + o_yw);
  d->tab.xcol = reinterpret_cast<const uint2*>(b + o_xcol);
  d->area2x2 = h.area2x2;
  return true;
}

// TF SAME/VALID output size + leading pad (reference kernels/padding.h:23-82)
static void conv_geom(int in, int k, int stride, int dil, int padding, int* out, int* pad_before) {
  const int eff = (k - 1) * dil + 1;
  const int o = padding == 0 ? (in + stride - 1) / stride : (in + stride - eff) / stride;
  int total = (o - 1) * stride + eff - in;
  if (total < 0) total = 0;
  *out = o; *pad_before = total / 2;
}

static int unary_act(int kind) {
  switch (kind) {
    case OP_HARD_SWISH: return ACT_HARD_SWISH;
    case OP_RELU: return ACT_RELU;
    case OP_RELU6: return ACT_RELU6;
    case OP_LOGISTIC: return ACT_LOGISTIC;
    default: return -1;
  }
}

// ---------------------------------------------------------------------------
// Planner: TFLite op list -> fused steps + arena layout
// ---------------------------------------------------------------------------
bool Engine::plan(std::string* err) {
  const int nt = (int)g_.tensors.size(), nops = (int)g_.ops.size();
  // Tensor cores for the 1x1 convs (3xTF32 split, ~1e-6 relative to the exact fp32 chain — the tolerance the reference's
  // own XNNPACK-vs-builtin conv tests use): on by default for the GEMM-dominated graphs (DeepLab 93 %, BodyPix 95 % of
  // their FLOPs are 1x1 convs), on request for the others (BSB_FLAG_TENSOR_CORES), never with BSB_FLAG_EXACT or when
  // every intermediate tensor is kept for bit-level comparison (BSB_FLAG_KEEP_TENSORS).
  tc_enabled_ = !(flags_ & (1u | 16u)) && ((flags_ & 4u) || model_type_ == MODEL_DEEPLAB || model_type_ == MODEL_BODYPIX);
  std::vector<std::vector<int>> consumers(nt);
  std::vector<int> produced_at(nt, -1);
  for (int i = 0; i < nops; ++i) {
    const GOp& O = g_.ops[i];
    if (O.kind == OP_DEQUANTIZE) continue;
    for (int t : O.in) if (t >= 0 && !g_.tensors[t].is_const) consumers[t].push_back(i);
    if (O.out >= 0) produced_at[O.out] = i;
  }
  tinfo_.assign(nt, TensorInfo());
  for (int t = 0; t < nt; ++t) {
    const GTensor& T = g_.tensors[t];
    if (T.is_const || T.dtype != 0) continue;
    if (T.dim4(0) != 1) { *err = "tensor '" + T.name + "': batch dimension must be 1"; return false; }
    TensorInfo& I = tinfo_[t];
    I.h = T.dim4(1); I.w = T.dim4(2); I.c = T.dim4(3); I.ld = I.c;
    I.frame_elems = (size_t)I.h * I.w * I.ld;
  }
  auto is_pw = [&](const GOp& O) {
    if (O.kind != OP_CONV_2D || O.in.size() < 2 || O.in[1] < 0) return false;
    const GTensor& w = g_.tensors[O.in[1]];
    return w.shape.size() == 4 && w.shape[1] == 1 && w.shape[2] == 1 && O.stride_w == 1 && O.stride_h == 1;
  };
  auto add_blob = [&](const std::vector<float>& v) {
    size_t off = (wblob_h_.size() + 63) / 64 * 64;
    wblob_h_.resize(off + v.size(), 0.f);
    std::copy(v.begin(), v.end(), wblob_h_.begin() + off);
    return off;
  };
  struct Fold { int x, s, add; };
  std::map<int, Fold> prologue;      // tensor consumed by a PW conv -> (x, scale, add)
  std::map<int, std::pair<int, int>> concat_for_pool;
  std::map<int, int> producer_step;  // tensor -> index in steps_ of the step that writes it
  std::map<int, std::vector<int>> aliased_into;
  std::vector<int> concat_virtual;
  std::vector<char> done(nops, 0);

  auto fold_unary = [&](int i, Step& st) {  // fold a single-consumer unary op after op i
    const int t = st.out;
    if (consumers[t].size() != 1) return;
    const int j = consumers[t][0];
    const int a = unary_act(g_.ops[j].kind);
    if (a < 0 || done[j] || j <= i) return;
    st.act2 = a; done[j] = 1; st.out = g_.ops[j].out;
  };
  auto fold_add = [&](int i, Step& st) {    // fold a single-consumer residual ADD
    const int t = st.out;
    if (t == g_.output || consumers[t].size() != 1) return;
    const int j = consumers[t][0];
    const GOp& A = g_.ops[j];
    if (A.kind != OP_ADD || done[j] || j <= i || A.in.size() != 2) return;
    const int other = A.in[0] == t ? A.in[1] : A.in[0];
    if (other == t || g_.tensors[other].is_const) return;
    if (g_.tensors[other].count() != g_.tensors[t].count()) return;
    if (other != g_.input && !(produced_at[other] >= 0 && produced_at[other] < i)) return;
    if (prologue.count(other)) return;
    st.residual = other; st.act3 = A.act; done[j] = 1; st.out = A.out;
  };
  auto pack_pw_weights = [&](const GTensor& w, int N, int K, Step& st) {
    st.K = K; st.N = N; st.n4 = (N + 3) / 4 * 4;
    std::vector<float> wt((size_t)K * st.n4, 0.f);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) wt[(size_t)k * st.n4 + n] = w.f32[(size_t)n * K + k];
    st.w_off = add_blob(wt);
  };
  // 3xTF32 operand split for the tensor-core path: w = hi + lo, hi = w with the 13 low mantissa bits cleared
  auto pack_tc_weights = [&](const GTensor& w, int N, int K, Step& st) {
    const int bn = pointwise_tc_tile_n(N);
    if (bn <= 0) return;
    st.kpad = (K + 31) / 32 * 32; st.npad = (N + bn - 1) / bn * bn;
    std::vector<float> hi((size_t)st.npad * st.kpad, 0.f), lo((size_t)st.npad * st.kpad, 0.f);
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
      const float v = w.f32[(size_t)n * K + k];
      uint32_t bits; std::memcpy(&bits, &v, 4); bits &= 0xffffe000u;
      float h; std::memcpy(&h, &bits, 4);
      hi[(size_t)n * st.kpad + k] = h; lo[(size_t)n * st.kpad + k] = v - h;
    }
    st.tc_hi_off = add_blob(hi); st.tc_lo_off = add_blob(lo); st.use_tc = true;
  };

  for (int i = 0; i < nops; ++i) {
    const GOp& O = g_.ops[i];
    if (done[i] || O.kind == OP_DEQUANTIZE) continue;
    Step st; st.op_index = i;
    auto need = [&](bool c, const char* what) { if (!c) { *err = std::string("unsupported model construct: ") + what; } return c; };
    switch (O.kind) {
      case OP_CONV_2D: {
        if (!need(O.in.size() >= 2 && O.in[1] >= 0 && g_.tensors[O.in[1]].is_const, "CONV_2D with non-constant weights")) return false;
        const GTensor& w = g_.tensors[O.in[1]];
        if (!need(w.shape.size() == 4 && w.f32.size() == w.count(), "CONV_2D weights that are not a rank-4 float tensor")) return false;
        if (!need(O.in.size() < 3 || O.in[2] < 0 || (int)g_.tensors[O.in[2]].f32.size() == w.shape[0], "CONV_2D bias size")) return false;
        if (!need(w.shape[3] == tinfo_[O.in[0]].c, "CONV_2D input depth")) return false;
        const int oc = w.shape[0], kh = w.shape[1], kw = w.shape[2], ic = w.shape[3];
        st.in = O.in[0]; st.out = O.out; st.act1 = O.act;
        if (O.in.size() > 2 && O.in[2] >= 0) { st.has_bias = true; st.b_off = add_blob(g_.tensors[O.in[2]].f32); }
        if (is_pw(O)) {
          st.kind = Step::PW;
          pack_pw_weights(w, oc, ic, st);
          auto it = prologue.find(st.in);
          if (it != prologue.end()) { st.in = it->second.x; st.scale = it->second.s; st.in_add = it->second.add; }
          // tensor cores (opt-in): plain GEMM-shaped layers with enough rows and depth to fill a 128 x N x 32 tile pipeline
          if (tc_enabled_ && st.scale < 0 && st.in_add < 0 && ic >= tuning().tc_min_k && ic % 4 == 0 && oc >= 8 &&
              tinfo_[st.in].h * tinfo_[st.in].w >= 1024 && tinfo_[st.in].ld % 4 == 0)
            pack_tc_weights(w, oc, ic, st);
        } else {
          st.kind = Step::CONV;
          st.kh = kh; st.kw = kw; st.sh = O.stride_h; st.sw = O.stride_w; st.dh = O.dil_h; st.dw = O.dil_w;
          st.K = ic; st.N = oc; st.n4 = (oc + 3) / 4 * 4;
          int o_h, o_w;
          conv_geom(tinfo_[st.in].h, kh, st.sh, st.dh, O.padding, &o_h, &st.pt);
          conv_geom(tinfo_[st.in].w, kw, st.sw, st.dw, O.padding, &o_w, &st.pl);
          if (!need(o_h == tinfo_[O.out].h && o_w == tinfo_[O.out].w, "CONV_2D output shape")) return false;
          if (!need((size_t)kh * kw * ic * st.n4 * 4 <= 96 * 1024, "dense KxK conv too large for the direct kernel")) return false;
          std::vector<float> wt((size_t)kh * kw * ic * st.n4, 0.f);   // [kh][kw][ic][oc4]
          for (int o = 0; o < oc; ++o) for (int y = 0; y < kh; ++y) for (int x = 0; x < kw; ++x) for (int c = 0; c < ic; ++c)
            wt[(((size_t)y * kw + x) * ic + c) * st.n4 + o] = w.f32[(((size_t)o * kh + y) * kw + x) * ic + c];
          st.w_off = add_blob(wt);
        }
        fold_unary(i, st); fold_add(i, st);
        break;
      }
      case OP_FULLY_CONNECTED: {
        if (!need(O.in.size() >= 2 && O.in[1] >= 0 && g_.tensors[O.in[1]].is_const, "FULLY_CONNECTED with non-constant weights")) return false;
        const GTensor& w = g_.tensors[O.in[1]];
        if (!need(w.shape.size() >= 2 && w.f32.size() == w.count(), "FULLY_CONNECTED weights that are not a float matrix")) return false;
        st.kind = Step::PW; st.in = O.in[0]; st.out = O.out; st.act1 = O.act;
        const int out_depth = w.shape[w.shape.size() - 2], in_depth = w.shape[w.shape.size() - 1];
        if (!need(tinfo_[st.in].c == in_depth, "FULLY_CONNECTED over flattened spatial dims")) return false;
        pack_pw_weights(w, out_depth, in_depth, st);
        if (O.in.size() > 2 && O.in[2] >= 0) { st.has_bias = true; st.b_off = add_blob(g_.tensors[O.in[2]].f32); }
        fold_unary(i, st);
        break;
      }
      case OP_DEPTHWISE_CONV_2D: {
        if (!need(O.depth_mult == 1, "depthwise multiplier != 1")) return false;
        if (!need(O.in.size() >= 2 && O.in[1] >= 0 && g_.tensors[O.in[1]].is_const, "DEPTHWISE_CONV_2D with non-constant weights")) return false;
        const GTensor& w = g_.tensors[O.in[1]];
        if (!need(w.shape.size() == 4 && w.f32.size() == w.count() && w.shape[3] == tinfo_[O.in[0]].c, "DEPTHWISE_CONV_2D weight shape")) return false;
        st.kind = Step::DW; st.in = O.in[0]; st.out = O.out; st.act1 = O.act;
        st.kh = w.shape[1]; st.kw = w.shape[2]; st.sh = O.stride_h; st.sw = O.stride_w; st.dh = O.dil_h; st.dw = O.dil_w;
        int o_h, o_w;
        conv_geom(tinfo_[st.in].h, st.kh, st.sh, st.dh, O.padding, &o_h, &st.pt);
        conv_geom(tinfo_[st.in].w, st.kw, st.sw, st.dw, O.padding, &o_w, &st.pl);
        if (!need(o_h == tinfo_[O.out].h && o_w == tinfo_[O.out].w, "DEPTHWISE_CONV_2D output shape")) return false;
        st.w_off = add_blob(w.f32);
        if (O.in.size() > 2 && O.in[2] >= 0) { st.has_bias = true; st.b_off = add_blob(g_.tensors[O.in[2]].f32); }
        fold_unary(i, st); fold_add(i, st);
        break;
      }
      case OP_AVERAGE_POOL_2D: {
        st.kind = Step::POOL; st.in = O.in[0]; st.out = O.out; st.act1 = O.act;
        int ph = tinfo_[st.in].h, pw = tinfo_[st.in].w, pc = tinfo_[st.in].c;
        if (!need(O.filter_h == ph && O.filter_w == pw && tinfo_[O.out].h == 1 && tinfo_[O.out].w == 1,
                  "AVERAGE_POOL_2D that is not global")) return false;
        if (!need(pc <= 512, "global pool over more than 512 channels")) return false;
        auto cp = concat_for_pool.find(st.in);
        if (cp != concat_for_pool.end()) { st.in = cp->second.first; st.in2 = cp->second.second; }
        rowsum_elems_ = std::max(rowsum_elems_, (size_t)ph * pc);
        // fold the squeeze-excite FC chain: pool -> {1x1 conv | FC} (+unary) [-> {1x1 conv | FC} (+unary)]
        int cur = st.out;
        while (st.n_fc < 2 && cur != g_.output && consumers[cur].size() == 1) {
          const int j = consumers[cur][0];
          const GOp& F = g_.ops[j];
          if (done[j] || j <= i || F.in.empty() || F.in[0] != cur || F.in.size() < 2 || !g_.tensors[F.in[1]].is_const) break;
          const GTensor& w = g_.tensors[F.in[1]];
          int K, N;
          if (is_pw(F)) { N = w.shape[0]; K = w.shape[3]; }
          else if (F.kind == OP_FULLY_CONNECTED) { N = w.shape[w.shape.size() - 2]; K = w.shape[w.shape.size() - 1]; }
          else break;
          // the FC weights are staged in shared memory by k_pool_fc: keep K x N4 inside the opt-in limit
          if (K != tinfo_[cur].c || K > 512 || N > 512 || (size_t)K * ((N + 3) / 4 * 4) * 4 > 192 * 1024 ||
              tinfo_[F.out].h != 1 || tinfo_[F.out].w != 1) break;
          Step tmp;
          pack_pw_weights(w, N, K, tmp);
          Step::Fc& fc = st.fc[st.n_fc++];
          fc.w_off = tmp.w_off; fc.K = K; fc.N = N; fc.n4 = tmp.n4; fc.act1 = F.act;
          if (F.in.size() > 2 && F.in[2] >= 0) { fc.has_bias = true; fc.b_off = add_blob(g_.tensors[F.in[2]].f32); }
          done[j] = 1; cur = F.out;
          if (consumers[cur].size() == 1) {
            const int u = consumers[cur][0];
            const int ua = unary_act(g_.ops[u].kind);
            if (ua >= 0 && !done[u] && u > j) { fc.act2 = ua; done[u] = 1; cur = g_.ops[u].out; }
          }
        }
        st.out = cur;
        break;
      }
      case OP_RESIZE_BILINEAR: {
        st.kind = Step::RESIZE; st.in = O.in[0]; st.out = O.out; st.align_corners = O.align_corners; st.half_pixel = O.half_pixel;
        if (!need(O.in.size() >= 2 && O.in[1] >= 0, "RESIZE_BILINEAR without a size tensor")) return false;
        const GTensor& sz = g_.tensors[O.in[1]];
        if (!need(sz.is_const && sz.i32.size() == 2 && sz.i32[0] == tinfo_[O.out].h && sz.i32[1] == tinfo_[O.out].w, "RESIZE_BILINEAR size")) return false;
        break;
      }
      case OP_CUSTOM: {
        if (!need(O.custom == "Convolution2DTransposeBias", "unknown custom op")) return false;
        if (!need(O.in.size() >= 3 && O.in[1] >= 0 && O.in[2] >= 0 && g_.tensors[O.in[1]].is_const && g_.tensors[O.in[2]].is_const,
                  "Convolution2DTransposeBias with non-constant weights / bias")) return false;
        const GTensor& w = g_.tensors[O.in[1]];
        if (!need(w.shape.size() == 4 && w.f32.size() == w.count() && w.shape[3] == tinfo_[O.in[0]].c &&
                  (int)g_.tensors[O.in[2]].f32.size() == w.shape[0], "Convolution2DTransposeBias weight shape")) return false;
        if (!need(w.shape[1] == 2 && w.shape[2] == 2 && O.stride_w == 2 && O.stride_h == 2 && O.tconv_same &&
                  tinfo_[O.out].h == 2 * tinfo_[O.in[0]].h && tinfo_[O.out].w == 2 * tinfo_[O.in[0]].w,
                  "Convolution2DTransposeBias other than k2 s2 SAME on even sizes")) return false;
        st.kind = Step::TCONV; st.in = O.in[0]; st.out = O.out; st.N = w.shape[0]; st.K = w.shape[3];
        st.w_off = add_blob(w.f32); st.has_bias = true; st.b_off = add_blob(g_.tensors[O.in[2]].f32);
        fold_unary(i, st);
        break;
      }
      case OP_MUL: {
        if (!need(O.in.size() == 2, "MUL arity")) return false;
        int x = O.in[0], s = O.in[1];
        if (g_.tensors[x].count() < g_.tensors[s].count()) std::swap(x, s);
        const bool bcast = g_.tensors[s].count() != g_.tensors[x].count();
        if (bcast && !need((int)g_.tensors[s].count() == tinfo_[x].c && !g_.tensors[s].is_const, "MUL broadcast other than per-channel")) return false;
        if (bcast && O.act == ACT_NONE && consumers[O.out].size() == 1 && O.out != g_.output) {
          const int j = consumers[O.out][0];
          const GOp& C = g_.ops[j];
          if (is_pw(C) && C.in[0] == O.out) { prologue[O.out] = Fold{x, s, -1}; done[i] = 1; continue; }
          if (C.kind == OP_ADD && C.act == ACT_NONE && C.in.size() == 2 && consumers[C.out].size() == 1 && C.out != g_.output) {
            const int other = C.in[0] == O.out ? C.in[1] : C.in[0];
            const GOp& P = g_.ops[consumers[C.out][0]];
            if (other != O.out && !g_.tensors[other].is_const && g_.tensors[other].count() == g_.tensors[x].count() &&
                is_pw(P) && P.in[0] == C.out && !prologue.count(other)) {
              prologue[C.out] = Fold{x, s, other}; done[i] = 1; done[j] = 1; continue;
            }
          }
        }
        st.kind = Step::ELT; st.in = x; st.out = O.out; st.act1 = O.act;
        if (bcast) { st.elt_mode = 3; st.scale = s; } else { st.elt_mode = 2; st.in2 = s; }
        break;
      }
      case OP_ADD: {
        if (!need(O.in.size() == 2 && g_.tensors[O.in[0]].count() == g_.tensors[O.in[1]].count(), "ADD with broadcasting")) return false;
        st.kind = Step::ELT; st.elt_mode = 1; st.in = O.in[0]; st.in2 = O.in[1]; st.out = O.out; st.act1 = O.act;
        break;
      }
      case OP_HARD_SWISH: case OP_RELU: case OP_RELU6: case OP_LOGISTIC:
        st.kind = Step::ELT; st.elt_mode = 0; st.in = O.in[0]; st.out = O.out; st.act1 = unary_act(O.kind);
        break;
      case OP_CONCATENATION: {
        const int rank = (int)g_.tensors[O.out].shape.size();
        const int axis = O.axis < 0 ? O.axis + rank : O.axis;
        if (!need(axis == rank - 1 && O.act == ACT_NONE, "CONCATENATION not on the channel axis")) return false;
        // (1) a concatenation that only feeds a global pool is never materialised: the pool's
        //     row-sum kernel reads both sources (pool(concat(a,b)) == concat(pool(a),pool(b)) exactly)
        if (O.in.size() == 2 && O.out != g_.output && consumers[O.out].size() == 1 &&
            g_.ops[consumers[O.out][0]].kind == OP_AVERAGE_POOL_2D && !prologue.count(O.in[0]) && !prologue.count(O.in[1])) {
          concat_for_pool[O.out] = std::make_pair(O.in[0], O.in[1]);
          continue;
        }
        // (2) otherwise producers write straight into the concatenated buffer when they can
        int off = 0;
        for (int t : O.in) {
          const bool can_alias = t != g_.input && consumers[t].size() == 1 && producer_step.count(t) && tinfo_[t].alias_parent < 0 &&
                                 off % 4 == 0 && tinfo_[O.out].c % 4 == 0 && !prologue.count(t);
          if (can_alias) {
            tinfo_[t].alias_parent = O.out; tinfo_[t].alias_off = off; tinfo_[t].ld = tinfo_[O.out].ld;
            tinfo_[t].frame_elems = tinfo_[O.out].frame_elems;
            aliased_into[O.out].push_back(t);
          } else {
            Step c; c.op_index = i; c.kind = Step::COPY; c.in = t; c.out = O.out; c.copy_off = off;
            steps_.push_back(c);
          }
          off += tinfo_[t].c;
        }
        if (aliased_into.count(O.out)) {   // a no-op marker step so the buffer counts as produced here
          bool any_copy = false;
          for (const Step& c : steps_) any_copy = any_copy || (c.kind == Step::COPY && c.out == O.out);
          if (!any_copy) concat_virtual.push_back(O.out);
        }
        continue;
      }
      default:
        *err = "unsupported TFLite operator code " + std::to_string(O.kind);
        return false;
    }
    producer_step[st.out] = (int)steps_.size();
    steps_.push_back(st);
  }

  // ---- fuse low-resolution inverted-residual blocks (expand -> depthwise -> SE -> project) into one kernel ----
  if ((flags_ & 8u) && !(flags_ & 1u)) {   // opt-in (BSB_FLAG_FUSE_BLOCKS); KEEP_TENSORS keeps every intermediate -> stand-alone kernels
    std::vector<int> readers(nt, 0);
    for (const Step& st : steps_) for (int t : {st.in, st.in2, st.scale, st.in_add, st.residual}) if (t >= 0) ++readers[t];
    std::vector<Step> fused;
    for (size_t i = 0; i < steps_.size(); ++i) {
      bool ok = i + 3 < steps_.size();
      if (ok) {
        const Step& e = steps_[i]; const Step& d = steps_[i + 1]; const Step& p = steps_[i + 2]; const Step& q = steps_[i + 3];
        ok = e.kind == Step::PW && !e.use_tc && e.scale < 0 && e.in_add < 0 && e.residual < 0 && e.out != g_.output &&
             d.kind == Step::DW && d.in == e.out && d.sh == 1 && d.sw == 1 && d.dh == 1 && d.dw == 1 && d.residual < 0 && d.kh == d.kw &&
             p.kind == Step::POOL && p.in == d.out && p.in2 < 0 && p.n_fc == 2 &&
             q.kind == Step::PW && !q.use_tc && q.in == d.out && q.scale == p.out && q.in_add < 0 && (q.residual < 0 || q.residual == e.in) &&
             readers[e.out] == 1 && readers[d.out] == 2 && readers[p.out] == 1 && p.out != g_.output && d.out != g_.output;
        if (ok) {
          const TensorInfo& X = tinfo_[e.in]; const TensorInfo& D = tinfo_[d.out];
          const int fc_max = std::max(std::max(p.fc[0].K * p.fc[0].n4, p.fc[1].K * p.fc[1].n4), q.K * q.n4);
          ok = X.h == D.h && X.w == D.w && p.fc[0].K == e.N && p.fc[1].N == e.N && p.fc[0].N <= 128 && e.N <= 128 &&
               (q.residual < 0 || X.c == q.N) && X.ld % 4 == 0 &&
               mb_block_smem_bytes(X.h, X.w, X.c, e.N, q.N, fc_max) > 0;
        }
      }
      if (!ok) { fused.push_back(steps_[i]); continue; }
      Step b; b.kind = Step::BLOCK; b.op_index = steps_[i].op_index; b.in = steps_[i].in; b.out = steps_[i + 3].out;
      b.residual = steps_[i + 3].residual; b.block = (int)blocks_.size();
      blocks_.push_back(FusedBlock{steps_[i], steps_[i + 1], steps_[i + 2], steps_[i + 3]});
      tinfo_[steps_[i].out].materialized = false; tinfo_[steps_[i + 1].out].materialized = false; tinfo_[steps_[i + 2].out].materialized = false;
      fused.push_back(b);
      i += 3;
    }
    steps_.swap(fused);
  }

  // ---- RESIZE_BILINEAR folded into the 1x1 conv that is its only reader ----
  if (tuning().up_pw) {
    for (size_t i = 0; i + 1 < steps_.size(); ++i) {
      const Step& r = steps_[i]; Step& q = steps_[i + 1];
      if (r.kind != Step::RESIZE || q.kind != Step::PW || q.in != r.out || q.scale >= 0 || q.in_add >= 0 || q.residual >= 0 || q.use_tc) continue;
      if (r.out == g_.output || consumers[r.out].size() != 1) continue;
      if (!upsample_pw_supported(q.K, q.N, q.n4, tinfo_[r.in].ld, tinfo_[q.out].ld)) continue;
      q.up_from = r.in; q.align_corners = r.align_corners; q.half_pixel = r.half_pixel;
      q.in = r.in;                                    // liveness: the conv now reads the low-resolution tensor
      tinfo_[r.out].materialized = false;
      steps_.erase(steps_.begin() + i);
    }
  }
  for (const Step& st : steps_) uses_tc_ = uses_tc_ || st.use_tc;
  // ---- the low-resolution middle of MobileNetV3-style graphs as one kernel ----
  if (tuning().cnn_chain && !(flags_ & (1u | 8u))) detect_chain();
  // ---- decoder stages (1x1 -> depthwise + residual [-> transposed conv]) as one kernel each ----
  if (tuning().head && !(flags_ & (1u | 8u))) detect_heads();
  // ---- encoder entry (16 -> 16 1x1 -> depthwise 3x3 stride 2) as one kernel ----
  if (tuning().pw_dws2 && !tuning().stem_pw && !(flags_ & (1u | 8u))) detect_pw_dws2();
  find_segments();
  dec_up_step_ = -1;
  if (tuning().dec_up && model_type_ == MODEL_DEEPLAB && !(flags_ & 1u) && !steps_.empty()) {
    const Step& last = steps_.back();
    if (last.kind == Step::RESIZE && last.out == g_.output && tinfo_[last.out].c == 21 && tinfo_[last.in].c == 21) dec_up_step_ = (int)steps_.size() - 1;
  }

  // ---- liveness + arena (floats; every tensor is max_batch frames) ----
  const int ns = (int)steps_.size();
  auto root = [&](int t) { return tinfo_[t].alias_parent >= 0 ? tinfo_[t].alias_parent : t; };
  auto use = [&](int t, int s) { if (t >= 0) { const int r = root(t); tinfo_[r].last_use = std::max(tinfo_[r].last_use, s); } };
  tinfo_[g_.input].materialized = true; tinfo_[g_.input].first_def = -1;
  for (int s = 0; s < ns; ++s) {
    const Step& st = steps_[s];
    use(st.in, s); use(st.in2, s); use(st.scale, s); use(st.in_add, s); use(st.residual, s);
    tinfo_[st.out].materialized = true;
    const int r = root(st.out);
    tinfo_[r].materialized = true;
    tinfo_[r].first_def = std::min(tinfo_[r].first_def, s);
    use(st.out, s);
  }
  // a sub-batched segment runs its steps once per frame group: every tensor it touches must stay allocated from the
  // segment's first step to its last (a buffer recycled inside the segment would be overwritten by group g while group
  // g + 1 still has to read it)
  for (int s0 = 0; s0 < ns; ++s0) {
    if (s0 >= (int)seg_len_.size() || seg_len_[s0] <= 0) continue;
    const int s1 = s0 + seg_len_[s0] - 1;
    for (int s = s0; s <= s1; ++s) {
      const Step& st = steps_[s];
      for (int t : {st.in, st.in2, st.scale, st.in_add, st.residual, st.up_from, st.out}) {
        if (t < 0) continue;
        const int r = root(t);
        tinfo_[r].last_use = std::max(tinfo_[r].last_use, s1);
        if (tinfo_[r].first_def > s0 && tinfo_[r].first_def <= s1) tinfo_[r].first_def = s0;
      }
    }
  }
  tinfo_[root(g_.output)].last_use = 1 << 30;
  tinfo_[g_.input].last_use = std::max(tinfo_[g_.input].last_use, 0);
  if (!tinfo_[g_.output].materialized) { *err = "graph output is never produced"; return false; }
  for (const Step& st : steps_)
    for (int t : {st.in, st.in2, st.scale, st.in_add, st.residual})
      if (t >= 0 && !tinfo_[t].materialized) { *err = "planner: step reads a tensor that was folded away"; return false; }

  struct Free { size_t off, size; };
  std::vector<Free> free_list;
  size_t top = 0;
  auto alloc = [&](size_t n) {
    n = (n + 63) / 64 * 64;
    for (size_t k = 0; k < free_list.size(); ++k)
      if (free_list[k].size >= n) {
        size_t off = free_list[k].off;
        free_list[k].off += n; free_list[k].size -= n;
        if (!free_list[k].size) free_list.erase(free_list.begin() + k);
        return off;
      }
    size_t off = top; top += n; return off;
  };
  auto release = [&](size_t off, size_t n) {
    n = (n + 63) / 64 * 64;
    free_list.push_back({off, n});
    std::sort(free_list.begin(), free_list.end(), [](const Free& a, const Free& b) { return a.off < b.off; });
    for (size_t k = 0; k + 1 < free_list.size();)
      if (free_list[k].off + free_list[k].size == free_list[k + 1].off) { free_list[k].size += free_list[k + 1].size; free_list.erase(free_list.begin() + k + 1); }
      else ++k;
  };
  const bool keep = (flags_ & 1u) != 0;
  std::vector<char> allocated(nt, 0);
  tinfo_[g_.input].offset = alloc(tinfo_[g_.input].frame_elems * max_batch_); allocated[g_.input] = 1;
  for (int s = 0; s < ns; ++s) {
    const int o = root(steps_[s].out);
    if (!allocated[o]) { tinfo_[o].offset = alloc(tinfo_[o].frame_elems * max_batch_); allocated[o] = 1; }
    if (keep) continue;
    for (int t = 0; t < nt; ++t)
      if (allocated[t] == 1 && tinfo_[t].alias_parent < 0 && tinfo_[t].materialized && tinfo_[t].last_use == s) {
        release(tinfo_[t].offset, tinfo_[t].frame_elems * max_batch_); allocated[t] = 2;
      }
  }
  arena_elems_ = top;

  {
    const Step& s0 = steps_[0];
    stem_u8_ok_ = s0.kind == Step::CONV && s0.in == g_.input && s0.K == 3 && s0.N == 16 && s0.dh == 1 && s0.dw == 1 &&
                  consumers[g_.input].size() == 1 && tinfo_[s0.out].ld % 4 == 0 && s0.residual < 0;
    stem_pw_ok_ = false;
    if (stem_u8_ok_ && steps_.size() > 1) {
      const Step& s1 = steps_[1];
      stem_pw_ok_ = s1.kind == Step::PW && s1.in == s0.out && s1.K == 16 && s1.N == 16 && s1.n4 == 16 && s1.scale < 0 && s1.in_add < 0 &&
                    s1.residual < 0 && !s1.use_tc && tinfo_[s1.out].ld % 4 == 0;
    }
  }
  // algorithmic FLOPs (2*MAC of conv / depthwise / fc / tconv), SURVEY.md Appendix A
  flops_ = 0;
  for (const GOp& O : g_.ops) {
    if (O.out < 0 || O.in.size() < 2) continue;
    const GTensor& out = g_.tensors[O.out]; const GTensor& w = g_.tensors[O.in[1]];
    if (O.kind == OP_CONV_2D) flops_ += 2.0 * out.count() * w.shape[1] * w.shape[2] * w.shape[3];
    else if (O.kind == OP_DEPTHWISE_CONV_2D) flops_ += 2.0 * out.count() * w.shape[1] * w.shape[2];
    else if (O.kind == OP_FULLY_CONNECTED) flops_ += 2.0 * out.count() * w.shape.back();
    else if (O.kind == OP_CUSTOM) flops_ += 2.0 * out.count() * w.shape[3];
  }
  return true;
}

void Engine::detect_heads() {
  auto reads = [&](const Step& r, int t) { return r.in == t || r.in2 == t || r.scale == t || r.in_add == t || r.residual == t; };
  for (size_t i = 0; i + 1 < steps_.size(); ++i) {
    const Step& P = steps_[i]; const Step& D = steps_[i + 1];
    if (P.kind != Step::PW || P.use_tc || P.up_from >= 0 || P.residual >= 0 || P.K != P.N || (P.N != 16 && P.N != 24) || P.n4 != P.N) continue;
    if (D.kind != Step::DW || D.in != P.out || D.residual != P.out || D.kh != 3 || D.kw != 3 || D.sh != 1 || D.sw != 1 || D.dh != 1 || D.dw != 1) continue;
    if (tinfo_[D.out].c != P.N || P.out == g_.output || D.out == g_.output) continue;
    bool only_d = true;                                         // t (the 1x1 output) must be private to the depthwise step
    for (size_t k = 0; k < steps_.size(); ++k) if (k != i + 1 && reads(steps_[k], P.out)) only_d = false;
    if (!only_d) continue;
    HeadPlan hp; hp.p = P; hp.d = D;
    if (i + 2 < steps_.size()) {
      const Step& T = steps_[i + 2];
      bool only_t = T.kind == Step::TCONV && T.in == D.out && T.K == P.N && T.N <= 2 && tinfo_[T.out].ld == T.N;
      for (size_t k = 0; k < steps_.size() && only_t; ++k) if (k != i + 2 && reads(steps_[k], D.out)) only_t = false;
      if (only_t) { hp.t = T; hp.has_t = true; }
    }
    const int ld_add = P.in_add >= 0 ? tinfo_[P.in_add].ld : 4;
    const int out_t = hp.has_t ? hp.t.out : D.out;
    if (!head_supported(P.N, tinfo_[P.in].ld, ld_add, tinfo_[out_t].ld, hp.has_t ? hp.t.N : 0, hp.has_t)) continue;
    Step h; h.kind = Step::HEAD; h.op_index = P.op_index; h.in = P.in; h.scale = P.scale; h.in_add = P.in_add; h.out = out_t;
    h.block = (int)heads_.size();
    heads_.push_back(hp);
    tinfo_[P.out].materialized = false;
    if (hp.has_t) tinfo_[D.out].materialized = false;
    const size_t n_erase = hp.has_t ? 3 : 2;
    steps_.erase(steps_.begin() + i, steps_.begin() + i + n_erase);
    steps_.insert(steps_.begin() + i, h);
  }
}

void Engine::detect_pw_dws2() {
  auto reads = [&](const Step& r, int t) { return r.in == t || r.in2 == t || r.scale == t || r.in_add == t || r.residual == t; };
  for (size_t i = 0; i + 1 < steps_.size(); ++i) {
    const Step& P = steps_[i]; const Step& D = steps_[i + 1];
    if (P.kind != Step::PW || P.use_tc || P.up_from >= 0 || P.residual >= 0 || P.scale >= 0 || P.in_add >= 0 || P.K != 16 || P.N != 16 || P.n4 != 16) continue;
    if (D.kind != Step::DW || D.in != P.out || D.residual >= 0 || D.kh != 3 || D.kw != 3 || D.sh != 2 || D.sw != 2 || D.dh != 1 || D.dw != 1) continue;
    if ((P.act1 != ACT_NONE && P.act2 != ACT_NONE) || (D.act1 != ACT_NONE && D.act2 != ACT_NONE)) continue;
    if (tinfo_[D.out].c != 16 || P.out == g_.output || D.out == g_.output) continue;
    bool only_d = true;                                         // the 1x1 output must be private to the depthwise step
    for (size_t k = 0; k < steps_.size(); ++k) if (k != i + 1 && reads(steps_[k], P.out)) only_d = false;
    if (!only_d || !pw_dws2_supported(P.K, P.N, tinfo_[P.in].ld, tinfo_[D.out].ld)) continue;
    HeadPlan hp; hp.p = P; hp.d = D; hp.s2 = true;
    Step h; h.kind = Step::HEAD; h.op_index = P.op_index; h.in = P.in; h.out = D.out;
    h.block = (int)heads_.size();
    heads_.push_back(hp);
    tinfo_[P.out].materialized = false;
    steps_.erase(steps_.begin() + i, steps_.begin() + i + 2);
    steps_.insert(steps_.begin() + i, h);
  }
}

// Find   DW(strided, from a larger tensor) -> SE -> PW(scaled)  { -> PW(expand) -> DW -> SE -> PW(scaled [+ residual]) }*
//        -> PW -> SE(1 FC, of the narrow tensor) -> MUL(channel scale)
// on tensors of <= 256 pixels and <= 128 channels, with every intermediate read only inside the run, and replace it
// by one CHAIN step (kernels_chain.cu).  Same arithmetic per output element, so results do not change.
bool Engine::detect_chain() {
  const int ns = (int)steps_.size();
  auto px = [&](int t) { return tinfo_[t].h * tinfo_[t].w; };
  auto plain_pw = [&](const Step& s) {
    return s.kind == Step::PW && !s.use_tc && s.in_add < 0 && s.K % 4 == 0 && (size_t)s.K * s.n4 <= 4096 && s.K <= 128 && s.N <= 128;
  };
  auto se_ok = [&](const Step& s, int in, int n_fc, int C) {
    if (s.kind != Step::POOL || s.in != in || s.in2 >= 0 || s.n_fc != n_fc || C > 128) return false;
    for (int k = 0; k < n_fc; ++k) if (s.fc[k].K > 128 || s.fc[k].N > 128) return false;
    return s.fc[0].K == C;
  };
  for (int i = 0; i + 6 < ns; ++i) {
    const Step& d0 = steps_[i];
    if (d0.kind != Step::DW || d0.residual >= 0 || d0.dh != 1 || d0.dw != 1 || d0.kh != d0.kw || (d0.kh != 3 && d0.kh != 5) || d0.sh != d0.sw) continue;
    const TensorInfo& T0 = tinfo_[d0.out];
    const int h = T0.h, w = T0.w, P = h * w;
    if (P > 256 || px(d0.in) <= P || T0.c > 128 || T0.c % 4 || tinfo_[d0.in].ld % 4 || chain_smem_bytes(h, w) == 0) continue;
    std::vector<int> types;
    int j = i + 1, curD = d0.out, curX = -1;
    types.push_back(CH_DWG);
    // entry: SE of D, scaled project into X
    if (!(j + 1 < ns && se_ok(steps_[j], curD, 2, T0.c) && plain_pw(steps_[j + 1]) && steps_[j + 1].in == curD && steps_[j + 1].scale == steps_[j].out &&
          steps_[j + 1].residual < 0 && steps_[j + 1].N <= 32)) continue;
    types.push_back(CH_SE); types.push_back(CH_PW);
    curX = steps_[j + 1].out; j += 2;
    // inverted-residual blocks
    while (j + 3 < ns) {
      const Step& e = steps_[j]; const Step& dd = steps_[j + 1]; const Step& p = steps_[j + 2]; const Step& q = steps_[j + 3];
      const bool ok = plain_pw(e) && e.in == curX && e.scale < 0 && e.residual < 0 && e.K <= 32 &&
                      dd.kind == Step::DW && dd.in == e.out && dd.sh == 1 && dd.sw == 1 && dd.dh == 1 && dd.dw == 1 && dd.residual < 0 &&
                      dd.kh == dd.kw && (dd.kh == 3 || dd.kh == 5) && px(dd.out) == P &&
                      se_ok(p, dd.out, 2, e.N) && plain_pw(q) && q.in == dd.out && q.scale == p.out && q.N <= 32 &&
                      (q.residual < 0 || q.residual == curX);
      if (!ok) break;
      types.push_back(CH_EXPAND_DW); types.push_back(-1); types.push_back(CH_SE); types.push_back(CH_PW);
      curX = q.out; curD = dd.out; j += 4;
    }
    // exit: PW of X into D and SE (one FC) of X, in either graph order, then D * sv -> global
    if (!(j + 2 < ns)) continue;
    const bool pw_first = steps_[j].kind == Step::PW;
    const Step& t0 = steps_[pw_first ? j : j + 1]; const Step& t1 = steps_[pw_first ? j + 1 : j]; const Step& t2 = steps_[j + 2];
    if (!(plain_pw(t0) && t0.in == curX && t0.scale < 0 && t0.residual < 0 && t0.N % 4 == 0 &&
          se_ok(t1, curX, 1, tinfo_[curX].c) && t1.fc[0].N == t0.N &&
          t2.kind == Step::ELT && t2.elt_mode == 3 && t2.in == t0.out && t2.scale == t1.out && tinfo_[t2.out].ld % 4 == 0)) continue;
    if (pw_first) { types.push_back(CH_PW | 256); types.push_back(CH_SE | 256); } else { types.push_back(CH_SE | 256); types.push_back(CH_PW | 256); }
    types.push_back(CH_SCALE_STORE);
    const int jend = j + 2;
    // every tensor produced inside the run (except the last) must be read only inside it
    bool contained = true;
    for (int a = i; a < jend && contained; ++a) {
      const int t = steps_[a].out;
      for (int b = 0; b < ns && contained; ++b) {
        if (b >= i && b <= jend) continue;
        const Step& r = steps_[b];
        for (int u : {r.in, r.in2, r.scale, r.in_add, r.residual}) if (u == t) contained = false;
      }
      if (t == g_.output) contained = false;
    }
    if (!contained) continue;
    ChainPlan cp; cp.h = h; cp.w = w;
    cp.seq.assign(steps_.begin() + i, steps_.begin() + jend + 1);
    cp.types = types;
    for (int a = i; a < jend; ++a) tinfo_[steps_[a].out].materialized = false;
    Step c; c.kind = Step::CHAIN; c.op_index = d0.op_index; c.in = d0.in; c.out = steps_[jend].out; c.block = (int)chains_.size();
    chains_.push_back(cp);
    steps_.erase(steps_.begin() + i, steps_.begin() + jend + 1);
    steps_.insert(steps_.begin() + i, c);
    return true;                                   // one chain per graph
  }
  return false;
}

// ---------------------------------------------------------------------------
// create / destroy
// ---------------------------------------------------------------------------
Engine* Engine::create(const std::string& model_path, int width, int height, int device, int max_batch,
                       unsigned flags, const Callbacks& cb, std::string* err) {
  Engine* e = new Engine();
  e->cb_ = cb; e->W_ = width; e->H_ = height; e->device_ = device; e->max_batch_ = std::max(1, max_batch); e->flags_ = flags;
  auto fail = [&](const std::string& m) { *err = m; delete e; return (Engine*)nullptr; };
  if (width <= 0 || height <= 0 || width > 16384 || height > 16384) return fail("invalid frame size");
  if (!load_tflite(model_path, &e->g_, err)) { delete e; return nullptr; }
  // lib/libbackscrub.cc:194-200
  e->model_type_ = model_type_from_name(model_path);
  if (e->model_type_ == MODEL_UNKNOWN) return fail("unknown model type '" + model_path + "'.");
  // lib/libbackscrub.cc:132-148
  if (e->model_type_ == MODEL_DEEPLAB) { e->scaling_ = (float)(1 / 127.5); e->offset_ = -1.f; }
  else { e->scaling_ = (float)(1 / 255.0); e->offset_ = 0.f; }
  const GTensor& in = e->g_.tensors[e->g_.input]; const GTensor& out = e->g_.tensors[e->g_.output];
  // lib/libbackscrub.cc:85-112: float32, batch 1
  if (in.dtype != 0 || out.dtype != 0) return fail("error: input/output tensor is not float32 type");
  if (in.shape.size() != 4 || out.shape.size() != 4 || in.shape[0] != 1 || out.shape[0] != 1) return fail("error: input/output tensor is not single vector");
  e->mh_ = in.shape[1]; e->mw_ = in.shape[2];
  if (in.shape[3] != 3) return fail("model input must have 3 channels");
  e->oh_ = out.shape[1]; e->ow_ = out.shape[2]; e->oc_ = out.shape[3];
  const int need_oc = e->model_type_ == MODEL_DEEPLAB ? 21 : (e->model_type_ == MODEL_MEET ? 2 : 1);
  if (e->oc_ != need_oc) return fail("model output channel count does not match the model family");
  // lib/libbackscrub.cc:231-246 (float arithmetic, truncating conversion)
  const float ratio = (float)e->mh_ / (float)e->mw_, frameratio = (float)height / (float)width;
  if (frameratio < ratio) {
    e->roidim_[0] = (int)(((float)width - (float)height / ratio) / 2); e->roidim_[1] = 0;
    e->roidim_[2] = (int)((float)height / ratio); e->roidim_[3] = height;
    e->in_roidim_[0] = 0; e->in_roidim_[1] = 0; e->in_roidim_[2] = e->mw_; e->in_roidim_[3] = e->mh_;
  } else {
    e->roidim_[0] = 0; e->roidim_[1] = 0; e->roidim_[2] = width; e->roidim_[3] = height;
    e->in_roidim_[0] = (int)(((float)e->mw_ - (float)e->mh_ / frameratio) / 2); e->in_roidim_[1] = 0;
    e->in_roidim_[2] = (int)((float)e->mh_ / frameratio); e->in_roidim_[3] = e->mh_;
  }
  // documented deviation for models whose output grid differs from the input grid (body-pix):
  // the reference slices ofinal with in_roidim and throws; scale the rectangle by out/in.
  if (e->oh_ == e->mh_ && e->ow_ == e->mw_) std::memcpy(e->out_roidim_, e->in_roidim_, sizeof(e->in_roidim_));
  else {
    e->out_roidim_[0] = e->in_roidim_[0] * e->ow_ / e->mw_; e->out_roidim_[1] = e->in_roidim_[1] * e->oh_ / e->mh_;
    e->out_roidim_[2] = e->in_roidim_[2] * e->ow_ / e->mw_; e->out_roidim_[3] = e->in_roidim_[3] * e->oh_ / e->mh_;
  }
  for (int k = 2; k < 4; ++k)
    if (e->roidim_[k] <= 0 || e->in_roidim_[k] <= 0 || e->out_roidim_[k] <= 0) return fail("degenerate ROI for this frame size");
  if (e->roidim_[0] < 0 || e->roidim_[0] + e->roidim_[2] > width || e->in_roidim_[0] < 0 || e->in_roidim_[0] + e->in_roidim_[2] > e->mw_)
    return fail("ROI outside the frame");
  if (!e->plan(err)) { delete e; return nullptr; }
  if (!e->upload(err)) { delete e; return nullptr; }
  return e;
}

bool Engine::upload(std::string* err) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { *err = "no CUDA device available (this library has no CPU path)"; return false; }
  if (device_ < 0 || device_ >= ndev) { *err = "CUDA device ordinal out of range"; return false; }
  CUDA_OK(cudaSetDevice(device_));
  CUDA_OK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  const size_t B = (size_t)max_batch_;
  CUDA_OK(cudaMalloc((void**)&wblob_, std::max<size_t>(wblob_h_.size(), 64) * 4));
  CUDA_OK(cudaMemcpy(wblob_, wblob_h_.data(), wblob_h_.size() * 4, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMalloc((void**)&arena_, std::max<size_t>(arena_elems_, 64) * 4));
  CUDA_OK(cudaMemset(arena_, 0, std::max<size_t>(arena_elems_, 64) * 4));
  CUDA_OK(cudaMalloc((void**)&rowsum_, std::max<size_t>(rowsum_elems_ * B, 64) * 4));
  CUDA_OK(cudaMalloc((void**)&pool_counters_, B * sizeof(unsigned)));
  CUDA_OK(cudaMemset(pool_counters_, 0, B * sizeof(unsigned)));
  // chain op lists (device pointers into the weight blob / arena are final from here on)
  for (ChainPlan& cp : chains_) {
    std::vector<ChainOp> ops;
    const int ne = (int)cp.seq.size();
    auto fc_of = [&](const Step::Fc& f) {
      FcLayer l; l.w = wblob_ + f.w_off; l.bias = f.has_bias ? wblob_ + f.b_off : nullptr; l.K = f.K; l.N = f.N; l.n4 = f.n4; l.act1 = f.act1; l.act2 = f.act2;
      return l;
    };
    for (int k = 0; k < ne; ++k) {
      if (cp.types[k] < 0) continue;
      const int ty = cp.types[k] & 255;
      const bool tail = (cp.types[k] & 256) != 0;           // the exit pair works on the narrow tensor X
      const Step& st = cp.seq[k];
      ChainOp o{};
      o.type = ty;
      switch (ty) {
        case CH_DWG:
          o.cin = tinfo_[st.out].c; o.wd = wblob_ + st.w_off; o.bd = st.has_bias ? wblob_ + st.b_off : nullptr;
          o.k = st.kh; o.s = st.sh; o.pt = st.pt; o.pl = st.pl; o.dact1 = st.act1; o.dact2 = st.act2;
          o.gin = tptr(st.in); o.gin_ld = tinfo_[st.in].ld; o.ih = tinfo_[st.in].h; o.iw = tinfo_[st.in].w; o.gin_frame = tinfo_[st.in].frame_elems;
          break;
        case CH_SE:
          o.src = tail ? 0 : 1; o.cin = st.fc[0].K; o.pool_act = st.act1; o.n_fc = st.n_fc;
          o.f0 = fc_of(st.fc[0]); if (st.n_fc > 1) o.f1 = fc_of(st.fc[1]);
          break;
        case CH_PW: {
          o.src = tail ? 0 : 1; o.dst = tail ? 1 : 0; o.cin = st.K; o.cout = st.N; o.w = wblob_ + st.w_off; o.b = st.has_bias ? wblob_ + st.b_off : nullptr;
          o.n4 = st.n4; o.act1 = st.act1; o.act2 = st.act2; o.use_scale = st.scale >= 0 ? 1 : 0; o.residual = st.residual >= 0 ? 1 : 0; o.act3 = st.act3;
          break;
        }
        case CH_EXPAND_DW: {
          const Step& dd = cp.seq[k + 1];
          o.cin = st.K; o.cout = st.N; o.w = wblob_ + st.w_off; o.b = st.has_bias ? wblob_ + st.b_off : nullptr; o.n4 = st.n4; o.act1 = st.act1; o.act2 = st.act2;
          o.wd = wblob_ + dd.w_off; o.bd = dd.has_bias ? wblob_ + dd.b_off : nullptr; o.k = dd.kh; o.s = 1; o.pt = dd.pt; o.pl = dd.pl;
          o.dact1 = dd.act1; o.dact2 = dd.act2;
          break;
        }
        case CH_SCALE_STORE:
          o.cin = tinfo_[st.in].c; o.act1 = st.act1; o.gout = tptr(st.out); o.gout_ld = tinfo_[st.out].ld; o.gout_frame = tinfo_[st.out].frame_elems;
          break;
      }
      ops.push_back(o);
    }
    cp.n_ops = (int)ops.size();
    CUDA_OK(cudaMalloc((void**)&cp.d_ops, ops.size() * sizeof(ChainOp)));
    CUDA_OK(cudaMemcpy(cp.d_ops, ops.data(), ops.size() * sizeof(ChainOp), cudaMemcpyHostToDevice));
  }
  // bilateral LUTs (cv::bilateralFilter d=5, sigma 100/100; oracle_img.c:or_bilateral_d5_u8c3)
  {
    std::vector<float> lut(768 + 16, 0.f);
    const double gc = -0.5 / (100.0 * 100.0), gs = -0.5 / (100.0 * 100.0);
    for (int i = 0; i < 768; ++i) lut[i] = (float)std::exp((double)i * i * gc);
    int k = 0;
    for (int i = -2; i <= 2; ++i) for (int j = -2; j <= 2; ++j) {
      const double r = std::sqrt((double)i * i + (double)j * j);
      if (r > 2) continue;
      lut[768 + k++] = (float)std::exp(r * r * gs);
    }
    CUDA_OK(cudaMalloc((void**)&lut_, lut.size() * 4));
    CUDA_OK(cudaMemcpy(lut_, lut.data(), lut.size() * 4, cudaMemcpyHostToDevice));
  }
  const size_t in_px = (size_t)mh_ * mw_, out_px = (size_t)oh_ * ow_, fpx = (size_t)W_ * H_;
  CUDA_OK(cudaMalloc((void**)&in_u8_, B * in_px * 3));
  CUDA_OK(cudaMemset(in_u8_, 0, B * in_px * 3));      // zero padding outside in_roidim stays zero
  CUDA_OK(cudaMalloc((void**)&filt_u8_, B * in_px * 3));
  CUDA_OK(cudaMalloc((void**)&state_, out_px));
  CUDA_OK(cudaMemset(state_, 0, out_px));
  opitch_ = (ow_ + 15) / 16 * 16;
  CUDA_OK(cudaMalloc((void**)&ofinal_, B * (size_t)oh_ * opitch_));
  CUDA_OK(cudaMemset(ofinal_, 0, B * (size_t)oh_ * opitch_));
  CUDA_OK(cudaMalloc((void**)&d_frames_, B * fpx * 3));
  CUDA_OK(cudaMalloc((void**)&d_out_, B * fpx * 3));
  CUDA_OK(cudaMalloc((void**)&d_yuyv_, B * fpx * 2));
  CUDA_OK(cudaMalloc((void**)&d_mask_, B * fpx));
  CUDA_OK(cudaMalloc((void**)&d_yuyv_in_, B * fpx * 2));
  {
    // app/deepseg.cc:603: until a background is set the frame is blended over plain green
    std::vector<uint8_t> green(fpx * 3);
    for (size_t i = 0; i < fpx; ++i) { green[3 * i] = 0; green[3 * i + 1] = 255; green[3 * i + 2] = 0; }
    CUDA_OK(cudaMalloc((void**)&d_bg_, fpx * 3));
    CUDA_OK(cudaMemcpy(d_bg_, green.data(), fpx * 3, cudaMemcpyHostToDevice));
    bg_cap_ = fpx * 3;
    CUDA_OK(cudaMalloc((void**)&d_bg_cursor_, sizeof(int)));
    CUDA_OK(cudaMemset(d_bg_cursor_, 0, sizeof(int)));
  }
  out_w_ = W_; out_h_ = H_;
  out_cap_ = B * fpx * 3; yuyv_cap_ = B * fpx * 2;
  CUDA_OK(cudaMallocHost((void**)&h_mask_, fpx));
  if (!upload_resize_tab(build_resize_tab(roidim_[2], roidim_[3], in_roidim_[2], in_roidim_[3]), &tab_in_, err)) return false;
  {
    const HostResizeTab up = build_resize_tab(out_roidim_[2], out_roidim_[3], roidim_[2], roidim_[3]);
    if (!upload_resize_tab(up, &tab_up_, err)) return false;
    // patch geometry of the post stage's tiles (k_post_tma): yofs / xofs are monotonic, so the extremes of a halo tile
    // give the rows / columns of the small mask it touches
    const int roi_x = roidim_[0], roi_y = roidim_[1], roi_w = roidim_[2], roi_h = roidim_[3], sw = out_roidim_[2];
    auto span = [&](int lo, int len, int n, const std::vector<int>& first, const std::vector<int>& last, int last_plus, int last_max) {
      const int hi = lo + len - 1;
      const int i0 = lo < 0 ? 0 : std::min(lo, n - 1), i1 = hi >= n ? n - 1 : std::max(hi, 0);
      const int a0 = first[i0], a1 = std::min(last[i1] + last_plus, last_max);
      return make_int2(a0, a1 - a0 + 1);
    };
    geo_nty_ = (H_ + 31) / 32; geo_ntx64_ = (W_ + 63) / 64;
    const int ntx128 = (W_ + 127) / 128;
    std::vector<int2> geo((size_t)geo_nty_ + geo_ntx64_ + ntx128);
    for (int ty = 0; ty < geo_nty_; ++ty) geo[ty] = span(ty * 32 - roi_y - 2, 36, roi_h, up.yofs0, up.yofs1, 0, 1 << 30);
    for (int tx = 0; tx < geo_ntx64_; ++tx) geo[geo_nty_ + tx] = span(tx * 64 - roi_x - 2, 68, roi_w, up.xofs, up.xofs, 1, sw - 1);
    for (int tx = 0; tx < ntx128; ++tx) geo[geo_nty_ + geo_ntx64_ + tx] = span(tx * 128 - roi_x - 2, 132, roi_w, up.xofs, up.xofs, 1, sw - 1);
    CUDA_OK(cudaMalloc((void**)&d_tile_geo_, geo.size() * sizeof(int2)));
    CUDA_OK(cudaMemcpy(d_tile_geo_, geo.data(), geo.size() * sizeof(int2), cudaMemcpyHostToDevice));
  }
  if (!refresh_bg_yuyv(err)) return false;
  CUDA_OK(cudaDeviceSynchronize());
  return true;
}

Engine::~Engine() {
#ifndef BSB_EMU
  for (auto& kv : graphs_) cudaGraphExecDestroy(kv.second);
#endif
  if (stream_) { cudaStreamSynchronize(stream_); }
  for (void* p : {(void*)wblob_, (void*)arena_, (void*)lut_, (void*)rowsum_, (void*)in_u8_, (void*)filt_u8_, (void*)state_, (void*)ofinal_,
                  (void*)d_frames_, (void*)d_out_, (void*)d_yuyv_, (void*)d_mask_, (void*)d_bg_, (void*)d_bg_raw_, (void*)d_yuyv_in_,
                  tab_in_.blob, tab_up_.blob, tab_bg_.blob, tab_out_.blob, (void*)d_bg_cursor_, (void*)d_bg_eff_, (void*)d_bg_frames_,
                  (void*)d_gauss_tmp_, (void*)d_stage_a_, (void*)d_stage_b_, (void*)d_stage_c_, (void*)d_bg_yuyv_, (void*)pool_counters_, (void*)d_tile_geo_})
    if (p) cudaFree(p);
  for (ChainPlan& cp : chains_) if (cp.d_ops) cudaFree(cp.d_ops);
  if (h_mask_) cudaFreeHost(h_mask_);
  if (stream_) cudaStreamDestroy(stream_);
}

// ---------------------------------------------------------------------------
// enqueue: the per-call kernel sequence (captured into a CUDA graph by run())
// ---------------------------------------------------------------------------
void Engine::enqueue_pre(int n, const uint8_t* d_frames, size_t pitch, size_t stride, const uint8_t* d_yuyv_in) {
  // camera YUYV frames are read in place (each bilinear tap converted like cv::COLOR_YUV2BGR_YUYV would have)
  if (d_yuyv_in)
    launch_resize_roi_swap(stream_, n, d_yuyv_in, (size_t)W_ * H_ * 2, (size_t)W_ * 2, roidim_[0], roidim_[1], roidim_[2], roidim_[3], tab_in_.tab,
                           in_u8_, mw_, mh_, in_roidim_[0], in_roidim_[1], in_roidim_[2], in_roidim_[3], tab_in_.area2x2, true);
  else
    launch_resize_roi_swap(stream_, n, d_frames, stride, pitch, roidim_[0], roidim_[1], roidim_[2], roidim_[3], tab_in_.tab,
                           in_u8_, mw_, mh_, in_roidim_[0], in_roidim_[1], in_roidim_[2], in_roidim_[3], tab_in_.area2x2);
  // with the fused stem the fp32 input tensor is only materialised for introspection (KEEP_TENSORS)
  float* f32 = (stem_u8_ok_ && !(flags_ & 1u)) ? nullptr : tptr(g_.input);
  launch_bilateral_norm(stream_, n, in_u8_, mw_, mh_, lut_, lut_ + 768, scaling_, offset_, f32, filt_u8_);
}

// Frames per pass of a sub-batched segment: the segment's largest tensor, in and out, should stay L2-resident
// (tuning().sub_batch_mb megabytes per tensor; 0 = off).
static int sub_batch_frames(size_t frame_bytes, int n) {
  const size_t budget = (size_t)tuning().sub_batch_mb << 20;
  if (!budget || !frame_bytes) return n;
  const size_t f = std::max<size_t>(1, budget / frame_bytes);
  return (int)std::min<size_t>(f, (size_t)n);
}

void Engine::enqueue_cnn(int n, bool from_u8) {
  bool first = true;
  bool skip_next = false;
  for (size_t si = 0; si < steps_.size(); ++si) {
    if (si < seg_len_.size() && seg_len_[si] > 0) {
      const int sub = sub_batch_frames(seg_frame_bytes_[si], n);
      if (sub < n) {
        const size_t end = si + (size_t)seg_len_[si];
        for (int f0 = 0; f0 < n; f0 += sub) {
          frame_off_ = f0;
          for (size_t k = si; k < end; ++k) run_step(k, std::min(sub, n - f0), from_u8, &first, &skip_next);
        }
        frame_off_ = 0;
        si = end - 1;
        continue;
      }
    }
    run_step(si, n, from_u8, &first, &skip_next);
  }
}

// Segments: maximal runs of plain per-frame steps (1x1 conv, depthwise, element-wise, copy, resize) that touch a tensor
// of more than 1 MB per frame.  Everything in such a run depends only on its own frame, so the run can be executed
// frame-group by frame-group.
void Engine::find_segments() {
  seg_len_.assign(steps_.size(), 0);
  seg_frame_bytes_.assign(steps_.size(), 0);
  if (tuning().sub_batch_mb <= 0) return;       // off: no segments, the arena recycles buffers step by step as before
  auto plain = [&](const Step& st) {
    return st.kind == Step::PW || st.kind == Step::DW || st.kind == Step::ELT || st.kind == Step::COPY || st.kind == Step::RESIZE;
  };
  auto fbytes = [&](const Step& st) {
    size_t m = 0;
    for (int t : {st.in, st.in2, st.in_add, st.residual, st.out, st.up_from})
      if (t >= 0) m = std::max(m, tinfo_[tinfo_[t].alias_parent >= 0 ? tinfo_[t].alias_parent : t].frame_elems * sizeof(float));
    return m;
  };
  for (size_t i = 0; i < steps_.size();) {
    if (!plain(steps_[i])) { ++i; continue; }
    size_t j = i, big = 0;
    while (j < steps_.size() && plain(steps_[j])) { big = std::max(big, fbytes(steps_[j])); ++j; }
    if (big > ((size_t)1 << 20) && j - i >= 2) { seg_len_[i] = (int)(j - i); seg_frame_bytes_[i] = big; }
    i = j;
  }
}

void Engine::run_step(size_t si, int n, bool from_u8, bool* first_p, bool* skip_next_p) {
  bool& first = *first_p;
  bool& skip_next = *skip_next_p;
  {
    const Step& st = steps_[si];
    if (skip_next) { skip_next = false; return; }        // the 1x1 conv that ran inside the stem kernel
    if (from_u8 && (int)si == dec_up_step_) return;      // pipeline call: the decision kernel interpolates on the fly
    const bool fused_stem = first && from_u8 && stem_u8_ok_;
    first = false;
    const TensorInfo& I = tinfo_[st.in];
    const TensorInfo& O = tinfo_[st.out];
    Epilogue e;
    e.bias = st.has_bias ? wblob_ + st.b_off : nullptr;
    e.act1 = st.act1; e.act2 = st.act2; e.act3 = st.act3;
    if (st.residual >= 0) { e.residual = tptr(st.residual); e.ld_res = tinfo_[st.residual].ld; }
    switch (st.kind) {
      case Step::CONV:
        if (fused_stem) {
          // the thread that produced a stem pixel holds its 16 channels: a following plain 16 -> 16 1x1 conv runs there too
          const Step* pw = (stem_pw_ok_ && tuning().stem_pw && si + 1 < steps_.size()) ? &steps_[si + 1] : nullptr;
          if (pw) {
            Epilogue e2;
            e2.bias = pw->has_bias ? wblob_ + pw->b_off : nullptr; e2.act1 = pw->act1; e2.act2 = pw->act2;
            launch_stem_u8(stream_, n, filt_u8_, I.h, I.w, scaling_, offset_, wblob_ + st.w_off, st.kh, st.kw, st.sh, st.sw, st.pt, st.pl,
                           tptr(st.out), O.h, O.w, O.ld, e, wblob_ + pw->w_off, tptr(pw->out), tinfo_[pw->out].ld, &e2);
            skip_next = true;
          } else {
            launch_stem_u8(stream_, n, filt_u8_, I.h, I.w, scaling_, offset_, wblob_ + st.w_off, st.kh, st.kw, st.sh, st.sw, st.pt, st.pl,
                           tptr(st.out), O.h, O.w, O.ld, e);
          }
          break;
        }
        launch_conv_direct(stream_, n, tptr(st.in), I.h, I.w, I.c, I.ld, wblob_ + st.w_off, st.N, st.kh, st.kw, st.sh, st.sw,
                           st.dh, st.dw, st.pt, st.pl, tptr(st.out), O.h, O.w, O.ld, e);
        break;
      case Step::PW:
        if (st.up_from >= 0) {
          const TensorInfo& S = tinfo_[st.up_from];
          launch_upsample_pw(stream_, n, tptr(st.up_from), S.h, S.w, st.K, S.ld, st.align_corners, st.half_pixel, wblob_ + st.w_off, st.n4, st.N,
                             tptr(st.out), O.h, O.w, O.ld, e);
          break;
        }
        if (st.use_tc && launch_pointwise_tc(stream_, n * I.h * I.w, st.K, st.N, tptr(st.in), I.ld, wblob_ + st.tc_hi_off, wblob_ + st.tc_lo_off,
                                             st.kpad, st.npad, tptr(st.out), O.ld, e))
          break;
        launch_pointwise(stream_, n * I.h * I.w, st.K, st.N, tptr(st.in), I.ld, wblob_ + st.w_off, st.n4, tptr(st.out), O.ld, e,
                         st.scale >= 0 ? tptr(st.scale) : nullptr, I.h * I.w,
                         st.in_add >= 0 ? tptr(st.in_add) : nullptr, st.in_add >= 0 ? tinfo_[st.in_add].ld : 0);
        break;
      case Step::DW:
        launch_depthwise(stream_, n, tptr(st.in), I.h, I.w, I.c, I.ld, wblob_ + st.w_off, st.kh, st.kw, st.sh, st.sw, st.dh, st.dw,
                         st.pt, st.pl, tptr(st.out), O.h, O.w, O.ld, e);
        break;
      case Step::POOL: {
        FcLayer fc[2];
        for (int k = 0; k < st.n_fc; ++k) {
          fc[k].w = wblob_ + st.fc[k].w_off; fc[k].bias = st.fc[k].has_bias ? wblob_ + st.fc[k].b_off : nullptr;
          fc[k].K = st.fc[k].K; fc[k].N = st.fc[k].N; fc[k].n4 = st.fc[k].n4; fc[k].act1 = st.fc[k].act1; fc[k].act2 = st.fc[k].act2;
        }
        const bool two = st.in2 >= 0;
        launch_pool_fc(stream_, n, tptr(st.in), I.c, I.ld, two ? tptr(st.in2) : nullptr, two ? tinfo_[st.in2].c : 0,
                       two ? tinfo_[st.in2].ld : 0, I.h, I.w, rowsum_, st.act1, nullptr, st.n_fc, fc, tptr(st.out), O.ld, pool_counters_);
        break;
      }
      case Step::RESIZE:
        launch_resize_bilinear(stream_, n, tptr(st.in), I.h, I.w, I.c, I.ld, tptr(st.out), O.h, O.w, O.ld, st.align_corners, st.half_pixel);
        break;
      case Step::TCONV:
        launch_tconv2x2(stream_, n, tptr(st.in), I.h, I.w, I.c, I.ld, wblob_ + st.w_off, wblob_ + st.b_off, st.N, tptr(st.out), O.h, O.w, O.ld, st.act2);
        break;
      case Step::ELT:
        launch_eltwise(stream_, st.elt_mode, n, I.h * I.w, I.c, tptr(st.in), I.ld, st.in2 >= 0 ? tptr(st.in2) : nullptr,
                       st.in2 >= 0 ? tinfo_[st.in2].ld : 0, st.scale >= 0 ? tptr(st.scale) : nullptr, tptr(st.out), O.ld, st.act1);
        break;
      case Step::BLOCK: {
        const FusedBlock& fb = blocks_[st.block];
        MbBlockArgs a{};
        a.x = tptr(st.in); a.ld_x = I.ld; a.y = tptr(st.out); a.ld_y = O.ld;
        a.h = I.h; a.w = I.w; a.cin = I.c; a.cexp = fb.expand.N; a.cout = fb.project.N;
        a.w1 = wblob_ + fb.expand.w_off; a.b1 = fb.expand.has_bias ? wblob_ + fb.expand.b_off : nullptr; a.n4_1 = fb.expand.n4;
        a.a1a = fb.expand.act1; a.a1b = fb.expand.act2;
        a.wd = wblob_ + fb.dw.w_off; a.bd = fb.dw.has_bias ? wblob_ + fb.dw.b_off : nullptr; a.k = fb.dw.kh; a.pt = fb.dw.pt; a.pl = fb.dw.pl;
        a.ada = fb.dw.act1; a.adb = fb.dw.act2;
        a.pool_act = fb.pool.act1;
        FcLayer* fl[2] = {&a.f0, &a.f1};
        for (int k = 0; k < 2; ++k) {
          fl[k]->w = wblob_ + fb.pool.fc[k].w_off; fl[k]->bias = fb.pool.fc[k].has_bias ? wblob_ + fb.pool.fc[k].b_off : nullptr;
          fl[k]->K = fb.pool.fc[k].K; fl[k]->N = fb.pool.fc[k].N; fl[k]->n4 = fb.pool.fc[k].n4; fl[k]->act1 = fb.pool.fc[k].act1; fl[k]->act2 = fb.pool.fc[k].act2;
        }
        a.w2 = wblob_ + fb.project.w_off; a.b2 = fb.project.has_bias ? wblob_ + fb.project.b_off : nullptr; a.n4_2 = fb.project.n4;
        a.a2a = fb.project.act1; a.a2b = fb.project.act2; a.residual = fb.project.residual >= 0 ? 1 : 0; a.a3 = fb.project.act3;
        launch_mb_block(stream_, n, a);
        break;
      }
      case Step::HEAD: {
        const HeadPlan& hp = heads_[st.block];
        const Step& P = hp.p; const Step& D = hp.d; const Step& T = hp.t;
        if (hp.s2) {
          launch_pw_dws2(stream_, tptr(P.in), I.ld, wblob_ + P.w_off, P.has_bias ? wblob_ + P.b_off : nullptr, P.act1 != ACT_NONE ? P.act1 : P.act2,
                         wblob_ + D.w_off, D.has_bias ? wblob_ + D.b_off : nullptr, D.act1 != ACT_NONE ? D.act1 : D.act2,
                         tptr(st.out), O.ld, n, I.h, I.w, O.h, O.w, D.pt, D.pl);
          break;
        }
        launch_head(stream_, P.N, tptr(P.in), I.ld, P.scale >= 0 ? tptr(P.scale) : nullptr, P.in_add >= 0 ? tptr(P.in_add) : nullptr,
                    P.in_add >= 0 ? tinfo_[P.in_add].ld : 0, wblob_ + P.w_off, P.has_bias ? wblob_ + P.b_off : nullptr, P.act1, P.act2,
                    wblob_ + D.w_off, D.has_bias ? wblob_ + D.b_off : nullptr, D.act1, D.act2, D.act3,
                    hp.has_t ? wblob_ + T.w_off : nullptr, hp.has_t ? wblob_ + T.b_off : nullptr, hp.has_t ? T.N : 0, hp.has_t ? T.act2 : 0,
                    tptr(st.out), O.ld, n, I.h, I.w, D.pt, D.pl);
        break;
      }
      case Step::CHAIN: {
        const ChainPlan& cp = chains_[st.block];
        launch_chain(stream_, n, cp.h, cp.w, cp.d_ops, cp.n_ops);
        break;
      }
      case Step::COPY:
        launch_copy_channels(stream_, n * I.h * I.w, I.c, tptr(st.in), I.ld, tptr(st.out) + st.copy_off, O.ld);
        break;
    }
  }
}

void Engine::enqueue_decision(int n) {
  if (dec_up_step_ >= 0) {
    const Step& st = steps_[dec_up_step_];
    const TensorInfo& I = tinfo_[st.in];
    launch_decision_up_iir(stream_, n, tptr(st.in), I.h, I.w, I.ld, st.align_corners, st.half_pixel, oh_, ow_, state_, ofinal_, opitch_);
    return;
  }
  launch_decision_iir(stream_, model_type_, n, tptr(g_.output), oh_, ow_, oc_, state_, ofinal_, opitch_);
}

// PostArgs of one call.  `yuyv_native` (optional out): the frames can stay in camera YUYV format for the post kernel.
PostArgs Engine::post_args(int n, const uint8_t* d_frames, size_t pitch, size_t stride, uint8_t* d_out, size_t out_stride,
                           uint8_t* d_yuyv, size_t yuyv_stride, uint8_t* d_mask, size_t mask_stride, const uint8_t* d_yuyv_in) const {
  const size_t fbytes = (size_t)W_ * H_ * 3;
  const bool flip = flip_h_ || flip_v_, resized = out_w_ != W_ || out_h_ != H_, tail = flip || resized;
  const bool want_rgb = d_out || d_yuyv;
  PostArgs a{};
  a.B = n; a.W = W_; a.H = H_;
  a.frames = d_frames; a.frame_pitch = pitch; a.frame_stride = stride;
  a.yuyv_in = d_yuyv_in; a.yuyv_in_stride = (size_t)W_ * H_ * 2;
  a.bg = d_bg_; a.bg_pitch = (size_t)W_ * 3; a.bg_stride = 0;
  if (has_bg_) {
    if (bgblur_k_) a.bg = d_bg_eff_;                       // blurred once, when the background / strength was set
    if (bg_count_ > 1) { a.bg_stride = fbytes; a.bg_cursor = d_bg_cursor_; a.bg_count = bg_count_; a.bg_advance = bg_advance_; }
  } else if (bgblur_k_ && want_rgb) {
    a.bg = d_bg_frames_; a.bg_stride = fbytes;            // the blurred camera frame itself (filled by enqueue_post)
  }
  if (bg_yuyv_valid_ && !(bgblur_k_ && !has_bg_)) a.bg_yuyv = d_bg_yuyv_;
  a.ofinal = ofinal_; a.ow = ow_; a.oh = oh_; a.opitch = opitch_;
  a.out_x = out_roidim_[0]; a.out_y = out_roidim_[1]; a.out_w = out_roidim_[2]; a.out_h = out_roidim_[3];
  a.roi_x = roidim_[0]; a.roi_y = roidim_[1]; a.roi_w = roidim_[2]; a.roi_h = roidim_[3];
  a.tab = tab_up_.tab; a.area2x2 = tab_up_.area2x2;
  a.out_pitch = (size_t)W_ * 3;
  if (!tail) {
    a.out = d_out; a.out_stride = out_stride;
    a.yuyv = d_yuyv; a.yuyv_stride = yuyv_stride;
  } else {
    a.out = want_rgb ? d_stage_a_ : nullptr; a.out_stride = fbytes;
    a.yuyv = nullptr; a.yuyv_stride = 0;
  }
  a.mask = d_mask; a.mask_stride = mask_stride;
  a.geo_rows = d_tile_geo_; a.geo_cols64 = d_tile_geo_ + geo_nty_; a.geo_cols128 = d_tile_geo_ + geo_nty_ + geo_ntx64_;
  return a;
}

void Engine::enqueue_post(int n, const uint8_t* d_frames, size_t pitch, size_t stride, uint8_t* d_out, size_t out_stride,
                          uint8_t* d_yuyv, size_t yuyv_stride, uint8_t* d_mask, size_t mask_stride, const uint8_t* d_yuyv_in) {
  const size_t fbytes = (size_t)W_ * H_ * 3;
  const bool flip = flip_h_ || flip_v_, resized = out_w_ != W_ || out_h_ != H_, tail = flip || resized;
  const bool want_rgb = d_out || d_yuyv;
  const bool ring = has_bg_ && bg_count_ > 1;
  // app/deepseg.cc:652-658: no background source => the background is the blurred camera frame itself
  if (!has_bg_ && bgblur_k_ && want_rgb)
    launch_gauss_blur(stream_, n, d_frames, pitch, stride, d_gauss_tmp_, d_bg_frames_, (size_t)W_ * 3, fbytes, W_, H_, taps_);
  const PostArgs a = post_args(n, d_frames, pitch, stride, d_out, out_stride, d_yuyv, yuyv_stride, d_mask, mask_stride, d_yuyv_in);
  launch_post(stream_, a);
  if (ring && bg_advance_) launch_advance_cursor(stream_, d_bg_cursor_, (int)(((long)n * bg_advance_) % bg_count_), bg_count_);
  if (!tail || !want_rgb) return;
  // app/deepseg.cc:667-679: flip, then scale to the virtual-camera geometry, then YUYV (:681)
  const size_t obytes = (size_t)out_w_ * out_h_ * 3;
  const uint8_t* cur = d_stage_a_; size_t cur_stride = fbytes;
  if (flip) {
    uint8_t* dst = (!resized && d_out) ? d_out : d_stage_b_;
    const size_t ds = (!resized && d_out) ? out_stride : fbytes;
    launch_flip_u8c3(stream_, n, cur, cur_stride, dst, ds, W_, H_, flip_h_, flip_v_);
    cur = dst; cur_stride = ds;
  }
  if (resized) {
    uint8_t* dst = d_out ? d_out : d_stage_c_;
    const size_t ds = d_out ? out_stride : obytes;
    launch_resize_u8c3(stream_, cur, W_, H_, (size_t)W_ * 3, dst, out_w_, out_h_, (size_t)out_w_ * 3, tab_out_.tab, tab_out_.area2x2, n, cur_stride, ds);
    cur = dst; cur_stride = ds;
  }
  if (d_yuyv) launch_rgb_to_yuyv(stream_, cur, d_yuyv, (size_t)out_w_ * out_h_, n, cur_stride, yuyv_stride);
}

// Camera YUYV input: the pre-processing kernel always reads the YUYV frames in place; the BGR frame is materialised
// (k_yuyv_to_bgr into `d_frames`) only when the post stage cannot read YUYV itself — no TMA path for this geometry, or
// the blur-my-background mode, whose Gaussian runs on the BGR frame.
bool Engine::yuyv_native(int n, const uint8_t* d_frames, size_t pitch, size_t stride, uint8_t* d_out, size_t out_stride,
                         uint8_t* d_yuyv, size_t yuyv_stride, uint8_t* d_mask, size_t mask_stride, const uint8_t* d_yuyv_in) const {
  if (!d_yuyv_in || (bgblur_k_ && !has_bg_)) return false;
  return post_tma_eligible(post_args(n, d_frames, pitch, stride, d_out, out_stride, d_yuyv, yuyv_stride, d_mask, mask_stride, d_yuyv_in));
}

void Engine::enqueue_all(int n, const uint8_t* d_frames, size_t pitch, size_t stride, uint8_t* d_out, size_t out_stride,
                         uint8_t* d_yuyv, size_t yuyv_stride, uint8_t* d_mask, size_t mask_stride, const uint8_t* d_yuyv_in,
                         bool native, bool sync_cbs) {
  const long l0 = thread_launch_count();
  if (d_yuyv_in && !native) launch_yuyv_to_bgr(stream_, d_yuyv_in, const_cast<uint8_t*>(d_frames), (size_t)n * W_ * H_);
  enqueue_pre(n, d_frames, pitch, stride, d_yuyv_in);
  if (sync_cbs && cb_.onprep) { cudaStreamSynchronize(stream_); cb_.onprep(cb_.caller_ctx); }
  enqueue_cnn(n, true);
  if (sync_cbs && cb_.oninfer) { cudaStreamSynchronize(stream_); cb_.oninfer(cb_.caller_ctx); }
  enqueue_decision(n);
  if (sync_cbs && cb_.onmask) { cudaStreamSynchronize(stream_); cb_.onmask(cb_.caller_ctx); }
  enqueue_post(n, d_frames, pitch, stride, d_out, out_stride, d_yuyv, yuyv_stride, d_mask, mask_stride, native ? d_yuyv_in : nullptr);
  last_call_launches_ = (int)(thread_launch_count() - l0);
}

bool Engine::run(int n, const uint8_t* d_frames, size_t pitch, size_t stride, uint8_t* d_out, size_t out_stride,
                 uint8_t* d_yuyv, size_t yuyv_stride, uint8_t* d_mask, size_t mask_stride, bool use_callbacks, std::string* err,
                 const uint8_t* d_yuyv_in) {
  if (n < 1 || n > max_batch_) { *err = "n_frames out of range (1..max_batch)"; return false; }
  if (d_yuyv_in && (W_ & 1)) { *err = "YUYV ingest needs an even width"; return false; }
  CUDA_OK(cudaSetDevice(device_));
  const bool cbs = use_callbacks && (cb_.onprep || cb_.oninfer || cb_.onmask);
  bool eager = cbs || (flags_ & 2u);
#ifdef BSB_EMU
  eager = true;
#endif
  if (eager) {
    const bool native = yuyv_native(n, d_frames, pitch, stride, d_out, out_stride, d_yuyv, yuyv_stride, d_mask, mask_stride, d_yuyv_in);
    last_native_ = native;
    enqueue_all(n, d_frames, pitch, stride, d_out, out_stride, d_yuyv, yuyv_stride, d_mask, mask_stride, d_yuyv_in, native, cbs);
    CUDA_OK(cudaGetLastError());
    return true;
  }
#ifndef BSB_EMU
  const GraphKey key{n, d_frames, pitch, stride, d_out, d_yuyv, d_mask, d_yuyv_in, d_out ? out_stride : 0, d_yuyv ? yuyv_stride : 0,
                     d_mask ? mask_stride : 0};
  auto it = graphs_.find(key);
  if (it == graphs_.end()) {
    if (graphs_.size() >= 64) drop_graphs();
    const bool native = yuyv_native(n, d_frames, pitch, stride, d_out, out_stride, d_yuyv, yuyv_stride, d_mask, mask_stride, d_yuyv_in);
    last_native_ = native;
    cudaGraph_t graph = nullptr;
    CUDA_OK(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
    enqueue_all(n, d_frames, pitch, stride, d_out, out_stride, d_yuyv, yuyv_stride, d_mask, mask_stride, d_yuyv_in, native, false);
    // the capture is always closed, also when a launch inside it failed: a stream left in capture mode would
    // poison every later call on this context
    const cudaError_t ce = cudaStreamEndCapture(stream_, &graph);
    if (ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      *err = std::string("CUDA graph capture failed: ") + cudaGetErrorString(ce);
      return false;
    }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t ie = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ie != cudaSuccess) { cudaGetLastError(); *err = std::string("cudaGraphInstantiate failed: ") + cudaGetErrorString(ie); return false; }
    it = graphs_.emplace(key, exec).first;
  }
  CUDA_OK(cudaGraphLaunch(it->second, stream_));
#endif
  return true;
}

bool Engine::run_yuyv(int n, const uint8_t* d_yuyv_in, uint8_t* d_out, size_t out_stride, uint8_t* d_yuyv, size_t yuyv_stride,
                      uint8_t* d_mask, size_t mask_stride, std::string* err) {
  const size_t row = (size_t)W_ * 3;
  return run(n, d_frames_, row, row * H_, d_out, out_stride, d_yuyv, yuyv_stride, d_mask, mask_stride, false, err, d_yuyv_in);
}

bool Engine::infer(int n, const float* h_in, float* h_out, std::string* err) {
  if (n < 1 || n > max_batch_) { *err = "n_frames out of range (1..max_batch)"; return false; }
  CUDA_OK(cudaSetDevice(device_));
  const TensorInfo& I = tinfo_[g_.input]; const TensorInfo& O = tinfo_[g_.output];
  CUDA_OK(cudaMemcpyAsync(tptr(g_.input), h_in, (size_t)n * I.frame_elems * 4, cudaMemcpyHostToDevice, stream_));
  enqueue_cnn(n, false);
  CUDA_OK(cudaMemcpyAsync(h_out, tptr(g_.output), (size_t)n * O.frame_elems * 4, cudaMemcpyDeviceToHost, stream_));
  CUDA_OK(cudaStreamSynchronize(stream_));
  CUDA_OK(cudaGetLastError());
  return true;
}

void Engine::drop_graphs() {
#ifndef BSB_EMU
  for (auto& kv : graphs_) cudaGraphExecDestroy(kv.second);
  graphs_.clear();
#endif
}

bool Engine::ensure(void** p, size_t* cap, size_t need, std::string* err) {
  if (need <= *cap && *p) return true;
  CUDA_OK(cudaStreamSynchronize(stream_));
  drop_graphs();                                  // captured launches hold the old pointer
  if (*p) cudaFree(*p);
  *p = nullptr; *cap = 0;
  CUDA_OK(cudaMalloc(p, need));
  *cap = need;
  return true;
}

bool Engine::set_background_ring(const uint8_t* frames, int count, int bw, int bh, size_t pitch, size_t frame_stride, int advance,
                                 std::string* err) {
  if (!frames || count < 1 || bw <= 0 || bh <= 0 || pitch < (size_t)bw * 3 || advance < 0 ||
      (count > 1 && frame_stride < pitch * (size_t)(bh - 1) + (size_t)bw * 3)) { *err = "invalid background"; return false; }
  CUDA_OK(cudaSetDevice(device_));
  const size_t raw = (size_t)bw * bh * 3, fbytes = (size_t)W_ * H_ * 3;
  if (!ensure((void**)&d_bg_raw_, &bg_raw_cap_, raw * count, err)) return false;
  if (!ensure((void**)&d_bg_, &bg_cap_, fbytes * count, err)) return false;
  if (bw != bg_w_ || bh != bg_h_) {
    CUDA_OK(cudaStreamSynchronize(stream_));
    if (!upload_resize_tab(build_resize_tab(bw, bh, W_, H_), &tab_bg_, err)) return false;
    bg_w_ = bw; bg_h_ = bh;
  }
  if (count != bg_count_ || advance != bg_advance_ || !has_bg_) { CUDA_OK(cudaStreamSynchronize(stream_)); drop_graphs(); }
  for (int i = 0; i < count; ++i)
    CUDA_OK(cudaMemcpy2DAsync(d_bg_raw_ + (size_t)i * raw, (size_t)bw * 3, frames + (size_t)i * frame_stride, pitch, (size_t)bw * 3, (size_t)bh,
                              cudaMemcpyHostToDevice, stream_));
  // app/background.cc:178-194: cv::resize(raw, out, Size(width, height))
  launch_resize_u8c3(stream_, d_bg_raw_, bw, bh, (size_t)bw * 3, d_bg_, W_, H_, (size_t)W_ * 3, tab_bg_.tab, tab_bg_.area2x2, count, raw, fbytes);
  CUDA_OK(cudaMemsetAsync(d_bg_cursor_, 0, sizeof(int), stream_));
  bg_count_ = count; bg_advance_ = advance;
  has_bg_ = true;
  if (!refresh_bg_blur(err)) return false;
  if (!refresh_bg_yuyv(err)) return false;
  CUDA_OK(cudaStreamSynchronize(stream_));
  CUDA_OK(cudaGetLastError());
  return true;
}

bool Engine::set_background_cursor(int index, std::string* err) {
  if (index < 0 || index >= bg_count_) { *err = "background index out of range"; return false; }
  CUDA_OK(cudaSetDevice(device_));
  CUDA_OK(cudaMemcpyAsync(d_bg_cursor_, &index, sizeof(int), cudaMemcpyHostToDevice, stream_));
  CUDA_OK(cudaStreamSynchronize(stream_));
  return true;
}

// (re)compute the blurred copy of the background ring; per-frame work only in camera-blur mode
bool Engine::refresh_bg_blur(std::string* err) {
  if (!bgblur_k_) return true;
  const size_t fbytes = (size_t)W_ * H_ * 3;
  if (!ensure((void**)&d_gauss_tmp_, &gauss_tmp_cap_, (size_t)max_batch_ * fbytes * sizeof(uint16_t), err)) return false;
  if (!has_bg_) return ensure((void**)&d_bg_frames_, &bg_frames_cap_, (size_t)max_batch_ * fbytes, err);
  if (!ensure((void**)&d_bg_eff_, &bg_eff_cap_, fbytes * bg_count_, err)) return false;
  for (int i = 0; i < bg_count_; i += max_batch_) {
    const int n = std::min(max_batch_, bg_count_ - i);
    launch_gauss_blur(stream_, n, d_bg_ + (size_t)i * fbytes, (size_t)W_ * 3, fbytes, d_gauss_tmp_, d_bg_eff_ + (size_t)i * fbytes, (size_t)W_ * 3,
                      fbytes, W_, H_, taps_);
  }
  CUDA_OK(cudaGetLastError());
  return true;
}

// YUYV of the effective background (ring): what convert_rgb_to_yuyv (app/deepseg.cc:87-106) yields wherever the composite
// equals the background, so all-background tiles of the post stage are copies.  Recomputed when the background changes.
bool Engine::refresh_bg_yuyv(std::string* err) {
  bg_yuyv_valid_ = false;
  if (W_ & 1) return true;
  const size_t npix = (size_t)W_ * H_;
  if (!ensure((void**)&d_bg_yuyv_, &bg_yuyv_cap_, npix * 2 * (size_t)bg_count_, err)) return false;
  const uint8_t* src = (has_bg_ && bgblur_k_) ? d_bg_eff_ : d_bg_;
  launch_rgb_to_yuyv(stream_, src, d_bg_yuyv_, npix, bg_count_, npix * 3, npix * 2);
  CUDA_OK(cudaGetLastError());
  bg_yuyv_valid_ = true;
  return true;
}

bool Engine::set_bgblur(int k, std::string* err) {
  GaussTaps t{};
  if (k != 0 && !gauss_taps(k, &t)) { *err = "strength value must be odd (1..255)"; return false; }   // app/deepseg.cc:423-426
  CUDA_OK(cudaSetDevice(device_));
  CUDA_OK(cudaStreamSynchronize(stream_));
  drop_graphs();
  bgblur_k_ = k; taps_ = t;
  if (!refresh_bg_blur(err)) return false;
  if (!refresh_bg_yuyv(err)) return false;
  CUDA_OK(cudaStreamSynchronize(stream_));
  return true;
}

bool Engine::set_output(bool flip_h, bool flip_v, int out_w, int out_h, std::string* err) {
  if (out_w <= 0) out_w = W_;
  if (out_h <= 0) out_h = H_;
  if (out_w > 16384 || out_h > 16384) { *err = "output size out of range"; return false; }
  CUDA_OK(cudaSetDevice(device_));
  CUDA_OK(cudaStreamSynchronize(stream_));
  drop_graphs();
  const size_t B = (size_t)max_batch_, fbytes = (size_t)W_ * H_ * 3, obytes = (size_t)out_w * out_h * 3;
  const bool flip = flip_h || flip_v, resized = out_w != W_ || out_h != H_;
  if (flip || resized) { if (!ensure((void**)&d_stage_a_, &stage_a_cap_, B * fbytes, err)) return false; }
  if (flip) { if (!ensure((void**)&d_stage_b_, &stage_b_cap_, B * fbytes, err)) return false; }
  if (resized) {
    if (!ensure((void**)&d_stage_c_, &stage_c_cap_, B * obytes, err)) return false;
    if (!upload_resize_tab(build_resize_tab(W_, H_, out_w, out_h), &tab_out_, err)) return false;
    // host-buffer API staging must hold the larger of the two geometries
    if (!ensure((void**)&d_out_, &out_cap_, B * std::max(fbytes, obytes), err)) return false;
    if (!ensure((void**)&d_yuyv_, &yuyv_cap_, B * std::max(fbytes, obytes) / 3 * 2, err)) return false;
  }
  flip_h_ = flip_h; flip_v_ = flip_v; out_w_ = out_w; out_h_ = out_h;
  return true;
}

bool Engine::sync(std::string* err) {
  CUDA_OK(cudaSetDevice(device_));
  CUDA_OK(cudaStreamSynchronize(stream_));
  CUDA_OK(cudaGetLastError());
  return true;
}

bool Engine::reset_state(std::string* err) {
  CUDA_OK(cudaSetDevice(device_));
  CUDA_OK(cudaMemsetAsync(state_, 0, (size_t)oh_ * ow_, stream_));
  CUDA_OK(cudaStreamSynchronize(stream_));
  return true;
}

double Engine::time_stage(int stage, int n, int iters, std::string* err) {
  if (n < 1 || n > max_batch_ || iters < 1 || stage < 0 || stage > 4) { *err = "bad arguments"; return -1.0; }
  if (cudaSetDevice(device_) != cudaSuccess) { *err = "cudaSetDevice failed"; return -1.0; }
  const size_t row = (size_t)W_ * 3, fbytes = row * H_, npix = (size_t)W_ * H_;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  // the last call's input format decides which kernels are timed (camera YUYV read in place, or BGR frames)
  const uint8_t* yin = last_native_ ? d_yuyv_in_ : nullptr;
  auto once = [&]() {
    if (stage == 0 || stage == 4) enqueue_pre(n, d_frames_, row, fbytes, yin);
    if (stage == 1 || stage == 4) enqueue_cnn(n, true);
    if (stage == 2 || stage == 4) enqueue_decision(n);
    if (stage == 3 || stage == 4) enqueue_post(n, d_frames_, row, fbytes, d_out_, fbytes, d_yuyv_, npix * 2, d_mask_, npix, yin);
  };
  once();                                   // warm-up (instruction cache, tables)
  cudaStreamSynchronize(stream_);
  cudaEventRecord(e0, stream_);
  for (int i = 0; i < iters; ++i) once();
  cudaEventRecord(e1, stream_);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (cudaGetLastError() != cudaSuccess) { *err = "CUDA error while timing"; return -1.0; }
  return (double)ms / iters;
}

long Engine::get_tensor(int t, float* out, long cap, std::string* err) {
  if (!(flags_ & 1u)) { *err = "bsb_get_tensor needs BSB_FLAG_KEEP_TENSORS"; return -1; }
  if (t < 0 || t >= (int)tinfo_.size()) { *err = "tensor index out of range"; return -1; }
  const TensorInfo& I = tinfo_[t];
  if (!I.materialized) return 0;
  const long n = (long)I.h * I.w * I.c;
  if (cap < n) { *err = "output buffer too small"; return -1; }
  if (cudaSetDevice(device_) != cudaSuccess || cudaStreamSynchronize(stream_) != cudaSuccess) { *err = "CUDA error in get_tensor"; return -1; }
  if (I.ld == I.c) {
    if (cudaMemcpy(out, tptr(t), (size_t)n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { *err = "CUDA memcpy failed"; return -1; }
  } else {
    for (long p = 0; p < (long)I.h * I.w; ++p)
      if (cudaMemcpy(out + p * I.c, tptr(t) + p * I.ld, (size_t)I.c * 4, cudaMemcpyDeviceToHost) != cudaSuccess) { *err = "CUDA memcpy failed"; return -1; }
  }
  return n;
}

long Engine::get_stage_u8(int which, int frame, uint8_t* out, long cap, std::string* err) {
  if (frame < 0 || frame >= max_batch_) { *err = "frame index out of range"; return -1; }
  const uint8_t* src = nullptr; long n = 0;
  if (which == 0) { n = (long)mh_ * mw_ * 3; src = in_u8_ + (size_t)frame * n; }
  else if (which == 1) { n = (long)mh_ * mw_ * 3; src = filt_u8_ ? filt_u8_ + (size_t)frame * n : nullptr; }
  else if (which == 2) { n = (long)oh_ * ow_; src = ofinal_ + (size_t)frame * oh_ * opitch_; }
  if (!src) { *err = which == 1 ? "stage buffer needs BSB_FLAG_KEEP_TENSORS" : "unknown stage"; return -1; }
  if (cap < n) { *err = "output buffer too small"; return -1; }
  if (cudaSetDevice(device_) != cudaSuccess || cudaStreamSynchronize(stream_) != cudaSuccess) { *err = "CUDA error in get_stage_u8"; return -1; }
  const cudaError_t ce = which == 2 ? cudaMemcpy2D(out, (size_t)ow_, src, (size_t)opitch_, (size_t)ow_, (size_t)oh_, cudaMemcpyDeviceToHost)
                                    : cudaMemcpy(out, src, (size_t)n, cudaMemcpyDeviceToHost);
  if (ce != cudaSuccess) { *err = "CUDA error in get_stage_u8"; return -1; }
  return n;
}

}  // namespace bsb
