// backscrub_b200/csrc/tflite_model.cpp — bounds-checked flatbuffer walk of a .tflite file.
#include "tflite_model.h"

#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "bsb_common.h"

namespace bsb {
namespace {

struct Malformed : std::runtime_error { using std::runtime_error::runtime_error; };

// Read-only cursor over the file with bounds checks on every access.
class Buf {
 public:
  Buf(const uint8_t* p, size_t n) : p_(p), n_(n) {}
  template <typename T> T rd(size_t o) const {
    if (o + sizeof(T) > n_) throw Malformed("offset out of range");
    T v; std::memcpy(&v, p_ + o, sizeof(T)); return v;
  }
  const uint8_t* ptr(size_t o, size_t len) const {
    if (o + len > n_ || o + len < o) throw Malformed("range out of file");
    return p_ + o;
  }
  size_t indirect(size_t o) const { return o + rd<uint32_t>(o); }
 private:
  const uint8_t* p_; size_t n_;
};

// A flatbuffer table: vtable lookup of field slots.
struct Table {
  const Buf* b; size_t pos;
  size_t field(int slot) const {
    size_t vt = pos - (size_t)(int64_t)b->rd<int32_t>(pos);
    uint16_t vtsize = b->rd<uint16_t>(vt);
    size_t fo = 4 + 2 * (size_t)slot;
    if (fo + 2 > vtsize) return 0;
    uint16_t off = b->rd<uint16_t>(vt + fo);
    return off ? pos + off : 0;
  }
  template <typename T> T scalar(int slot, T def) const { size_t f = field(slot); return f ? b->rd<T>(f) : def; }
  // vector field: returns element start, sets n
  size_t vec(int slot, uint32_t* n) const {
    size_t f = field(slot);
    if (!f) { *n = 0; return 0; }
    size_t v = b->indirect(f);
    *n = b->rd<uint32_t>(v);
    return v + 4;
  }
  Table sub(int slot) const { size_t f = field(slot); return Table{b, f ? b->indirect(f) : 0}; }
  std::string str(int slot) const {
    size_t f = field(slot);
    if (!f) return std::string();
    size_t s = b->indirect(f);
    uint32_t n = b->rd<uint32_t>(s);
    return std::string(reinterpret_cast<const char*>(b->ptr(s + 4, n)), n);
  }
  std::vector<int> ints(int slot) const {
    uint32_t n; size_t s = vec(slot, &n);
    std::vector<int> v(n);
    for (uint32_t i = 0; i < n; ++i) v[i] = b->rd<int32_t>(s + 4 * (size_t)i);
    return v;
  }
  Table elem(size_t vec_start, uint32_t i) const { return Table{b, b->indirect(vec_start + 4 * (size_t)i)}; }
};

// IEEE binary16 -> binary32 (exact); reference dequantize.h:31 widens fp16 weights the same way
float half_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  const uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
  uint32_t bits;
  if (e == 0) {
    if (m == 0) bits = sign;
    else {  // subnormal: normalise
      int shift = 0; uint32_t mm = m;
      while (!(mm & 0x400u)) { mm <<= 1; ++shift; }
      bits = sign | ((uint32_t)(113 - shift) << 23) | ((mm & 0x3ffu) << 13);
    }
  } else if (e == 31) bits = sign | 0x7f800000u | (m << 13);
  else bits = sign | ((e + 112u) << 23) | (m << 13);
  float f; std::memcpy(&f, &bits, 4); return f;
}

}  // namespace

int model_type_from_name(const std::string& path) {
  if (path.find("body-pix") != std::string::npos) return MODEL_BODYPIX;
  if (path.find("deeplab") != std::string::npos) return MODEL_DEEPLAB;
  if (path.find("segm_") != std::string::npos) return MODEL_MEET;
  if (path.find("selfie") != std::string::npos) return MODEL_MLKIT;
  return MODEL_UNKNOWN;
}

bool load_tflite(const std::string& path, Graph* g, std::string* err) {
  std::vector<uint8_t> file;
  {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { *err = "unable to load model from file: '" + path + "'"; return false; }
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    file.resize(n > 0 ? (size_t)n : 0);
    size_t got = file.empty() ? 0 : std::fread(file.data(), 1, file.size(), f);
    std::fclose(f);
    if (got != file.size() || file.size() < 16) { *err = "model file too short: '" + path + "'"; return false; }
  }
  try {
    Buf b(file.data(), file.size());
    if (std::memcmp(b.ptr(4, 4), "TFL3", 4) != 0) throw Malformed("not a TFL3 flatbuffer");
    Table model{&b, b.indirect(0)};
    // operator codes (schema.fbs:1108): max(deprecated_builtin_code, builtin_code)
    uint32_t n_codes; size_t codes = model.vec(1, &n_codes);
    std::vector<int> code(n_codes); std::vector<std::string> custom(n_codes);
    for (uint32_t i = 0; i < n_codes; ++i) {
      Table oc = model.elem(codes, i);
      int dep = oc.scalar<int8_t>(0, 0), nw = oc.scalar<int32_t>(3, 0);
      code[i] = dep > nw ? dep : nw;
      custom[i] = oc.str(1);
    }
    uint32_t n_buf; size_t bufs = model.vec(4, &n_buf);
    uint32_t n_sg; size_t sgs = model.vec(2, &n_sg);
    if (n_sg < 1) throw Malformed("no subgraph");
    Table sg = model.elem(sgs, 0);

    uint32_t n_t; size_t tens = sg.vec(0, &n_t);
    g->tensors.assign(n_t, GTensor());
    for (uint32_t i = 0; i < n_t; ++i) {
      Table tt = sg.elem(tens, i);
      GTensor& T = g->tensors[i];
      T.shape = tt.ints(0);
      if (T.shape.size() > 4) throw Malformed("tensor rank > 4");
      for (int d : T.shape) if (d < 0 || d > (1 << 20)) throw Malformed("bad tensor dimension");
      T.dtype = tt.scalar<int8_t>(1, 0);
      T.name = tt.str(3);
      uint32_t bidx = tt.scalar<uint32_t>(2, 0);
      if (bidx >= n_buf) throw Malformed("buffer index out of range");
      Table bt = model.elem(bufs, bidx);
      uint32_t nbytes; size_t data = bt.vec(0, &nbytes);
      if (nbytes) {
        T.is_const = true;
        const size_t cnt = T.count();
        if (T.dtype == 0) {
          if (nbytes < cnt * 4) throw Malformed("short f32 buffer");
          T.f32.resize(cnt); std::memcpy(T.f32.data(), b.ptr(data, cnt * 4), cnt * 4);
        } else if (T.dtype == 1) {
          if (nbytes < cnt * 2) throw Malformed("short f16 buffer");
          T.f32.resize(cnt);
          for (size_t k = 0; k < cnt; ++k) T.f32[k] = half_to_float(b.rd<uint16_t>(data + 2 * k));
        } else if (T.dtype == 2) {
          if (nbytes < cnt * 4) throw Malformed("short i32 buffer");
          T.i32.resize(cnt); std::memcpy(T.i32.data(), b.ptr(data, cnt * 4), cnt * 4);
        } else throw Malformed("unsupported constant tensor type");
      }
    }
    std::vector<int> ins = sg.ints(1), outs = sg.ints(2);
    if (ins.empty() || outs.empty()) throw Malformed("missing graph input/output");
    g->input = ins[0]; g->output = outs[0];
    auto check_t = [&](int t) { if (t < -1 || t >= (int)n_t) throw Malformed("tensor index out of range"); };
    check_t(g->input); check_t(g->output);
    if (g->input < 0 || g->output < 0) throw Malformed("graph input/output is an optional (-1) tensor");

    uint32_t n_ops; size_t ops = sg.vec(3, &n_ops);
    g->ops.assign(n_ops, GOp());
    for (uint32_t i = 0; i < n_ops; ++i) {
      Table ot = sg.elem(ops, i);
      GOp& O = g->ops[i];
      uint32_t ci = ot.scalar<uint32_t>(0, 0);
      if (ci >= n_codes) throw Malformed("opcode index out of range");
      O.kind = code[ci]; O.custom = custom[ci];
      O.in = ot.ints(1);
      std::vector<int> o = ot.ints(2);
      O.out = o.empty() ? -1 : o[0];
      for (int t : O.in) check_t(t);
      check_t(O.out);
      if (O.out < 0) throw Malformed("operator without an output tensor");
      if (!O.in.empty() && O.in[0] < 0) throw Malformed("operator whose first input is optional (-1)");
      Table bo = ot.sub(4);
      const bool has = bo.pos != 0;
      switch (O.kind) {
        case OP_CONV_2D:            // Conv2DOptions :521
          if (has) { O.padding = bo.scalar<int8_t>(0, 0); O.stride_w = bo.scalar<int32_t>(1, 1); O.stride_h = bo.scalar<int32_t>(2, 1);
                     O.act = bo.scalar<int8_t>(3, 0); O.dil_w = bo.scalar<int32_t>(4, 1); O.dil_h = bo.scalar<int32_t>(5, 1); }
          break;
        case OP_DEPTHWISE_CONV_2D:  // DepthwiseConv2DOptions :551
          if (has) { O.padding = bo.scalar<int8_t>(0, 0); O.stride_w = bo.scalar<int32_t>(1, 1); O.stride_h = bo.scalar<int32_t>(2, 1);
                     O.depth_mult = bo.scalar<int32_t>(3, 1); O.act = bo.scalar<int8_t>(4, 0);
                     O.dil_w = bo.scalar<int32_t>(5, 1); O.dil_h = bo.scalar<int32_t>(6, 1); }
          break;
        case OP_AVERAGE_POOL_2D:    // Pool2DOptions :542
          if (has) { O.padding = bo.scalar<int8_t>(0, 0); O.stride_w = bo.scalar<int32_t>(1, 1); O.stride_h = bo.scalar<int32_t>(2, 1);
                     O.filter_w = bo.scalar<int32_t>(3, 1); O.filter_h = bo.scalar<int32_t>(4, 1); O.act = bo.scalar<int8_t>(5, 0); }
          break;
        case OP_RESIZE_BILINEAR:    // ResizeBilinearOptions :721
          if (has) { O.align_corners = bo.scalar<uint8_t>(2, 0) != 0; O.half_pixel = bo.scalar<uint8_t>(3, 0) != 0; }
          break;
        case OP_FULLY_CONNECTED: if (has) O.act = bo.scalar<int8_t>(0, 0); break;
        case OP_ADD: case OP_MUL: if (has) O.act = bo.scalar<int8_t>(0, 0); break;
        case OP_CONCATENATION: if (has) { O.axis = bo.scalar<int32_t>(0, 0); O.act = bo.scalar<int8_t>(1, 0); } break;
        case OP_CUSTOM: {
          uint32_t nc; size_t cv = ot.vec(5, &nc);
          if (O.custom == "Convolution2DTransposeBias") {
            if (nc < 12) throw Malformed("Convolution2DTransposeBias: short custom_options");
            // TfLiteTransposeConvParams {padding (C enum, 1 = Same), stride_width, stride_height}
            O.tconv_same = b.rd<int32_t>(cv) == 1;
            O.stride_w = b.rd<int32_t>(cv + 4); O.stride_h = b.rd<int32_t>(cv + 8);
          }
          break;
        }
        default: break;
      }
      if (O.stride_w < 1 || O.stride_h < 1 || O.dil_w < 1 || O.dil_h < 1) throw Malformed("bad stride/dilation");
    }
    // fp16 weight storage: DEQUANTIZE(const f16) becomes a constant f32 tensor
    for (GOp& O : g->ops) {
      if (O.kind != OP_DEQUANTIZE) continue;
      if (O.in.empty() || O.in[0] < 0 || O.out < 0) throw Malformed("bad DEQUANTIZE");
      GTensor& s = g->tensors[O.in[0]]; GTensor& d = g->tensors[O.out];
      if (!s.is_const || s.f32.empty()) throw Malformed("DEQUANTIZE of a non-constant tensor is unsupported");
      d.f32 = s.f32; d.is_const = true;
    }
  } catch (const std::exception& e) {
    *err = std::string("malformed model '") + path + "': " + e.what();
    return false;
  }
  return true;
}

}  // namespace bsb
