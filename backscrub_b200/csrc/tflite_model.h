// backscrub_b200/csrc/tflite_model.h — .tflite (schema v3) reader -> graph IR.
//
// Replaces, for this path, what the reference gets from
// tflite::FlatBufferModel::BuildFromFile + InterpreterBuilder
// (lib/libbackscrub.cc:191-209).  Format: reference tensorflow/lite/schema/schema.fbs.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace bsb {

// TFLite builtin operator codes in use (schema.fbs enum BuiltinOperator :229)
enum OpKind : int {
  OP_ADD = 0, OP_AVERAGE_POOL_2D = 1, OP_CONCATENATION = 2, OP_CONV_2D = 3, OP_DEPTHWISE_CONV_2D = 4,
  OP_DEQUANTIZE = 6, OP_FULLY_CONNECTED = 9, OP_LOGISTIC = 14, OP_MUL = 18, OP_RELU = 19, OP_RELU6 = 21,
  OP_RESIZE_BILINEAR = 23, OP_CUSTOM = 32, OP_HARD_SWISH = 117,
};

struct GTensor {
  std::vector<int> shape;
  int dtype = 0;              // 0 f32, 1 f16, 2 i32 (schema.fbs TensorType :36)
  bool is_const = false;
  std::vector<float> f32;     // constant payload widened to fp32
  std::vector<int32_t> i32;
  std::string name;
  size_t count() const { size_t n = 1; for (int d : shape) n *= (size_t)d; return n; }
  // NHWC accessors with rank right-aligned to 4 (RuntimeShape::ExtendedShape)
  int dim4(int i) const { int r = (int)shape.size(); int k = i - (4 - r); return k < 0 ? 1 : shape[k]; }
};

struct GOp {
  int kind = -1;
  std::vector<int> in;
  int out = -1;
  int padding = 0;            // 0 SAME, 1 VALID
  int stride_w = 1, stride_h = 1, dil_w = 1, dil_h = 1, depth_mult = 1;
  int act = 0;                // fused activation (ActivationFunctionType)
  int filter_w = 1, filter_h = 1;
  bool align_corners = false, half_pixel = false;
  int axis = 0;
  std::string custom;
  bool tconv_same = true;
};

struct Graph {
  std::vector<GTensor> tensors;
  std::vector<GOp> ops;
  int input = -1, output = -1;
};

// Returns false and fills `err` on any malformed / unsupported file.
bool load_tflite(const std::string& path, Graph* g, std::string* err);

// lib/libbackscrub.cc:116-130 — model family from the file name
int model_type_from_name(const std::string& path);

}  // namespace bsb
