// oracle/ref_tconv_driver.cc — thin driver around the REFERENCE's own
// lib/transpose_conv_bias.cc (compiled in place from /root/reference by
// oracle/Makefile; nothing from the reference is copied into this repo).
// TEST INFRASTRUCTURE ONLY: produces oracle/_ref/libref_tconv.so, used by
// tests/ to validate oracle_nn.c:or_tconv_bias and the CUDA kernel against the
// genuine reference op.
//
// The driver hand-builds the TfLiteContext / TfLiteNode / TfLiteTensor that
// RegisterConvolution2DTransposeBias()->prepare / ->invoke expect
// (lib/transpose_conv_bias.cc:118-256).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "lib/transpose_conv_bias.h"
#include "tensorflow/lite/c/builtin_op_data.h"
#include "tensorflow/lite/c/common.h"

namespace {

TfLiteStatus ResizeTensorCb(TfLiteContext*, TfLiteTensor* tensor, TfLiteIntArray* new_size) {
  size_t n = 1;
  for (int i = 0; i < new_size->size; ++i) n *= static_cast<size_t>(new_size->data[i]);
  if (tensor->dims) TfLiteIntArrayFree(tensor->dims);
  tensor->dims = new_size;
  tensor->bytes = n * sizeof(float);
  std::free(tensor->data.raw);
  tensor->data.raw = static_cast<char*>(std::calloc(n, sizeof(float)));
  return kTfLiteOk;
}

void ReportErrorCb(TfLiteContext*, const char*, ...) {}

TfLiteIntArray* MakeDims(std::initializer_list<int> d) {
  TfLiteIntArray* a = TfLiteIntArrayCreate(static_cast<int>(d.size()));
  int i = 0;
  for (int v : d) a->data[i++] = v;
  return a;
}

}  // namespace

// in: [ih, iw, ic]; w: OHWI [oc, kh, kw, ic]; bias: [oc]; padding_same: 1 = kTfLitePaddingSame.
// out must hold oh*ow*oc floats where (oh, ow) are returned by the op's own Prepare;
// returns 0 on success, and writes the output size to *oh / *ow.  If out == nullptr
// only the size is computed.
extern "C" int ref_tconv_bias(const float* in, int ih, int iw, int ic,
                              const float* w, int oc, int kh, int kw, const float* bias,
                              int stride_h, int stride_w, int padding_same,
                              float* out, int* oh, int* ow) {
  TfLiteTensor tensors[4];
  std::memset(tensors, 0, sizeof(tensors));
  auto set = [&](int idx, const float* data, TfLiteIntArray* dims, size_t count) {
    tensors[idx].type = kTfLiteFloat32;
    tensors[idx].dims = dims;
    tensors[idx].bytes = count * sizeof(float);
    tensors[idx].data.raw = static_cast<char*>(std::malloc(count * sizeof(float) + 4));
    if (data) std::memcpy(tensors[idx].data.raw, data, count * sizeof(float));
    tensors[idx].allocation_type = kTfLiteArenaRw;
  };
  set(0, in, MakeDims({1, ih, iw, ic}), static_cast<size_t>(ih) * iw * ic);
  set(1, w, MakeDims({oc, kh, kw, ic}), static_cast<size_t>(oc) * kh * kw * ic);
  set(2, bias, MakeDims({oc}), static_cast<size_t>(oc));
  set(3, nullptr, MakeDims({1, 1, 1, 1}), 1);

  TfLiteContext ctx;
  std::memset(&ctx, 0, sizeof(ctx));
  ctx.tensors_size = 4;
  ctx.tensors = tensors;
  ctx.ResizeTensor = ResizeTensorCb;
  ctx.ReportError = ReportErrorCb;

  TfLiteNode node;
  std::memset(&node, 0, sizeof(node));
  node.inputs = TfLiteIntArrayCreate(3);
  node.inputs->data[0] = 0; node.inputs->data[1] = 1; node.inputs->data[2] = 2;
  node.outputs = TfLiteIntArrayCreate(1);
  node.outputs->data[0] = 3;
  TfLiteTransposeConvParams params;
  params.padding = padding_same ? kTfLitePaddingSame : kTfLitePaddingValid;
  params.stride_width = stride_w;
  params.stride_height = stride_h;
  node.custom_initial_data = &params;
  node.custom_initial_data_size = sizeof(params);

  TfLiteRegistration* reg = mediapipe::tflite_operations::RegisterConvolution2DTransposeBias();
  int rc = 0;
  if (reg->prepare(&ctx, &node) != kTfLiteOk) rc = 1;
  if (!rc) {
    *oh = tensors[3].dims->data[1];
    *ow = tensors[3].dims->data[2];
    if (out) {
      if (reg->invoke(&ctx, &node) != kTfLiteOk) rc = 2;
      else std::memcpy(out, tensors[3].data.raw, tensors[3].bytes);
    }
  }
  for (auto& t : tensors) { std::free(t.data.raw); if (t.dims) TfLiteIntArrayFree(t.dims); }
  TfLiteIntArrayFree(node.inputs);
  TfLiteIntArrayFree(node.outputs);
  return rc;
}
