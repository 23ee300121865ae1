/* oracle/oracle_model.c — .tflite (schema v3) loader + sequential fp32 interpreter.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * On-disk format: TF/lite/schema/schema.fbs (Model :1231, SubGraph :1169,
 * Tensor :195, Buffer :1191, OperatorCode :1108, Operator :1134, option tables
 * :510-:721).  Execution order = operator order, one op at a time, every
 * intermediate kept (TF/lite/core/subgraph.cc:1139-1215 walks the plan the
 * same way); Invoke runs with denormals flushed (TF/lite/interpreter.cc:226).
 */
#include "oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#if defined(__SSE__)
#include <xmmintrin.h>
#endif

typedef struct {
  int rank;
  int shape[4];
  int type;          /* 0 f32, 1 f16, 2 i32 */
  size_t count;      /* elements */
  int is_const;
  float* f32;        /* f32 view: constants (widened) or activation storage */
  int32_t* i32;      /* for INT32 constants */
} or_tensor;

typedef struct {
  int kind;
  int n_in, in[4];
  int out;
  /* options */
  int padding, stride_w, stride_h, dil_w, dil_h, act, depth_mult;
  int fw, fh;
  int align_corners, half_pixel;
  int axis;
  char custom[48];
  int tc_padding_same;
} or_op;

struct or_model {
  uint8_t* buf;
  size_t len;
  int n_tensors, n_ops;
  or_tensor* tensors;
  or_op* ops;
  int input, output;
};

/* ---- flatbuffer accessors ---- */
static uint32_t rd_u32(const uint8_t* b, size_t o) { uint32_t v; memcpy(&v, b + o, 4); return v; }
static int32_t rd_i32(const uint8_t* b, size_t o) { int32_t v; memcpy(&v, b + o, 4); return v; }
static uint16_t rd_u16(const uint8_t* b, size_t o) { uint16_t v; memcpy(&v, b + o, 2); return v; }
static size_t fb_indirect(const uint8_t* b, size_t o) { return o + rd_u32(b, o); }
static size_t fb_field(const uint8_t* b, size_t table, int slot) {
  size_t vt = table - rd_i32(b, table);
  uint16_t vtsize = rd_u16(b, vt);
  size_t fo = 4 + 2 * (size_t)slot;
  if (fo >= vtsize) return 0;
  uint16_t off = rd_u16(b, vt + fo);
  return off ? table + off : 0;
}
static size_t fb_vec(const uint8_t* b, size_t table, int slot, uint32_t* n) {
  size_t f = fb_field(b, table, slot);
  if (!f) { *n = 0; return 0; }
  size_t v = fb_indirect(b, f);
  *n = rd_u32(b, v);
  return v + 4;
}
static int32_t fb_i32(const uint8_t* b, size_t table, int slot, int32_t def) {
  size_t f = fb_field(b, table, slot);
  return f ? rd_i32(b, f) : def;
}
static int fb_i8(const uint8_t* b, size_t table, int slot, int def) {
  size_t f = fb_field(b, table, slot);
  return f ? (int)(int8_t)b[f] : def;
}
static size_t fb_table_at(const uint8_t* b, size_t vec_start, uint32_t i) {
  return fb_indirect(b, vec_start + 4 * (size_t)i);
}

static void seterr(char* err, size_t n, const char* msg) {
  if (err && n) { snprintf(err, n, "%s", msg); }
}

void or_model_free(or_model* m) {
  if (!m) return;
  if (m->tensors) {
    for (int i = 0; i < m->n_tensors; ++i) { free(m->tensors[i].f32); free(m->tensors[i].i32); }
    free(m->tensors);
  }
  free(m->ops);
  free(m->buf);
  free(m);
}

or_model* or_model_load(const char* path, char* err, size_t errlen) {
  FILE* f = fopen(path, "rb");
  if (!f) { seterr(err, errlen, "cannot open model file"); return NULL; }
  fseek(f, 0, SEEK_END);
  long len = ftell(f);
  fseek(f, 0, SEEK_SET);
  or_model* m = (or_model*)calloc(1, sizeof(or_model));
  m->buf = (uint8_t*)malloc((size_t)len);
  m->len = (size_t)len;
  if (fread(m->buf, 1, (size_t)len, f) != (size_t)len) { fclose(f); or_model_free(m); seterr(err, errlen, "short read"); return NULL; }
  fclose(f);
  const uint8_t* b = m->buf;
  if (len < 8 || memcmp(b + 4, "TFL3", 4) != 0) { or_model_free(m); seterr(err, errlen, "not a TFL3 flatbuffer"); return NULL; }
  size_t model = fb_indirect(b, 0);

  /* operator codes */
  uint32_t n_codes; size_t codes = fb_vec(b, model, 1, &n_codes);
  int* code_val = (int*)calloc(n_codes ? n_codes : 1, sizeof(int));
  char (*code_custom)[48] = calloc(n_codes ? n_codes : 1, 48);
  for (uint32_t i = 0; i < n_codes; ++i) {
    size_t oc = fb_table_at(b, codes, i);
    int dep = fb_i8(b, oc, 0, 0);
    int nw = fb_i32(b, oc, 3, 0);
    code_val[i] = dep > nw ? dep : nw;
    size_t cf = fb_field(b, oc, 1);
    if (cf) {
      size_t s = fb_indirect(b, cf);
      uint32_t n = rd_u32(b, s);
      if (n > 47) n = 47;
      memcpy(code_custom[i], b + s + 4, n);
    }
  }
  /* buffers */
  uint32_t n_buf; size_t bufs = fb_vec(b, model, 4, &n_buf);
  /* subgraph 0 */
  uint32_t n_sg; size_t sgs = fb_vec(b, model, 2, &n_sg);
  if (n_sg < 1) { free(code_val); free(code_custom); or_model_free(m); seterr(err, errlen, "no subgraph"); return NULL; }
  size_t sg = fb_table_at(b, sgs, 0);

  uint32_t n_t; size_t tens = fb_vec(b, sg, 0, &n_t);
  m->n_tensors = (int)n_t;
  m->tensors = (or_tensor*)calloc(n_t, sizeof(or_tensor));
  for (uint32_t i = 0; i < n_t; ++i) {
    size_t tt = fb_table_at(b, tens, i);
    or_tensor* T = &m->tensors[i];
    uint32_t nd; size_t sh = fb_vec(b, tt, 0, &nd);
    T->rank = (int)nd > 4 ? 4 : (int)nd;
    T->count = 1;
    for (int d = 0; d < T->rank; ++d) { T->shape[d] = rd_i32(b, sh + 4 * (size_t)d); T->count *= (size_t)T->shape[d]; }
    T->type = fb_i8(b, tt, 1, 0);
    uint32_t bidx = 0; { size_t bf = fb_field(b, tt, 2); if (bf) bidx = rd_u32(b, bf); }
    if (bidx < n_buf) {
      size_t bt = fb_table_at(b, bufs, bidx);
      uint32_t nbytes; size_t data = fb_vec(b, bt, 0, &nbytes);
      if (nbytes) {
        T->is_const = 1;
        if (T->type == 0) {
          T->f32 = (float*)malloc(T->count * 4);
          memcpy(T->f32, b + data, T->count * 4);
        } else if (T->type == 1) {
          T->f32 = (float*)malloc(T->count * 4);
          for (size_t k = 0; k < T->count; ++k) T->f32[k] = or_half_to_float(rd_u16(b, data + 2 * k));
        } else if (T->type == 2) {
          T->i32 = (int32_t*)malloc(T->count * 4);
          memcpy(T->i32, b + data, T->count * 4);
        }
      }
    }
  }
  uint32_t n_in; size_t ins = fb_vec(b, sg, 1, &n_in);
  uint32_t n_out; size_t outs = fb_vec(b, sg, 2, &n_out);
  m->input = n_in ? rd_i32(b, ins) : -1;
  m->output = n_out ? rd_i32(b, outs) : -1;

  uint32_t n_ops; size_t ops = fb_vec(b, sg, 3, &n_ops);
  m->n_ops = (int)n_ops;
  m->ops = (or_op*)calloc(n_ops, sizeof(or_op));
  for (uint32_t i = 0; i < n_ops; ++i) {
    size_t ot = fb_table_at(b, ops, i);
    or_op* O = &m->ops[i];
    uint32_t ci = 0; { size_t cf = fb_field(b, ot, 0); if (cf) ci = rd_u32(b, cf); }
    O->kind = code_val[ci];
    memcpy(O->custom, code_custom[ci], 48);
    uint32_t ni; size_t iv = fb_vec(b, ot, 1, &ni);
    O->n_in = ni > 4 ? 4 : (int)ni;
    for (int k = 0; k < O->n_in; ++k) O->in[k] = rd_i32(b, iv + 4 * (size_t)k);
    uint32_t no; size_t ov = fb_vec(b, ot, 2, &no);
    O->out = no ? rd_i32(b, ov) : -1;
    O->dil_w = O->dil_h = 1; O->stride_w = O->stride_h = 1; O->depth_mult = 1;
    size_t bo = fb_field(b, ot, 4);
    if (bo) bo = fb_indirect(b, bo);
    switch (O->kind) {
      case OR_CONV_2D:
        if (bo) { O->padding = fb_i8(b, bo, 0, 0); O->stride_w = fb_i32(b, bo, 1, 1); O->stride_h = fb_i32(b, bo, 2, 1);
                  O->act = fb_i8(b, bo, 3, 0); O->dil_w = fb_i32(b, bo, 4, 1); O->dil_h = fb_i32(b, bo, 5, 1); }
        break;
      case OR_DEPTHWISE_CONV_2D:
        if (bo) { O->padding = fb_i8(b, bo, 0, 0); O->stride_w = fb_i32(b, bo, 1, 1); O->stride_h = fb_i32(b, bo, 2, 1);
                  O->depth_mult = fb_i32(b, bo, 3, 1); O->act = fb_i8(b, bo, 4, 0);
                  O->dil_w = fb_i32(b, bo, 5, 1); O->dil_h = fb_i32(b, bo, 6, 1); }
        break;
      case OR_AVERAGE_POOL_2D:
        if (bo) { O->padding = fb_i8(b, bo, 0, 0); O->stride_w = fb_i32(b, bo, 1, 1); O->stride_h = fb_i32(b, bo, 2, 1);
                  O->fw = fb_i32(b, bo, 3, 1); O->fh = fb_i32(b, bo, 4, 1); O->act = fb_i8(b, bo, 5, 0); }
        break;
      case OR_RESIZE_BILINEAR:
        if (bo) { O->align_corners = fb_i8(b, bo, 2, 0); O->half_pixel = fb_i8(b, bo, 3, 0); }
        break;
      case OR_FULLY_CONNECTED:
        if (bo) O->act = fb_i8(b, bo, 0, 0);
        break;
      case OR_ADD: case OR_MUL:
        if (bo) O->act = fb_i8(b, bo, 0, 0);
        break;
      case OR_CONCATENATION:
        if (bo) { O->axis = fb_i32(b, bo, 0, 0); O->act = fb_i8(b, bo, 1, 0); }
        break;
      case OR_CUSTOM: {
        uint32_t nc; size_t cv = fb_vec(b, ot, 5, &nc);
        if (nc >= 12) {
          /* TfLiteTransposeConvParams {padding (C enum: 1 = Same, 2 = Valid), stride_width,
           * stride_height}, TF/lite/c/builtin_op_data.h:412-416 */
          O->tc_padding_same = rd_i32(b, cv) == 1;
          O->stride_w = rd_i32(b, cv + 4);
          O->stride_h = rd_i32(b, cv + 8);
        }
        break;
      }
      default: break;
    }
  }
  free(code_val); free(code_custom);

  /* fold DEQUANTIZE of constants (fp16 weight storage): output becomes a constant f32 view */
  for (int i = 0; i < m->n_ops; ++i) {
    or_op* O = &m->ops[i];
    if (O->kind != OR_DEQUANTIZE) continue;
    or_tensor* src = &m->tensors[O->in[0]];
    or_tensor* dst = &m->tensors[O->out];
    if (!src->is_const || !src->f32) { or_model_free(m); seterr(err, errlen, "DEQUANTIZE of non-constant"); return NULL; }
    dst->f32 = (float*)malloc(src->count * 4);
    memcpy(dst->f32, src->f32, src->count * 4);
    dst->is_const = 1;
  }
  /* allocate activations */
  for (int i = 0; i < m->n_tensors; ++i) {
    or_tensor* T = &m->tensors[i];
    if (!T->is_const && T->type == 0 && T->count) T->f32 = (float*)calloc(T->count, 4);
  }
  if (m->input < 0 || m->output < 0) { or_model_free(m); seterr(err, errlen, "missing input/output"); return NULL; }
  /* same checks as lib/libbackscrub.cc:85-112: float32, batch 1 */
  if (m->tensors[m->input].type != 0 || m->tensors[m->input].shape[0] != 1 ||
      m->tensors[m->output].type != 0 || m->tensors[m->output].shape[0] != 1) {
    or_model_free(m); seterr(err, errlen, "input/output must be float32 with batch 1"); return NULL;
  }
  return m;
}

int or_model_num_tensors(const or_model* m) { return m->n_tensors; }
int or_model_num_ops(const or_model* m) { return m->n_ops; }
int or_model_input(const or_model* m) { return m->input; }
int or_model_output(const or_model* m) { return m->output; }
int or_model_tensor_shape(const or_model* m, int t, int shape[4]) {
  for (int d = 0; d < 4; ++d) shape[d] = d < m->tensors[t].rank ? m->tensors[t].shape[d] : 1;
  return m->tensors[t].rank;
}
int or_model_tensor_is_const(const or_model* m, int t) { return m->tensors[t].is_const; }
const float* or_model_tensor_data(const or_model* m, int t) { return m->tensors[t].f32; }
int or_model_op(const or_model* m, int op, int* kind, int inputs[4], int* n_inputs, int* output) {
  if (op < 0 || op >= m->n_ops) return -1;
  *kind = m->ops[op].kind; *n_inputs = m->ops[op].n_in; *output = m->ops[op].out;
  for (int k = 0; k < 4; ++k) inputs[k] = k < m->ops[op].n_in ? m->ops[op].in[k] : -1;
  return 0;
}

static int shape4(const or_tensor* T, int* n, int* h, int* w, int* c) {
  int s[4] = {1, 1, 1, 1};
  /* right-align like RuntimeShape::ExtendedShape */
  for (int d = 0; d < T->rank; ++d) s[4 - T->rank + d] = T->shape[d];
  *n = s[0]; *h = s[1]; *w = s[2]; *c = s[3];
  return 0;
}

int or_model_invoke(or_model* m, const float* input) {
#if defined(__SSE__)
  unsigned int old_csr = _mm_getcsr();
  _mm_setcsr(old_csr | 0x8040u); /* FTZ | DAZ */
#endif
  int rc = 0;
  or_tensor* Tin = &m->tensors[m->input];
  memcpy(Tin->f32, input, Tin->count * 4);
  for (int i = 0; i < m->n_ops && rc == 0; ++i) {
    const or_op* O = &m->ops[i];
    or_tensor* out = O->out >= 0 ? &m->tensors[O->out] : NULL;
    const or_tensor* a = O->n_in > 0 && O->in[0] >= 0 ? &m->tensors[O->in[0]] : NULL;
    const or_tensor* b = O->n_in > 1 && O->in[1] >= 0 ? &m->tensors[O->in[1]] : NULL;
    const or_tensor* c = O->n_in > 2 && O->in[2] >= 0 ? &m->tensors[O->in[2]] : NULL;
    int n, h, w, ch, on, oh, ow, oc;
    switch (O->kind) {
      case OR_DEQUANTIZE: break; /* folded at load */
      case OR_CONV_2D:
        shape4(a, &n, &h, &w, &ch); shape4(out, &on, &oh, &ow, &oc);
        or_conv2d(a->f32, h, w, ch, b->f32, b->shape[0], b->shape[1], b->shape[2], c ? c->f32 : NULL,
                  O->stride_h, O->stride_w, O->dil_h, O->dil_w, O->padding, O->act, out->f32, oh, ow);
        break;
      case OR_DEPTHWISE_CONV_2D:
        shape4(a, &n, &h, &w, &ch); shape4(out, &on, &oh, &ow, &oc);
        or_depthwise_conv2d(a->f32, h, w, ch, b->f32, b->shape[1], b->shape[2], c ? c->f32 : NULL,
                            O->stride_h, O->stride_w, O->dil_h, O->dil_w, O->padding, O->depth_mult, O->act,
                            out->f32, oh, ow);
        break;
      case OR_AVERAGE_POOL_2D:
        shape4(a, &n, &h, &w, &ch); shape4(out, &on, &oh, &ow, &oc);
        or_average_pool(a->f32, h, w, ch, O->fh, O->fw, O->stride_h, O->stride_w, O->padding, O->act, out->f32, oh, ow);
        break;
      case OR_FULLY_CONNECTED: {
        int in_depth = b->shape[b->rank - 1], out_depth = b->shape[b->rank - 2];
        or_fully_connected(a->f32, (int)(a->count / (size_t)in_depth), in_depth, b->f32, out_depth,
                           c ? c->f32 : NULL, O->act, out->f32);
        break;
      }
      case OR_RESIZE_BILINEAR:
        shape4(a, &n, &h, &w, &ch); shape4(out, &on, &oh, &ow, &oc);
        or_resize_bilinear(a->f32, h, w, ch, out->f32, oh, ow, O->align_corners, O->half_pixel);
        break;
      case OR_HARD_SWISH: or_hard_swish(a->f32, out->f32, out->count); break;
      case OR_LOGISTIC: or_logistic(a->f32, out->f32, out->count); break;
      case OR_RELU: or_relu(a->f32, out->f32, out->count, OR_ACT_RELU); break;
      case OR_RELU6: or_relu(a->f32, out->f32, out->count, OR_ACT_RELU6); break;
      case OR_ADD:
        if (a->count != b->count) { rc = -2; break; }
        or_add(a->f32, b->f32, out->f32, out->count, O->act);
        break;
      case OR_MUL: {
        shape4(out, &on, &oh, &ow, &oc);
        const or_tensor* big = a->count >= b->count ? a : b;
        const or_tensor* small = a->count >= b->count ? b : a;
        if (small->count == big->count) or_mul(big->f32, small->f32, out->f32, out->count / (size_t)oc, oc, 0, O->act);
        else if (small->count == (size_t)oc) or_mul(big->f32, small->f32, out->f32, out->count / (size_t)oc, oc, 1, O->act);
        else rc = -3;
        break;
      }
      case OR_CONCATENATION: {
        /* TF/lite/kernels/internal/reference/concatenation.h:28 — channel axis only */
        shape4(out, &on, &oh, &ow, &oc);
        int axis = O->axis < 0 ? O->axis + out->rank : O->axis;
        if (axis != out->rank - 1) { rc = -4; break; }
        size_t outer = out->count / (size_t)oc;
        int off = 0;
        for (int k = 0; k < O->n_in; ++k) {
          const or_tensor* s = &m->tensors[O->in[k]];
          int sc = s->shape[s->rank - 1];
          for (size_t p = 0; p < outer; ++p) memcpy(out->f32 + p * oc + off, s->f32 + p * sc, (size_t)sc * 4);
          off += sc;
        }
        break;
      }
      case OR_CUSTOM:
        if (strcmp(O->custom, "Convolution2DTransposeBias") != 0) { rc = -5; break; }
        shape4(a, &n, &h, &w, &ch); shape4(out, &on, &oh, &ow, &oc);
        or_tconv_bias(a->f32, h, w, ch, b->f32, b->shape[0], b->shape[1], b->shape[2], c->f32,
                      O->stride_h, O->stride_w, O->tc_padding_same, out->f32, oh, ow);
        break;
      default: rc = -100 - O->kind; break;
    }
  }
#if defined(__SSE__)
  _mm_setcsr(old_csr);
#endif
  return rc;
}

double or_model_flops(const or_model* m) {
  double fl = 0;
  for (int i = 0; i < m->n_ops; ++i) {
    const or_op* O = &m->ops[i];
    if (O->out < 0) continue;
    const or_tensor* out = &m->tensors[O->out];
    if (O->kind == OR_CONV_2D) {
      const or_tensor* w = &m->tensors[O->in[1]];
      fl += 2.0 * (double)out->count * w->shape[1] * w->shape[2] * w->shape[3];
    } else if (O->kind == OR_DEPTHWISE_CONV_2D) {
      const or_tensor* w = &m->tensors[O->in[1]];
      fl += 2.0 * (double)out->count * w->shape[1] * w->shape[2];
    } else if (O->kind == OR_FULLY_CONNECTED) {
      const or_tensor* w = &m->tensors[O->in[1]];
      fl += 2.0 * (double)out->count * w->shape[w->rank - 1];
    } else if (O->kind == OR_CUSTOM) {
      const or_tensor* w = &m->tensors[O->in[1]];
      fl += 2.0 * (double)out->count * w->shape[3];
    }
  }
  return fl;
}
