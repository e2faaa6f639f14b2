/*
 * yf_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See yf_oracle.h.
 *
 * Restates, op by op, the TensorFlow-Lite *reference* int8 kernels (tensorflow==2.10.0, the
 * version pinned by the reference at yoloface/tensorflow/requirements.txt:2; default build,
 * i.e. double rounding, TFLITE_SINGLE_ROUNDING off).  TFLite itself is a third-party dependency
 * that is absent from /root/reference, so each function names the published TFLite routine it
 * restates and the reference call site that depends on it:
 *   inference call site ......... yoloface/tflite/tflite_prediction.py:23-41
 *   graph / parameters .......... yoloface/tflite/yoloface_int8.tflite (54 ops)
 *   ST's fused view of the graph  stm32/X-CUBE-AI/App/network.c:2193-2938
 */
#define _GNU_SOURCE
#include "yf_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* Minimal FlatBuffer reader (schema subset: SURVEY.md Appendix A)                             */
/* ------------------------------------------------------------------------------------------ */
static char g_err[256];
const char* yfo_last_error(void) { return g_err; }
#define FAIL(...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return 0; } while (0)

typedef struct { const uint8_t* b; size_t n; } fb_t;
static uint32_t rd_u32(const fb_t* f, size_t o) { uint32_t v = 0; if (o + 4 <= f->n) memcpy(&v, f->b + o, 4); return v; }
static int32_t rd_i32(const fb_t* f, size_t o) { return (int32_t)rd_u32(f, o); }
static uint16_t rd_u16(const fb_t* f, size_t o) { uint16_t v = 0; if (o + 2 <= f->n) memcpy(&v, f->b + o, 2); return v; }
static int8_t rd_i8(const fb_t* f, size_t o) { return o < f->n ? (int8_t)f->b[o] : 0; }
/* absolute position of field `slot` of table `t`, or 0 if absent */
static size_t fb_field(const fb_t* f, size_t t, int slot) {
  size_t vt = t - (size_t)(int64_t)rd_i32(f, t);
  uint16_t vlen = rd_u16(f, vt);
  if (4 + 2 * slot >= vlen) return 0;
  uint16_t off = rd_u16(f, vt + 4 + 2 * (size_t)slot);
  return off ? t + off : 0;
}
static size_t fb_indirect(const fb_t* f, size_t o) { return o + rd_u32(f, o); }
/* vector at field position `o` (o may be 0): returns length, *data = first element */
static uint32_t fb_vec(const fb_t* f, size_t o, size_t* data) {
  if (!o) { *data = 0; return 0; }
  size_t v = fb_indirect(f, o);
  *data = v + 4;
  return rd_u32(f, v);
}

/* ------------------------------------------------------------------------------------------ */
/* Model                                                                                       */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
  int rank, shape[4], type, qdim, nscale, nzp;
  const float* scale;      /* points into the flatbuffer (little-endian host assumed) */
  const int64_t* zp;
  const uint8_t* data; size_t data_len;
  const char* name; uint32_t name_len;
} tensor_t;

typedef struct {
  int opcode, nin, in[3], out;
  int padding, stride_w, stride_h, filter_w, filter_h, depth_mult, fused_act, axis;
  float alpha;
} op_t;

struct yfo_model {
  fb_t fb;
  int ntensors, nops, input, output;
  tensor_t* t;
  op_t* op;
  char** names;
};

void yfo_free(yfo_model* m) {
  if (!m) return;
  if (m->names) { for (int i = 0; i < m->ntensors; ++i) free(m->names[i]); free(m->names); }
  free(m->t); free(m->op); free(m);
}

yfo_model* yfo_load(const uint8_t* buf, size_t len) {
  if (!buf || len < 16 || memcmp(buf + 4, "TFL3", 4) != 0) FAIL("not a TFL3 flatbuffer");
  yfo_model* m = (yfo_model*)calloc(1, sizeof *m);
  m->fb.b = buf; m->fb.n = len;
  const fb_t* f = &m->fb;
  size_t root = fb_indirect(f, 0);

  /* operator codes: max(deprecated_builtin_code, builtin_code) */
  size_t oc_data; uint32_t noc = fb_vec(f, fb_field(f, root, 1), &oc_data);
  int* codes = (int*)calloc(noc ? noc : 1, sizeof(int));
  for (uint32_t i = 0; i < noc; ++i) {
    size_t t = fb_indirect(f, oc_data + 4 * i);
    size_t a = fb_field(f, t, 0), c = fb_field(f, t, 3);
    int dep = a ? rd_i8(f, a) : 0, cur = c ? rd_i32(f, c) : 0;
    codes[i] = dep > cur ? dep : cur;
  }
  /* buffers */
  size_t bf_data; uint32_t nbuf = fb_vec(f, fb_field(f, root, 4), &bf_data);
  /* subgraph 0 */
  size_t sg_data; uint32_t nsg = fb_vec(f, fb_field(f, root, 2), &sg_data);
  if (nsg < 1) { free(codes); yfo_free(m); FAIL("no subgraph"); }
  size_t sg = fb_indirect(f, sg_data);

  size_t tv; m->ntensors = (int)fb_vec(f, fb_field(f, sg, 0), &tv);
  m->t = (tensor_t*)calloc((size_t)m->ntensors, sizeof(tensor_t));
  m->names = (char**)calloc((size_t)m->ntensors, sizeof(char*));
  for (int i = 0; i < m->ntensors; ++i) {
    tensor_t* T = &m->t[i];
    size_t t = fb_indirect(f, tv + 4 * (size_t)i);
    size_t sd; uint32_t r = fb_vec(f, fb_field(f, t, 0), &sd);
    T->rank = r > 4 ? 4 : (int)r;
    for (int k = 0; k < T->rank; ++k) T->shape[k] = rd_i32(f, sd + 4 * (size_t)k);
    size_t ty = fb_field(f, t, 1); T->type = ty ? rd_i8(f, ty) : 0;
    size_t bu = fb_field(f, t, 2); uint32_t bidx = bu ? rd_u32(f, bu) : 0;
    if (bidx < nbuf) {
      size_t bt = fb_indirect(f, bf_data + 4 * (size_t)bidx), dd;
      uint32_t dl = fb_vec(f, fb_field(f, bt, 0), &dd);
      if (dl) { T->data = f->b + dd; T->data_len = dl; }
    }
    size_t nd; T->name_len = fb_vec(f, fb_field(f, t, 3), &nd);
    m->names[i] = (char*)calloc(T->name_len + 1, 1);
    if (T->name_len) memcpy(m->names[i], f->b + nd, T->name_len);
    T->name = m->names[i];
    size_t q = fb_field(f, t, 4);
    if (q) {
      q = fb_indirect(f, q);
      size_t d;
      T->nscale = (int)fb_vec(f, fb_field(f, q, 2), &d); T->scale = (const float*)(f->b + d);
      T->nzp = (int)fb_vec(f, fb_field(f, q, 3), &d);    T->zp = (const int64_t*)(f->b + d);
      size_t qd = fb_field(f, q, 6); T->qdim = qd ? rd_i32(f, qd) : 0;
    }
  }
  size_t iv; uint32_t ni = fb_vec(f, fb_field(f, sg, 1), &iv);
  size_t ov; uint32_t no = fb_vec(f, fb_field(f, sg, 2), &ov);
  m->input = ni ? rd_i32(f, iv) : -1;
  m->output = no ? rd_i32(f, ov) : -1;

  size_t opv; m->nops = (int)fb_vec(f, fb_field(f, sg, 3), &opv);
  m->op = (op_t*)calloc((size_t)m->nops, sizeof(op_t));
  for (int i = 0; i < m->nops; ++i) {
    op_t* O = &m->op[i];
    size_t t = fb_indirect(f, opv + 4 * (size_t)i);
    size_t oi = fb_field(f, t, 0); uint32_t cidx = oi ? rd_u32(f, oi) : 0;
    O->opcode = cidx < noc ? codes[cidx] : -1;
    size_t d; uint32_t n = fb_vec(f, fb_field(f, t, 1), &d);
    O->nin = n > 3 ? 3 : (int)n;
    for (int k = 0; k < O->nin; ++k) O->in[k] = rd_i32(f, d + 4 * (size_t)k);
    n = fb_vec(f, fb_field(f, t, 2), &d);
    O->out = n ? rd_i32(f, d) : -1;
    size_t opt = fb_field(f, t, 4);
    size_t ot = opt ? fb_indirect(f, opt) : 0;
    O->stride_w = O->stride_h = 1; O->depth_mult = 1; O->axis = 3;
#define OPT_I8(slot, dflt) (ot && fb_field(f, ot, slot) ? rd_i8(f, fb_field(f, ot, slot)) : (dflt))
#define OPT_I32(slot, dflt) (ot && fb_field(f, ot, slot) ? rd_i32(f, fb_field(f, ot, slot)) : (dflt))
    switch (O->opcode) {
      case YFO_OP_CONV_2D:
        O->padding = OPT_I8(0, 0); O->stride_w = OPT_I32(1, 1); O->stride_h = OPT_I32(2, 1);
        O->fused_act = OPT_I8(3, 0); break;
      case YFO_OP_DEPTHWISE_CONV_2D:
        O->padding = OPT_I8(0, 0); O->stride_w = OPT_I32(1, 1); O->stride_h = OPT_I32(2, 1);
        O->depth_mult = OPT_I32(3, 1); O->fused_act = OPT_I8(4, 0); break;
      case YFO_OP_MAX_POOL_2D:
        O->padding = OPT_I8(0, 0); O->stride_w = OPT_I32(1, 1); O->stride_h = OPT_I32(2, 1);
        O->filter_w = OPT_I32(3, 1); O->filter_h = OPT_I32(4, 1); O->fused_act = OPT_I8(5, 0); break;
      case YFO_OP_ADD: O->fused_act = OPT_I8(0, 0); break;
      case YFO_OP_CONCATENATION: O->axis = OPT_I32(0, 0); O->fused_act = OPT_I8(1, 0); break;
      case YFO_OP_LEAKY_RELU: {
        size_t a = ot ? fb_field(f, ot, 0) : 0; float al = 0.f;
        if (a) memcpy(&al, f->b + a, 4);
        O->alpha = al; break; }
      default: break;
    }
    if (O->fused_act != 0) { free(codes); yfo_free(m); FAIL("op %d: fused activation %d unsupported", i, O->fused_act); }
  }
  free(codes);
  return m;
}

int yfo_num_tensors(const yfo_model* m) { return m->ntensors; }
int yfo_num_ops(const yfo_model* m) { return m->nops; }
int yfo_input_tensor(const yfo_model* m) { return m->input; }
int yfo_output_tensor(const yfo_model* m) { return m->output; }
int yfo_tensor_info(const yfo_model* m, int t, int shape[4], int* type, int* nscale, int* qdim,
                    const uint8_t** data, size_t* data_len) {
  if (t < 0 || t >= m->ntensors) return -1;
  const tensor_t* T = &m->t[t];
  for (int k = 0; k < 4; ++k) shape[k] = k < T->rank ? T->shape[k] : 1;
  if (type) *type = T->type;
  if (nscale) *nscale = T->nscale;
  if (qdim) *qdim = T->qdim;
  if (data) *data = T->data;
  if (data_len) *data_len = T->data_len;
  return T->rank;
}
float yfo_tensor_scale(const yfo_model* m, int t, int i) { return m->t[t].scale[i]; }
int64_t yfo_tensor_zp(const yfo_model* m, int t, int i) { return m->t[t].zp[i]; }
const char* yfo_tensor_name(const yfo_model* m, int t) { return m->t[t].name; }
int yfo_op_info(const yfo_model* m, int op, int* opcode, int inputs[3], int* output) {
  if (op < 0 || op >= m->nops) return -1;
  const op_t* O = &m->op[op];
  *opcode = O->opcode; *output = O->out;
  for (int k = 0; k < 3; ++k) inputs[k] = k < O->nin ? O->in[k] : -1;
  return O->nin;
}

/* ------------------------------------------------------------------------------------------ */
/* Fixed-point primitives                                                                      */
/* ------------------------------------------------------------------------------------------ */
/* tensorflow/lite/kernels/internal/quantization_util.cc :: QuantizeMultiplier */
void yfo_quantize_multiplier(double d, int32_t* mult, int* shift) {
  if (d == 0.) { *mult = 0; *shift = 0; return; }
  const double q = frexp(d, shift);
  int64_t q_fixed = (int64_t)round(q * (double)(1LL << 31));   /* TfLiteRound: half away from zero */
  if (q_fixed == (1LL << 31)) { q_fixed /= 2; ++*shift; }
  if (*shift < -31) { *shift = 0; q_fixed = 0; }
  *mult = (int32_t)q_fixed;
}
/* gemmlowp fixedpoint.h :: SaturatingRoundingDoublingHighMul (int32 specialisation) */
int32_t yfo_srdhm(int32_t a, int32_t b) {
  if (a == b && a == INT32_MIN) return INT32_MAX;
  int64_t ab = (int64_t)a * (int64_t)b;
  int32_t nudge = ab >= 0 ? (1 << 30) : (1 - (1 << 30));
  return (int32_t)((ab + nudge) / (1LL << 31));                /* C '/' truncates toward zero */
}
/* gemmlowp fixedpoint.h :: RoundingDivideByPOT */
int32_t yfo_rdivpot(int32_t x, int exponent) {
  const int32_t mask = (int32_t)((1LL << exponent) - 1);
  const int32_t remainder = x & mask;
  const int32_t threshold = (mask >> 1) + (x < 0 ? 1 : 0);
  return (x >> exponent) + (remainder > threshold ? 1 : 0);    /* arithmetic shift */
}
/* tensorflow/lite/kernels/internal/common.h :: MultiplyByQuantizedMultiplier (double rounding) */
int32_t yfo_mbqm(int32_t x, int32_t mult, int shift) {
  int left = shift > 0 ? shift : 0, right = shift > 0 ? 0 : -shift;
  return yfo_rdivpot(yfo_srdhm((int32_t)((uint32_t)x << left), mult), right);
}
static inline int8_t clamp8(int32_t v) { return (int8_t)(v < -128 ? -128 : (v > 127 ? 127 : v)); }

/* ------------------------------------------------------------------------------------------ */
/* Run-time tensors (shapes propagated from the actual input size)                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int h, w, c; int8_t* d; } act_t;

/* kernel_util.h :: ComputePaddingHeightWidth / ComputeOutSize */
static int out_size(int padding_same, int in, int k, int stride) {
  return padding_same ? (in + stride - 1) / stride : (in - k + stride) / stride;
}
static int pad_before(int in, int k, int stride, int out) {
  int total = (out - 1) * stride + k - in;
  return total > 0 ? total / 2 : 0;
}

/* reference_ops::Pad (int8; pad value = output zero point, pad.cc) -- paddings tensor is int32 [4,2] */
static void op_pad(const yfo_model* m, const op_t* O, const act_t* x, act_t* y) {
  const int32_t* p = (const int32_t*)m->t[O->in[1]].data;
  int pt = p[2], pb = p[3], pl = p[4], pr = p[5];
  int8_t zp = (int8_t)m->t[O->out].zp[0];
  y->h = x->h + pt + pb; y->w = x->w + pl + pr; y->c = x->c;
  y->d = (int8_t*)malloc((size_t)y->h * y->w * y->c);
  memset(y->d, zp, (size_t)y->h * y->w * y->c);
  for (int i = 0; i < x->h; ++i)
    memcpy(y->d + ((size_t)(i + pt) * y->w + pl) * y->c, x->d + (size_t)i * x->w * x->c, (size_t)x->w * x->c);
}

/* kernel_util.cc :: PopulateConvolutionQuantizationParams (per-channel) */
static void conv_multipliers(const yfo_model* m, const op_t* O, int cout, int32_t* mult, int* shift) {
  const tensor_t* in = &m->t[O->in[0]]; const tensor_t* flt = &m->t[O->in[1]]; const tensor_t* out = &m->t[O->out];
  for (int c = 0; c < cout; ++c) {
    float fs = flt->nscale > 1 ? flt->scale[c] : flt->scale[0];
    double eff = (double)in->scale[0] * (double)fs / (double)out->scale[0];
    yfo_quantize_multiplier(eff, &mult[c], &shift[c]);
  }
}

/* reference_integer_ops::ConvPerChannel (conv.h); filter OHWI */
static void op_conv(const yfo_model* m, const op_t* O, const act_t* x, act_t* y) {
  const tensor_t* flt = &m->t[O->in[1]];
  int cout = flt->shape[0], kh = flt->shape[1], kw = flt->shape[2], cin = flt->shape[3];
  const int8_t* w = (const int8_t*)flt->data;
  const int32_t* bias = O->nin > 2 && O->in[2] >= 0 ? (const int32_t*)m->t[O->in[2]].data : NULL;
  int same = O->padding == 0;
  y->h = out_size(same, x->h, kh, O->stride_h); y->w = out_size(same, x->w, kw, O->stride_w); y->c = cout;
  int ph = same ? pad_before(x->h, kh, O->stride_h, y->h) : 0, pw = same ? pad_before(x->w, kw, O->stride_w, y->w) : 0;
  int32_t in_off = -(int32_t)m->t[O->in[0]].zp[0], out_off = (int32_t)m->t[O->out].zp[0];
  int32_t mult[256]; int shift[256];
  conv_multipliers(m, O, cout, mult, shift);
  y->d = (int8_t*)malloc((size_t)y->h * y->w * cout);
  for (int oy = 0; oy < y->h; ++oy) for (int ox = 0; ox < y->w; ++ox) for (int oc = 0; oc < cout; ++oc) {
    int32_t acc = 0;
    for (int ky = 0; ky < kh; ++ky) {
      int iy = oy * O->stride_h - ph + ky; if (iy < 0 || iy >= x->h) continue;
      for (int kx = 0; kx < kw; ++kx) {
        int ix = ox * O->stride_w - pw + kx; if (ix < 0 || ix >= x->w) continue;
        const int8_t* xi = x->d + ((size_t)iy * x->w + ix) * cin;
        const int8_t* wi = w + (((size_t)oc * kh + ky) * kw + kx) * cin;
        for (int ic = 0; ic < cin; ++ic) acc += (int32_t)wi[ic] * ((int32_t)xi[ic] + in_off);
      }
    }
    if (bias) acc += bias[oc];
    acc = yfo_mbqm(acc, mult[oc], shift[oc]) + out_off;
    y->d[((size_t)oy * y->w + ox) * cout + oc] = clamp8(acc);
  }
}

/* reference_integer_ops::DepthwiseConvPerChannel (depthwise_conv.h); filter [1,KH,KW,C*dm] */
static void op_dwconv(const yfo_model* m, const op_t* O, const act_t* x, act_t* y) {
  const tensor_t* flt = &m->t[O->in[1]];
  int kh = flt->shape[1], kw = flt->shape[2], cout = flt->shape[3], dm = O->depth_mult;
  const int8_t* w = (const int8_t*)flt->data;
  const int32_t* bias = O->nin > 2 && O->in[2] >= 0 ? (const int32_t*)m->t[O->in[2]].data : NULL;
  int same = O->padding == 0;
  y->h = out_size(same, x->h, kh, O->stride_h); y->w = out_size(same, x->w, kw, O->stride_w); y->c = cout;
  int ph = same ? pad_before(x->h, kh, O->stride_h, y->h) : 0, pw = same ? pad_before(x->w, kw, O->stride_w, y->w) : 0;
  int32_t in_off = -(int32_t)m->t[O->in[0]].zp[0], out_off = (int32_t)m->t[O->out].zp[0];
  int32_t mult[256]; int shift[256];
  conv_multipliers(m, O, cout, mult, shift);
  y->d = (int8_t*)malloc((size_t)y->h * y->w * cout);
  for (int oy = 0; oy < y->h; ++oy) for (int ox = 0; ox < y->w; ++ox)
    for (int ic = 0; ic < x->c; ++ic) for (int q = 0; q < dm; ++q) {
      int oc = ic * dm + q; int32_t acc = 0;
      for (int ky = 0; ky < kh; ++ky) {
        int iy = oy * O->stride_h - ph + ky; if (iy < 0 || iy >= x->h) continue;
        for (int kx = 0; kx < kw; ++kx) {
          int ix = ox * O->stride_w - pw + kx; if (ix < 0 || ix >= x->w) continue;
          acc += (int32_t)w[((size_t)ky * kw + kx) * cout + oc] * ((int32_t)x->d[((size_t)iy * x->w + ix) * x->c + ic] + in_off);
        }
      }
      if (bias) acc += bias[oc];
      acc = yfo_mbqm(acc, mult[oc], shift[oc]) + out_off;
      y->d[((size_t)oy * y->w + ox) * cout + oc] = clamp8(acc);
    }
}

/* activations.cc :: LeakyReluPrepare -- note the *float32* expressions before widening */
static void leaky_params(const yfo_model* m, const op_t* O, int32_t* mi, int* si, int32_t* ma, int* sa) {
  float s_in = m->t[O->in[0]].scale[0], s_out = m->t[O->out].scale[0];
  double alpha_multiplier = (double)(float)(s_in * O->alpha / s_out);
  double identity_multiplier = (double)(float)(s_in / s_out);
  yfo_quantize_multiplier(alpha_multiplier, ma, sa);
  yfo_quantize_multiplier(identity_multiplier, mi, si);
}
/* reference_ops::QuantizeLeakyRelu (leaky_relu.h) on one value */
static inline int8_t leaky_one(int32_t q, int32_t zin, int32_t zout, int32_t mi, int si, int32_t ma, int sa) {
  int32_t v = q - zin;
  int32_t u = zout + (v >= 0 ? yfo_mbqm(v, mi, si) : yfo_mbqm(v, ma, sa));
  return clamp8(u);
}
int yfo_leaky_lut(const yfo_model* m, int op, int8_t lut[256]) {
  if (op < 0 || op >= m->nops || m->op[op].opcode != YFO_OP_LEAKY_RELU) return -1;
  const op_t* O = &m->op[op];
  int32_t mi, ma; int si, sa; leaky_params(m, O, &mi, &si, &ma, &sa);
  int32_t zin = (int32_t)m->t[O->in[0]].zp[0], zout = (int32_t)m->t[O->out].zp[0];
  for (int q = -128; q < 128; ++q) lut[q + 128] = leaky_one(q, zin, zout, mi, si, ma, sa);
  return 0;
}
static void op_leaky(const yfo_model* m, const op_t* O, const act_t* x, act_t* y) {
  int32_t mi, ma; int si, sa; leaky_params(m, O, &mi, &si, &ma, &sa);
  int32_t zin = (int32_t)m->t[O->in[0]].zp[0], zout = (int32_t)m->t[O->out].zp[0];
  size_t n = (size_t)x->h * x->w * x->c;
  *y = *x; y->d = (int8_t*)malloc(n);
  for (size_t i = 0; i < n; ++i) y->d[i] = leaky_one(x->d[i], zin, zout, mi, si, ma, sa);
}

/* reference_integer_ops::MaxPool (pooling.h): max over the in-bounds part of the window */
static void op_maxpool(const yfo_model* m, const op_t* O, const act_t* x, act_t* y) {
  (void)m;
  int same = O->padding == 0, kh = O->filter_h, kw = O->filter_w;
  y->h = out_size(same, x->h, kh, O->stride_h); y->w = out_size(same, x->w, kw, O->stride_w); y->c = x->c;
  int ph = same ? pad_before(x->h, kh, O->stride_h, y->h) : 0, pw = same ? pad_before(x->w, kw, O->stride_w, y->w) : 0;
  y->d = (int8_t*)malloc((size_t)y->h * y->w * y->c);
  for (int oy = 0; oy < y->h; ++oy) for (int ox = 0; ox < y->w; ++ox) for (int c = 0; c < x->c; ++c) {
    int y0 = oy * O->stride_h - ph, x0 = ox * O->stride_w - pw;
    int ys = y0 < 0 ? 0 : y0, ye = y0 + kh > x->h ? x->h : y0 + kh;
    int xs = x0 < 0 ? 0 : x0, xe = x0 + kw > x->w ? x->w : x0 + kw;
    int32_t mx = -128;
    for (int iy = ys; iy < ye; ++iy) for (int ix = xs; ix < xe; ++ix) {
      int32_t v = x->d[((size_t)iy * x->w + ix) * x->c + c]; if (v > mx) mx = v;
    }
    y->d[((size_t)oy * y->w + ox) * y->c + c] = clamp8(mx);
  }
}

/* add.cc :: Prepare + reference_integer_ops::AddElementwise (add.h), left_shift 20 */
static void op_add(const yfo_model* m, const op_t* O, const act_t* a, const act_t* b, act_t* y) {
  const tensor_t* t1 = &m->t[O->in[0]]; const tensor_t* t2 = &m->t[O->in[1]]; const tensor_t* to = &m->t[O->out];
  const int left_shift = 20;
  float mx = t1->scale[0] > t2->scale[0] ? t1->scale[0] : t2->scale[0];
  const double twice_max = (double)(2 * mx);
  int32_t m1, m2, mo; int s1, s2, so;
  yfo_quantize_multiplier((double)t1->scale[0] / twice_max, &m1, &s1);
  yfo_quantize_multiplier((double)t2->scale[0] / twice_max, &m2, &s2);
  yfo_quantize_multiplier(twice_max / (double)((float)(1 << left_shift) * to->scale[0]), &mo, &so);
  int32_t o1 = -(int32_t)t1->zp[0], o2 = -(int32_t)t2->zp[0], oo = (int32_t)to->zp[0];
  size_t n = (size_t)a->h * a->w * a->c;
  *y = *a; y->d = (int8_t*)malloc(n);
  for (size_t i = 0; i < n; ++i) {
    int32_t v1 = (o1 + a->d[i]) * (1 << left_shift), v2 = (o2 + b->d[i]) * (1 << left_shift);
    int32_t sum = yfo_mbqm(v1, m1, s1) + yfo_mbqm(v2, m2, s2);
    y->d[i] = clamp8(yfo_mbqm(sum, mo, so) + oo);
  }
}

/* quantize.cc :: Prepare + reference_ops::Requantize<int8,int8> (requantize.h) */
static void op_quantize(const yfo_model* m, const op_t* O, const act_t* x, act_t* y) {
  const tensor_t* ti = &m->t[O->in[0]]; const tensor_t* to = &m->t[O->out];
  int32_t mult; int shift;
  yfo_quantize_multiplier((double)ti->scale[0] / (double)to->scale[0], &mult, &shift);
  int32_t zin = (int32_t)ti->zp[0], zout = (int32_t)to->zp[0];
  size_t n = (size_t)x->h * x->w * x->c;
  *y = *x; y->d = (int8_t*)malloc(n);
  for (size_t i = 0; i < n; ++i) y->d[i] = clamp8(yfo_mbqm(x->d[i] - zin, mult, shift) + zout);
}

/* reference_ops::Concatenation along the channel axis (inputs share the output's quant params) */
static void op_concat(const act_t* a, const act_t* b, act_t* y) {
  y->h = a->h; y->w = a->w; y->c = a->c + b->c;
  y->d = (int8_t*)malloc((size_t)y->h * y->w * y->c);
  for (size_t p = 0; p < (size_t)a->h * a->w; ++p) {
    memcpy(y->d + p * y->c, a->d + p * a->c, (size_t)a->c);
    memcpy(y->d + p * y->c + a->c, b->d + p * b->c, (size_t)b->c);
  }
}

/* Walk the 54 operators in flatbuffer order (what tf.lite.Interpreter.invoke() does). */
static int run_graph(const yfo_model* m, const int8_t* in, int H, int W, act_t* acts /*[ntensors]*/) {
  const tensor_t* ti = &m->t[m->input];
  acts[m->input].h = H; acts[m->input].w = W; acts[m->input].c = ti->shape[3];
  size_t n = (size_t)H * W * ti->shape[3];
  acts[m->input].d = (int8_t*)malloc(n); memcpy(acts[m->input].d, in, n);
  for (int i = 0; i < m->nops; ++i) {
    const op_t* O = &m->op[i];
    const act_t* x = &acts[O->in[0]];
    act_t* y = &acts[O->out];
    if (!x->d) FAIL("op %d: input tensor %d not produced", i, O->in[0]);
    switch (O->opcode) {
      case YFO_OP_PAD: op_pad(m, O, x, y); break;
      case YFO_OP_CONV_2D: op_conv(m, O, x, y); break;
      case YFO_OP_DEPTHWISE_CONV_2D: op_dwconv(m, O, x, y); break;
      case YFO_OP_LEAKY_RELU: op_leaky(m, O, x, y); break;
      case YFO_OP_MAX_POOL_2D: op_maxpool(m, O, x, y); break;
      case YFO_OP_ADD: op_add(m, O, x, &acts[O->in[1]], y); break;
      case YFO_OP_QUANTIZE: op_quantize(m, O, x, y); break;
      case YFO_OP_CONCATENATION:
        if (O->axis != 3 && O->axis != -1) FAIL("op %d: concat axis %d unsupported", i, O->axis);
        op_concat(x, &acts[O->in[1]], y); break;
      default: FAIL("op %d: opcode %d unsupported", i, O->opcode);
    }
  }
  return 1;
}

long yfo_op_out_elems(const yfo_model* m, int op, int H, int W, int shape_out[4]) {
  if (op < 0 || op >= m->nops) return -1;
  /* cheap shape propagation: run the size rules only */
  int (*hw)[3] = calloc((size_t)m->ntensors, sizeof *hw);
  hw[m->input][0] = H; hw[m->input][1] = W; hw[m->input][2] = m->t[m->input].shape[3];
  for (int i = 0; i <= op; ++i) {
    const op_t* O = &m->op[i]; int* x = hw[O->in[0]]; int* y = hw[O->out];
    int same = O->padding == 0;
    switch (O->opcode) {
      case YFO_OP_PAD: { const int32_t* p = (const int32_t*)m->t[O->in[1]].data;
        y[0] = x[0] + p[2] + p[3]; y[1] = x[1] + p[4] + p[5]; y[2] = x[2]; break; }
      case YFO_OP_CONV_2D: { const tensor_t* f = &m->t[O->in[1]];
        y[0] = out_size(same, x[0], f->shape[1], O->stride_h); y[1] = out_size(same, x[1], f->shape[2], O->stride_w); y[2] = f->shape[0]; break; }
      case YFO_OP_DEPTHWISE_CONV_2D: { const tensor_t* f = &m->t[O->in[1]];
        y[0] = out_size(same, x[0], f->shape[1], O->stride_h); y[1] = out_size(same, x[1], f->shape[2], O->stride_w); y[2] = f->shape[3]; break; }
      case YFO_OP_MAX_POOL_2D:
        y[0] = out_size(same, x[0], O->filter_h, O->stride_h); y[1] = out_size(same, x[1], O->filter_w, O->stride_w); y[2] = x[2]; break;
      case YFO_OP_CONCATENATION: y[0] = x[0]; y[1] = x[1]; y[2] = x[2] + hw[O->in[1]][2]; break;
      default: y[0] = x[0]; y[1] = x[1]; y[2] = x[2]; break;
    }
  }
  int* y = hw[m->op[op].out];
  long n = (long)y[0] * y[1] * y[2];
  if (shape_out) { shape_out[0] = 1; shape_out[1] = y[0]; shape_out[2] = y[1]; shape_out[3] = y[2]; }
  free(hw);
  return n;
}

int yfo_run(const yfo_model* m, const int8_t* in, int H, int W, int8_t* out, int8_t** op_out) {
  act_t* acts = (act_t*)calloc((size_t)m->ntensors, sizeof(act_t));
  int ok = run_graph(m, in, H, W, acts);
  if (ok) {
    const act_t* o = &acts[m->output];
    if (out) memcpy(out, o->d, (size_t)o->h * o->w * o->c);
    if (op_out) for (int i = 0; i < m->nops; ++i) if (op_out[i]) {
      const act_t* a = &acts[m->op[i].out];
      memcpy(op_out[i], a->d, (size_t)a->h * a->w * a->c);
    }
  }
  for (int i = 0; i < m->ntensors; ++i) free(acts[i].d);
  free(acts);
  return ok ? 0 : -1;
}

typedef struct { const yfo_model* m; const int8_t* in; int8_t* out; int n, H, W, tid, nthreads, rc; } job_t;
static void* batch_worker(void* p) {
  job_t* j = (job_t*)p;
  size_t isz = (size_t)j->H * j->W * 3, osz = (size_t)(j->H / 8) * (j->W / 8) * 18;
  for (int i = j->tid; i < j->n; i += j->nthreads)
    if (yfo_run(j->m, j->in + isz * i, j->H, j->W, j->out + osz * i, NULL)) j->rc = -1;
  return NULL;
}
int yfo_run_batch(const yfo_model* m, const int8_t* in, int n, int H, int W, int8_t* out, int threads) {
  if (threads < 1) threads = 1;
  if (threads > n) threads = n > 0 ? n : 1;
  pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof *th);
  job_t* jobs = (job_t*)calloc((size_t)threads, sizeof *jobs);
  for (int t = 0; t < threads; ++t) {
    jobs[t] = (job_t){m, in, out, n, H, W, t, threads, 0};
    if (t) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  batch_worker(&jobs[0]);
  int rc = jobs[0].rc;
  for (int t = 1; t < threads; ++t) { pthread_join(th[t], NULL); rc |= jobs[t].rc; }
  free(th); free(jobs);
  return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* Head decode + NMS                                                                           */
/* ------------------------------------------------------------------------------------------ */
/* yoloface.c:98-102 */
static float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

/* One candidate (cell i, anchor j) of a [gh,gw,3*6] head: yoloface.c:116-138 with the canonical (un-swapped)
 * box form of tflite_prediction.py:5-11,43-57.  stride = input pixels per head cell (8 for this model). */
static void decode_one(const int8_t* head, int gw, float out_scale, int out_zp, const float anchors[3][2], float stride,
                       int i, int j, yfo_det* d) {
  const int8_t* q = head + (size_t)i * 18 + j * 6;                     /* yoloface.c:116 */
  float conf = sigmoidf_(((float)q[4] - (float)out_zp) * out_scale);
  int gx = i % gw, gy = i / gw;                                         /* yoloface.c:129-130 */
  float x = ((float)q[0] - (float)out_zp) * out_scale, y = ((float)q[1] - (float)out_zp) * out_scale;
  float w = ((float)q[2] - (float)out_zp) * out_scale, h = ((float)q[3] - (float)out_zp) * out_scale;
  x = (sigmoidf_(x) + (float)gx) * stride; y = (sigmoidf_(y) + (float)gy) * stride;   /* :135-136 */
  w = expf(w) * anchors[j][0]; h = expf(h) * anchors[j][1];            /* :137-138 */
  /* tflite_prediction.py:5-11 xywh2xyxy (the firmware's x/y swap + clamp is an LCD quirk) */
  d->x1 = x - w / 2; d->y1 = y - h / 2; d->x2 = x + w / 2; d->y2 = y + h / 2; d->conf = conf;
}

static const float kAnchors[3][2] = {{9, 14}, {12, 17}, {22, 21}};      /* yoloface.c:20 */

void yfo_decode_all(const int8_t* head, int gh, int gw, float out_scale, int out_zp, const float* anchors6, float stride,
                    yfo_det* cands) {
  const float (*an)[2] = anchors6 ? (const float (*)[2])anchors6 : kAnchors;
  for (int i = 0; i < gh * gw; ++i) for (int j = 0; j < 3; ++j)
    decode_one(head, gw, out_scale, out_zp, an, stride, i, j, &cands[i * 3 + j]);
}

int yfo_decode_nms_ex(const int8_t* head, int gh, int gw, float out_scale, int out_zp, const float* anchors6, float stride,
                      float conf_thr, float iou_thr, int plus_one, yfo_det* dets, int max_det) {
  const float (*an)[2] = anchors6 ? (const float (*)[2])anchors6 : kAnchors;
  int ncand = gh * gw * 3, n = 0;
  yfo_det* c = (yfo_det*)malloc(sizeof(yfo_det) * (size_t)ncand);
  int* idx = (int*)malloc(sizeof(int) * (size_t)ncand);
  for (int i = 0; i < gh * gw; ++i) for (int j = 0; j < 3; ++j) {
    const int8_t* q = head + (size_t)i * 18 + j * 6;
    float conf = sigmoidf_(((float)q[4] - (float)out_zp) * out_scale);
    if (!(conf >= conf_thr)) continue;                                  /* yoloface.c:123 */
    decode_one(head, gw, out_scale, out_zp, an, stride, i, j, &c[n]);
    idx[n] = i * 3 + j; ++n;
  }
  /* stable insertion sort: conf desc, candidate index asc */
  for (int a = 1; a < n; ++a) {
    yfo_det d = c[a]; int id = idx[a], b = a - 1;
    while (b >= 0 && c[b].conf < d.conf) { c[b + 1] = c[b]; idx[b + 1] = idx[b]; --b; }
    c[b + 1] = d; idx[b + 1] = id;
  }
  int kept = 0;
  char* dead = (char*)calloc((size_t)(n ? n : 1), 1);
  for (int a = 0; a < n && kept < max_det; ++a) {
    if (dead[a]) continue;
    dets[kept++] = c[a];
    if (iou_thr < 0) continue;
    float one = plus_one ? 1.f : 0.f;
    float ax1 = c[a].x1, ay1 = c[a].y1, ax2 = c[a].x2, ay2 = c[a].y2;
    if (plus_one) { ax1 = truncf(ax1); ay1 = truncf(ay1); ax2 = truncf(ax2); ay2 = truncf(ay2); }
    float area_a = (ax2 - ax1 + one) * (ay2 - ay1 + one);
    for (int b = a + 1; b < n; ++b) {                                    /* yoloface_test.py:177-199 */
      if (dead[b]) continue;
      float bx1 = c[b].x1, by1 = c[b].y1, bx2 = c[b].x2, by2 = c[b].y2;
      if (plus_one) { bx1 = truncf(bx1); by1 = truncf(by1); bx2 = truncf(bx2); by2 = truncf(by2); }
      float area_b = (bx2 - bx1 + one) * (by2 - by1 + one);
      float xx1 = fmaxf(ax1, bx1), yy1 = fmaxf(ay1, by1), xx2 = fminf(ax2, bx2), yy2 = fminf(ay2, by2);
      float iw = fmaxf(0.f, xx2 - xx1 + one), ih = fmaxf(0.f, yy2 - yy1 + one);
      float inter = iw * ih, uni = area_a + area_b - inter;
      float iou = inter / uni;
      if (!(iou <= iou_thr)) dead[b] = 1;
    }
  }
  free(dead); free(c); free(idx);
  return kept;
}

int yfo_decode_nms(const int8_t* head, int gh, int gw, float out_scale, int out_zp,
                   float conf_thr, float iou_thr, int plus_one, yfo_det* dets, int max_det) {
  return yfo_decode_nms_ex(head, gh, gw, out_scale, out_zp, NULL, 8.f, conf_thr, iou_thr, plus_one, dets, max_det);
}

/* ------------------------------------------------------------------------------------------ */
/* Camera-side pre-processing: yoloface.c:26-71 (box average) + :73-93 (expand, -128)          */
/* ------------------------------------------------------------------------------------------ */
void yfo_rgb565_to_input(const uint8_t* src, int8_t* dst) {
  for (int y = 0; y < 56; ++y) for (int x = 0; x < 56; ++x) {
    uint32_t sr = 0, sg = 0, sb = 0;
    for (int dy = 0; dy < 2; ++dy) for (int dx = 0; dx < 2; ++dx) {
      size_t o = ((size_t)(2 * y + dy) * 112 + (2 * x + dx)) * 2;
      uint16_t px = (uint16_t)((src[o] << 8) | src[o + 1]);
      sr += (px >> 11) & 0x1F; sg += (px >> 5) & 0x3F; sb += px & 0x1F;
    }
    uint8_t r = (uint8_t)((sr >> 2) << 3), g = (uint8_t)((sg >> 2) << 2), b = (uint8_t)((sb >> 2) << 3);
    int8_t* o = dst + ((size_t)y * 56 + x) * 3;
    /* yoloface.c:88-90: (int8_t)r - 128 stored into ai_i8 (wraps modulo 256) */
    o[0] = (int8_t)(uint8_t)((int)(int8_t)r - 128);
    o[1] = (int8_t)(uint8_t)((int)(int8_t)g - 128);
    o[2] = (int8_t)(uint8_t)((int)(int8_t)b - 128);
  }
}
