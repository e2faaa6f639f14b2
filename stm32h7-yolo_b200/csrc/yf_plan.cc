// yf_plan.cc -- .tflite reader + lowering to fused device steps.  See yf_plan.h.
#include "yf_plan.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>

namespace yf {

// ------------------------------------------------------------------------------------------
// FlatBuffer access (TFL3 schema subset; field slots per SURVEY.md Appendix A)
// ------------------------------------------------------------------------------------------
namespace {
struct Cursor {
  const uint8_t* base; size_t len; size_t pos;   // pos = table position (0 = invalid)
  template <class T> T at(size_t o) const { T v{}; if (o + sizeof(T) <= len) std::memcpy(&v, base + o, sizeof(T)); return v; }
  size_t slot(int i) const {
    if (!pos) return 0;
    size_t vt = pos - static_cast<size_t>(static_cast<int64_t>(at<int32_t>(pos)));
    uint16_t vlen = at<uint16_t>(vt);
    if (4 + 2 * i >= vlen) return 0;
    uint16_t off = at<uint16_t>(vt + 4 + 2 * static_cast<size_t>(i));
    return off ? pos + off : 0;
  }
  template <class T> T scalar(int i, T dflt) const { size_t p = slot(i); return p ? at<T>(p) : dflt; }
  Cursor table(int i) const { size_t p = slot(i); return Cursor{base, len, p ? p + at<uint32_t>(p) : 0}; }
  // vector field: returns element count, *first = position of element 0
  uint32_t vec(int i, size_t* first) const {
    size_t p = slot(i); if (!p) { *first = 0; return 0; }
    size_t v = p + at<uint32_t>(p); *first = v + 4; return at<uint32_t>(v);
  }
  Cursor elem_table(size_t first, uint32_t k) const {
    size_t p = first + 4 * static_cast<size_t>(k); return Cursor{base, len, p + at<uint32_t>(p)};
  }
};
}  // namespace

bool TflModel::parse(const uint8_t* buf, size_t len, std::string* err) {
  if (!buf || len < 16 || std::memcmp(buf + 4, "TFL3", 4) != 0) { if (err) *err = "not a TFL3 flatbuffer"; return false; }
  bytes.assign(buf, buf + len);
  const uint8_t* b = bytes.data();
  Cursor root{b, len, static_cast<size_t>(Cursor{b, len, 0}.at<uint32_t>(0))};
  size_t f; uint32_t n = root.vec(1, &f);
  std::vector<int> codes(n);
  for (uint32_t i = 0; i < n; ++i) {
    Cursor c = root.elem_table(f, i);
    codes[i] = std::max<int>(c.scalar<int8_t>(0, 0), c.scalar<int32_t>(3, 0));
  }
  size_t bf; uint32_t nbuf = root.vec(4, &bf);
  size_t sf; if (root.vec(2, &sf) < 1) { if (err) *err = "no subgraph"; return false; }
  Cursor sg = root.elem_table(sf, 0);
  size_t tf; uint32_t nt = sg.vec(0, &tf);
  tensors.resize(nt);
  for (uint32_t i = 0; i < nt; ++i) {
    Cursor t = sg.elem_table(tf, i); TflTensor& T = tensors[i];
    size_t p; uint32_t r = t.vec(0, &p);
    for (uint32_t k = 0; k < r; ++k) T.shape.push_back(t.at<int32_t>(p + 4 * k));
    T.type = t.scalar<int8_t>(1, 0);
    uint32_t bi = t.scalar<uint32_t>(2, 0);
    if (bi < nbuf) {
      Cursor bt = root.elem_table(bf, bi); size_t d; uint32_t dl = bt.vec(0, &d);
      if (dl) { T.data = b + d; T.size = dl; }
    }
    uint32_t nl = t.vec(3, &p); T.name.assign(reinterpret_cast<const char*>(b + p), nl);
    Cursor q = t.table(4);
    if (q.pos) {
      uint32_t ns = q.vec(2, &p); for (uint32_t k = 0; k < ns; ++k) T.scale.push_back(q.at<float>(p + 4 * k));
      uint32_t nz = q.vec(3, &p); for (uint32_t k = 0; k < nz; ++k) T.zp.push_back(q.at<int64_t>(p + 8 * k));
      T.qdim = q.scalar<int32_t>(6, 0);
    }
  }
  size_t p; uint32_t c = sg.vec(1, &p); input = c ? sg.at<int32_t>(p) : -1;
  c = sg.vec(2, &p); output = c ? sg.at<int32_t>(p) : -1;
  size_t of; uint32_t no = sg.vec(3, &of);
  ops.resize(no);
  for (uint32_t i = 0; i < no; ++i) {
    Cursor o = sg.elem_table(of, i); TflOperator& O = ops[i];
    uint32_t ci = o.scalar<uint32_t>(0, 0); O.opcode = ci < codes.size() ? codes[ci] : -1;
    uint32_t k = o.vec(1, &p); for (uint32_t j = 0; j < k; ++j) O.in.push_back(o.at<int32_t>(p + 4 * j));
    k = o.vec(2, &p); O.out = k ? o.at<int32_t>(p) : -1;
    Cursor x = o.table(4);
    switch (O.opcode) {
      case OP_CONV_2D:
        O.padding_same = x.scalar<int8_t>(0, 0) == 0; O.stride_w = x.scalar<int32_t>(1, 1); O.stride_h = x.scalar<int32_t>(2, 1);
        O.fused_act = x.scalar<int8_t>(3, 0); break;
      case OP_DEPTHWISE_CONV_2D:
        O.padding_same = x.scalar<int8_t>(0, 0) == 0; O.stride_w = x.scalar<int32_t>(1, 1); O.stride_h = x.scalar<int32_t>(2, 1);
        O.depth_mult = x.scalar<int32_t>(3, 1); O.fused_act = x.scalar<int8_t>(4, 0); break;
      case OP_MAX_POOL_2D:
        O.padding_same = x.scalar<int8_t>(0, 0) == 0; O.stride_w = x.scalar<int32_t>(1, 1); O.stride_h = x.scalar<int32_t>(2, 1);
        O.filter_w = x.scalar<int32_t>(3, 1); O.filter_h = x.scalar<int32_t>(4, 1); O.fused_act = x.scalar<int8_t>(5, 0); break;
      case OP_ADD: O.fused_act = x.scalar<int8_t>(0, 0); break;
      case OP_CONCATENATION: O.axis = x.scalar<int32_t>(0, 0); O.fused_act = x.scalar<int8_t>(1, 0); break;
      case OP_LEAKY_RELU: O.alpha = x.scalar<float>(0, 0.f); break;
      default: break;
    }
    if (O.fused_act != 0) { if (err) *err = "fused activation on op " + std::to_string(i) + " is not supported"; return false; }
  }
  if (input < 0 || output < 0) { if (err) *err = "model without input/output"; return false; }
  return true;
}

// ------------------------------------------------------------------------------------------
// Fixed-point helpers
// ------------------------------------------------------------------------------------------
void quantize_multiplier(double d, int32_t* mult, int* shift) {
  if (d == 0.) { *mult = 0; *shift = 0; return; }
  const double q = std::frexp(d, shift);
  int64_t qf = static_cast<int64_t>(std::round(q * static_cast<double>(1LL << 31)));
  if (qf == (1LL << 31)) { qf /= 2; ++*shift; }
  if (*shift < -31) { *shift = 0; qf = 0; }
  *mult = static_cast<int32_t>(qf);
}
int32_t mbqm_host(int32_t x, int32_t mult, int shift) {
  int ls = shift > 0 ? shift : 0, rs = shift > 0 ? 0 : -shift;
  int64_t ab = static_cast<int64_t>(static_cast<int32_t>(static_cast<uint32_t>(x) << ls)) * mult;
  int32_t t = static_cast<int32_t>((ab + (1LL << 30)) >> 31);   // == SRDHM for |x*m| < 2^62
  if (rs == 0) return t;
  int32_t half = 1 << (rs - 1);
  return (t + half + (t >> 31)) >> rs;                            // == RoundingDivideByPOT
}

bool epi_lean_words(const EpiCh& k, int32_t out[4]) {
  if (k.ls != 0 || k.e < 1 || k.e > 13 || k.sgn_mask != -1 || k.mult <= (1 << 30)) return false;
  if ((k.add64 - (1LL << 30)) % k.mult != 0) return false;
  const long long b = (k.add64 - (1LL << 30)) / k.mult;           // bias'
  if (k.acc_bound >= (1 << 22) || b >= (1 << 22) || b <= -(1 << 22)) return false;
  const long long c2p = static_cast<long long>(k.c2) + (128LL << k.e);
  const long long kc = 128 + 256 * c2p;
  if (kc >= (1LL << 30) || kc <= -(1LL << 30)) return false;
  out[0] = static_cast<int32_t>(b * 512); out[1] = k.mult; out[2] = static_cast<int32_t>(kc); out[3] = 8 + k.e;
  return true;
}

static int round_up(int v, int a) { return (v + a - 1) / a * a; }
static int out_dim(bool same, int in, int k, int s) { return same ? (in + s - 1) / s : (in - k + s) / s; }
static int pad_before(int in, int k, int s, int out) { int t = (out - 1) * s + k - in; return t > 0 ? t / 2 : 0; }

std::vector<BlobPiece> st_blob_layout(const TflModel& m, size_t* total) {
  std::vector<BlobPiece> v; size_t off = 0;
  for (size_t i = 0; i < m.ops.size(); ++i) {
    const TflOperator& O = m.ops[i];
    if (O.opcode != OP_CONV_2D && O.opcode != OP_DEPTHWISE_CONV_2D) continue;
    BlobPiece p; p.op = static_cast<int>(i);
    p.w_len = m.tensors[O.in[1]].size; p.b_len = m.tensors[O.in[2]].size;
    p.w_off = off; off = (off + p.w_len + 3) & ~size_t(3);
    p.b_off = off; off = (off + p.b_len + 3) & ~size_t(3);
    v.push_back(p);
  }
  if (total) *total = off;
  return v;
}
std::vector<uint8_t> st_blob_from_model(const TflModel& m) {
  size_t total; auto lay = st_blob_layout(m, &total);
  std::vector<uint8_t> blob(total, 0);
  for (const BlobPiece& p : lay) {
    const TflOperator& O = m.ops[p.op];
    std::memcpy(blob.data() + p.w_off, m.tensors[O.in[1]].data, p.w_len);
    std::memcpy(blob.data() + p.b_off, m.tensors[O.in[2]].data, p.b_len);
  }
  return blob;
}

// ------------------------------------------------------------------------------------------
// Lowering
// ------------------------------------------------------------------------------------------
namespace {

struct Builder {
  const TflModel& m; Plan& P; std::string& err;
  const uint8_t* blob; size_t blob_len; int slot_align; bool st_act;
  std::vector<std::vector<int>> consumers;           // tensor -> ops
  std::vector<int> th, tw, tc;                       // propagated tensor shapes
  std::vector<char> done;                            // op already folded
  std::map<int, BlobPiece> pieces;

  bool fail(const std::string& s) { err = s; return false; }
  int sole_consumer(int t, int opcode) const {
    if (consumers[t].size() != 1) return -1;
    int o = consumers[t][0]; return m.ops[o].opcode == opcode ? o : -1;
  }
  const int8_t* weights(int op) const {
    auto it = pieces.find(op);
    if (blob && it != pieces.end()) return reinterpret_cast<const int8_t*>(blob + it->second.w_off);
    return reinterpret_cast<const int8_t*>(m.tensors[m.ops[op].in[1]].data);
  }
  const int32_t* bias(int op) const {
    auto it = pieces.find(op);
    if (blob && it != pieces.end()) return reinterpret_cast<const int32_t*>(blob + it->second.b_off);
    return reinterpret_cast<const int32_t*>(m.tensors[m.ops[op].in[2]].data);
  }

  bool propagate_shapes() {
    size_t nt = m.tensors.size(); th.assign(nt, 0); tw.assign(nt, 0); tc.assign(nt, 0);
    th[m.input] = P.H; tw[m.input] = P.W; tc[m.input] = m.tensors[m.input].shape.size() == 4 ? m.tensors[m.input].shape[3] : 0;
    for (const TflOperator& O : m.ops) {
      int x = O.in[0], y = O.out;
      switch (O.opcode) {
        case OP_PAD: { const int32_t* p = reinterpret_cast<const int32_t*>(m.tensors[O.in[1]].data);
          if (!p || m.tensors[O.in[1]].size < 32) return fail("PAD without constant paddings");
          th[y] = th[x] + p[2] + p[3]; tw[y] = tw[x] + p[4] + p[5]; tc[y] = tc[x]; break; }
        case OP_CONV_2D: { const auto& f = m.tensors[O.in[1]].shape;
          th[y] = out_dim(O.padding_same, th[x], f[1], O.stride_h); tw[y] = out_dim(O.padding_same, tw[x], f[2], O.stride_w); tc[y] = f[0]; break; }
        case OP_DEPTHWISE_CONV_2D: { const auto& f = m.tensors[O.in[1]].shape;
          th[y] = out_dim(O.padding_same, th[x], f[1], O.stride_h); tw[y] = out_dim(O.padding_same, tw[x], f[2], O.stride_w); tc[y] = f[3]; break; }
        case OP_MAX_POOL_2D:
          th[y] = out_dim(O.padding_same, th[x], O.filter_h, O.stride_h); tw[y] = out_dim(O.padding_same, tw[x], O.filter_w, O.stride_w); tc[y] = tc[x]; break;
        case OP_CONCATENATION: th[y] = th[x]; tw[y] = tw[x]; tc[y] = 0; for (int t : O.in) tc[y] += tc[t]; break;
        default: th[y] = th[x]; tw[y] = tw[x]; tc[y] = tc[x]; break;
      }
    }
    return true;
  }

  int new_buffer(int H, int W, int C, bool observer_only) {
    PBuffer b; b.H = H; b.W = W; b.C = C; b.CP = round_up(C, 16); b.observer_only = observer_only;
    P.buffers.push_back(b); return static_cast<int>(P.buffers.size()) - 1;
  }
  // give TFLite tensor t a home (unless it already has one, e.g. a concat slot)
  void place(int t, bool observer_only = false) {
    if (P.loc[t].buf >= 0) return;
    if (t == m.output) { P.loc[t] = TensorLoc{P.output_buf, 0, tc[t]}; return; }
    int b = new_buffer(th[t], tw[t], tc[t], observer_only);
    P.loc[t] = TensorLoc{b, 0, tc[t]};
  }

  int add_lut(const int8_t* table) {
    P.luts.insert(P.luts.end(), reinterpret_cast<const uint8_t*>(table), reinterpret_cast<const uint8_t*>(table) + 256);
    return static_cast<int>(P.luts.size() / 256) - 1;
  }
  int leaky_lut(int op) {   // activations.cc::LeakyReluPrepare + reference_ops::QuantizeLeakyRelu
    const TflOperator& O = m.ops[op];
    float s_in = m.tensors[O.in[0]].scale[0], s_out = m.tensors[O.out].scale[0];
    int32_t mi, ma; int si, sa;
    quantize_multiplier(static_cast<double>(static_cast<float>(s_in * O.alpha / s_out)), &ma, &sa);
    quantize_multiplier(static_cast<double>(static_cast<float>(s_in / s_out)), &mi, &si);
    int32_t zin = static_cast<int32_t>(m.tensors[O.in[0]].zp[0]), zout = static_cast<int32_t>(m.tensors[O.out].zp[0]);
    int8_t tab[256];
    for (int q = -128; q < 128; ++q) {
      int32_t u;
      if (st_act) {   // ST: float32 de-quantise, leak, re-quantise with round-half-even (nl_func_array_integer tables)
        volatile float v = (static_cast<float>(q) - static_cast<float>(zin)) * s_in;
        if (q < zin) v = v * 0.1f;
        v = v / s_out;
        u = static_cast<int32_t>(std::nearbyintf(v)) + zout;
      } else {
        int32_t v = q - zin;
        u = zout + (v >= 0 ? mbqm_host(v, mi, si) : mbqm_host(v, ma, sa));
      }
      tab[q + 128] = static_cast<int8_t>(std::min(127, std::max(-128, u)));
    }
    return add_lut(tab);
  }
  int quantize_lut(int op) {   // quantize.cc::Prepare + reference_ops::Requantize
    const TflOperator& O = m.ops[op];
    int32_t mult; int shift;
    quantize_multiplier(static_cast<double>(m.tensors[O.in[0]].scale[0]) / static_cast<double>(m.tensors[O.out].scale[0]), &mult, &shift);
    int32_t zin = static_cast<int32_t>(m.tensors[O.in[0]].zp[0]), zout = static_cast<int32_t>(m.tensors[O.out].zp[0]);
    int8_t tab[256];
    for (int q = -128; q < 128; ++q) {
      int32_t u;
      if (st_act) {
        // ST folds the QUANTIZE operators into forward_concat (network.c:2307-2313, 2631-2637), whose arithmetic is
        // inside the closed library.  The one int8 -> int8 rescaling rule of ST's generator that IS visible -- the
        // activation tables, network.c:2218..2902 -- is float32 with round-half-to-even; the ST mode applies the same
        // rule here.  A plausible reading, not a pinned one.
        volatile float v = (static_cast<float>(q) - static_cast<float>(zin)) * m.tensors[O.in[0]].scale[0];
        v = v / m.tensors[O.out].scale[0];
        u = static_cast<int32_t>(std::nearbyintf(v)) + zout;
      } else {
        u = mbqm_host(q - zin, mult, shift) + zout;
      }
      tab[q + 128] = static_cast<int8_t>(std::min(127, std::max(-128, u)));
    }
    return add_lut(tab);
  }
  int compose_luts(int a, int b) {   // b o a
    int8_t tab[256];
    for (int i = 0; i < 256; ++i) {
      int8_t mid = static_cast<int8_t>(P.luts[static_cast<size_t>(a) * 256 + i]);
      tab[i] = static_cast<int8_t>(P.luts[static_cast<size_t>(b) * 256 + (mid + 128)]);
    }
    return add_lut(tab);
  }

  // kernel_util.cc::PopulateConvolutionQuantizationParams, folded for the epilogue (yf_plan.h)
  bool make_epi(int op, int cout, const std::vector<int64_t>& wsum, const std::vector<int64_t>& wabs, Step* s) {
    const TflOperator& O = m.ops[op];
    const TflTensor& tin = m.tensors[O.in[0]]; const TflTensor& tf = m.tensors[O.in[1]]; const TflTensor& tout = m.tensors[O.out];
    const int32_t* bs = bias(op);
    int32_t zin = static_cast<int32_t>(tin.zp[0]), zout = static_cast<int32_t>(tout.zp[0]);
    s->epi_base = static_cast<int>(P.epi.size());
    for (int c = 0; c < cout; ++c) {
      float fs = tf.scale.size() > 1 ? tf.scale[c] : tf.scale[0];
      double eff = static_cast<double>(tin.scale[0]) * static_cast<double>(fs) / static_cast<double>(tout.scale[0]);
      int32_t mult; int shift; quantize_multiplier(eff, &mult, &shift);
      EpiCh e{}; e.mult = mult; e.ls = shift > 0 ? shift : 0; e.e = shift > 0 ? 0 : -shift;
      if (e.e > 24 || e.ls > 8) return fail("requant shift out of the folded-epilogue range on op " + std::to_string(op));
      int64_t biasf = static_cast<int64_t>(bs ? bs[c] : 0) - static_cast<int64_t>(zin) * wsum[c];
      if (std::llabs(biasf) > (1LL << 28)) return fail("folded bias too large on op " + std::to_string(op));
      // (bias' << ls) * mult must stay inside int64 with room for the accumulator's share: |bias' << ls| < 2^31
      if ((std::llabs(biasf) << e.ls) >= (1LL << 31)) return fail("left-shifted folded bias overflows on op " + std::to_string(op));
      e.add64 = (biasf << e.ls) * static_cast<int64_t>(mult) + (1LL << 30);
      e.c2 = (e.e > 0 ? (1 << (e.e - 1)) : 0) + zout * (1 << e.e);
      e.sgn_mask = e.e > 0 ? -1 : 0;
      e.acc_bound = static_cast<int32_t>(std::min<int64_t>(0x7fffffff, 128 * wabs[c] + std::llabs(biasf)));
      P.epi.push_back(e);
    }
    return true;
  }

  size_t push_weights(const std::vector<uint8_t>& img) {
    size_t off = (P.wblob.size() + 127) & ~size_t(127);
    P.wblob.resize(off, 0);
    P.wblob.insert(P.wblob.end(), img.begin(), img.end());
    return off;
  }

  // Epilogue chain shared by CONV / DEPTHWISE / MAX_POOL: optional LEAKY_RELU table, optional fused
  // ADD, optional QUANTIZE table, destination possibly a concat slot.  `t` = output tensor of the
  // main op; returns the tensor the step finally produces.
  bool fuse_tail(Step* s, int t, bool allow_add) {
    int cur = t;
    int lk = sole_consumer(cur, OP_LEAKY_RELU);
    if (lk >= 0) {
      place(cur, true); s->raw_buf = P.loc[cur].buf;
      s->lut1 = leaky_lut(lk); done[lk] = 1; s->ops.push_back(lk); cur = m.ops[lk].out;
    } else if (allow_add) {
      int ad = sole_consumer(cur, OP_ADD);
      if (ad >= 0) {
        const TflOperator& A = m.ops[ad];
        int other = A.in[0] == cur ? A.in[1] : A.in[0];
        if (P.loc[other].buf < 0) return fail("ADD operand not produced before op " + std::to_string(ad));
        const TflTensor& t1 = m.tensors[A.in[0]]; const TflTensor& t2 = m.tensors[A.in[1]]; const TflTensor& to = m.tensors[A.out];
        AddParams ap{}; ap.enabled = 1;
        // the kernel computes operand "x" = skip tensor, operand "y" = this conv's output
        bool conv_is_second = (A.in[1] == cur);
        float mx = std::max(t1.scale[0], t2.scale[0]);
        double twice = static_cast<double>(2 * mx);
        int32_t m1, m2, mo; int s1, s2, so;
        quantize_multiplier(static_cast<double>(t1.scale[0]) / twice, &m1, &s1);
        quantize_multiplier(static_cast<double>(t2.scale[0]) / twice, &m2, &s2);
        quantize_multiplier(twice / static_cast<double>(static_cast<float>(1 << 20) * to.scale[0]), &mo, &so);
        if (s1 > 0 || s2 > 0 || so > 0) return fail("ADD multipliers >= 1 unsupported");
        if (conv_is_second) { ap.zp1 = (int32_t)t1.zp[0]; ap.m1 = m1; ap.s1 = s1; ap.zp2 = (int32_t)t2.zp[0]; ap.m2 = m2; ap.s2 = s2; }
        else                { ap.zp1 = (int32_t)t2.zp[0]; ap.m1 = m2; ap.s1 = s2; ap.zp2 = (int32_t)t1.zp[0]; ap.m2 = m1; ap.s2 = s1; }
        ap.mo = mo; ap.so = so; ap.zp_out = (int32_t)to.zp[0];
        s->add = ap; s->add_buf = P.loc[other].buf; s->add_coff = P.loc[other].coff;
        place(cur, true); s->pre_add_buf = P.loc[cur].buf;
        done[ad] = 1; s->ops.push_back(ad); cur = A.out;
      }
    }
    int qz = sole_consumer(cur, OP_QUANTIZE);
    if (qz >= 0) {
      place(cur, true);
      if (s->lut1 >= 0) s->mid_buf = P.loc[cur].buf; else s->raw_buf = P.loc[cur].buf;
      int l = quantize_lut(qz);
      if (s->lut1 >= 0) { s->lut2 = l; s->lut_fused = compose_luts(s->lut1, l); } else { s->lut1 = l; }
      done[qz] = 1; s->ops.push_back(qz); cur = m.ops[qz].out;
    }
    if (s->lut_fused < 0) s->lut_fused = s->lut1;
    place(cur);
    s->out_buf = P.loc[cur].buf; s->out_coff = P.loc[cur].coff;
    return true;
  }

  // resolve the data input of a conv/depthwise/pool: fold an explicit PAD (top/left only here)
  bool resolve_input(const TflOperator& O, Step* s, int* src_tensor) {
    int x = O.in[0]; s->pad_t = s->pad_l = 0;
    // producer of x a PAD?
    for (size_t i = 0; i < m.ops.size(); ++i) if (m.ops[i].out == x && m.ops[i].opcode == OP_PAD) {
      const int32_t* p = reinterpret_cast<const int32_t*>(m.tensors[m.ops[i].in[1]].data);
      if (consumers[x].size() != 1) return fail("PAD output with several consumers");
      if (p[0] || p[1] || p[6] || p[7]) return fail("PAD on batch/channel axis");
      s->pad_t = p[2]; s->pad_l = p[4];           // bottom/right padding = plain out-of-bounds reads
      done[i] = 1; s->ops.push_back(static_cast<int>(i)); x = m.ops[i].in[0];
      break;
    }
    if (P.loc[x].buf < 0) return fail("input tensor " + std::to_string(x) + " not materialised");
    *src_tensor = x; s->in_buf = P.loc[x].buf; s->in_coff = P.loc[x].coff;
    s->Hin = th[x]; s->Win = tw[x]; s->Cin = tc[x];
    s->in_zp = static_cast<int>(m.tensors[x].zp[0]);
    return true;
  }

  bool lower_conv(int i) {
    const TflOperator& O = m.ops[i]; Step s{}; s.op_first = i; s.ops.push_back(i);
    int src; if (!resolve_input(O, &s, &src)) return false;
    const auto& fs = m.tensors[O.in[1]].shape;            // OHWI
    int cout = fs[0], kh = fs[1], kw = fs[2], cin = fs[3];
    s.kh = kh; s.kw = kw; s.stride = O.stride_h; s.Cout = cout;
    if (O.stride_h != O.stride_w) return fail("anisotropic stride");
    s.Hout = th[O.out]; s.Wout = tw[O.out];
    if (O.padding_same) { s.pad_t += pad_before(s.Hin, kh, s.stride, s.Hout); s.pad_l += pad_before(s.Win, kw, s.stride, s.Wout); }
    const int8_t* w = weights(i);
    std::vector<int64_t> wsum(cout, 0), wabs(cout, 0);
    for (int o = 0; o < cout; ++o) for (int k = 0; k < kh * kw * cin; ++k) {
      const int v = w[static_cast<size_t>(o) * kh * kw * cin + k]; wsum[o] += v; wabs[o] += v < 0 ? -v : v;
    }
    if (!make_epi(i, cout, wsum, wabs, &s)) return false;
    s.Npad = round_up(cout, 16);
    const PBuffer& ib = P.buffers[s.in_buf];
    std::vector<uint8_t> img;
    if (kh == 1 && kw == 1 && s.stride == 1) {
      s.kind = STEP_CONV1X1; s.name = "conv1x1_" + std::to_string(i);
      if (ib.is_input) return fail("1x1 conv directly on the network input is not supported");
      if (s.in_coff != 0) return fail("1x1 conv reading a concat slot");
      int cp = ib.CP; s.Kpad = round_up(cp, 32);
      // physical channel p of the input buffer -> logical channel of the conv input tensor
      std::vector<int> p2l(s.Kpad, -1);
      if (consumers_are_concat_slots(src)) {
        int l = 0; for (const auto& sl : concat_slots[src]) { for (int k = 0; k < sl.second; ++k) p2l[sl.first + k] = l++; }
      } else for (int c = 0; c < cin; ++c) p2l[c] = c;
      img.assign(static_cast<size_t>(s.Kpad / 16) * s.Npad * 16, 0);
      for (int p = 0; p < s.Kpad; ++p) if (p2l[p] >= 0) for (int o = 0; o < cout; ++o)
        img[(static_cast<size_t>(p / 16) * s.Npad + o) * 16 + p % 16] = static_cast<uint8_t>(w[static_cast<size_t>(o) * cin + p2l[p]]);
    } else {
      s.kind = STEP_CONV_IM2COL; s.name = "conv" + std::to_string(kh) + "x" + std::to_string(kw) + "_" + std::to_string(i);
      int K = kh * kw * cin; s.Kpad = round_up(K, 32);
      if (s.Kpad > 32) return fail("im2col conv with K > 32 not supported (op " + std::to_string(i) + ")");
      if (!ib.is_input) return fail("im2col conv is only implemented for the dense network input");
      if (kh != 3 || kw != 3 || s.stride != 2 || s.pad_t != 1 || s.pad_l != 1) return fail("im2col conv: only 3x3 stride 2 pad(1,1,0,0)");
      img.assign(static_cast<size_t>(s.Kpad / 16) * s.Npad * 16, 0);
      for (int k = 0; k < K; ++k) for (int o = 0; o < cout; ++o)
        img[(static_cast<size_t>(k / 16) * s.Npad + o) * 16 + k % 16] = static_cast<uint8_t>(w[static_cast<size_t>(o) * K + k]);
      // bands of output rows: input band (2*rows+2 input rows) must fit a 16 KB smem slot
      int rows = s.Hout; while (rows > 1 && static_cast<size_t>(2 * rows + 2) * s.Win * cin > 16384) rows = (rows + 1) / 2;
      s.band_rows = rows; s.bands = (s.Hout + rows - 1) / rows;
    }
    s.w_off = push_weights(img); s.w_bytes = img.size();
    P.macs_per_image += static_cast<long>(s.Hout) * s.Wout * cout * kh * kw * cin;
    if (!fuse_tail(&s, O.out, true)) return false;
    P.steps.push_back(s); done[i] = 1; return true;
  }

  bool lower_dw(int i) {
    const TflOperator& O = m.ops[i]; Step s{}; s.op_first = i; s.ops.push_back(i); s.kind = STEP_DW;
    s.name = "dwconv_" + std::to_string(i);
    int src; if (!resolve_input(O, &s, &src)) return false;
    const auto& fs = m.tensors[O.in[1]].shape;            // [1,KH,KW,C]
    int kh = fs[1], kw = fs[2], c = fs[3];
    if (O.depth_mult != 1 || kh != 3 || kw != 3 || O.stride_h != O.stride_w) return fail("depthwise: only 3x3, multiplier 1");
    if (P.buffers[s.in_buf].is_input || s.in_coff != 0) return fail("depthwise on input/concat slot unsupported");
    if (!concat_view_is_dense(src)) return fail("depthwise on a concat output whose slots are padded apart (op " + std::to_string(i) + ")");
    s.kh = kh; s.kw = kw; s.stride = O.stride_h; s.Cout = c; s.Hout = th[O.out]; s.Wout = tw[O.out];
    if (O.padding_same) { s.pad_t += pad_before(s.Hin, kh, s.stride, s.Hout); s.pad_l += pad_before(s.Win, kw, s.stride, s.Wout); }
    const int8_t* w = weights(i);
    std::vector<int64_t> wsum(c, 0), wabs(c, 0);
    for (int t = 0; t < 9; ++t) for (int ch = 0; ch < c; ++ch) { const int v = w[static_cast<size_t>(t) * c + ch]; wsum[ch] += v; wabs[ch] += v < 0 ? -v : v; }
    if (!make_epi(i, c, wsum, wabs, &s)) return false;
    // one-hot dp4a words: word[tap][ch] = (uint8)w << 8*(ch%4)
    int cp = P.buffers[s.in_buf].CP;
    std::vector<uint8_t> img(static_cast<size_t>(9) * cp * 4, 0);
    for (int t = 0; t < 9; ++t) for (int ch = 0; ch < c; ++ch) {
      uint32_t word = static_cast<uint32_t>(static_cast<uint8_t>(w[static_cast<size_t>(t) * c + ch])) << (8 * (ch % 4));
      std::memcpy(&img[(static_cast<size_t>(t) * cp + ch) * 4], &word, 4);
    }
    s.w_off = push_weights(img); s.w_bytes = img.size();
    P.macs_per_image += static_cast<long>(s.Hout) * s.Wout * c * 9;
    if (!fuse_tail(&s, O.out, false)) return false;
    P.steps.push_back(s); done[i] = 1; return true;
  }

  bool lower_pool(int i) {
    const TflOperator& O = m.ops[i]; Step s{}; s.op_first = i; s.ops.push_back(i); s.kind = STEP_MAXPOOL;
    s.name = "maxpool_" + std::to_string(i);
    int src; if (!resolve_input(O, &s, &src)) return false;
    if (P.buffers[s.in_buf].is_input || s.in_coff != 0) return fail("maxpool on input/concat slot unsupported");
    if (!concat_view_is_dense(src)) return fail("maxpool on a concat output whose slots are padded apart (op " + std::to_string(i) + ")");
    if (O.stride_h != O.stride_w) return fail("anisotropic stride");
    s.kh = O.filter_h; s.kw = O.filter_w; s.stride = O.stride_h; s.Cout = s.Cin; s.Hout = th[O.out]; s.Wout = tw[O.out];
    if (O.padding_same) { s.pad_t = pad_before(s.Hin, s.kh, s.stride, s.Hout); s.pad_l = pad_before(s.Win, s.kw, s.stride, s.Wout); }
    if (!fuse_tail(&s, O.out, false)) return false;
    P.steps.push_back(s); done[i] = 1; return true;
  }

  bool lower_lut(int i) {   // stand-alone LEAKY_RELU / QUANTIZE (producer had several consumers)
    const TflOperator& O = m.ops[i]; Step s{}; s.op_first = i; s.ops.push_back(i); s.kind = STEP_LUT;
    s.name = "lut_" + std::to_string(i);
    int x = O.in[0]; if (P.loc[x].buf < 0) return fail("LUT input not materialised");
    s.in_buf = P.loc[x].buf; s.in_coff = P.loc[x].coff; s.Hin = s.Hout = th[x]; s.Win = s.Wout = tw[x]; s.Cin = s.Cout = tc[x];
    s.lut1 = O.opcode == OP_LEAKY_RELU ? leaky_lut(i) : quantize_lut(i); s.lut_fused = s.lut1;
    place(O.out); s.out_buf = P.loc[O.out].buf; s.out_coff = P.loc[O.out].coff;
    P.steps.push_back(s); done[i] = 1; return true;
  }

  // concat bookkeeping: output tensor -> list of (physical channel offset, channels)
  std::map<int, std::vector<std::pair<int, int>>> concat_slots;
  bool consumers_are_concat_slots(int t) const { return concat_slots.count(t) != 0; }
  // Only the 1x1 conv maps physical to logical channels (p2l); every other reader needs the concat view to be the
  // plain channel sequence, i.e. slot k must start where slot k-1 ends.
  bool concat_view_is_dense(int t) const {
    auto it = concat_slots.find(t);
    if (it == concat_slots.end()) return true;
    int next = 0;
    for (const auto& sl : it->second) { if (sl.first != next) return false; next = sl.first + sl.second; }
    return true;
  }
  bool plan_concats() {
    for (size_t i = 0; i < m.ops.size(); ++i) {
      const TflOperator& O = m.ops[i]; if (O.opcode != OP_CONCATENATION) continue;
      if (O.axis != 3 && O.axis != -1) return fail("concat axis must be channels");
      int off = 0; std::vector<std::pair<int, int>> slots;
      for (int t : O.in) { off = round_up(off, slot_align); slots.push_back({off, tc[t]}); off += tc[t]; }
      int b = new_buffer(th[O.out], tw[O.out], off, false);
      for (size_t k = 0; k < O.in.size(); ++k) {
        int t = O.in[k];
        if (consumers[t].size() != 1) return fail("concat input with other consumers");
        const TflTensor& ti = m.tensors[t]; const TflTensor& to = m.tensors[O.out];
        if (ti.scale[0] != to.scale[0] || ti.zp[0] != to.zp[0]) return fail("concat input with different quantisation");
        P.loc[t] = TensorLoc{b, slots[k].first, tc[t]};
      }
      P.loc[O.out] = TensorLoc{b, 0, off, 1};
      concat_slots[O.out] = slots;
      done[i] = 1;
    }
    return true;
  }

  bool run() {
    size_t nt = m.tensors.size();
    consumers.assign(nt, {});
    for (size_t i = 0; i < m.ops.size(); ++i) for (int t : m.ops[i].in) if (t >= 0 && !m.tensors[t].data) consumers[t].push_back(static_cast<int>(i));
    done.assign(m.ops.size(), 0);
    if (blob) {
      size_t total; for (const BlobPiece& p : st_blob_layout(m, &total)) pieces[p.op] = p;
      if (blob_len < total) return fail("weights blob shorter than the ST layout");
    }
    if (P.H % 8 || P.W % 8 || P.H < 8 || P.W < 8) return fail("input height/width must be multiples of 8");
    if (!propagate_shapes()) return false;
    P.loc.assign(nt, TensorLoc{});
    // caller-visible dense buffers
    { PBuffer b; b.H = P.H; b.W = P.W; b.C = tc[m.input]; b.CP = tc[m.input]; b.is_input = true; P.buffers.push_back(b);
      P.input_buf = 0; P.loc[m.input] = TensorLoc{0, 0, tc[m.input]}; }
    { PBuffer b; b.H = th[m.output]; b.W = tw[m.output]; b.C = tc[m.output]; b.CP = tc[m.output]; b.is_output = true; P.buffers.push_back(b);
      P.output_buf = 1; }
    P.GH = th[m.output]; P.GW = tw[m.output];
    P.out_scale = m.tensors[m.output].scale[0]; P.out_zp = static_cast<int>(m.tensors[m.output].zp[0]);
    if (!plan_concats()) return false;
    for (size_t i = 0; i < m.ops.size(); ++i) {
      if (done[i]) continue;
      bool ok = true;
      switch (m.ops[i].opcode) {
        case OP_PAD: {   // must be folded by its consumer, which comes later
          int y = m.ops[i].out;
          if (consumers[y].size() != 1) return fail("PAD with several consumers");
          int c = m.ops[consumers[y][0]].opcode;
          if (c != OP_CONV_2D && c != OP_DEPTHWISE_CONV_2D) return fail("PAD not followed by a convolution");
          break; }
        case OP_CONV_2D: ok = lower_conv(static_cast<int>(i)); break;
        case OP_DEPTHWISE_CONV_2D: ok = lower_dw(static_cast<int>(i)); break;
        case OP_MAX_POOL_2D: ok = lower_pool(static_cast<int>(i)); break;
        case OP_LEAKY_RELU: case OP_QUANTIZE: ok = lower_lut(static_cast<int>(i)); break;
        case OP_ADD: return fail("stand-alone ADD (op " + std::to_string(i) + ") is not supported");
        default: return fail("operator code " + std::to_string(m.ops[i].opcode) + " is not supported");
      }
      if (!ok) return false;
    }
    for (size_t i = 0; i < m.ops.size(); ++i) if (!done[i]) return fail("op " + std::to_string(i) + " was not lowered");
    if (P.loc[m.output].buf != P.output_buf) return fail("network output was not produced into the head buffer");
    // arena layout (bytes per image); observer-only buffers go last
    size_t off = 0;
    for (int pass = 0; pass < 2; ++pass) {
      for (PBuffer& b : P.buffers) {
        if (b.is_input || b.is_output || b.observer_only != (pass == 1)) continue;
        b.offset = off; off += (static_cast<size_t>(b.H) * b.W * b.CP + 127) & ~size_t(127);
      }
      if (pass == 0) P.arena_bytes_per_image = off; else P.arena_bytes_per_image_observer = off;
    }
    return true;
  }
};

}  // namespace

bool build_plan(const TflModel& m, int H, int W, const uint8_t* blob, size_t blob_len, Plan* plan, std::string* err,
                int slot_align, bool st_activations) {
  *plan = Plan{}; plan->H = H; plan->W = W;
  std::string e;
  Builder b{m, *plan, e, blob, blob_len, slot_align, st_activations};
  bool ok = b.run();
  if (!ok && err) *err = e;
  return ok;
}

// ------------------------------------------------------------------------------------------
// Fused program: shared-memory placement by liveness + per-phase parameter blocks
// ------------------------------------------------------------------------------------------
bool build_fused(const Plan& P, FusedProgram* F, int threads, int cluster) {
  *F = FusedProgram{};
  if (threads != kFusedWorkerThreads && threads != kFusedLatThreads) { F->ok = false; F->why = "unsupported CTA shape"; return false; }
  if (cluster < 1 || cluster > kFusedMaxCluster || (cluster & (cluster - 1)) || (cluster > 1 && threads != kFusedLatThreads)) {
    F->ok = false; F->why = "unsupported cluster size"; return false;
  }
  F->cluster = cluster;
  const int wgs = threads / 128, ctrl_warp = threads / 32 - 1;
  const int tmem_cols = threads == kFusedLatThreads ? kFusedLatTmemCols : kFusedTmemCols;
  F->threads = threads; F->tmem_cols = tmem_cols;
  auto no = [&](const std::string& w) { F->ok = false; F->why = w; return false; };
  const int ns = static_cast<int>(P.steps.size());
  if (ns > kFusedMaxPhases) return no("too many steps");
  for (const EpiCh& e : P.epi) {
    int32_t w4[4];
    if (!epi_lean_words(e, w4)) return no("a channel's requantisation is outside the lean epilogue's range (shift, multiplier or accumulator bound)");
  }
  // live range of every buffer: first writer .. last reader (in step order)
  const int nb = static_cast<int>(P.buffers.size());
  std::vector<int> birth(nb, 1 << 30), death(nb, -1);
  for (int i = 0; i < ns; ++i) {
    const Step& s = P.steps[i];
    if (s.kind == STEP_LUT) return no("stand-alone table step");
    if (s.kind == STEP_CONV_IM2COL && i != 0) return no("im2col conv is not the first step");
    if (s.kind != STEP_CONV_IM2COL && P.buffers[s.in_buf].is_input) return no("network input read by a non-im2col step");
    if (P.buffers[s.out_buf].is_output && (s.kind != STEP_CONV1X1 || i != ns - 1)) return no("head not produced by the last 1x1 conv");
    if (s.out_coff % 16 || s.in_coff != 0 || s.add_coff % 16) return no("unaligned channel slot");
    if ((s.kind == STEP_CONV1X1 || s.kind == STEP_CONV_IM2COL) && s.Npad > 64) return no("N > 64");
    if (s.kind == STEP_CONV_IM2COL && s.Npad != 16) return no("first conv with more than 16 output channels");
    death[s.in_buf] = std::max(death[s.in_buf], i);
    if (s.add_buf >= 0) death[s.add_buf] = std::max(death[s.add_buf], i);
    birth[s.out_buf] = std::min(birth[s.out_buf], i); death[s.out_buf] = std::max(death[s.out_buf], i);
  }
  // padded buffers: produced by a conv step, consumed only by depthwise / pool steps
  std::vector<char> padded(nb, 0), has_prod_conv(nb, 0), bad_cons(nb, 0), has_cons(nb, 0);
  for (int i = 0; i < ns; ++i) {
    const Step& s = P.steps[i];
    if (s.kind == STEP_DW || s.kind == STEP_MAXPOOL) has_cons[s.in_buf] = 1; else bad_cons[s.in_buf] = 1;
    if (s.add_buf >= 0) bad_cons[s.add_buf] = 1;
    if ((s.kind == STEP_CONV1X1 || s.kind == STEP_CONV_IM2COL) && s.out_coff == 0) has_prod_conv[s.out_buf] = 1; else bad_cons[s.out_buf] = 1;
  }
  for (int b = 0; b < nb; ++b) padded[b] = has_prod_conv[b] && has_cons[b] && !bad_cons[b] && !P.buffers[b].is_input && !P.buffers[b].is_output;
  for (int i = 0; i < ns; ++i) if ((P.steps[i].kind == STEP_DW || P.steps[i].kind == STEP_MAXPOOL) && !padded[P.steps[i].in_buf])
    return no("depthwise/pool input " + P.steps[i].name + " is not a conv-produced, dw/pool-only buffer");

  // ---- image pairs: the trailing run of resolution-preserving steps at the head's resolution ("back") is executed
  //      once for two images stacked into a tall image of 2H+2 rows (two separator rows between them) ----
  int split = ns;
  while (split > 1) {
    const Step& s = P.steps[split - 1];
    const bool keeps = (s.kind == STEP_CONV1X1 || (s.kind == STEP_DW && s.stride == 1 && s.pad_t == 1 && s.pad_l == 1)) &&
                       s.Hin == P.GH && s.Win == P.GW && s.Hout == P.GH && s.Wout == P.GW;
    if (!keeps) break;
    --split;
  }
  std::vector<char> in_back(nb, 0), crossing(nb, 0);
  auto mark_back = [&]() {
    std::fill(in_back.begin(), in_back.end(), 0); std::fill(crossing.begin(), crossing.end(), 0);
    for (int i = split; i < ns; ++i) {
      const Step& s = P.steps[i];
      in_back[s.in_buf] = 1; in_back[s.out_buf] = 1; if (s.add_buf >= 0) in_back[s.add_buf] = 1;
    }
    for (int i = 0; i < split; ++i) {
      const Step& s = P.steps[i];
      if (in_back[s.out_buf]) crossing[s.out_buf] = 1;
      // a front step READING a buffer the back phases also touch would need the pair addressing too: give up pairing
      if (in_back[s.in_buf] || (s.add_buf >= 0 && in_back[s.add_buf])) return false;
    }
    for (int b = 0; b < nb; ++b) if (crossing[b] && (padded[b] || P.buffers[b].is_output)) return false;
    return true;
  };
  if (split < ns && (ns - split < 2 || !mark_back())) split = ns;
  if (split == ns) { std::fill(in_back.begin(), in_back.end(), 0); std::fill(crossing.begin(), crossing.end(), 0); }
  const int Hs = P.GH, Ws = P.GW, Hp = 2 * Hs + 2;        // single image / tall pair image at the head's resolution

  auto cells_of = [&](int b) {
    const PBuffer& B = P.buffers[b];
    const int H = in_back[b] ? Hp : B.H;
    return padded[b] ? (H + 2) * (B.W + 2) : H * B.W;
  };
  // word-plane stride: cells * 4 bytes, skewed so that the planes of the words one warp touches start
  // 32 / nw banks apart (a warp of the depthwise / pool phases covers ~32 / nw cells of each of its nw words)
  auto plane_stride = [](int cells, int nw) { int w = cells; while (w % 32 != (32 / std::max(nw, 1)) % 32) ++w; return w * 4; };
  auto words_of = [&](int b) { return (P.buffers[b].C + 3) / 4; };
  auto bytes_of = [&](int b) {
    const PBuffer& B = P.buffers[b];
    return ((padded[b] ? plane_stride(cells_of(b), words_of(b)) * words_of(b) : cells_of(b) * B.CP) + 127) & ~127;
  };
  // scratch regions live for exactly one phase: the first conv's A stages, the pools' row maxima
  std::vector<int> scratch_size(ns, 0);
  for (int i = 0; i < ns; ++i) {
    const Step& s = P.steps[i];
    if (s.kind == STEP_CONV_IM2COL) scratch_size[i] = 2 * wgs * 6144 + 2048;   // A stages: two rounds of one tile per warpgroup
    if (s.kind == STEP_MAXPOOL) scratch_size[i] = ((s.Cout + 3) / 4) * plane_stride(s.Hin * s.Wout, (s.Cout + 3) / 4);
    if (s.kind == STEP_CONV_IM2COL && ((s.Hout * s.Wout + 127) / 128) * s.Npad > tmem_cols)
      return no("accumulator tiles of the first conv exceed TMEM");
  }
  // first-fit placement in birth order (scratch of phase i is born with the buffers written at i).  Buffers that
  // cross from the front into the back phases hold image A while image B's front phases run: they are live throughout.
  struct Item { int id, birth, death, size; };          // id >= 0: buffer, id < 0: scratch of phase -id-1
  std::vector<Item> items;
  for (int b = 0; b < nb; ++b) if (!P.buffers[b].is_input && !P.buffers[b].is_output && !P.buffers[b].observer_only && death[b] >= 0)
    items.push_back(Item{b, crossing[b] ? 0 : birth[b], crossing[b] ? ns - 1 : death[b], bytes_of(b)});
  for (int i = 0; i < ns; ++i) if (scratch_size[i]) items.push_back(Item{-i - 1, i, i, (scratch_size[i] + 127) & ~127});
  // Offsets: place the items one by one at the lowest offset that clears every already placed item whose live range
  // overlaps.  The result depends on the order; the problem is tiny (~30 items), so several orders are tried -- by
  // birth, by size, by live-range length, then a fixed pseudo-random sequence of permutations -- and the smallest
  // arena wins (deterministic: the same plan always yields the same map, which the generated kernel relies on).
  std::vector<int> off(nb, -1), scratch_off(ns, -1);
  int arena = 1 << 30;
  {
    const int ni = static_cast<int>(items.size());
    auto place = [&](const std::vector<int>& order, std::vector<int>* offs) {
      offs->assign(ni, -1);
      int top = 0;
      for (int oi = 0; oi < ni; ++oi) {
        const Item& it = items[order[oi]];
        std::vector<std::pair<int, int>> busy;               // [off, end) of placed items alive at the same time
        for (int oj = 0; oj < oi; ++oj) {
          const Item& ot = items[order[oj]];
          if (ot.birth <= it.death && it.birth <= ot.death) busy.push_back({(*offs)[order[oj]], (*offs)[order[oj]] + ot.size});
        }
        std::sort(busy.begin(), busy.end());
        int c = 0;
        for (const auto& b : busy) { if (c + it.size <= b.first) break; c = std::max(c, b.second); }
        (*offs)[order[oi]] = c; top = std::max(top, c + it.size);
      }
      return top;
    };
    std::vector<int> order(ni), offs, best_offs;
    auto consider = [&]() { const int t = place(order, &offs); if (t < arena) { arena = t; best_offs = offs; } };
    for (int i = 0; i < ni; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return items[x].birth != items[y].birth ? items[x].birth < items[y].birth : items[x].id > items[y].id; });
    consider();
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return items[x].size > items[y].size; });
    consider();
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return items[x].death - items[x].birth > items[y].death - items[y].birth; });
    consider();
    uint32_t rs = 0x9E3779B9u;
    for (int trial = 0; trial < 2000; ++trial) {
      for (int i = ni - 1; i > 0; --i) { rs = rs * 1664525u + 1013904223u; std::swap(order[i], order[(rs >> 8) % static_cast<uint32_t>(i + 1)]); }
      consider();
    }
    if (ni == 0) arena = 0;
    for (int i = 0; i < ni; ++i) { if (items[i].id >= 0) off[items[i].id] = best_offs[i]; else scratch_off[-items[i].id - 1] = best_offs[i]; }
  }
  // smem map: [input image][arena][parameter slots][phase descriptors][barriers]
  const Step& s0 = P.steps[0];
  F->in_off = 16;                                // the im2col builder reads up to 4 bytes before the image
  F->in_bytes = s0.Hin * s0.Win * s0.Cin;
  if (F->in_bytes % 16) return no("input image size is not a multiple of 16 bytes");
  F->arena_off = (F->in_off + F->in_bytes + 16 + 127) & ~127;
  F->arena_bytes = (arena + 127) & ~127;
  F->slot_off = F->arena_off + F->arena_bytes;
  F->split = split;
  int overread_end = 0;                          // a 128-row MMA tile reads past the rows (and the K chunks) its buffer holds
  // parameter blocks
  int slot = 0;
  for (int i = 0; i < ns; ++i) {
    const Step& s = P.steps[i];
    const bool back = i >= split;
    FusedPhase ph{}; ph.kind = s.kind;
    ph.Hin = back ? Hp : s.Hin; ph.Win = s.Win; ph.Hout = back ? Hp : s.Hout; ph.Wout = s.Wout;
    ph.rows_in = ph.Hin * ph.Win; ph.rows_out = ph.Hout * ph.Wout;
    ph.pair = back ? 1 : 0;
    if (back) { ph.sep_y = Hs; ph.rows_a = Hs * Ws; ph.row_b0 = (Hs + 2) * Ws; ph.rows_single = (Hs + 2) * Ws; }
    else ph.rows_single = ph.rows_out;
    ph.stride = s.stride; ph.pad_t = s.pad_t; ph.pad_l = s.pad_l; ph.ksize = s.kh; ph.in_zp = s.in_zp;
    const PBuffer& ib = P.buffers[s.in_buf]; const PBuffer& ob = P.buffers[s.out_buf];
    ph.in_off = ib.is_input ? F->in_off : F->arena_off + off[s.in_buf];
    ph.in_cs = (ib.is_input ? ph.rows_in : cells_of(s.in_buf)) * 16;
    ph.in_wp = (!ib.is_input && padded[s.in_buf]) ? s.Win + 2 : 0;
    ph.in_ws = ph.in_wp ? plane_stride(cells_of(s.in_buf), words_of(s.in_buf)) : 0;
    ph.to_global = ob.is_output ? 1 : 0;
    ph.out_cs = (ob.is_output ? ph.rows_out : cells_of(s.out_buf)) * 16;
    ph.out_wp = (!ob.is_output && padded[s.out_buf]) ? s.Wout + 2 : 0;
    ph.out_ws = ph.out_wp ? plane_stride(cells_of(s.out_buf), words_of(s.out_buf)) : 0;
    if (ph.out_wp) { for (int j = i + 1; j < ns; ++j) if (P.steps[j].in_buf == s.out_buf) { ph.out_zp = P.steps[j].in_zp; break; } }
    ph.out_off = ob.is_output ? 0 : F->arena_off + off[s.out_buf] + (s.out_coff / 16) * ph.out_cs;
    if (!back && !ob.is_output && crossing[s.out_buf]) ph.out_pair_shift = (Hs + 2) * Ws * 16;   // image B's rows of the tall buffer
    ph.add_off = -1; ph.scratch_off = scratch_off[i] >= 0 ? F->arena_off + scratch_off[i] : -1;
    if (s.add.enabled) { ph.add = s.add; ph.add_cs = cells_of(s.add_buf) * 16; ph.add_off = F->arena_off + off[s.add_buf] + (s.add_coff / 16) * ph.add_cs; }
    ph.cout = s.Cout; ph.chunks_out = (s.Cout + 15) / 16; ph.epi_base = s.epi_base; ph.has_lut = s.lut_fused >= 0;
    ph.npad = s.Npad;
    ph.ntiles = (ph.rows_out + 127) / 128;
    ph.nw = (s.Cout + 3) / 4;
    const bool shared_front = cluster > 1 && !back;       // this phase's work is dealt over the whole cluster
    ph.per = ((shared_front && s.kind == STEP_DW) ? cluster * threads : threads) / ph.nw;
    ph.dy = ph.per / ph.Wout; ph.dx = ph.per % ph.Wout;
    ph.tpg = s.Npad ? std::max(1, std::min(ph.ntiles, tmem_cols / s.Npad)) : 0;
    if (ph.tpg && ph.ntiles > ph.tpg) {              // balance the groups (7 tiles: 4 + 3, not 4 + 3 -> same; 9: 3 x 3)
      const int groups = (ph.ntiles + ph.tpg - 1) / ph.tpg;
      ph.tpg = (ph.ntiles + groups - 1) / groups;
    }
    ph.scratch_ws = s.kind == STEP_MAXPOOL ? plane_stride(s.Hin * s.Wout, ph.nw) : 0;
    if (s.kind == STEP_CONV1X1) {
      if ((ph.ntiles + ph.tpg - 1) / ph.tpg > 4) return no("more than four tile groups in step " + s.name);
      auto meet_counts = [&](int rows, uint32_t (&own)[2]) {
        uint32_t v = 0;
        own[0] = own[1] = 0;
        for (int t0 = 0, g = 0; t0 < ph.ntiles; t0 += ph.tpg, ++g) {
          const int nt = std::min(ph.tpg, ph.ntiles - t0);
          int n = fused_has_rows(ctrl_warp, t0, nt, rows, ph.chunks_out, wgs) ? 0 : 1;
          for (int w = 0; w < 4 * wgs; ++w)
            if (fused_has_rows(w, t0, nt, rows, ph.chunks_out, wgs)) { ++n; own[g >> 1] |= 1u << (16 * (g & 1) + w); }
          v |= static_cast<uint32_t>(n) << (8 * g);
        }
        return v;
      };
      ph.grp_warps = meet_counts(ph.rows_out, ph.own);
      ph.grp_warps_single = meet_counts(ph.rows_single, ph.own_single);
      if (shared_front) {
        // one tile group; byte r / mask r = the warps of CTA rank r (virtual warpgroup wg * cluster + r of wgs * cluster)
        if (ph.ntiles > ph.tpg) return no("cluster shape: step " + s.name + " needs more than one tile group");
        ph.grp_warps = 0; ph.own[0] = ph.own[1] = 0;
        for (int r = 0; r < cluster; ++r) {
          int n = 1;                                     // the control warp always meets, as a row owner or only to release
          for (int w = 0; w < 4 * wgs; ++w) {
            const int vw = (((w >> 2) * cluster + r) << 2) | (w & 3);
            if (fused_has_rows(vw, 0, ph.ntiles, ph.rows_out, ph.chunks_out, wgs * cluster)) {
              if (w != ctrl_warp) ++n;
              ph.own[r >> 1] |= 1u << (16 * (r & 1) + w);
            }
          }
          ph.grp_warps |= static_cast<uint32_t>(n) << (8 * r);
        }
        ph.grp_warps_single = ph.grp_warps; ph.own_single[0] = ph.own[0]; ph.own_single[1] = ph.own[1];
      }
    }
    {
      // reciprocal multipliers; every quotient the kernel forms has x < 4096
      auto rcp = [&](int d, int xmax, uint32_t* out) {
        if (d <= 0) { *out = 0; return true; }
        const uint32_t m = static_cast<uint32_t>(((1u << 20) + d - 1) / d);
        for (int x = 0; x <= xmax; ++x) if (((static_cast<uint64_t>(x) * m) >> 20) != static_cast<uint64_t>(x / d) || static_cast<uint64_t>(x) * m >= (1ull << 32)) return false;
        *out = m; return true;
      };
      const int ncell = ph.out_wp ? 2 * ph.out_wp + 2 * ph.Hout : 0;
      const int xmax = std::max({cluster * threads, ph.ntiles * 128, ph.Hin * ph.Wout + threads, ncell * ph.nw + threads, ph.rows_out});
      if (xmax >= 4096 || !rcp(ph.nw, xmax, &ph.rcp_nw) || !rcp(ph.Wout, xmax, &ph.rcp_wout) || !rcp(ncell, xmax, &ph.rcp_ncell) || !rcp(ph.per, xmax, &ph.rcp_per))
        return no("index range of step " + s.name + " exceeds the kernel's mul-shift division");
    }
    ph.idesc = static_cast<int32_t>((2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(s.Npad >> 3) << 17) | (8u << 24));
    ph.adesc_lo = static_cast<uint32_t>(((s.kind == STEP_CONV_IM2COL ? 2048 : ph.in_cs) >> 4) & 0x3FFF) << 16;
    ph.bdesc_lo = static_cast<uint32_t>(((s.Npad * 16) >> 4) & 0x3FFF) << 16;
    // block: [weights][table][depthwise constants]
    std::vector<uint8_t> blk;
    auto put = [&](const void* src, size_t n) { size_t o = blk.size(); blk.resize((o + n + 15) & ~size_t(15), 0); std::memcpy(blk.data() + o, src, n); return static_cast<int>(o); };
    auto put_epi = [&]() {     // lean requant constants (epi_lean_words) per output channel, padded to whole chunks
      std::vector<uint8_t> e(static_cast<size_t>(ph.chunks_out) * 16 * 16, 0);
      for (int c = 0; c < s.Cout; ++c) {
        int32_t w4[4]; epi_lean_words(P.epi[s.epi_base + c], w4);
        std::memcpy(e.data() + static_cast<size_t>(c) * 16, w4, 16);
      }
      return put(e.data(), e.size());
    };
    if (s.kind == STEP_CONV1X1) {
      ph.nk = s.Kpad / 32;
      overread_end = std::max(overread_end, ph.in_off + (2 * ph.nk - 1) * ph.in_cs + ph.ntiles * 2048);
      ph.w_off = put(P.wblob.data() + s.w_off, s.w_bytes);
      ph.epi_off = put_epi();
    } else if (s.kind == STEP_CONV_IM2COL) {
      // K = 64 layout for the fused builder: chunk ky holds the 9 bytes (kx, c) of input row ky,
      // bytes 9..15 and chunk 3 carry zero weights (the A bytes there are don't-care)
      if (s.kh != 3 || s.kw != 3 || s.Cin != 3) return no("im2col conv shape");
      ph.nk = 2;
      const int8_t* w32 = reinterpret_cast<const int8_t*>(P.wblob.data() + s.w_off);   // [2][Npad][16], k = (ky*3+kx)*3+c
      std::vector<uint8_t> w64(static_cast<size_t>(4) * s.Npad * 16, 0);
      for (int o = 0; o < s.Cout; ++o) for (int ky = 0; ky < 3; ++ky) for (int j = 0; j < 9; ++j) {
        const int k = ky * 9 + j;
        w64[(static_cast<size_t>(ky) * s.Npad + o) * 16 + j] = static_cast<uint8_t>(w32[(static_cast<size_t>(k / 16) * s.Npad + o) * 16 + k % 16]);
      }
      ph.w_off = put(w64.data(), w64.size());
      ph.epi_off = put_epi();
    } else if (s.kind == STEP_DW) {
      ph.dw_off = put(P.wblob.data() + s.w_off, s.w_bytes);             // [9][CP] one-hot words
      // four arrays [bias9 | mult | kc | sh] (epi_lean_words), each [nw words][4 channels] int32: the threads of
      // a warp (consecutive words) read consecutive 16-byte groups of one array
      const size_t arr = static_cast<size_t>(ph.nw) * 16;
      std::vector<uint8_t> e(arr * 4, 0);
      for (int c = 0; c < s.Cout; ++c) {
        int32_t w4[4]; epi_lean_words(P.epi[s.epi_base + c], w4);
        uint8_t* b = e.data() + static_cast<size_t>(c) * 4;
        for (int f = 0; f < 4; ++f) std::memcpy(b + f * arr, &w4[f], 4);
      }
      ph.dwepi_off = put(e.data(), e.size());
    }
    if (ph.has_lut) ph.lut_off = put(P.luts.data() + static_cast<size_t>(s.lut_fused) * 256, 256);
    if (blk.empty()) blk.resize(16, 0);
    ph.param_off = static_cast<int>(F->params.size()); ph.param_bytes = static_cast<int>(blk.size());
    F->params.insert(F->params.end(), blk.begin(), blk.end());
    slot = std::max(slot, ph.param_bytes);
    F->phases.push_back(ph);
  }
  F->slot_bytes = (slot + 127) & ~127;
  F->desc_off = F->slot_off + kFusedParamSlots * F->slot_bytes;
  F->smem_bytes = F->desc_off + ((static_cast<int>(F->phases.size() * sizeof(FusedPhase)) + 127) & ~127) + 384;   // + barriers, TMEM slot, parameter-block table
  F->smem_bytes = std::max(F->smem_bytes, overread_end);      // the over-read bytes meet zero weights; they only have to exist
  F->smem_bytes_spec = std::max(F->desc_off + 384, overread_end);
  F->in_pf_phase = 1;                                         // prefetch of the next image rides on the first 1x1 conv phase
  for (int i = 1; i < ns; ++i) if (P.steps[i].kind == STEP_CONV1X1) { F->in_pf_phase = i; break; }
  if (F->in_pf_phase >= split) return no("no front 1x1 conv phase to carry the input prefetch");
  F->head_bytes = P.GH * P.GW * P.buffers[P.output_buf].C;
  if (F->smem_bytes > 200 * 1024) return no("activations do not fit shared memory (" + std::to_string(F->smem_bytes) + " bytes)");
  F->ok = true;
  return true;
}

}  // namespace yf
