// yf_requant.cuh -- the folded TFLite requantisation as both kernel families use it (device code only).
//
//   y = clamp(RoundingDivideByPOT(SaturatingRoundingDoublingHighMul(acc + bias', m), e) + zp_out)
// with the identities of DESIGN.md section 2:  t = ((acc + bias') * m + 2^30) >> 31;  y = (t + c2 + (t >> 31)) >> e.
// Valid for e >= 1 and no left shift (checked on the host before a kernel is given this form).
#pragma once
#include <stdint.h>

#include "yf_plan.h"

namespace yf {

// per-channel requant constants as they travel in the parameter blocks (yf_plan.cc::build_fused)
struct alignas(16) EpiChF { int32_t bias; int32_t mult; int32_t c2p; int32_t e; };
static_assert(sizeof(EpiChF) == 16, "EpiChF layout");

// ---- fixed-point pieces -------------------------------------------------------------------------
// returns the int8 result + 128 clamped to [0,255]; acc already holds the folded bias; c2p = half + (zp_out + 128) << e; needs e >= 1
__device__ __forceinline__ int32_t requant_idx(int32_t acc, int32_t mult, int32_t c2p, int32_t e) {
  const long long p = static_cast<long long>(acc) * static_cast<long long>(mult) + (1ll << 30);
  const int32_t t = static_cast<int32_t>(p >> 31);
  return __vimin_s32_relu((t + c2p + (t >> 31)) >> e, 255);
}
__device__ __forceinline__ int32_t mbqm_f(int32_t x, int32_t m, int s) {
  const long long ab = static_cast<long long>(x) * static_cast<long long>(m);
  const int32_t t = static_cast<int32_t>((ab + (1ll << 30)) >> 31);
  const int rs = -s;
  if (rs == 0) return t;
  return (t + (1 << (rs - 1)) + (t >> 31)) >> rs;
}
// reference_integer_ops::AddElementwise; x = skip operand, y = this conv's int8 output
static __device__ __noinline__ uint32_t add_word(uint32_t skipw, uint32_t yw, const AddParams a) {
  uint32_t o = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int32_t x = static_cast<int8_t>((skipw >> (8 * j)) & 0xff), y = static_cast<int8_t>((yw >> (8 * j)) & 0xff);
    const int32_t sx = mbqm_f((x - a.zp1) << 20, a.m1, a.s1), sy = mbqm_f((y - a.zp2) << 20, a.m2, a.s2);
    o |= static_cast<uint32_t>(max(-128, min(127, mbqm_f(sx + sy, a.mo, a.so) + a.zp_out)) & 0xff) << (8 * j);
  }
  return o;
}

// requantise NW 4-channel words of one accumulator row in ONE basic block (ILP across 4*NW chains)
template <int NW, bool LUT>
__device__ __forceinline__ void requant_words(const uint32_t (&v)[16], const EpiChF* ek, const uint8_t* lut, uint32_t (&w)[4]) {
  int32_t idx[NW * 4];
#pragma unroll
  for (int c = 0; c < NW * 4; ++c) {
    const EpiChF k = ek[c];
    idx[c] = requant_idx(static_cast<int32_t>(v[c]) + k.bias, k.mult, k.c2p, k.e);
  }
#pragma unroll
  for (int wi = 0; wi < NW; ++wi) {
    if (LUT)
      w[wi] = static_cast<uint32_t>(lut[idx[4 * wi]]) | (static_cast<uint32_t>(lut[idx[4 * wi + 1]]) << 8) |
              (static_cast<uint32_t>(lut[idx[4 * wi + 2]]) << 16) | (static_cast<uint32_t>(lut[idx[4 * wi + 3]]) << 24);
    else
      w[wi] = (static_cast<uint32_t>(idx[4 * wi]) | (static_cast<uint32_t>(idx[4 * wi + 1]) << 8) | (static_cast<uint32_t>(idx[4 * wi + 2]) << 16) |
               (static_cast<uint32_t>(idx[4 * wi + 3]) << 24)) ^ 0x80808080u;         // index -> int8
  }
}


// the same for the 1..4 real words of a 16-channel chunk (nwords is warp-uniform)
template <bool LUT>
__device__ __forceinline__ void requant_chunk(const uint32_t (&v)[16], const EpiChF* ek, const uint8_t* lut, int nwords, uint32_t (&w)[4]) {
  switch (nwords) {
    case 1: requant_words<1, LUT>(v, ek, lut, w); break;
    case 2: requant_words<2, LUT>(v, ek, lut, w); break;
    case 3: requant_words<3, LUT>(v, ek, lut, w); break;
    default: requant_words<4, LUT>(v, ek, lut, w); break;
  }
}

}  // namespace yf
