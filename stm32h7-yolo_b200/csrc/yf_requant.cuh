// yf_requant.cuh -- the folded TFLite requantisation as both kernel families use it (device code only).
//
//   y = clamp(RoundingDivideByPOT(SaturatingRoundingDoublingHighMul(acc + bias', m), e) + zp_out)
// With x = acc + bias' (|x| < 2^22, checked on the host from the weights: x * 512 fits an int32) and 2^30 < m < 2^31:
//   t  = (x*m + 2^30) >> 31                      SRDHM
//      = (h + 2^7) >> 8        with h = (x * 512 * m) >> 32 = mulhi(x << 9, m)       (nested floor divisions)
//   y' = (t + c2p - [t < 0]) >> e                RoundingDivideByPOT + zp_out + 128 (c2p = 2^(e-1) + (zp_out + 128) << e)
//      = (h + 2^7 + 256 * c2p - 256 * [x < 0]) >> (8 + e)                            ([t < 0] == [x < 0] because m > 2^30)
// i.e. IMAD (x << 9 from the raw accumulator and the pre-shifted bias), IMAD.HI with the constant folded into its addend,
// one sign correction, one shift, one clamp -- and no 64-bit product (IMAD.WIDE issues at a quarter of the IMAD rate on
// this part, IMAD.HI at half: tools/microbench/pipes.cu).  tests/test_oracle_golden.py proves the identity against
// the literal TFLite form with Python integers.  Valid for 1 <= e <= 13 and no left shift (epi_lean_form, on the host).
#pragma once
#include <stdint.h>

#include "yf_plan.h"

namespace yf {

// per-channel requant constants as they travel in the parameter blocks (yf_plan.cc::build_fused)
struct alignas(16) EpiChF { int32_t bias9; int32_t mult; int32_t kc; int32_t sh; };   // bias' << 9 | m | 2^7 + 256 * c2p | 8 + e
static_assert(sizeof(EpiChF) == 16, "EpiChF layout");

// ---- fixed-point pieces -------------------------------------------------------------------------
// returns the int8 result + 128 clamped to [0,255] (a table index); acc is the raw accumulator (no bias)
// (written in PTX so that the sign correction stays folded into the IMAD.HI addend: IMAD, SHF, IMAD, IMAD.HI, SHF, VIMNMX --
//  left to itself the compiler re-associates it into eight instructions)
__device__ __forceinline__ int32_t requant_idx(int32_t acc, int32_t bias9, int32_t mult, int32_t kc, int32_t sh) {
  int32_t y;
  asm("{\n\t.reg .s32 a, sg, hi, u;\n\t"
      "mad.lo.s32 a, %1, 512, %2;\n\t"         // a = (acc + bias') << 9
      "shr.s32 sg, a, 31;\n\t"                 // -[a < 0]
      "mad.lo.s32 hi, sg, 256, %4;\n\t"        // 2^7 + 256 * c2p - 256 * [a < 0]
      "mad.hi.s32 u, a, %3, hi;\n\t"           // mulhi(a, m) + ...
      "shr.s32 %0, u, %5;\n\t}"
      : "=r"(y) : "r"(acc), "r"(bias9), "r"(mult), "r"(kc), "r"(sh));
  return __vimin_s32_relu(y, 255);
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
    idx[c] = requant_idx(static_cast<int32_t>(v[c]), k.bias9, k.mult, k.kc, k.sh);
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
