// yf_pool.cuh -- the sliding-window line maximum both kernel families use for MAX_POOL_2D (device code only).
#pragma once
#include <stdint.h>

namespace yf {

// One line of a separable max-pool with window K, stride 2: a thread slides the window along a row (pass 1) or a
// column (pass 2), keeping the K taps in registers (a ring whose slots are compile-time: the loop body is unrolled over
// one ring revolution), so every input is loaded once per line -- the window-per-output form loads it K / 2 times.
// Positions outside [0, len) contribute the identity (0 in the biased form), i.e. the maximum is over in-bounds cells.
// BIASED_IN: the input already holds x ^ 0x80 per byte (pass 2 reads pass 1's row maxima).  emit(o, biased word).
template <int K, bool BIASED_IN, class Emit>
__device__ __forceinline__ void pool_line(const uint8_t* in, int in_step, int len, int pad, int o_begin, int o_end, Emit emit) {
  constexpr int S = 2, G = K / S;
  uint32_t ev[K], od[K];
  auto ld = [&](int x, uint32_t& e, uint32_t& o) {
    uint32_t v = BIASED_IN ? 0u : 0x80808080u;               // identity
    if (static_cast<unsigned>(x) < static_cast<unsigned>(len)) v = *reinterpret_cast<const uint32_t*>(in + x * in_step);
    if (!BIASED_IN) v ^= 0x80808080u;
    e = v & 0x00ff00ffu; o = v & 0xff00ff00u;
  };
  int x = o_begin * S - pad;                                  // first tap of the first window
#pragma unroll
  for (int j = 0; j < K - S; ++j) ld(x + j, ev[j], od[j]);
#pragma unroll 1
  for (int o = o_begin; o < o_end; o += G) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int t = 0; t < S; ++t) ld(x + K - S + t, ev[(K - S + g * S + t) % K], od[(K - S + g * S + t) % K]);
      if (o + g < o_end) {
        uint32_t me = ev[0], mo = od[0];
#pragma unroll
        for (int j = 1; j + 1 < K; j += 2) { me = __vimax3_u16x2(me, ev[j], ev[j + 1]); mo = __vimax3_u16x2(mo, od[j], od[j + 1]); }
        if ((K & 1) == 0) { me = __vmaxu2(me, ev[K - 1]); mo = __vmaxu2(mo, od[K - 1]); }
        emit(o + g, me | mo);
      }
      x += S;
    }
  }
}

}  // namespace yf
