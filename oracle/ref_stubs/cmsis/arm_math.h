/* Host stand-in for CMSIS-DSP's arm_math.h (absent from the reference tree) -- just enough for
 * stm32/Drivers/CMSIS/NN/Include/arm_nnsupportfunctions.h to compile with host gcc.  Test infrastructure. */
#ifndef YF_STUB_ARM_MATH_H
#define YF_STUB_ARM_MATH_H
#include <stdint.h>
typedef int8_t q7_t;
typedef int16_t q15_t;
typedef int32_t q31_t;
typedef int64_t q63_t;
#define __STATIC_FORCEINLINE static inline __attribute__((always_inline, unused))
#define __STATIC_INLINE static inline
#define __SIMD32(addr) (*(int32_t**)&(addr))
static inline uint32_t __ROR(uint32_t v, uint32_t s) { s &= 31; return s ? (v >> s) | (v << (32 - s)) : v; }
static inline uint32_t __SXTB16(uint32_t x) { return ((uint32_t)(int32_t)(int8_t)x & 0xffffu) | ((uint32_t)(int32_t)(int8_t)(x >> 16) << 16); }
#define __PKHBT(a, b, s) ((((uint32_t)(a)) & 0x0000ffffu) | ((((uint32_t)(b)) << (s)) & 0xffff0000u))
#endif
