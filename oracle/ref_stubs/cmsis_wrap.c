/* Exports the reference tree's own CMSIS-NN fixed-point primitives
 * (stm32/Drivers/CMSIS/NN/Include/arm_nnsupportfunctions.h:210-263 -- the CMSIS-NN statement of TFLite's
 * SaturatingRoundingDoublingHighMul / RoundingDivideByPOT) so tests can pin the oracle's a11 primitives to them.
 * CAVEAT: the header targets ILP32 Cortex-M.  On this LP64 host `mult / (1UL << 31)` (line 225) divides UNSIGNED and
 * `Q31_MIN (0x80000000L)` is positive, so the doubling-high-mult is only meaningful here for non-negative products;
 * the tests restrict themselves accordingly.  The divide-by-power-of-two is exact on any host. */
#include "arm_nnsupportfunctions.h"
int32_t ref_sat_doubling_high_mult(int32_t a, int32_t b) { return arm_nn_sat_doubling_high_mult(a, b); }
int32_t ref_divide_by_power_of_two(int32_t x, int32_t e) { return arm_nn_divide_by_power_of_two(x, e); }
