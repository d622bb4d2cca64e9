// Double-precision multiply on the integer pipes.
//
// Why: on sm_100a DMUL / DFMA / DMMA share ONE FP64 pipe per SM sub-partition (tools/microbench/fp64_pipes.cu).  While
// the tile kernels keep that pipe full of DMMAs, every DMUL of the Phi-tile builder queues behind them for hundreds of
// cycles (ncu: 32 % of warp samples are builder DMULs in stall_math).  The builder's multiplies are only ~2 % of the
// pipe's work, so moving them to the idle integer / FMA pipes costs nothing there and frees the builder from the queue.
//
// dmul_emu(a, b): 64 x 64 -> high 64 bit product of the two significands, rounded to nearest (ties away from zero: the
// only difference from IEEE round-to-nearest-even, and only on exact ties), exponent and sign reassembled with integer
// ops.  Zeros, denormals, infinities, NaNs and results outside the normal range take the (rare, exact) a * b path.
#pragma once
#include <cstdint>

namespace grief {

template <bool kExactCarry = true>
__device__ __forceinline__ double dmul_emu(double a, double b) {
  const uint32_t ah = (uint32_t)__double2hiint(a), al = (uint32_t)__double2loint(a);
  const uint32_t bh = (uint32_t)__double2hiint(b), bl = (uint32_t)__double2loint(b);
  const uint32_t ea = (ah >> 20) & 0x7ffu, eb = (bh >> 20) & 0x7ffu;
  // significands with the implicit bit at bit 63
  const uint32_t Ah = (ah << 11) | (al >> 21) | 0x80000000u, Al = al << 11;
  const uint32_t Bh = (bh << 11) | (bl >> 21) | 0x80000000u, Bl = bl << 11;
  uint64_t hi;                                                   // top 64 bits of the 128-bit product, in [2^62, 2^64)
  if (kExactCarry) {
    hi = __umul64hi(((uint64_t)Ah << 32) | Al, ((uint64_t)Bh << 32) | Bl);
  } else {                                                       // drops < 3 units of bit 0: error <= 0.503 ulp
    hi = (uint64_t)Ah * Bh + __umulhi(Ah, Bl) + __umulhi(Al, Bh);
  }
  const uint32_t top = (uint32_t)(hi >> 63);
  const uint64_t hs = hi << (top ^ 1u);                          // normalised: bit 63 set
  const uint64_t m = (hs + 0x400ull) >> 11;                      // 53-bit significand, 2^53 on a rounding carry
  const int e = (int)(ea + eb + top) - 1023;                     // biased exponent
  if ((ea - 1u) < 2046u && (eb - 1u) < 2046u && (uint32_t)(e - 1) < 2045u) {
    const uint64_t bits = (((uint64_t)((ah ^ bh) & 0x80000000u)) << 32) | ((((uint64_t)(uint32_t)(e - 1)) << 52) + m);
    return __longlong_as_double((long long)bits);
  }
  return a * b;
}

}  // namespace grief
