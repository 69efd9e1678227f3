// 64-bit RNS modular arithmetic for the sm_100a kernels.
//
// Semantics follow SEAL 4.0 util/uintarithsmallmod.h as used by the reference runtime
// (reference: lib/Runtime/SEAL_HEVM.cpp delegates all arithmetic to seal::Evaluator):
// every value that leaves a kernel is canonical in [0,q); inside kernels values are
// kept lazily in [0,2q) / [0,4q) (Harvey butterflies), which does not change the bits
// of the canonical result.  All functions are __host__ __device__ so the warp-level
// NTT schedule can be replayed on the CPU by the test-only warp emulator
// (tests/emul/): the product never runs them on the host.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

typedef uint64_t u64;
typedef uint32_t u32;

struct Tw { // Shoup operand: value and floor(value * 2^64 / q)   (SEAL MultiplyUIntModOperand)
  u64 w, wq;
};

struct ModQ { // one RNS prime q = 2^60 - delta (SEAL-style 60-bit prime) with its Barrett ratio floor(2^128 / q)
  u64 q, ratio_lo, ratio_hi;
  u64 delta; // 2^60 - q  (< 2^32)
};

HD u64 mulhi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}

// x * w mod q, result in [0, 2q) for ANY 64-bit x (Shoup / Harvey lazy product)
// written as x*w + Q*(2^64 - q) so that both low products chain into one multiply-add sequence
HD u64 shoup_lazy(u64 x, Tw t, u64 q) { return x * t.w + mulhi64(x, t.wq) * (0 - q); }
HD u64 csub(u64 x, u64 m) { // x in [0,2m) -> [0,m)
  u64 y = x - m;
  return y < x ? y : x; // unsigned wrap: x < m  =>  y > x
}
HD u64 shoup_mul(u64 x, Tw t, u64 q) { return csub(shoup_lazy(x, t, q), q); }

// Lazy range control for q = 2^60 - delta, delta < 2^32:  v = k*2^60 + r  ==  r + k*delta (mod q).
// Any 64-bit v (k <= 15) is mapped into [0, 2^60 + 15*delta) which is inside [0, 2q): three integer
// instructions (shift, mask, 32x32->64 multiply-add) instead of a 6-instruction 64-bit compare/select.
HD u64 fold60(u64 v, u64 delta) {
  const u32 k = (u32)(v >> 60);
  return (v & 0x0FFFFFFFFFFFFFFFull) + (u64)k * (u32)delta;
}
// Cooley-Tukey butterfly without range correction: x' = x + w*y, y' = x - w*y + 2q.
// Each application grows the bound of its outputs by 2q; callers fold60() before 16q is reached.
HD void ct_bfly_lazy(u64 &x, u64 &y, Tw t, u64 q, u64 q2) {
  const u64 v = shoup_lazy(y, t, q);
  const u64 u = x;
  x = u + v;
  y = u + q2 - v;
}
// Gentleman-Sande butterfly, inputs in [0,2q): x' = fold(x + y) in [0,2q), y' = w*(x - y) in [0,2q)
HD void gs_bfly_fold(u64 &x, u64 &y, Tw t, u64 q, u64 q2, u64 delta) {
  const u64 u = x, v = y;
  x = fold60(u + v, delta);
  y = shoup_lazy(u + q2 - v, t, q);
}
// canonical representative of a lazily bounded value (any 64-bit input)
HD u64 canon60(u64 v, u64 q, u64 delta) { return csub(fold60(v, delta), q); }

// Harvey forward (Cooley-Tukey) butterfly: x,y in [0,4q) -> [0,4q)
HD void ct_bfly(u64 &x, u64 &y, Tw t, u64 q, u64 q2) {
  u64 u = csub(x, q2);
  u64 v = shoup_lazy(y, t, q);
  x = u + v;
  y = u + q2 - v;
}
// Harvey inverse (Gentleman-Sande) butterfly: x,y in [0,2q) -> [0,2q)
HD void gs_bfly(u64 &x, u64 &y, Tw t, u64 q, u64 q2) {
  u64 u = x, v = y;
  x = csub(u + v, q2);
  y = shoup_lazy(u + q2 - v, t, q);
}

// Barrett reduction of a 64-bit value (SEAL barrett_reduce_64)
HD u64 reduce64(u64 x, const ModQ &m) {
  u64 r = x - mulhi64(x, m.ratio_hi) * m.q;
  return csub(r, m.q);
}
// Barrett reduction of a 128-bit value (SEAL barrett_reduce_128)
HD u64 reduce128(u64 lo, u64 hi, const ModQ &m) {
  u64 carry = mulhi64(lo, m.ratio_lo);
  u64 t_lo = lo * m.ratio_hi, t_hi = mulhi64(lo, m.ratio_hi);
  u64 tmp1 = t_lo + carry;
  u64 tmp3 = t_hi + (tmp1 < t_lo ? 1 : 0);
  u64 u_lo = hi * m.ratio_lo, u_hi = mulhi64(hi, m.ratio_lo);
  u64 s = tmp1 + u_lo;
  carry = u_hi + (s < tmp1 ? 1 : 0);
  u64 qhat = hi * m.ratio_hi + tmp3 + carry;
  return csub(lo - qhat * m.q, m.q);
}
HD u64 mulmod(u64 a, u64 b, const ModQ &m) { return reduce128(a * b, mulhi64(a, b), m); }
HD u64 addmod(u64 a, u64 b, u64 q) { return csub(a + b, q); }
HD u64 submod(u64 a, u64 b, u64 q) { return csub(a + q - b, q); }
HD u64 negmod(u64 a, u64 q) { return a ? q - a : 0; }

// 128-bit lazy multiply-accumulate (SEAL switch_key_inplace inner loop: multiply_uint64 + add_uint128)
HD void mac128(u64 &lo, u64 &hi, u64 a, u64 b) {
  u64 pl = a * b, ph = mulhi64(a, b);
  lo += pl;
  hi += ph + (lo < pl ? 1 : 0);
}
