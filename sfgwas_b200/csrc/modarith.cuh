// modarith.cuh -- 64-bit RNS modular arithmetic for sm_100a (device side).
//
// Conventions (all values are residues modulo an odd prime q < 2^62):
//  * "Montgomery form" means x*2^64 mod q, exactly the representation the reference keeps its cached
//    plaintext diagonals and Lattigo keeps its switching keys in (gwas/matmult.go:401-440).
//  * mred128() is the reduction of gwas/matmult.go:291-324 (ReduceAndAddUint128) followed by the
//    canonical `eval.Reduce` of :357-359.
//  * Shoup multiplication is used for constants known ahead of time (NTT twiddles, N^-1, P^-1).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "types.h"

namespace sfg {

struct u128 {
    uint64_t lo, hi;
};

__device__ __forceinline__ uint64_t add_mod(uint64_t a, uint64_t b, uint64_t q) {
    uint64_t r = a + b;
    return r >= q ? r - q : r;
}
__device__ __forceinline__ uint64_t sub_mod(uint64_t a, uint64_t b, uint64_t q) {
    return a >= b ? a - b : a + q - b;
}
__device__ __forceinline__ uint64_t csub(uint64_t a, uint64_t q) { return a >= q ? a - q : a; }

// x * w mod q for w < q with precomputed wsh = floor(w*2^64/q); x may be any 64-bit value. Result in [0, 2q).
__device__ __forceinline__ uint64_t mul_shoup_lazy(uint64_t x, uint64_t w, uint64_t wsh, uint64_t q) {
    uint64_t qe = __umul64hi(x, wsh);
    return x * w - qe * q;
}
__device__ __forceinline__ uint64_t mul_shoup(uint64_t x, uint64_t w, uint64_t wsh, uint64_t q) {
    return csub(mul_shoup_lazy(x, w, wsh, q), q);
}

// 128-bit accumulate: acc += a*b  (gwas/matmult.go:247-289 MulCoeffsAndAdd128; wraps silently mod 2^128)
__device__ __forceinline__ void mac128(u128 &acc, uint64_t a, uint64_t b) {
    uint64_t lo = a * b, hi = __umul64hi(a, b);
    acc.lo += lo;
    acc.hi += hi + (acc.lo < lo);
}

// gwas/matmult.go:291-324 ReduceAndAddUint128 into a zero-initialised residue, then eval.Reduce (:357-359):
//   y = hi - hi64((lo*qinv)*q) + q   (u64 wrap-around as written) ;  return y mod q  (canonical)
__device__ __forceinline__ uint64_t mred128(u128 acc, const LimbConst &c) {
    uint64_t hhi = __umul64hi(acc.lo * c.qinv, c.q);
    uint64_t y = acc.hi - hhi + c.q;
    // BRedAdd (Lattigo): y - hi64(y*bred_hi)*q, one conditional subtraction
    uint64_t r = y - __umul64hi(y, c.bred_hi) * c.q;
    return csub(r, c.q);
}

// Lattigo ring.MRed(x, y): x*y*2^-64 mod q, canonical
__device__ __forceinline__ uint64_t mred(uint64_t x, uint64_t y, const LimbConst &c) {
    uint64_t lo = x * y, hi = __umul64hi(x, y);
    uint64_t hhi = __umul64hi(lo * c.qinv, c.q);
    uint64_t r = hi - hhi + c.q;
    return csub(r, c.q);
}

// gwas/matmult.go:433-440 MForm(a, q, u) = a*2^64 mod q (canonical), u = BredParams {hi, lo}
__device__ __forceinline__ uint64_t mform(uint64_t a, const LimbConst &c) {
    uint64_t mhi = __umul64hi(a, c.bred_lo);
    uint64_t r = (0 - (a * c.bred_hi + mhi)) * c.q;
    return csub(r, c.q);
}

// a mod q for any 64-bit a (Lattigo ring.BRedAdd)
__device__ __forceinline__ uint64_t bred_add(uint64_t a, const LimbConst &c) {
    uint64_t r = a - __umul64hi(a, c.bred_hi) * c.q;
    return csub(r, c.q);
}

// exact (a*b) mod q for canonical a, b via Barrett on the 128-bit product (used off the hot loops)
__device__ __forceinline__ uint64_t mul_mod(uint64_t a, uint64_t b, const LimbConst &c) {
    // a*b*2^64 * 2^-64: to Montgomery (mform) then Montgomery-multiply
    return mred(mform(a, c), b, c);
}

}  // namespace sfg
