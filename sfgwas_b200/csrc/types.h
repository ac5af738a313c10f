// types.h -- plain C++ types shared by host-only (g++) and device (nvcc) translation units.
#pragma once
#include <cstdint>

namespace sfg {

struct alignas(16) LimbConst {
    uint64_t q;        // modulus
    uint64_t qinv;     // q^-1 mod 2^64             (Lattigo ring.MredParams)
    uint64_t bred_hi;  // floor(2^128/q) >> 64      (Lattigo ring.BredParams[0])
    uint64_t bred_lo;  // floor(2^128/q) & (2^64-1) (Lattigo ring.BredParams[1])
    uint64_t ninv;     // N^-1 mod q
    uint64_t ninv_sh;  // floor(ninv * 2^64 / q)
    uint64_t r64;      // 2^64 mod q
    uint64_t r64_sh;   // floor(r64 * 2^64 / q)
};

// ---- host number theory (hostmath.cpp) ----
uint64_t h_mulmod(uint64_t a, uint64_t b, uint64_t q);
uint64_t h_powmod(uint64_t a, uint64_t e, uint64_t q);
uint64_t h_invmod(uint64_t a, uint64_t q);
uint64_t h_shoup(uint64_t w, uint64_t q);
uint64_t h_primitive_root(uint64_t q);
uint64_t h_bitrev(uint64_t x, int bits);
LimbConst h_limb_const(uint64_t q, int N);
uint64_t h_galois_element(int logN, int k);
void h_permute_ntt_index(int logN, uint64_t galEl, uint32_t *index);
void h_trig_tables(int M, double *roots_re_im, double *ddcos);

}  // namespace sfg
