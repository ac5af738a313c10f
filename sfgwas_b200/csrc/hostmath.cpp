// hostmath.cpp -- host number theory + high-precision table generation (compiled by g++, links libquadmath).
#include <quadmath.h>

#include <cmath>
#include <cstring>

#include "types.h"

namespace sfg {

typedef unsigned __int128 u128h;

uint64_t h_mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128h)a * b) % q); }
uint64_t h_powmod(uint64_t a, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q;
    a %= q;
    while (e) {
        if (e & 1) r = h_mulmod(r, a, q);
        a = h_mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}
uint64_t h_invmod(uint64_t a, uint64_t q) { return h_powmod(a, q - 2, q); }
uint64_t h_shoup(uint64_t w, uint64_t q) { return (uint64_t)((((u128h)w) << 64) / q); }
uint64_t h_bitrev(uint64_t x, int bits) {
    uint64_t r = 0;
    for (int i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}

// Same root selection as Lattigo ring.primitiveRoot: candidates g = 3, 4, 5, ... ; the first g with
// g^((q-1)/f) != 1 for every prime factor f of q-1.  The caller may instead pass psi explicitly (sfg_ctx_create).
uint64_t h_primitive_root(uint64_t q) {
    uint64_t factors[64];
    int nf = 0;
    uint64_t m = q - 1;
    for (uint64_t f = 2; f * f <= m; f += (f == 2 ? 1 : 2)) {
        if (m % f == 0) {
            factors[nf++] = f;
            while (m % f == 0) m /= f;
        }
    }
    if (m > 1) factors[nf++] = m;
    for (uint64_t g = 3;; g++) {
        bool ok = true;
        for (int i = 0; i < nf && ok; i++) ok = h_powmod(g, (q - 1) / factors[i], q) != 1;
        if (ok) return g;
    }
}

LimbConst h_limb_const(uint64_t q, int N) {
    LimbConst c;
    c.q = q;
    uint64_t qi = 1, qq = q;
    for (int i = 0; i < 63; i++) {  // Lattigo ring.MRedParams
        qi *= qq;
        qq *= qq;
    }
    c.qinv = qi;
    u128h u = (~(u128h)0) / q;  // floor(2^128/q) for odd q > 1 (Lattigo ring.BRedParams)
    c.bred_hi = (uint64_t)(u >> 64);
    c.bred_lo = (uint64_t)u;
    c.ninv = h_invmod((uint64_t)N % q, q);
    c.ninv_sh = h_shoup(c.ninv, q);
    c.r64 = (uint64_t)((((u128h)1) << 64) % q);
    c.r64_sh = h_shoup(c.r64, q);
    return c;
}

uint64_t h_galois_element(int logN, int k) {  // Lattigo GaloisElementForColumnRotationBy: 5^(k mod 2N) mod 2N
    uint64_t twoN = 2ULL << logN;
    uint64_t e = (uint64_t)((int64_t)k & (int64_t)(twoN - 1));
    uint64_t r = 1, b = 5;
    while (e) {
        if (e & 1) r = (r * b) % twoN;
        b = (b * b) % twoN;
        e >>= 1;
    }
    return r;
}

void h_permute_ntt_index(int logN, uint64_t galEl, uint32_t *index) {  // Lattigo ring.PermuteNTTIndex (App. B.4)
    uint64_t N = 1ULL << logN, mask = (N << 1) - 1;
    for (uint64_t i = 0; i < N; i++) {
        uint64_t t1 = 2 * h_bitrev(i, logN) + 1;
        uint64_t t2 = (((galEl * t1) & mask) - 1) >> 1;
        index[i] = (uint32_t)h_bitrev(t2, logN);
    }
}

// exp(2 pi i k / M) for k in [0, M] rounded to double, and cos(2 pi t / M) for t in [0, M) as double-double.
void h_trig_tables(int M, double *roots_re_im /* 2*(M+1) */, double *ddcos /* 2*M (hi, lo) */) {
    for (int k = 0; k <= M; k++) {
        __float128 ang = (__float128)2 * M_PIq * (__float128)k / (__float128)M;
        __float128 cr = cosq(ang), sr = sinq(ang);
        // exact values at multiples of M/4 (avoid tiny residues)
        if ((k * 4) % M == 0) {
            int qd = (k * 4) / M % 4;
            cr = (qd == 0) ? 1 : (qd == 2 ? -1 : 0);
            sr = (qd == 1) ? 1 : (qd == 3 ? -1 : 0);
        }
        roots_re_im[2 * k] = (double)cr;
        roots_re_im[2 * k + 1] = (double)sr;
        if (k < M) {
            double hi = (double)cr;
            double lo = (double)(cr - (__float128)hi);
            ddcos[2 * k] = hi;
            ddcos[2 * k + 1] = lo;
        }
    }
}

}  // namespace sfg
