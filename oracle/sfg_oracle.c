/*
 * sfg_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See sfg_oracle.h.
 *
 * Restates, in plain C:
 *   - gwas/matmult.go (hhcho/sfgwas): lazy 128-bit MAC, Montgomery reduce-and-add, MForm,
 *     generalized-diagonal extraction, BSGS stream matmult (Preprocess / Compute / fused Stream);
 *   - the Lattigo v2.1 (fork hcholab/lattigo@e8d68c24b94a, NOT on disk -> [UNVERIFIED])
 *     ring NTT, hybrid key-switch, NTT-domain automorphism and CKKS big-float encoder.
 * Parity: "unpinned" for the Lattigo parts (no fork source, no Go toolchain, no reference tests).
 */
#define _GNU_SOURCE
#include "sfg_oracle.h"
#include <math.h>
#include <pthread.h>
#include <quadmath.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------------------------ */
/* scalar arithmetic                                                                          */
/* ------------------------------------------------------------------------------------------ */
static inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128)a * b) % q); }
static inline uint64_t addmod(uint64_t a, uint64_t b, uint64_t q) { uint64_t r = a + b; return r >= q ? r - q : r; }
static inline uint64_t submod(uint64_t a, uint64_t b, uint64_t q) { return a >= b ? a - b : a + q - b; }
static uint64_t powmod(uint64_t a, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q; a %= q;
    while (e) { if (e & 1) r = mulmod(r, a, q); a = mulmod(a, a, q); e >>= 1; }
    return r;
}
static uint64_t invmod(uint64_t a, uint64_t q) { return powmod(a, q - 2, q); } /* q prime */

/* Lattigo ring.MRedParams: q^-1 mod 2^64 (positive inverse, cf. gwas/matmult.go:300-301 sign) */
static uint64_t mred_params(uint64_t q) {
    uint64_t qInv = 1;
    for (int i = 0; i < 63; i++) { qInv *= q; q *= q; }
    return qInv;
}
/* Lattigo ring.BRedParams: floor(2^128/q) as {hi, lo} (consistent with MForm, gwas/matmult.go:433-440) */
static void bred_params(uint64_t q, uint64_t u[2]) {
    /* floor((2^128 - 1)/q) == floor(2^128/q) because q is odd > 1 (q does not divide 2^128) */
    u128 all = ~(u128)0;
    u128 v = all / q;
    u[0] = (uint64_t)(v >> 64);
    u[1] = (uint64_t)v;
}

/* Lattigo ring.MRed(x, y, q, qInv) = x*y*2^-64 mod q, canonical */
uint64_t orc_mred(uint64_t x, uint64_t y, uint64_t q, uint64_t qInv) {
    u128 m = (u128)x * y;
    uint64_t mhi = (uint64_t)(m >> 64), mlo = (uint64_t)m;
    uint64_t hhi = (uint64_t)(((u128)(mlo * qInv) * q) >> 64);
    uint64_t r = mhi - hhi + q;
    if (r >= q) r -= q;
    return r;
}

/* gwas/matmult.go:433-440  MForm(a, q, u): r = -(a*u[0] + hi64(a*u[1])) * q; if r >= q: r -= q  (= a*2^64 mod q) */
uint64_t orc_mform(uint64_t a, uint64_t q, const uint64_t u[2]) {
    uint64_t mhi = (uint64_t)(((u128)a * u[1]) >> 64);
    uint64_t r = (uint64_t)(0 - (a * u[0] + mhi)) * q;
    if (r >= q) r -= q;
    return r;
}

/* Lattigo ring.BRedAdd(a, q, u): a mod q for any 64-bit a */
uint64_t orc_bred_add(uint64_t a, uint64_t q, const uint64_t u[2]) {
    uint64_t mhi = (uint64_t)(((u128)a * u[0]) >> 64);
    uint64_t r = a - mhi * q;
    if (r >= q) r -= q;
    return r;
}

/* ------------------------------------------------------------------------------------------ */
/* context                                                                                    */
/* ------------------------------------------------------------------------------------------ */
struct orc_ctx {
    int logN, N, slots, nQ, nP, nQP;
    double scale;
    uint64_t mod[ORC_MAXMOD], mred[ORC_MAXMOD], bred[ORC_MAXMOD][2];
    uint64_t psi[ORC_MAXMOD], psiInv[ORC_MAXMOD];
    uint64_t *nttPsi[ORC_MAXMOD], *nttPsiInv[ORC_MAXMOD]; /* bit-reversed, Montgomery form */
    uint64_t nInvMont[ORC_MAXMOD];
    /* encoder tables (quad precision) */
    int *rotGroup;        /* 5^j mod 2N, j < slots */
    __float128 *rootsRe, *rootsIm; /* exp(2 pi i k / 2N), k <= 2N */
};

static uint64_t bitrev(uint64_t x, int bits) {
    uint64_t r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* Lattigo ring.primitiveRoot: smallest g >= 3 (search starts at g=2, increments first) such that
 * g^((q-1)/f) != 1 for every prime factor f of q-1.  [UNVERIFIED vs fork] */
uint64_t orc_primitive_root(uint64_t q) {
    uint64_t factors[64]; int nf = 0;
    uint64_t m = q - 1;
    for (uint64_t f = 2; f * f <= m; f += (f == 2 ? 1 : 2)) {
        if (m % f == 0) { factors[nf++] = f; while (m % f == 0) m /= f; }
    }
    if (m > 1) factors[nf++] = m;
    uint64_t g = 2;
    for (;;) {
        g++;
        int ok = 1;
        for (int i = 0; i < nf; i++) if (powmod(g, (q - 1) / factors[i], q) == 1) { ok = 0; break; }
        if (ok) return g;
    }
}

orc_ctx *orc_ctx_new(int logN, const uint64_t *qi, int nQ, const uint64_t *pi, int nP, double scale) {
    if (nQ + nP > ORC_MAXMOD) return NULL;
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
    c->logN = logN; c->N = 1 << logN; c->slots = c->N >> 1; c->nQ = nQ; c->nP = nP; c->nQP = nQ + nP;
    c->scale = scale;
    int N = c->N;
    for (int i = 0; i < c->nQP; i++) {
        uint64_t q = i < nQ ? qi[i] : pi[i - nQ];
        c->mod[i] = q;
        c->mred[i] = mred_params(q);
        bred_params(q, c->bred[i]);
        /* Lattigo ring.genNTTParams (App. B.3): psi = g^((q-1)/2N) */
        uint64_t g = orc_primitive_root(q);
        uint64_t twoN = (uint64_t)N << 1;
        uint64_t power = (q - 1) / twoN, powerInv = (q - 1) - power;
        c->psi[i] = powmod(g, power, q);
        c->psiInv[i] = powmod(g, powerInv, q);
        uint64_t psiMont = orc_mform(c->psi[i], q, c->bred[i]);
        uint64_t psiInvMont = orc_mform(c->psiInv[i], q, c->bred[i]);
        c->nttPsi[i] = (uint64_t *)malloc(sizeof(uint64_t) * N);
        c->nttPsiInv[i] = (uint64_t *)malloc(sizeof(uint64_t) * N);
        c->nttPsi[i][0] = orc_mform(1, q, c->bred[i]);
        c->nttPsiInv[i][0] = c->nttPsi[i][0];
        for (int j = 1; j < N; j++) {
            uint64_t prev = bitrev(j - 1, logN), next = bitrev(j, logN);
            c->nttPsi[i][next] = orc_mred(c->nttPsi[i][prev], psiMont, q, c->mred[i]);
            c->nttPsiInv[i][next] = orc_mred(c->nttPsiInv[i][prev], psiInvMont, q, c->mred[i]);
        }
        c->nInvMont[i] = orc_mform(invmod((uint64_t)N, q), q, c->bred[i]);
    }
    /* encoder tables (Lattigo ckks encoder: m = 2N, rotGroup = 5^j mod m, roots = exp(2 pi i k/m)) */
    int M = 2 * N;
    c->rotGroup = (int *)malloc(sizeof(int) * c->slots);
    uint64_t five = 1;
    for (int j = 0; j < c->slots; j++) { c->rotGroup[j] = (int)five; five = (five * 5) % (uint64_t)M; }
    c->rootsRe = (__float128 *)malloc(sizeof(__float128) * (M + 1));
    c->rootsIm = (__float128 *)malloc(sizeof(__float128) * (M + 1));
    for (int k = 0; k <= M; k++) {
        __float128 ang = 2.0Q * M_PIq * (__float128)k / (__float128)M;
        c->rootsRe[k] = cosq(ang); c->rootsIm[k] = sinq(ang);
    }
    return c;
}
void orc_ctx_free(orc_ctx *c) {
    if (!c) return;
    for (int i = 0; i < c->nQP; i++) { free(c->nttPsi[i]); free(c->nttPsiInv[i]); }
    free(c->rotGroup); free(c->rootsRe); free(c->rootsIm); free(c);
}
int orc_ctx_N(const orc_ctx *c) { return c->N; }
int orc_ctx_nQ(const orc_ctx *c) { return c->nQ; }
int orc_ctx_nP(const orc_ctx *c) { return c->nP; }
uint64_t orc_ctx_modulus(const orc_ctx *c, int i) { return c->mod[i]; }
uint64_t orc_ctx_mred(const orc_ctx *c, int i) { return c->mred[i]; }
void orc_ctx_bred(const orc_ctx *c, int i, uint64_t out[2]) { out[0] = c->bred[i][0]; out[1] = c->bred[i][1]; }
uint64_t orc_ctx_psi(const orc_ctx *c, int i) { return c->psi[i]; }

/* ------------------------------------------------------------------------------------------ */
/* NTT: Lattigo ring.NTT / ring.InvNTT (App. B.3). Cooley-Tukey forward, Gentleman-Sande      */
/* inverse, twiddles bit-reversed in Montgomery form, outputs canonical.                      */
/* ------------------------------------------------------------------------------------------ */
/* Evaluated like Lattigo's NTTLazy / InvNTTLazy: Harvey butterflies on values kept in [0, 4q) (forward) / [0, 2q) (inverse) with
 * MRedConstant (no conditional subtraction inside the transform), one canonicalising pass at the end.  Outputs are the same canonical
 * residues as the textbook butterflies (exact arithmetic mod q); only the CPU time differs, which matters for the CPU-baseline arm. */
static inline uint64_t mred_lazy(uint64_t x, uint64_t y, uint64_t q, uint64_t qInv) { /* x*y*2^-64 mod q in (0, 2q), x*y < q*2^64 */
    u128 m = (u128)x * y;
    uint64_t hhi = (uint64_t)(((u128)((uint64_t)m * qInv) * q) >> 64);
    return (uint64_t)(m >> 64) - hhi + q;
}
static inline uint64_t csub(uint64_t x, uint64_t m) { return x - (m & (0 - (uint64_t)(x >= m))); } /* branch-free x >= m ? x - m : x */
void orc_ntt(const orc_ctx *c, int idx, uint64_t *a) {
    const uint64_t q = c->mod[idx], qInv = c->mred[idx], q2 = 2 * q;
    const uint64_t *psi = c->nttPsi[idx];
    int N = c->N, t = N;
    for (int m = 1; m < N; m <<= 1) {
        t >>= 1;
        for (int i = 0; i < m; i++) {
            int j1 = 2 * i * t, j2 = j1 + t;
            uint64_t F = psi[m + i];
            for (int j = j1; j < j2; j++) {
                uint64_t U = csub(a[j], q2), V = mred_lazy(a[j + t], F, q, qInv); /* U, V in [0, 2q) */
                a[j] = U + V;
                a[j + t] = U - V + q2;
            }
        }
    }
    for (int j = 0; j < N; j++) a[j] = csub(csub(a[j], q2), q);
}
void orc_intt(const orc_ctx *c, int idx, uint64_t *a) {
    const uint64_t q = c->mod[idx], qInv = c->mred[idx], q2 = 2 * q;
    const uint64_t *psi = c->nttPsiInv[idx];
    int N = c->N, t = 1;
    for (int m = N; m > 1; m >>= 1) {
        int h = m >> 1, j1 = 0;
        for (int i = 0; i < h; i++) {
            int j2 = j1 + t;
            uint64_t F = psi[h + i];
            for (int j = j1; j < j2; j++) {
                uint64_t U = a[j], V = a[j + t]; /* in [0, 2q) */
                a[j] = csub(U + V, q2);
                a[j + t] = mred_lazy(U - V + q2, F, q, qInv);
            }
            j1 += t << 1;
        }
        t <<= 1;
    }
    for (int j = 0; j < N; j++) a[j] = orc_mred(a[j], c->nInvMont[idx], q, qInv);
}

/* ------------------------------------------------------------------------------------------ */
/* K1 / K2 / K3 : gwas/matmult.go:247-324, 411-440                                            */
/* ------------------------------------------------------------------------------------------ */
/* gwas/matmult.go:247-289 MulCoeffsAndAdd128: c[j] += a[j]*b[j] (128-bit, wraps silently) */
void orc_mul_coeffs_and_add128(const uint64_t *a, const uint64_t *b, orc_u128 *c, size_t n) {
    for (size_t j = 0; j < n; j++) {
        u128 p = (u128)a[j] * b[j];
        uint64_t hi = (uint64_t)(p >> 64), lo = (uint64_t)p;
        uint64_t nlo = c[j].lo + lo;
        uint64_t carry = nlo < lo;
        c[j].lo = nlo;
        c[j].hi += hi + carry;
    }
}
/* gwas/matmult.go:291-324 ReduceAndAddUint128: out[j] += in.hi - hi64((in.lo*qInv)*q) + q  (NOT range-reduced) */
void orc_reduce_and_add_uint128(const orc_u128 *in, uint64_t *out, uint64_t qInv, uint64_t q, size_t n) {
    for (size_t j = 0; j < n; j++) {
        uint64_t hhi = (uint64_t)(((u128)(in[j].lo * qInv) * q) >> 64);
        out[j] += in[j].hi - hhi + q;
    }
}
/* gwas/matmult.go:411-431 MFormLvl over limbs 0..level */
void orc_mform_lvl(const orc_ctx *c, int level, uint64_t *p) {
    for (int i = 0; i <= level; i++)
        for (int j = 0; j < c->N; j++) p[(size_t)i * c->N + j] = orc_mform(p[(size_t)i * c->N + j], c->mod[i], c->bred[i]);
}
/* fork-specific eval.Reduce (gwas/matmult.go:357-359): canonical x mod q on every coefficient [UNVERIFIED: assumed BRedAdd] */
void orc_reduce_canonical(const orc_ctx *c, int nlimbs, uint64_t *p) {
    for (int i = 0; i < nlimbs; i++)
        for (int j = 0; j < c->N; j++) p[(size_t)i * c->N + j] = orc_bred_add(p[(size_t)i * c->N + j], c->mod[i], c->bred[i]);
}

/* ------------------------------------------------------------------------------------------ */
/* diagonals : gwas/matmult.go:627-664                                                        */
/* ------------------------------------------------------------------------------------------ */
static inline int imod(int n, int m) { n %= m; return n < 0 ? n + m : n; } /* crypto/utilities.go:26-32 */

/* gwas/matmult.go:627-631 */
int orc_get_diag_bool(int r, int cdim, int dim, int index) {
    index = imod(index, dim);
    return (dim + 1 - r) <= index || index <= cdim - 1;
}
/* gwas/matmult.go:636-664. X is an r x cdim int8 block with leading dimension ld. dst has `dim` entries. */
int orc_get_diag(double *dst, const int8_t *X, size_t ld, int r, int cdim, int dim, int index) {
    index = imod(index, dim);
    if ((dim + 1 - r) <= index || index <= cdim - 1) {
        int i = imod(-index, dim);
        for (int j = 0; j < dim; j++) {
            dst[j] = (i < r && j < cdim) ? (double)X[(size_t)i * ld + j] : 0.0;
            i = imod(i + 1, dim);
        }
        return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* encoder: Lattigo ckks EncoderBig.EncodeNTT (App. B.6) at quad precision (113-bit mantissa). */
/* The reference uses 256-bit big.Float; both are "correctly rounded" for these inputs with     */
/* overwhelming probability (distance of scale*w_k to a half-integer vs. 1e-24 abs error).      */
/* ------------------------------------------------------------------------------------------ */
static void special_invfft_q(const orc_ctx *c, __float128 *re, __float128 *im) {
    int n = c->slots, M = 2 * c->N;
    for (int len = n; len >= 1; len >>= 1) {
        int lenh = len >> 1, lenq = len << 2, gap = M / lenq;
        for (int i = 0; i < n; i += len) {
            for (int j = 0; j < lenh; j++) {
                int idx = (lenq - (c->rotGroup[j] % lenq)) * gap;
                __float128 ur = re[i + j] + re[i + j + lenh], ui = im[i + j] + im[i + j + lenh];
                __float128 vr = re[i + j] - re[i + j + lenh], vi = im[i + j] - im[i + j + lenh];
                __float128 wr = c->rootsRe[idx], wi = c->rootsIm[idx];
                re[i + j] = ur; im[i + j] = ui;
                re[i + j + lenh] = vr * wr - vi * wi;
                im[i + j + lenh] = vr * wi + vi * wr;
            }
        }
    }
    int bits = c->logN - 1;
    for (int i = 0; i < n; i++) { re[i] /= (__float128)n; im[i] /= (__float128)n; }
    for (int i = 0; i < n; i++) {
        int j = (int)bitrev((uint64_t)i, bits);
        if (i < j) { __float128 t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
    }
}

/* round half away from zero (Lattigo scaleUpVecExactBigFloat: x +- 0.5 then truncate) */
static int64_t round_away_q(__float128 x) {
    if (x < 0) return -(int64_t)floorq(-x + 0.5Q);
    return (int64_t)floorq(x + 0.5Q);
}

void orc_encode_coeffs(const orc_ctx *c, const double *values, int nrot, int64_t *out) {
    int n = c->slots;
    __float128 *re = (__float128 *)calloc(n, sizeof(__float128));
    __float128 *im = (__float128 *)calloc(n, sizeof(__float128));
    /* gwas/matmult.go:666-672 convertToComplex128WithRot: res[(i+nrot) mod n] = v[i] */
    for (int i = 0; i < n; i++) re[imod(i + nrot, n)] = (__float128)values[i];
    special_invfft_q(c, re, im);
    __float128 sc = (__float128)c->scale;
    for (int k = 0; k < n; k++) {
        out[k] = round_away_q(re[k] * sc);
        out[k + n] = round_away_q(im[k] * sc);
    }
    free(re); free(im);
}

void orc_encode_ntt(const orc_ctx *c, const double *values, int nrot, int level, uint64_t *out) {
    int N = c->N;
    int64_t *m = (int64_t *)malloc(sizeof(int64_t) * N);
    orc_encode_coeffs(c, values, nrot, m);
    for (int l = 0; l <= level; l++) {
        uint64_t q = c->mod[l];
        uint64_t *o = out + (size_t)l * N;
        for (int k = 0; k < N; k++) {
            int64_t v = m[k];
            o[k] = v >= 0 ? (uint64_t)v % q : (q - ((uint64_t)(-v) % q)) % q;
        }
        orc_ntt(c, l, o);
    }
    free(m);
}

/* ------------------------------------------------------------------------------------------ */
/* test-only CKKS key material (keys are INPUTS of the path in production)                     */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint64_t s; } rng_t;
static uint64_t rng_next(rng_t *r) { /* splitmix64 */
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static uint64_t rng_uniform(rng_t *r, uint64_t q) {
    uint64_t lim = UINT64_MAX - (UINT64_MAX % q);
    uint64_t x;
    do x = rng_next(r); while (x >= lim);
    return x % q;
}
static double rng_gauss(rng_t *r, double sigma) {
    double u1 = ((rng_next(r) >> 11) + 1.0) / 9007199254740993.0, u2 = (rng_next(r) >> 11) / 9007199254740992.0;
    return sigma * sqrt(-2.0 * log(u1)) * cos(2.0 * M_PI * u2);
}
/* small error polynomial, NTT domain over limbs [0,nl) given by index list */
static void sample_error_ntt(const orc_ctx *c, rng_t *r, const int *limbs, int nl, uint64_t *e) {
    int N = c->N;
    int64_t *ev = (int64_t *)malloc(sizeof(int64_t) * N);
    for (int j = 0; j < N; j++) { double g = rng_gauss(r, 3.2); if (g > 19) g = 19; if (g < -19) g = -19; ev[j] = (int64_t)llround(g); }
    for (int k = 0; k < nl; k++) {
        uint64_t q = c->mod[limbs[k]];
        uint64_t *o = e + (size_t)k * N;
        for (int j = 0; j < N; j++) o[j] = ev[j] >= 0 ? (uint64_t)ev[j] : q - (uint64_t)(-ev[j]);
        orc_ntt(c, limbs[k], o);
    }
    free(ev);
}

void orc_keygen_secret(const orc_ctx *c, uint64_t seed, uint64_t *sk) {
    rng_t r = { seed };
    int N = c->N;
    int8_t *t = (int8_t *)malloc(N);
    for (int j = 0; j < N; j++) t[j] = (int8_t)(rng_next(&r) % 3) - 1;
    for (int i = 0; i < c->nQP; i++) {
        uint64_t q = c->mod[i];
        uint64_t *o = sk + (size_t)i * N;
        for (int j = 0; j < N; j++) o[j] = t[j] == 0 ? 0 : (t[j] > 0 ? 1 : q - 1);
        orc_ntt(c, i, o);
    }
    free(t);
}

int orc_beta(const orc_ctx *c) { return (c->nQ + c->nP - 1) / c->nP; }

/* Lattigo keygen.newSwitchingKey(skIn, skOut): swk[i] = (-a_i*skOut + e_i + P*skIn on the limbs of digit i, a_i),
 * NTT domain, then MForm. Layout [beta][2][nQ+nP][N]. */
void orc_gen_switching_key(const orc_ctx *c, const uint64_t *skIn, const uint64_t *skOut, uint64_t seed, uint64_t *swk) {
    rng_t r = { seed ^ 0xA5A5A5A5DEADBEEFULL };
    int N = c->N, nQP = c->nQP, alpha = c->nP, beta = orc_beta(c);
    int limbs[ORC_MAXMOD];
    for (int i = 0; i < nQP; i++) limbs[i] = i;
    uint64_t *e = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)nQP * N);
    for (int i = 0; i < beta; i++) {
        uint64_t *k0 = swk + ((size_t)i * 2 + 0) * nQP * N;
        uint64_t *k1 = swk + ((size_t)i * 2 + 1) * nQP * N;
        sample_error_ntt(c, &r, limbs, nQP, e);
        for (int l = 0; l < nQP; l++) {
            uint64_t q = c->mod[l];
            /* P mod q_l (0 on P limbs) */
            uint64_t Pmod = 1;
            for (int p = 0; p < c->nP; p++) Pmod = mulmod(Pmod, c->mod[c->nQ + p] % q, q);
            int inDigit = (l >= i * alpha && l < (i + 1) * alpha && l < c->nQ);
            for (int j = 0; j < N; j++) {
                size_t o = (size_t)l * N + j;
                uint64_t a = rng_uniform(&r, q);
                uint64_t v = submod(e[o], mulmod(a, skOut[o], q), q);
                if (inDigit) v = addmod(v, mulmod(Pmod, skIn[o], q), q);
                k0[o] = orc_mform(v, q, c->bred[l]);
                k1[o] = orc_mform(a, q, c->bred[l]);
            }
        }
    }
    free(e);
}

/* Lattigo Parameters.GaloisElementForColumnRotationBy(k): 5^(k mod 2N) mod 2N */
uint64_t orc_galois_element(const orc_ctx *c, int k) {
    uint64_t twoN = (uint64_t)c->N << 1;
    uint64_t e = (uint64_t)(k & (int)(twoN - 1));
    uint64_t r = 1, b = 5;
    while (e) { if (e & 1) r = (r * b) % twoN; b = (b * b) % twoN; e >>= 1; }
    return r;
}

/* Lattigo ring.PermuteNTTIndex (App. B.4) */
void orc_permute_ntt_index(int logN, uint64_t galEl, uint32_t *index) {
    uint64_t N = 1ULL << logN, mask = (N << 1) - 1;
    for (uint64_t i = 0; i < N; i++) {
        uint64_t tmp1 = 2 * bitrev(i, logN) + 1;
        uint64_t tmp2 = (((galEl * tmp1) & mask) - 1) >> 1;
        index[i] = (uint32_t)bitrev(tmp2, logN);
    }
}

/* Lattigo keygen.genrotKey: skOut = permute(sk, galEl^-1); swk = newSwitchingKey(sk, skOut) */
void orc_gen_rotation_key(const orc_ctx *c, const uint64_t *sk, uint64_t galEl, uint64_t seed, uint64_t *swk) {
    int N = c->N;
    uint64_t twoN = (uint64_t)N << 1;
    /* inverse of galEl mod 2N */
    uint64_t inv = 1;
    { uint64_t b = galEl % twoN, e = twoN / 2 - 1; /* group exponent of (Z/2N)^* divides N => g^(N-1) = g^-1 */
      e = (uint64_t)N - 1; while (e) { if (e & 1) inv = (inv * b) % twoN; b = (b * b) % twoN; e >>= 1; } }
    uint32_t *index = (uint32_t *)malloc(sizeof(uint32_t) * N);
    orc_permute_ntt_index(c->logN, inv, index);
    uint64_t *skOut = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)c->nQP * N);
    for (int l = 0; l < c->nQP; l++)
        for (int j = 0; j < N; j++) skOut[(size_t)l * N + j] = sk[(size_t)l * N + index[j]];
    orc_gen_switching_key(c, sk, skOut, seed ^ (galEl * 0x9E3779B97F4A7C15ULL), swk);
    free(index); free(skOut);
}

void orc_encrypt_sk(const orc_ctx *c, const uint64_t *sk, const uint64_t *pt, int level, uint64_t seed, uint64_t *ct) {
    rng_t r = { seed ^ 0x1234567811223344ULL };
    int N = c->N, nl = level + 1;
    int limbs[ORC_MAXMOD] = { 0 };
    for (int i = 0; i < nl; i++) limbs[i] = i;
    uint64_t *e = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)nl * N);
    sample_error_ntt(c, &r, limbs, nl, e);
    uint64_t *c0 = ct, *c1 = ct + (size_t)nl * N;
    for (int l = 0; l < nl; l++) {
        uint64_t q = c->mod[l];
        for (int j = 0; j < N; j++) {
            size_t o = (size_t)l * N + j;
            uint64_t a = rng_uniform(&r, q);
            c1[o] = a;
            c0[o] = addmod(submod(e[o], mulmod(a, sk[o], q), q), pt ? pt[o] : 0, q);
        }
    }
    free(e);
}

void orc_decrypt_coeffs(const orc_ctx *c, const uint64_t *sk, const uint64_t *ct, int level, uint64_t *out) {
    int N = c->N, nl = level + 1;
    const uint64_t *c0 = ct, *c1 = ct + (size_t)nl * N;
    for (int l = 0; l < nl; l++) {
        uint64_t q = c->mod[l];
        for (int j = 0; j < N; j++) {
            size_t o = (size_t)l * N + j;
            out[o] = addmod(c0[o], mulmod(c1[o], sk[o], q), q);
        }
        orc_intt(c, l, out + (size_t)l * N);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* key-switch: Lattigo v2.1 evaluator.switchKeysInPlace (App. B.5) [UNVERIFIED vs fork]        */
/* ------------------------------------------------------------------------------------------ */
/* fast exact base conversion of Lattigo (ring.modUpExact / Decomposer reconstructRNS+multSum):
 * src residues x_k mod s_k (k < ns), coefficient-wise; y_k = x_k*(S/s_k)^-1 mod s_k;
 * v = (uint64) sum_k float64(y_k)/float64(s_k)  (float64, index order);
 * out_t = sum_k y_k*(S/s_k mod t) - v*S mod t.  */
static void base_convert(const orc_ctx *c, const uint64_t *const *src, const int *sidx, int ns, int N, int tidx, uint64_t *dst) {
    /* constants in Montgomery form so that every product is one MRed (as Lattigo's ring arithmetic does) instead of a 128/64 division;
     * the float64 quotient estimate is evaluated exactly as before (division and sum in index order) */
    uint64_t tmod = c->mod[tidx], tInv = c->mred[tidx];
    uint64_t smod[ORC_MAXMOD], sOverSkInvM[ORC_MAXMOD], sOverSkModTM[ORC_MAXMOD], SmodT = 1;
    for (int k = 0; k < ns; k++) smod[k] = c->mod[sidx[k]];
    for (int k = 0; k < ns; k++) {
        uint64_t prod = 1, prodT = 1;
        for (int j = 0; j < ns; j++) if (j != k) { prod = mulmod(prod, smod[j] % smod[k], smod[k]); prodT = mulmod(prodT, smod[j] % tmod, tmod); }
        sOverSkInvM[k] = orc_mform(invmod(prod, smod[k]), smod[k], c->bred[sidx[k]]);
        sOverSkModTM[k] = orc_mform(prodT, tmod, c->bred[tidx]);
        SmodT = mulmod(SmodT, smod[k] % tmod, tmod);
    }
    uint64_t SmodTM = orc_mform(SmodT, tmod, c->bred[tidx]);
    for (int x = 0; x < N; x++) {
        double vi = 0.0;
        uint64_t acc = 0;
        for (int k = 0; k < ns; k++) {
            uint64_t y = orc_mred(src[k][x], sOverSkInvM[k], smod[k], c->mred[sidx[k]]); /* src is canonical mod s_k */
            vi += (double)y / (double)smod[k];
            acc = addmod(acc, orc_mred(orc_bred_add(y, tmod, c->bred[tidx]), sOverSkModTM[k], tmod, tInv), tmod);
        }
        uint64_t v = (uint64_t)vi;
        dst[x] = submod(acc, orc_mred(orc_bred_add(v, tmod, c->bred[tidx]), SmodTM, tmod, tInv), tmod);
    }
}

void orc_keyswitch(const orc_ctx *c, int level, const uint64_t *c1, const uint64_t *swk, uint64_t *out0, uint64_t *out1) {
    int N = c->N, nQ = c->nQ, nP = c->nP, nQP = c->nQP, alpha = nP;
    int nl = level + 1;
    int beta = (nl + alpha - 1) / alpha; /* ceil((level+1)/alpha) */
    size_t PN = (size_t)N;
    uint64_t *c2 = (uint64_t *)malloc(sizeof(uint64_t) * nl * PN);      /* INTT(c1) */
    uint64_t *d = (uint64_t *)malloc(sizeof(uint64_t) * PN);
    uint64_t *acc0 = (uint64_t *)calloc((size_t)(nl + nP) * PN, sizeof(uint64_t));
    uint64_t *acc1 = (uint64_t *)calloc((size_t)(nl + nP) * PN, sizeof(uint64_t));
    memcpy(c2, c1, sizeof(uint64_t) * nl * PN);
    for (int l = 0; l < nl; l++) orc_intt(c, l, c2 + l * PN);

    for (int i = 0; i < beta; i++) {
        int st = i * alpha, ed = st + alpha; if (ed > nl) ed = nl;
        int cnt = ed - st;
        const uint64_t *src[ORC_MAXMOD]; int sidx[ORC_MAXMOD];
        for (int k = 0; k < cnt; k++) { src[k] = c2 + (size_t)(st + k) * PN; sidx[k] = st + k; }
        const uint64_t *k0 = swk + ((size_t)i * 2 + 0) * nQP * PN;
        const uint64_t *k1 = swk + ((size_t)i * 2 + 1) * nQP * PN;
        for (int t = 0; t < nl + nP; t++) {
            int midx = t < nl ? t : nQ + (t - nl);  /* modulus index in QP */
            uint64_t q = c->mod[midx], qInv = c->mred[midx];
            const uint64_t *dn;
            if (t >= st && t < ed) {
                dn = c1 + (size_t)t * PN; /* decomposeAndSplitNTT: reuse the NTT-domain input limb */
            } else {
                if (cnt == 1) { /* Decomposer.DecomposeAndSplit single-modulus path: BRedAdd of the representative */
                    for (int x = 0; x < N; x++) d[x] = orc_bred_add(src[0][x], q, c->bred[midx]);
                } else {
                    base_convert(c, src, sidx, cnt, N, midx, d);
                }
                orc_ntt(c, midx, d);
                dn = d;
            }
            const uint64_t *kk0 = k0 + (size_t)midx * PN, *kk1 = k1 + (size_t)midx * PN;
            uint64_t *a0 = acc0 + (size_t)t * PN, *a1 = acc1 + (size_t)t * PN;
            for (int x = 0; x < N; x++) {
                a0[x] = addmod(a0[x], orc_mred(dn[x], kk0[x], q, qInv), q);
                a1[x] = addmod(a1[x], orc_mred(dn[x], kk1[x], q, qInv), q);
            }
        }
    }
    /* ModDownSplitNTTPQ: out = (accQ - ext_{P->Q}(INTT(accP))) * P^-1 mod q */
    for (int comp = 0; comp < 2; comp++) {
        uint64_t *acc = comp ? acc1 : acc0, *out = comp ? out1 : out0;
        const uint64_t *psrc[ORC_MAXMOD]; int pidx[ORC_MAXMOD];
        for (int p = 0; p < nP; p++) { orc_intt(c, nQ + p, acc + (size_t)(nl + p) * PN); psrc[p] = acc + (size_t)(nl + p) * PN; pidx[p] = nQ + p; }
        for (int l = 0; l < nl; l++) {
            uint64_t q = c->mod[l];
            base_convert(c, psrc, pidx, nP, N, l, d);
            orc_ntt(c, l, d);
            uint64_t Pmod = 1;
            for (int p = 0; p < nP; p++) Pmod = mulmod(Pmod, c->mod[nQ + p] % q, q);
            uint64_t PinvM = orc_mform(invmod(Pmod, q), q, c->bred[l]);
            for (int x = 0; x < N; x++) out[(size_t)l * PN + x] = orc_mred(submod(acc[(size_t)l * PN + x], d[x], q), PinvM, q, c->mred[l]);
        }
    }
    free(c2); free(d); free(acc0); free(acc1);
}

/* crypto/basics.go:201-210 RotateRightWithEvaluator -> Lattigo RotateNew(ct, slots - nrot) -> permuteNTT:
 * key-switch c1, add c0, then permute both with the NTT index of galEl. */
void orc_rotate_right(const orc_ctx *c, int level, const uint64_t *ct, int nrot, const uint64_t *swk, uint64_t *ctOut) {
    int N = c->N, nl = level + 1;
    size_t PS = (size_t)nl * N;
    nrot = imod(nrot, c->slots);
    if (nrot == 0) { memcpy(ctOut, ct, sizeof(uint64_t) * 2 * PS); return; }
    uint64_t galEl = orc_galois_element(c, c->slots - nrot);
    uint64_t *t0 = (uint64_t *)malloc(sizeof(uint64_t) * PS), *t1 = (uint64_t *)malloc(sizeof(uint64_t) * PS);
    orc_keyswitch(c, level, ct + PS, swk, t0, t1);
    for (int l = 0; l < nl; l++)
        for (int j = 0; j < N; j++) t0[(size_t)l * N + j] = addmod(t0[(size_t)l * N + j], ct[(size_t)l * N + j], c->mod[l]);
    uint32_t *index = (uint32_t *)malloc(sizeof(uint32_t) * N);
    orc_permute_ntt_index(c->logN, galEl, index);
    for (int l = 0; l < nl; l++)
        for (int j = 0; j < N; j++) {
            ctOut[(size_t)l * N + j] = t0[(size_t)l * N + index[j]];
            ctOut[PS + (size_t)l * N + j] = t1[(size_t)l * N + index[j]];
        }
    free(t0); free(t1); free(index);
}

/* ------------------------------------------------------------------------------------------ */
/* the hot path: gwas/matmult.go:914-1505                                                     */
/* ------------------------------------------------------------------------------------------ */
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
static int isqrt_ceil(int slots) { int d = (int)ceil(sqrt((double)slots)); return d; } /* matmult.go:918,1047 */

struct orc_diag_cache {
    int slots, d, m_ct, numBlockRows, nlimbs, N;
    uint8_t *babyTable, *giantTable, *shiftTable; /* [numBlockRows][d], [numBlockRows][d], [numBlockRows][slots] */
    uint64_t **pt;                                /* [numBlockRows][slots][m_ct] -> [nlimbs][N] or NULL */
};

typedef struct {
    const orc_ctx *c; const int8_t *X; size_t nrows, ncols; int maxLevel; int bi; int nproc, tid;
    orc_diag_cache *dc; int square_unused;
} prep_job;

/* one encoder goroutine of MatMult4StreamPreprocess (matmult.go:1013-1034): shifts with shift % nproc == tid */
static void encode_shift(const orc_ctx *c, const int8_t *X, size_t nrows, size_t ncols, int bi, int shift, int maxLevel,
                         uint64_t **slot /* [m_ct] */, double *buf) {
    int slots = c->slots, d = isqrt_ceil(slots), N = c->N;
    int m_ct = (int)((ncols - 1) / slots) + 1;
    int nr = (size_t)(bi + 1) * slots < nrows ? slots : (int)(nrows - (size_t)bi * slots);
    int giant = shift / d;
    for (int bj = 0; bj < m_ct; bj++) {
        int j1 = bj * slots, j2 = (size_t)(bj + 1) * slots < ncols ? (bj + 1) * slots : (int)ncols;
        int nc = j2 - j1;
        const int8_t *blk = X + (size_t)bi * slots * ncols + j1;
        /* EncodeDiagWithEncoder(blockVec, -shift, d*giant, maxLevel, enc) matmult.go:711-731,1024 */
        if (orc_get_diag(buf, blk, ncols, nr, nc, slots, -shift)) {
            uint64_t *p = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(maxLevel + 1) * N);
            orc_encode_ntt(c, buf, d * giant, maxLevel, p);
            orc_mform_lvl(c, maxLevel, p); /* ToMontgomeryForm matmult.go:1026 */
            slot[bj] = p;
        } else {
            slot[bj] = NULL;
        }
    }
}

static void *prep_worker(void *arg) {
    prep_job *j = (prep_job *)arg;
    orc_diag_cache *dc = j->dc;
    double *buf = (double *)malloc(sizeof(double) * dc->slots);
    for (int shift = 0; shift < dc->slots; shift++) {
        if (!dc->shiftTable[(size_t)j->bi * dc->slots + shift]) continue;
        if (shift % j->nproc != j->tid) continue; /* jobChannels[shift%nproc] matmult.go:986 */
        encode_shift(j->c, j->X, j->nrows, j->ncols, j->bi, shift, j->maxLevel,
                     dc->pt + ((size_t)j->bi * dc->slots + shift) * dc->m_ct, buf);
    }
    free(buf);
    return NULL;
}

static void active_tables(const orc_ctx *c, size_t nrows, size_t ncols, int bi, uint8_t *baby, uint8_t *giant, uint8_t *shiftT,
                          int shift_lo, int shift_hi) {
    int slots = c->slots, d = isqrt_ceil(slots);
    int m_ct = (int)((ncols - 1) / slots) + 1;
    int nr = (size_t)(bi + 1) * slots < nrows ? slots : (int)(nrows - (size_t)bi * slots);
    memset(baby, 0, d); memset(giant, 0, d); memset(shiftT, 0, slots);
    for (int shift = 0; shift < slots; shift++) {
        if (shift < shift_lo || shift >= shift_hi) continue;
        int any = 0;
        for (int bj = 0; bj < m_ct && !any; bj++) { /* EncodeDiagBool matmult.go:675-682 */
            int j1 = bj * slots, j2 = (size_t)(bj + 1) * slots < ncols ? (bj + 1) * slots : (int)ncols;
            any = orc_get_diag_bool(nr, j2 - j1, slots, -shift);
        }
        if (any) { baby[shift % d] = 1; giant[shift / d] = 1; shiftT[shift] = 1; } /* matmult.go:962-974 */
    }
}

orc_diag_cache *orc_matmult4_stream_preprocess(const orc_ctx *c, const int8_t *X, size_t nrows, size_t ncols, int maxLevel,
                                               int nproc, int shift_lo, int shift_hi) {
    int slots = c->slots, d = isqrt_ceil(slots);
    orc_diag_cache *dc = (orc_diag_cache *)calloc(1, sizeof(*dc));
    dc->slots = slots; dc->d = d; dc->N = c->N; dc->nlimbs = maxLevel + 1;
    dc->m_ct = (int)((ncols - 1) / slots) + 1;          /* matmult.go:920 */
    dc->numBlockRows = (int)((nrows - 1) / slots) + 1;  /* matmult.go:921 */
    dc->babyTable = (uint8_t *)calloc((size_t)dc->numBlockRows * d, 1);
    dc->giantTable = (uint8_t *)calloc((size_t)dc->numBlockRows * d, 1);
    dc->shiftTable = (uint8_t *)calloc((size_t)dc->numBlockRows * slots, 1);
    dc->pt = (uint64_t **)calloc((size_t)dc->numBlockRows * slots * dc->m_ct, sizeof(uint64_t *));
    if (nproc < 1) nproc = 1;
    if (shift_hi <= 0 || shift_hi > slots) shift_hi = slots;
    for (int bi = 0; bi < dc->numBlockRows; bi++) {
        active_tables(c, nrows, ncols, bi, dc->babyTable + (size_t)bi * d, dc->giantTable + (size_t)bi * d,
                      dc->shiftTable + (size_t)bi * slots, shift_lo, shift_hi);
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nproc);
        prep_job *jobs = (prep_job *)malloc(sizeof(prep_job) * nproc);
        for (int t = 0; t < nproc; t++) {
            jobs[t] = (prep_job){ c, X, nrows, ncols, maxLevel, bi, nproc, t, dc, 0 };
            pthread_create(&th[t], NULL, prep_worker, &jobs[t]);
        }
        for (int t = 0; t < nproc; t++) pthread_join(th[t], NULL);
        free(th); free(jobs);
    }
    return dc;
}
void orc_diag_cache_free(orc_diag_cache *dc) {
    if (!dc) return;
    size_t n = (size_t)dc->numBlockRows * dc->slots * dc->m_ct;
    for (size_t i = 0; i < n; i++) free(dc->pt[i]);
    free(dc->pt); free(dc->babyTable); free(dc->giantTable); free(dc->shiftTable); free(dc);
}
size_t orc_diag_cache_num_polys(const orc_diag_cache *dc) {
    size_t n = (size_t)dc->numBlockRows * dc->slots * dc->m_ct, k = 0;
    for (size_t i = 0; i < n; i++) k += dc->pt[i] != NULL;
    return k;
}
int orc_diag_cache_mct(const orc_diag_cache *dc) { return dc->m_ct; }
const uint64_t *orc_diag_cache_get(const orc_diag_cache *dc, int bi, int shift, int bj) {
    return dc->pt[((size_t)bi * dc->slots + shift) * dc->m_ct + bj];
}

/* gwas/filestream.go:140-233 WriteDiag + header (App. D.2). Coefficients big-endian u64 (ring.WriteCoeffsTo, App. B.8). */
int orc_diag_cache_write_files(const orc_ctx *c, const orc_diag_cache *dc, const char *prefix) {
    for (int bi = 0; bi < dc->numBlockRows; bi++) {
        char fn[4096];
        snprintf(fn, sizeof fn, "%s_%d.bin", prefix, bi);
        FILE *f = fopen(fn, "wb");
        if (!f) return -1;
        uint64_t dataLen = (uint64_t)dc->N * dc->nlimbs * 8;
        uint64_t hdr[6] = { (uint64_t)dc->m_ct, (uint64_t)(dc->nlimbs - 1), 0, (uint64_t)dc->N, (uint64_t)dc->nlimbs,
                            4 + (1 + dataLen) * (uint64_t)dc->m_ct };
        memcpy(&hdr[2], &c->scale, 8);
        fwrite(hdr, 8, 6, f);
        fwrite(dc->babyTable + (size_t)bi * dc->d, 1, dc->d, f);
        fwrite(dc->giantTable + (size_t)bi * dc->d, 1, dc->d, f);
        uint8_t *buf = (uint8_t *)malloc(hdr[5]);
        for (int shift = 0; shift < dc->slots; shift++) {
            if (!dc->shiftTable[(size_t)bi * dc->slots + shift]) continue;
            uint64_t ptr = 4;
            uint32_t sh = (uint32_t)shift;
            memcpy(buf, &sh, 4);
            for (int bj = 0; bj < dc->m_ct; bj++) {
                const uint64_t *p = orc_diag_cache_get(dc, bi, shift, bj);
                buf[ptr++] = p == NULL;
                if (p) for (size_t k = 0; k < (size_t)dc->N * dc->nlimbs; k++) {
                    uint64_t v = p[k];
                    for (int b = 0; b < 8; b++) buf[ptr++] = (uint8_t)(v >> (56 - 8 * b));
                }
            }
            fwrite(&ptr, 8, 1, f);
            fwrite(buf, 1, ptr, f);
        }
        free(buf);
        fclose(f);
    }
    return 0;
}

/* accumulators: CipherVectorAccV2 (matmult.go:208-245) -- `level` is a limb COUNT (App. A.4) */
typedef struct { orc_u128 *acc0, *acc1; } acc_ct; /* each [nlimbAcc][N] */

typedef struct {
    const orc_ctx *c; const orc_diag_cache *dc; int bi; int nproc, tid; int s; int nlimbAcc; int levelA;
    uint64_t **rot; /* [s][d] -> ct [2][levelA'+1][N] (already dropped to maxLevel) */
    acc_ct **acc;   /* [s][d] -> array of m_ct acc_ct, or NULL */
    pthread_mutex_t *mux; /* [s][d] */
    int nlA; /* limbs of rot cts */
} mac_job;

/* CPMultAccWithoutMRedV2({rot}, plainVec, acc) matmult.go:380-399 for one (i, shift) */
static void cpmult_acc(const orc_ctx *c, const uint64_t *rot, int nlA, uint64_t *const *plainVec, int m_ct, int nlimbPt,
                       acc_ct *acc, int nlimbAcc) {
    size_t N = c->N;
    (void)nlimbPt;
    for (int n = 0; n < m_ct; n++) {
        if (!plainVec[n]) continue;
        for (int l = 0; l < nlimbAcc; l++) {
            orc_mul_coeffs_and_add128(rot + (size_t)l * N, plainVec[n] + (size_t)l * N, acc[n].acc0 + (size_t)l * N, N);
            orc_mul_coeffs_and_add128(rot + ((size_t)nlA + l) * N, plainVec[n] + (size_t)l * N, acc[n].acc1 + (size_t)l * N, N);
        }
    }
}

static void *mac_worker(void *arg) { /* data processors matmult.go:1154-1168 */
    mac_job *j = (mac_job *)arg;
    const orc_diag_cache *dc = j->dc;
    int d = dc->d;
    for (int shift = 0; shift < dc->slots; shift++) {
        if (!dc->shiftTable[(size_t)j->bi * dc->slots + shift]) continue;
        if (shift % j->nproc != j->tid) continue; /* diagChannels[shift%nproc] matmult.go:1147 */
        int baby = shift % d, giant = shift / d;
        uint64_t *const *pv = dc->pt + ((size_t)j->bi * dc->slots + shift) * dc->m_ct;
        for (int i = 0; i < j->s; i++) {
            pthread_mutex_lock(&j->mux[i * d + giant]);
            cpmult_acc(j->c, j->rot[i * d + baby], j->nlA, pv, dc->m_ct, dc->nlimbs, j->acc[i * d + giant], j->nlimbAcc);
            pthread_mutex_unlock(&j->mux[i * d + giant]);
        }
    }
    return NULL;
}

typedef struct {
    const orc_ctx *c; int nproc, tid; int s, d; int levelIn; const uint64_t *A; int numBlockRows, bi; int nlA_in;
    const uint8_t *babyTable; const uint64_t *const *swk; uint64_t **rot; int maxLevel;
} rot_job;

static void *rot_worker(void *arg) { /* rotation cache workers matmult.go:1098-1119 */
    rot_job *j = (rot_job *)arg;
    const orc_ctx *c = j->c;
    size_t N = c->N;
    int nl = j->maxLevel + 1;
    for (int baby = 0; baby < j->d; baby++) {
        if (!j->babyTable[baby] || baby % j->nproc != j->tid) continue;
        for (int i = 0; i < j->s; i++) {
            /* A[i][bi] dropped to maxLevel: DropLevelNew truncates limbs (crypto/basics.go:806-824) */
            const uint64_t *ctin = j->A + ((size_t)i * j->numBlockRows + j->bi) * 2 * j->nlA_in * N;
            uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * 2 * nl * N);
            memcpy(tmp, ctin, sizeof(uint64_t) * nl * N);
            memcpy(tmp + nl * N, ctin + (size_t)j->nlA_in * N, sizeof(uint64_t) * nl * N);
            uint64_t *o = (uint64_t *)malloc(sizeof(uint64_t) * 2 * nl * N);
            const uint64_t *key = baby ? j->swk[baby] : NULL;
            orc_rotate_right(c, j->maxLevel, tmp, -baby, key, o); /* RotateRightWithEvaluator(.., -baby) matmult.go:1114 */
            free(tmp);
            j->rot[i * j->d + baby] = o;
        }
    }
    return NULL;
}

typedef struct {
    const orc_ctx *c; int nproc, tid; int i, d, m_ct, nlimbAcc; acc_ct **acc; const uint64_t *const *swk;
    uint64_t *S_i; pthread_mutex_t *outMux;
} post_job;

/* ModularReduceV2 (matmult.go:343-366) + giant-step rotation (:1203-1209) + Add into out (:1223-1227) */
static void *post_worker(void *arg) {
    post_job *j = (post_job *)arg;
    const orc_ctx *c = j->c;
    size_t N = c->N;
    int nl = j->nlimbAcc, level = nl - 1;
    uint64_t *ct = (uint64_t *)malloc(sizeof(uint64_t) * 2 * nl * N), *r = (uint64_t *)malloc(sizeof(uint64_t) * 2 * nl * N);
    for (int l = 0; l < j->d; l++) {
        if (!j->acc[j->i * j->d + l] || l % j->nproc != j->tid) continue;
        for (int n = 0; n < j->m_ct; n++) {
            memset(ct, 0, sizeof(uint64_t) * 2 * nl * N);
            for (int lv = 0; lv < nl; lv++) {
                orc_reduce_and_add_uint128(j->acc[j->i * j->d + l][n].acc0 + (size_t)lv * N, ct + (size_t)lv * N, c->mred[lv], c->mod[lv], N);
                orc_reduce_and_add_uint128(j->acc[j->i * j->d + l][n].acc1 + (size_t)lv * N, ct + ((size_t)nl + lv) * N, c->mred[lv], c->mod[lv], N);
            }
            orc_reduce_canonical(c, nl, ct);
            orc_reduce_canonical(c, nl, ct + (size_t)nl * N);
            const uint64_t *src = ct;
            if (l > 0) { orc_rotate_right(c, level, ct, -l * j->d, j->swk[(l * j->d) % c->slots], r); src = r; }
            uint64_t *o = j->S_i + (size_t)n * 2 * nl * N;
            pthread_mutex_lock(j->outMux);
            for (int comp = 0; comp < 2; comp++)
                for (int lv = 0; lv < nl; lv++)
                    for (size_t x = 0; x < N; x++) {
                        size_t off = ((size_t)comp * nl + lv) * N + x;
                        o[off] = addmod(o[off], src[off], c->mod[lv]);
                    }
            pthread_mutex_unlock(j->outMux);
        }
    }
    free(ct); free(r);
    return NULL;
}

static acc_ct *new_acc(int m_ct, int nlimbAcc, size_t N) { /* NewCipherVectorAccV2 matmult.go:231-245 */
    acc_ct *a = (acc_ct *)malloc(sizeof(acc_ct) * m_ct);
    for (int n = 0; n < m_ct; n++) {
        a[n].acc0 = (orc_u128 *)calloc((size_t)nlimbAcc * N, sizeof(orc_u128));
        a[n].acc1 = (orc_u128 *)calloc((size_t)nlimbAcc * N, sizeof(orc_u128));
    }
    return a;
}

static void run_threads(void *(*fn)(void *), void *jobs, size_t jobsz, int nproc) {
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nproc);
    for (int t = 0; t < nproc; t++) pthread_create(&th[t], NULL, fn, (char *)jobs + jobsz * t);
    for (int t = 0; t < nproc; t++) pthread_join(th[t], NULL);
    free(th);
}

void orc_matmult4_stream_compute(const orc_ctx *c, const uint64_t *A, int s, int numBlockRows, int levelA, int maxLevel,
                                 const orc_diag_cache *dc, const uint64_t *const *swk, int nproc, uint64_t *S,
                                 double *t_rot_baby, double *t_mac, double *t_post) {
    int d = dc->d, m_ct = dc->m_ct;
    size_t N = c->N;
    int nlimbAcc = maxLevel; /* the limb-COUNT quirk, matmult.go:1125 -> :231 (App. A.4) */
    int nlOut = nlimbAcc;    /* output level = nlimbAcc - 1 (matmult.go:350) */
    if (nproc < 1) nproc = 1;
    /* matmult.go:1053-1056: A is dropped only when Level() > maxLevel; an input at maxLevel-1 still has the nlimbAcc limbs the MAC
     * reads (:393) and is rotated at its own level; fewer limbs index past Coeffs[] in the reference (panic) */
    if (levelA < maxLevel - 1) { fprintf(stderr, "input level %d has fewer than %d limbs (index out of range)\n", levelA, maxLevel); abort(); }
    int lvRot = levelA < maxLevel ? levelA : maxLevel;
    double tr = 0, tm = 0, tp = 0;

    acc_ct **acc = (acc_ct **)calloc((size_t)s * d, sizeof(acc_ct *));
    pthread_mutex_t *mux = (pthread_mutex_t *)malloc(sizeof(pthread_mutex_t) * s * d);
    for (int k = 0; k < s * d; k++) pthread_mutex_init(&mux[k], NULL);
    uint64_t **rot = (uint64_t **)calloc((size_t)s * d, sizeof(uint64_t *));

    for (int bi = 0; bi < numBlockRows; bi++) {
        double t0 = now_s();
        rot_job *rj = (rot_job *)malloc(sizeof(rot_job) * nproc);
        for (int t = 0; t < nproc; t++)
            rj[t] = (rot_job){ c, nproc, t, s, d, levelA, A, numBlockRows, bi, levelA + 1, dc->babyTable + (size_t)bi * d, swk, rot, lvRot };
        run_threads(rot_worker, rj, sizeof(rot_job), nproc);
        free(rj);
        double t1 = now_s(); tr += t1 - t0;
        for (int g = 0; g < d; g++)
            if (dc->giantTable[(size_t)bi * d + g])
                for (int i = 0; i < s; i++)
                    if (!acc[i * d + g]) acc[i * d + g] = new_acc(m_ct, nlimbAcc, N);
        mac_job *mj = (mac_job *)malloc(sizeof(mac_job) * nproc);
        for (int t = 0; t < nproc; t++) mj[t] = (mac_job){ c, dc, bi, nproc, t, s, nlimbAcc, levelA, rot, acc, mux, lvRot + 1 };
        run_threads(mac_worker, mj, sizeof(mac_job), nproc);
        free(mj);
        tm += now_s() - t1;
        for (int k = 0; k < s * d; k++) { free(rot[k]); rot[k] = NULL; }
    }
    double t2 = now_s();
    /* post-processing. out starts as CZeroMat = fresh Enc(0) in the reference (matmult.go:1174): the deterministic
     * part S is accumulated from 0 here (App. A.5). */
    memset(S, 0, sizeof(uint64_t) * (size_t)s * m_ct * 2 * nlOut * N);
    for (int i = 0; i < s; i++) {
        pthread_mutex_t om; pthread_mutex_init(&om, NULL);
        post_job *pj = (post_job *)malloc(sizeof(post_job) * nproc);
        for (int t = 0; t < nproc; t++)
            pj[t] = (post_job){ c, nproc, t, i, d, m_ct, nlimbAcc, acc, swk, S + (size_t)i * m_ct * 2 * nlOut * N, &om };
        run_threads(post_worker, pj, sizeof(post_job), nproc);
        free(pj);
        pthread_mutex_destroy(&om);
    }
    tp = now_s() - t2;
    for (int k = 0; k < s * d; k++) {
        if (acc[k]) { for (int n = 0; n < m_ct; n++) { free(acc[k][n].acc0); free(acc[k][n].acc1); } free(acc[k]); }
        pthread_mutex_destroy(&mux[k]);
    }
    free(acc); free(mux); free(rot);
    if (t_rot_baby) *t_rot_baby = tr;
    if (t_mac) *t_mac = tm;
    if (t_post) *t_post = tp;
}

/* gwas/matmult.go:1238-1505: fused variant = same result as Preprocess+Compute on the (missing->0, optionally squared)
 * matrix, plus optional per-column sum / sqSum (computed BEFORE squaring, on the missing->0 value, :1292-1304). */
void orc_matmult4_stream(const orc_ctx *c, const uint64_t *A, int s, int levelA, const int8_t *X, size_t nrows, size_t ncols,
                         int maxLevel, int computeSquaredSum, int square, const uint64_t *const *swk, int nproc, uint64_t *S,
                         double *sum, double *sqSum) {
    int8_t *Y = (int8_t *)malloc(nrows * ncols);
    if (computeSquaredSum) { memset(sum, 0, sizeof(double) * ncols); memset(sqSum, 0, sizeof(double) * ncols); }
    for (size_t r = 0; r < nrows; r++)
        for (size_t j = 0; j < ncols; j++) {
            int8_t v = X[r * ncols + j];
            if (v < 0) v = 0;
            if (computeSquaredSum) { sqSum[j] += (double)(int8_t)(v * v); sum[j] += (double)v; }
            if (square) v = (int8_t)(v * v);
            Y[r * ncols + j] = v;
        }
    int numBlockRows = (int)((nrows - 1) / c->slots) + 1;
    orc_diag_cache *dc = orc_matmult4_stream_preprocess(c, Y, nrows, ncols, maxLevel, nproc, 0, c->slots);
    orc_matmult4_stream_compute(c, A, s, numBlockRows, levelA, maxLevel, dc, swk, nproc, S, NULL, NULL, NULL);
    orc_diag_cache_free(dc);
    free(Y);
}

/* ------------------------------------------------------------------------------------------ */
/* CPU baseline micro-benchmark of the K1 loop with the reference's locking pattern            */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int N, limbs, s, ndiag, nthreads, tid; uint64_t *rot, *pt; orc_u128 *acc; pthread_mutex_t *mux; } bm_job;
static void *bm_worker(void *arg) {
    bm_job *j = (bm_job *)arg;
    size_t N = j->N;
    for (int dgi = j->tid; dgi < j->ndiag; dgi += j->nthreads) {
        const uint64_t *pt = j->pt + (size_t)(dgi % 8) * j->limbs * N;
        for (int i = 0; i < j->s; i++) {
            pthread_mutex_lock(&j->mux[i]);
            for (int l = 0; l < j->limbs; l++) {
                orc_mul_coeffs_and_add128(j->rot + ((size_t)i * 2 * j->limbs + l) * N, pt + (size_t)l * N, j->acc + ((size_t)i * 2 * j->limbs + l) * N, N);
                orc_mul_coeffs_and_add128(j->rot + ((size_t)i * 2 * j->limbs + j->limbs + l) * N, pt + (size_t)l * N, j->acc + ((size_t)i * 2 * j->limbs + j->limbs + l) * N, N);
            }
            pthread_mutex_unlock(&j->mux[i]);
        }
    }
    return NULL;
}
double orc_bench_mac(int N, int limbs, int s, int ndiag, int nthreads) {
    uint64_t *rot = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)s * 2 * limbs * N);
    uint64_t *pt = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)8 * limbs * N);
    orc_u128 *acc = (orc_u128 *)calloc((size_t)s * 2 * limbs * N, sizeof(orc_u128));
    rng_t r = { 42 };
    for (size_t k = 0; k < (size_t)s * 2 * limbs * N; k++) rot[k] = rng_next(&r) >> 31;
    for (size_t k = 0; k < (size_t)8 * limbs * N; k++) pt[k] = rng_next(&r) >> 31;
    pthread_mutex_t *mux = (pthread_mutex_t *)malloc(sizeof(pthread_mutex_t) * s);
    for (int i = 0; i < s; i++) pthread_mutex_init(&mux[i], NULL);
    bm_job *jobs = (bm_job *)malloc(sizeof(bm_job) * nthreads);
    for (int t = 0; t < nthreads; t++) jobs[t] = (bm_job){ N, limbs, s, ndiag, nthreads, t, rot, pt, acc, mux };
    double t0 = now_s();
    run_threads(bm_worker, jobs, sizeof(bm_job), nthreads);
    double dt = now_s() - t0;
    free(jobs); free(mux); free(rot); free(pt); free(acc);
    return dt;
}

/* ------------------------------------------------------------------------------------------ */
/* ct x ct / ct x pt algebra of the callers around the path:            */
/* crypto.CMult / CMultScalar / MaskTrunc / InnerSumAll / Sub as used by QXLazyNormStream and    */
/* QXtLazyNormStream (gwas/matmult.go:27-116, crypto/basics.go:110-127,236-293,386-427,553-566). */
/* Lattigo v2.1 evaluator.mulRelin / Rescale / ring.DivRoundByLastModulusNTT, [UNVERIFIED vs the  */
/* fork]: every step is exact arithmetic mod q on canonical residues except the key-switch.       */
/* ------------------------------------------------------------------------------------------ */

/* evaluator.MulRelinNew(ctA, ctB) for two degree-1 ciphertexts at `level`, relinearised with rlk = swk(s^2 -> s):
 *   d0 = a0*b0, d1 = a0*b1 + a1*b0, d2 = a1*b1 ; (p0, p1) = switchKeys(d2, rlk) ; out = (d0 + p0, d1 + p1).
 * The reference multiplies MForm(a) by b with MRed, i.e. the plain modular product. */
void orc_mul_relin(const orc_ctx *c, int level, const uint64_t *ctA, const uint64_t *ctB, const uint64_t *rlk, uint64_t *out) {
    int N = c->N, nl = level + 1;
    size_t PS = (size_t)nl * N;
    uint64_t *d2 = (uint64_t *)malloc(sizeof(uint64_t) * PS), *p0 = (uint64_t *)malloc(sizeof(uint64_t) * PS), *p1 = (uint64_t *)malloc(sizeof(uint64_t) * PS);
    for (int l = 0; l < nl; l++) {
        uint64_t q = c->mod[l], qInv = c->mred[l];
        for (int j = 0; j < N; j++) {
            size_t o = (size_t)l * N + j;
            uint64_t a0 = orc_mform(ctA[o], q, c->bred[l]), a1 = orc_mform(ctA[PS + o], q, c->bred[l]);
            out[o] = orc_mred(a0, ctB[o], q, qInv);
            out[PS + o] = addmod(orc_mred(a0, ctB[PS + o], q, qInv), orc_mred(a1, ctB[o], q, qInv), q);
            d2[o] = orc_mred(a1, ctB[PS + o], q, qInv);
        }
    }
    orc_keyswitch(c, level, d2, rlk, p0, p1);
    for (int l = 0; l < nl; l++)
        for (int j = 0; j < N; j++) {
            size_t o = (size_t)l * N + j;
            out[o] = addmod(out[o], p0[o], c->mod[l]);
            out[PS + o] = addmod(out[PS + o], p1[o], c->mod[l]);
        }
    free(d2); free(p0); free(p1);
}

/* evaluator.MulRelinNew(plaintext, ct): both components multiplied by the NTT-domain plaintext (crypto/basics.go:121) */
void orc_mul_plain(const orc_ctx *c, int level, const uint64_t *pt, const uint64_t *ct, uint64_t *out) {
    int N = c->N, nl = level + 1;
    size_t PS = (size_t)nl * N;
    for (int comp = 0; comp < 2; comp++)
        for (int l = 0; l < nl; l++)
            for (int j = 0; j < N; j++) {
                size_t o = (size_t)l * N + j;
                out[comp * PS + o] = orc_mred(orc_mform(pt[o], c->mod[l], c->bred[l]), ct[comp * PS + o], c->mod[l], c->mred[l]);
            }
}

/* ring.DivRoundByLastModulusNTT on both components: ct [2][level+1][N] -> out [2][level][N]  (one step of evaluator.Rescale):
 *   t = InvNTT(x_L); t = (t + (qL-1)/2) mod qL; for l < L: z = NTT_l(t - (qL-1)/2 mod q_l); out_l = (x_l - z) * qL^-1 mod q_l */
void orc_rescale_once(const orc_ctx *c, int level, const uint64_t *ct, uint64_t *out) {
    int N = c->N, nl = level + 1;
    uint64_t qL = c->mod[level], half = (qL - 1) >> 1;
    uint64_t *t = (uint64_t *)malloc(sizeof(uint64_t) * N), *z = (uint64_t *)malloc(sizeof(uint64_t) * N);
    for (int comp = 0; comp < 2; comp++) {
        const uint64_t *x = ct + (size_t)comp * nl * N;
        uint64_t *o = out + (size_t)comp * level * N;
        memcpy(t, x + (size_t)level * N, sizeof(uint64_t) * N);
        orc_intt(c, level, t);
        for (int j = 0; j < N; j++) { t[j] += half; if (t[j] >= qL) t[j] -= qL; }
        for (int l = 0; l < level; l++) {
            uint64_t q = c->mod[l];
            uint64_t halfNeg = q - orc_bred_add(half, q, c->bred[l]);
            uint64_t inv = invmod(qL % q, q);
            for (int j = 0; j < N; j++) z[j] = orc_bred_add(t[j] + halfNeg, q, c->bred[l]);
            orc_ntt(c, l, z);
            for (int j = 0; j < N; j++) o[(size_t)l * N + j] = mulmod(submod(x[(size_t)l * N + j], z[j], q), inv, q);
        }
    }
    free(t); free(z);
}

/* evaluator.Add / Sub on equal-scale operands: limb-wise, nl limbs of each of the two components */
void orc_ct_addsub(const orc_ctx *c, int nl, const uint64_t *a, const uint64_t *b, int sub, uint64_t *out) {
    int N = c->N;
    for (int comp = 0; comp < 2; comp++)
        for (int l = 0; l < nl; l++)
            for (int j = 0; j < N; j++) {
                size_t o = ((size_t)comp * nl + l) * N + j;
                out[o] = sub ? submod(a[o], b[o], c->mod[l]) : addmod(a[o], b[o], c->mod[l]);
            }
}
