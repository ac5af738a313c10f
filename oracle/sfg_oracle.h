/*
 * sfg_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the genotype x CKKS-ciphertext MatMult hot path of
 * hhcho/sfgwas (gwas/matmult.go) and of the Lattigo-fork arithmetic it calls.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path
 * (sfgwas_b200/, include/sfgwas_b200.h) never links or calls it.
 *
 * PARITY STATUS: "parity unpinned" for everything that goes through Lattigo
 * internals (NTT root choice, key-switch, encoder): the fork
 * github.com/hcholab/lattigo/v2 v2.1.2-0.20230123224332-e8d68c24b94a is not
 * vendored in /root/reference, there is no Go toolchain here, and the reference
 * ships no tests / golden vectors.  Those parts restate the published Lattigo
 * v2.1 algorithms and are pinned only by math-defined
 * known-answer tests (tests/test_oracle_*.py).  The lazy-MAC primitives
 * (MulCoeffsAndAdd128, ReduceAndAddUint128, MForm) are fully specified by
 * gwas/matmult.go:247-324,433-440 and ARE pinned (closed-form bignum KATs).
 *
 * Every function cites the reference file:line it follows.
 */
#ifndef SFG_ORACLE_H
#define SFG_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAXMOD 48

typedef struct { uint64_t hi, lo; } orc_u128; /* gwas/matmult.go:196-199 (hi first, then lo) */

typedef struct orc_ctx orc_ctx;

/* ---- parameters / ring (Lattigo ring.NewRing + ckks.Parameters) ---- */
orc_ctx *orc_ctx_new(int logN, const uint64_t *qi, int nQ, const uint64_t *pi, int nP, double scale);
void orc_ctx_free(orc_ctx *c);
int orc_ctx_N(const orc_ctx *c);
int orc_ctx_nQ(const orc_ctx *c);
int orc_ctx_nP(const orc_ctx *c);
uint64_t orc_ctx_modulus(const orc_ctx *c, int idx);          /* Q then P */
uint64_t orc_ctx_mred(const orc_ctx *c, int idx);             /* ring.MredParams */
void orc_ctx_bred(const orc_ctx *c, int idx, uint64_t out[2]); /* ring.BredParams {hi,lo} */
uint64_t orc_ctx_psi(const orc_ctx *c, int idx);              /* 2N-th root psi (plain, not Montgomery) */
uint64_t orc_primitive_root(uint64_t q);                       /* Lattigo ring.primitiveRoot */

/* ---- scalar primitives ---- */
uint64_t orc_mred(uint64_t x, uint64_t y, uint64_t q, uint64_t qInv);   /* Lattigo ring.MRed */
uint64_t orc_mform(uint64_t a, uint64_t q, const uint64_t u[2]);        /* gwas/matmult.go:433-440 */
uint64_t orc_bred_add(uint64_t a, uint64_t q, const uint64_t u[2]);     /* Lattigo ring.BRedAdd */

/* ---- NTT (Lattigo ring.NTT / InvNTT, App. B.3): in-place on one limb ---- */
void orc_ntt(const orc_ctx *c, int idx, uint64_t *a);
void orc_intt(const orc_ctx *c, int idx, uint64_t *a);

/* ---- K1/K2/K3 exactly as written in the reference ---- */
void orc_mul_coeffs_and_add128(const uint64_t *a, const uint64_t *b, orc_u128 *c, size_t n); /* matmult.go:247-289 */
void orc_reduce_and_add_uint128(const orc_u128 *in, uint64_t *out, uint64_t qInv, uint64_t q, size_t n); /* :291-324 */
void orc_mform_lvl(const orc_ctx *c, int level, uint64_t *p); /* :411-431, poly [level+1][N] in place */
void orc_reduce_canonical(const orc_ctx *c, int nlimbs, uint64_t *p); /* fork eval.Reduce: x mod q per limb */

/* ---- diagonals (gwas/matmult.go:573-672) ---- */
int orc_get_diag_bool(int r, int cdim, int dim, int index);                      /* :627-631 */
int orc_get_diag(double *dst, const int8_t *X, size_t ld, int r, int cdim, int dim, int index); /* :636-664 */

/* ---- encoder: EncodeNTT of a real slot vector, right-rotated by nrot
 *      (convertToComplex128WithRot + EncoderBig.EncodeNTT, matmult.go:666-731, App. B.6).
 *      out: [level+1][N] NTT-domain residues (NOT Montgomery form). ---- */
void orc_encode_ntt(const orc_ctx *c, const double *values, int nrot, int level, uint64_t *out);
/* coefficient-domain integer message (before RNS/NTT), as int64: out[N] */
void orc_encode_coeffs(const orc_ctx *c, const double *values, int nrot, int64_t *out);

/* ---- test-only CKKS (keys are INPUTS of the path; generated here for tests) ---- */
void orc_keygen_secret(const orc_ctx *c, uint64_t seed, uint64_t *sk /* [nQ+nP][N] NTT, plain */);
/* switching key skIn -> skOut, layout [beta][2][nQ+nP][N], NTT + Montgomery form (Lattigo SwitchingKey) */
int  orc_beta(const orc_ctx *c);
void orc_gen_switching_key(const orc_ctx *c, const uint64_t *skIn, const uint64_t *skOut, uint64_t seed, uint64_t *swk);
/* rotation key for left-rotation by k: galEl = 5^k mod 2N, skOut = pi_{galEl^-1}(sk) */
uint64_t orc_galois_element(const orc_ctx *c, int k);
void orc_gen_rotation_key(const orc_ctx *c, const uint64_t *sk, uint64_t galEl, uint64_t seed, uint64_t *swk);
/* symmetric encryption of an NTT-domain plaintext pt[level+1][N] -> ct[2][level+1][N] */
void orc_encrypt_sk(const orc_ctx *c, const uint64_t *sk, const uint64_t *pt, int level, uint64_t seed, uint64_t *ct);
/* decrypt: m = c0 + c1*s, INTT -> coefficient domain residues [level+1][N] */
void orc_decrypt_coeffs(const orc_ctx *c, const uint64_t *sk, const uint64_t *ct, int level, uint64_t *out);

/* ---- rotation = key-switch + automorphism (Lattigo v2.1 permuteNTT / switchKeysInPlace, App. B.4-B.5) ---- */
void orc_permute_ntt_index(int logN, uint64_t galEl, uint32_t *index /* [N] */);
void orc_keyswitch(const orc_ctx *c, int level, const uint64_t *c1 /* [level+1][N] */, const uint64_t *swk,
                   uint64_t *out0, uint64_t *out1 /* [level+1][N] each */);
/* crypto.RotateRightWithEvaluator (crypto/basics.go:201-210): ctOut = RotR_nrot(ct) ; swk for galEl 5^(slots-nrot mod slots) */
void orc_rotate_right(const orc_ctx *c, int level, const uint64_t *ct, int nrot, const uint64_t *swk, uint64_t *ctOut);

/* ---- the hot path (gwas/matmult.go:914-1505) ----
 * Genotype matrix X: int8 row-major nrows x ncols (App. D.1).
 * A: [s][numBlockRows] ciphertexts at level `levelA` (>= maxLevel), each [2][levelA+1][N].
 * rotation keys: callback table swk_for_rot[k] for left rotations k (NULL if absent): array of
 *   `slots` pointers indexed by k.
 * Output S (deterministic part): [s][m_ct] ciphertexts at level maxLevel-1: [2][maxLevel][N].
 */
typedef struct orc_diag_cache orc_diag_cache;

orc_diag_cache *orc_matmult4_stream_preprocess(const orc_ctx *c, const int8_t *X, size_t nrows, size_t ncols,
                                               int maxLevel, int nproc,
                                               int shift_lo, int shift_hi /* sample window: [lo,hi) of shifts; 0,slots = all */);
void orc_diag_cache_free(orc_diag_cache *dc);
size_t orc_diag_cache_num_polys(const orc_diag_cache *dc);
int orc_diag_cache_mct(const orc_diag_cache *dc);
/* fetch one cached plaintext (NULL if nil) : [maxLevel+1][N], NTT + Montgomery form */
const uint64_t *orc_diag_cache_get(const orc_diag_cache *dc, int bi, int shift, int bj);
/* write / read the reference's on-disk format (gwas/filestream.go:42-282, App. D.2) */
int orc_diag_cache_write_files(const orc_ctx *c, const orc_diag_cache *dc, const char *prefix);

void orc_matmult4_stream_compute(const orc_ctx *c, const uint64_t *A, int s, int numBlockRows, int levelA,
                                 int maxLevel, const orc_diag_cache *dc, const uint64_t *const *swk_for_rot,
                                 int nproc, uint64_t *S /* [s][m_ct][2][maxLevel][N] */,
                                 double *t_rot_baby, double *t_mac, double *t_post);

void orc_matmult4_stream(const orc_ctx *c, const uint64_t *A, int s, int levelA, const int8_t *X, size_t nrows,
                         size_t ncols, int maxLevel, int computeSquaredSum, int square,
                         const uint64_t *const *swk_for_rot, int nproc, uint64_t *S, double *sum, double *sqSum);


/* ---- ct x ct / ct x pt algebra of the callers (QXLazyNormStream / QXtLazyNormStream, gwas/matmult.go:27-116;
 *      crypto.CMult / CMultScalar / MaskTrunc / InnerSumAll, crypto/basics.go) -- Lattigo v2.1 mulRelin, Rescale ---- */
void orc_mul_relin(const orc_ctx *c, int level, const uint64_t *ctA, const uint64_t *ctB, const uint64_t *rlk, uint64_t *out);
void orc_mul_plain(const orc_ctx *c, int level, const uint64_t *pt, const uint64_t *ct, uint64_t *out);
void orc_rescale_once(const orc_ctx *c, int level, const uint64_t *ct /* [2][level+1][N] */, uint64_t *out /* [2][level][N] */);
void orc_ct_addsub(const orc_ctx *c, int nl, const uint64_t *a, const uint64_t *b, int sub, uint64_t *out);

/* pure MAC micro-benchmark used by bench.py cpu_baseline: nthreads, returns seconds */
double orc_bench_mac(int N, int limbs, int s, int ndiag, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
