/*
 * sfgwas_b200.h -- C ABI of libsfgwas_b200.so: the B200 (sm_100a) implementation of the genotype x CKKS-ciphertext
 * MatMult hot path of hhcho/sfgwas.  Plain pointers and sizes only; no Go, torch or C++ types cross this boundary.
 *
 * The reference has no FFI for this path: the boundary is the set of exported Go functions of package `gwas`
 * (SURVEY.md 8b).  Each entry point below names the reference function whose body it replaces; the cgo shim that
 * binds them (go/gwas/matmult_b200.go) keeps the Go signatures verbatim.  See INTEGRATION.md.
 *
 * Conventions
 *  - Every function returns 0 on success, non-zero on failure; sfg_last_error(ctx) gives the message.  The Go shim
 *    panics on non-zero, matching the reference's panic/log.Fatal behaviour (gwas/matmult.go:360-362).
 *  - Polynomials are uint64 residues, limb-major: a ciphertext of L limbs is [2][L][N], a plaintext [L][N], exactly the
 *    order of Lattigo's Poly.Coeffs[l][j] (gwas/matmult.go:372-375,394-395).  "flat" entry points take one contiguous
 *    buffer; the "_ptrs" variants take one pointer per limb (cgo: one Go slice per limb).
 *  - Data pointers may be host (pageable or pinned) or device pointers (unified addressing, cudaMemcpyDefault).
 *  - Host buffers are never retained after a call returns.  Handles are owned by the library and must be destroyed.
 *  - The library is re-entrant across contexts; calls on ONE context are serialised by the caller or by distinct
 *    sfg_ctx handles (one per goroutine, like the reference's per-goroutine ckks.Evaluator, gwas/matmult.go:1110).
 *  - There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef SFGWAS_B200_H
#define SFGWAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sfg_ctx sfg_ctx;       /* ckks.Parameters + ring tables + Galois keys on one GPU (crypto.CryptoParams subset) */
typedef struct sfg_geno sfg_geno;     /* device-resident int8 genotype matrix (what a GenoFileStream yields, App. D.1) */
typedef struct sfg_cache sfg_cache;   /* device-resident diagonal cache (replaces the DiagCacheStream files, App. D.2) */

/* ---- context: replaces ring.NewRing / ckks.NewEvaluator setup done inside the path (gwas/matmult.go:328,345,403,1110) ---- */
/* qi/pi: params.Qi()/Pi(); scale: params.Scale(); psi: optional nQ+nP primitive 2N-th roots (NULL = Lattigo's rule). */
int sfg_ctx_create(int device, int logN, const uint64_t *qi, int nQ, const uint64_t *pi, int nP, double scale,
                   const uint64_t *psi, sfg_ctx **out);
void sfg_ctx_destroy(sfg_ctx *ctx);
const char *sfg_last_error(const sfg_ctx *ctx);
int sfg_version(void);
/* HBM budget for cached diagonals in bytes (0 = 70% of free memory at preprocess time). */
int sfg_ctx_set_cache_budget(sfg_ctx *ctx, size_t bytes);
/* counters: kernels launched so far by this context; encoder statistics {rechecked coefficients, unresolved ties}. */
unsigned long long sfg_ctx_launch_count(const sfg_ctx *ctx);
int sfg_ctx_encoder_stats(sfg_ctx *ctx, unsigned long long out[2]);
int sfg_ctx_psi(const sfg_ctx *ctx, uint64_t *psi_out /* nQ+nP */);

/* Galois (rotation) keys: cryptoParams.RotKs.Keys[galEl].Value[i][0|1].Coeffs[*] (crypto/crypto.go:232-275, gwas/matmult.go:1110).
 * key: [beta][2][nQ+nP][N] contiguous, NTT + Montgomery form as Lattigo stores it.  rot_left: the left-rotation amount k
 * with galEl = 5^k mod 2N. */
int sfg_ctx_set_rotation_key(sfg_ctx *ctx, int rot_left, const uint64_t *key);
int sfg_ctx_set_rotation_key_ptrs(sfg_ctx *ctx, int rot_left, const uint64_t *const *limbs /* beta*2*(nQ+nP) pointers */);
int sfg_ctx_has_rotation_key(const sfg_ctx *ctx, int rot_left);

/* ---- lattice primitives (each mirrors one reference / Lattigo function; used by the parity tests and the NTT sweep) ---- */
/* ring.NTT / ring.InvNTT on npoly polynomials, polynomial p uses modulus index limb_idx[p % nsel] (Q then P). In place. */
int sfg_ntt(sfg_ctx *ctx, uint64_t *polys, int npoly, const int *limb_idx, int nsel, int inverse);
/* MulCoeffsAndAdd128 (gwas/matmult.go:247-289): acc is n x {hi, lo} like the reference's uint128 struct */
int sfg_mul_coeffs_and_add128(sfg_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t *acc_hi_lo, size_t n);
/* ReduceAndAddUint128 (gwas/matmult.go:291-324) for modulus index `limb` */
int sfg_reduce_and_add_uint128(sfg_ctx *ctx, const uint64_t *acc_hi_lo, uint64_t *out, int limb, size_t n);
/* MFormLvl (gwas/matmult.go:411-431): p is [level+1][N], in place */
int sfg_mform_lvl(sfg_ctx *ctx, int level, uint64_t *p);
/* crypto.RotateRightWithEvaluator (crypto/basics.go:201-210) on nct ciphertexts [nct][2][level+1][N] */
int sfg_rotate_right(sfg_ctx *ctx, int level, const uint64_t *cts, int nct, int nrot, uint64_t *out);

/* ---- genotype matrix: what GenoFileStream.NextRow (gwas/filestream.go:414-426) delivers, kept in HBM ---- */
int sfg_geno_create(sfg_ctx *ctx, size_t nrows, size_t ncols, sfg_geno **out);
/* append `nrows_chunk` consecutive rows (row-major int8, ncols each; filtered rows/cols already removed by the Go stream) */
int sfg_geno_push_rows(sfg_geno *g, const int8_t *rows, size_t nrows_chunk);
void sfg_geno_destroy(sfg_geno *g);
/* EncodeDiagWithEncoder + ToMontgomeryForm for one (block row, shift) (gwas/matmult.go:711-731,401-409):
 * out [m_ct][level+1][N] (Montgomery form iff mont != 0), present[m_ct] = 0 for nil plaintexts. nrot = right rotation. */
int sfg_encode_diag(sfg_ctx *ctx, const sfg_geno *g, int block_row, int shift, int nrot, int level, int mont, uint64_t *out,
                    uint8_t *present, int64_t *coeffs_out /* optional [m_ct][N], NULL to skip */);

/* ---- the three stream entry points (gwas/matmult.go:914, 1043, 1238) ---- */
/* MatMult4StreamPreprocess(cryptoParams, gfs, maxLevel, cacheFilePrefix): builds the diagonal cache in HBM (or, when it
 * exceeds the budget, a handle that regenerates diagonals on the fly; results are identical). */
int sfg_matmult4_stream_preprocess(sfg_ctx *ctx, const sfg_geno *g, int max_level, sfg_cache **out);
/* block-row sharding: the cache of a rank holds the diagonals of its own block rows [bi_lo, bi_hi) only (HBM and
 * preprocessing time proportional to the rank's share); it serves sfg_matmult4_partial over that range and sfg_matmult4_finish */
int sfg_matmult4_stream_preprocess_rows(sfg_ctx *ctx, const sfg_geno *g, int max_level, int bi_lo, int bi_hi, sfg_cache **out);
/* giant-step sharding (strong scaling of ONE product over the GPUs of a box; SURVEY 8e, north_star "modular-add allreduce"): the
 * cache of share `part` of `nparts` holds the diagonals P[bi][g*d+b][bj] of its own contiguous share of the active giant steps g only
 * (HBM and preprocessing time / nparts).  sfg_matmult4_stream_compute* on it returns the partial sum over those giant steps,
 * S_part[i][bj] = sum_{g in share} RotL_{g d}(reduce(acc[i][g])[bj]) (gwas/matmult.go:1203-1227 is a mod-q sum over g, so the shares
 * add up bit-exactly): combine with an integer SUM all-reduce of the canonical residues + sfg_ct_mod_reduce. */
int sfg_matmult4_stream_preprocess_giants(sfg_ctx *ctx, const sfg_geno *g, int max_level, int part, int nparts, sfg_cache **out);
/* baby-step sharding on top of it (the baby rotations RotL_b(A[i][bi]), gwas/matmult.go:1083-1119, are the part every rank would otherwise
 * repeat): rank `part` of `nparts` rotates its share of the K = (block row, baby step) entries into the DEVICE buffer d_R, laid out as
 * nparts chunks of sfg_matmult4_baby_chunk_bytes() -- chunk r is exactly rank r's share, so ONE all-gather over NVLink completes the buffer --
 * and sfg_matmult4_stream_compute_r_dev runs the rest of Compute (MAC + giant-step sums over this rank's cache) on the complete buffer.
 * d_A [s][nbr][2][level_a+1][N] is a DEVICE pointer; s <= 16. */
size_t sfg_matmult4_baby_chunk_bytes(const sfg_ctx *ctx, const sfg_cache *cache, int s, int nparts);
int sfg_matmult4_baby_dev(sfg_ctx *ctx, const uint64_t *d_A, int s, int num_block_rows, int level_a, int max_level, const sfg_cache *cache, int part,
                          int nparts, void *d_R);
int sfg_matmult4_stream_compute_r_dev(sfg_ctx *ctx, const void *d_R, int s, int max_level, const sfg_cache *cache, uint64_t *d_out);
/* x mod q_l on npoly DEVICE polynomials [nl][N] (sums of up to 2^8 canonical residues after an integer all-reduce) */
int sfg_ct_mod_reduce(sfg_ctx *ctx, uint64_t *d_polys, size_t npoly, int nl);
void sfg_cache_destroy(sfg_cache *cache);
/* number of non-nil diagonal polynomials, bytes resident, and whether they are materialised */
int sfg_cache_info(const sfg_cache *cache, size_t *num_polys, size_t *bytes, int *materialised, int *m_ct, int *num_block_rows);
/* copy one cached plaintext to the host: out [max_level][N] (the used limbs, NTT + Montgomery form); *present = 0 if nil */
int sfg_cache_get_diag(sfg_ctx *ctx, const sfg_cache *cache, int block_row, int shift, int block_col, uint64_t *out, int *present);
/* The reference's on-disk cache, `<prefix>_<bi>.bin` per block row (DiagCacheStream, gwas/filestream.go:19-282):
 * write = what MatMult4StreamPreprocess leaves on disk (all maxLevel+1 limbs, Montgomery form, big-endian coefficients), so a CPU run
 * of the reference can consume a GPU preprocess; load = build the HBM cache from files the reference wrote for an nrows x ncols
 * matrix (records in any order; nrows = ncols = 0: take the block structure from the files alone, which is all the reference's
 * MatMult4StreamCompute has in hand).  A missing file fails like NewDiagCacheStream's panic. */
int sfg_cache_write_files(sfg_ctx *ctx, const sfg_cache *cache, const char *prefix);
int sfg_cache_load_files(sfg_ctx *ctx, const char *prefix, size_t nrows, size_t ncols, int max_level, sfg_cache **out);

/* crypto.SaveCipherMatrixToFile / LoadCipherMatrixFromFile (crypto/utilities.go:82-141): the on-disk CipherMatrix of
 * `assoc_cache_mult.%d.bin` (gwas/assoc.go:317-333,434-437) and `Qcomb.bin` -- {u32 nrows, u32 ncols, u64 len(sizes), sizes, u64 len(blob),
 * blob = concatenated Ciphertext.MarshalBinary()}, coefficients big-endian -- so MatMult outputs of a GPU run and of a CPU run of the
 * reference interoperate.  Host-only (no context, no device): cts [nrows][ncols][2][level+1][N], scales [nrows][ncols] (ct.Scale());
 * all ciphertexts of degree 1 at one level.  Errors: non-zero return, message from sfg_last_error(NULL). */
int sfg_cipher_matrix_save(const char *filename, int logN, const uint64_t *cts, const double *scales, int nrows, int ncols, int level);
int sfg_cipher_matrix_info(const char *filename, int *nrows, int *ncols, int *level, int *logN);
int sfg_cipher_matrix_load(const char *filename, int logN, uint64_t *cts, double *scales, int nrows, int ncols, int level);

/* MatMult4StreamCompute(cryptoParams, A, maxLevel, cacheFilePrefix) (gwas/matmult.go:1043-1236).
 * A: [s][num_block_rows][2][level_a+1][N]; out: [s][m_ct][2][max_level][N] = the deterministic sum
 * S[i][bj] = sum_g RotL_{g d}(reduce(acc[i][g])[bj]) at level max_level-1; the Go shim adds it into
 * crypto.CZeroMat with eva.Add to keep the reference's randomised-zero semantics. */
int sfg_matmult4_stream_compute(sfg_ctx *ctx, const uint64_t *A, int s, int num_block_rows, int level_a, int max_level,
                                const sfg_cache *cache, uint64_t *out);
/* same, A and out given as one pointer per limb: A_limbs[((i*nbr+bi)*2+c)*(level_a+1)+l], out_limbs[((i*m_ct+bj)*2+c)*max_level+l] */
int sfg_matmult4_stream_compute_ptrs(sfg_ctx *ctx, const uint64_t *const *A_limbs, int s, int num_block_rows, int level_a,
                                     int max_level, const sfg_cache *cache, uint64_t *const *out_limbs);

/* MatMult4Stream(cryptoParams, A, gfs, maxLevel, computeSquaredSum, square, nproc) (gwas/matmult.go:1238-1505).
 * sum / sq_sum: ncols float64 each (may be NULL when compute_squared_sum == 0). The genotype handle is not modified. */
int sfg_matmult4_stream(sfg_ctx *ctx, const uint64_t *A, int s, int level_a, const sfg_geno *g, int max_level,
                        int compute_squared_sum, int square, uint64_t *out, double *sum, double *sq_sum);

/* ---- multi-GPU building blocks: block-row sharding with a modular-add reduce between K2 and K6 ---- */
/* size in uint64 of the per-call accumulator image cv[d][m_ct][s][2][max_level][N] */
size_t sfg_cv_elems(const sfg_ctx *ctx, const sfg_cache *cache, int s, int max_level);
/* partial products over block rows [bi_lo, bi_hi): canonical residues written to the DEVICE buffer d_cv */
int sfg_matmult4_partial(sfg_ctx *ctx, const uint64_t *A, int s, int num_block_rows, int level_a, int max_level,
                         const sfg_cache *cache, int bi_lo, int bi_hi, uint64_t *d_cv);
/* after an integer SUM all-reduce / reduce-scatter of up to 2^8 partial images: x mod q per limb, on device */
int sfg_cv_mod_reduce(sfg_ctx *ctx, const sfg_cache *cache, int s, int max_level, uint64_t *d_cv, size_t first_elem, size_t num_elems);
/* giant-step rotations and sum for giant indices [g_lo, g_hi): out [s][m_ct][2][max_level][N] on the HOST (partial over g) */
int sfg_matmult4_finish(sfg_ctx *ctx, const sfg_cache *cache, int s, int max_level, const uint64_t *d_cv, int g_lo, int g_hi,
                        uint64_t *out);
/* device-resident variant for a reduce-scattered image: d_cv_share holds the giants [g_lo, g_hi) ONLY (what this rank received), d_out
 * [s][m_ct][2][max_level][N] is a DEVICE buffer (partial over g; combine with an integer all-reduce + sfg_ct_mod_reduce) */
int sfg_matmult4_finish_dev(sfg_ctx *ctx, const sfg_cache *cache, int s, int max_level, const uint64_t *d_cv_share, int g_lo, int g_hi,
                            uint64_t *d_out);
/* out = (a + b) mod q limb-wise on host buffers of ncts ciphertexts [2][nl][N] (combining per-rank partial sums) */
int sfg_ct_add(sfg_ctx *ctx, const uint64_t *a, const uint64_t *b, int ncts, int nl, uint64_t *out);

/* ---- genotype scan of the randomized-PCA sketch (gwas/pca.go:124-162) on the HBM-resident matrix ----
 * localSketch[rand_index[i]][j] += sgn[i] * row_i[j];  xsum[j] += row_i[j];  x2sum[j] += row_i[j]^2   (rows with dosages 0/1/2).
 * rand_index[nrows] in [0, kp), sgn[nrows] = +-1 (drawn by the caller's PRG); sketch [kp][ncols] float64 (exact integers), xsum /
 * x2sum [ncols] uint64.  scan_ms (optional): device time of the scan. */
int sfg_geno_count_sketch(sfg_ctx *ctx, const sfg_geno *g, const int32_t *rand_index, const int8_t *sgn, int kp, double *sketch,
                          uint64_t *xsum, uint64_t *x2sum, float *scan_ms);

/* ---- ciphertext algebra of the callers around the path --------------------------------------------
 * QXLazyNormStream / QXtLazyNormStream (gwas/matmult.go:27-77,83-116) wrap MatMult4StreamCompute in crypto.CMult, InnerProd,
 * InnerSumAll, CMultScalar, MaskTrunc and eval.Sub (crypto/basics.go:110-127,236-293,386-427,553-566); the network bootstrap
 * between them (mpcObj.Network.BootstrapMatAll) stays in Go.  Host buffers; ciphertext k of an operand is [2][nl][N] uint64 with
 * nl >= level+1 stored limbs (limbs above `level` are ignored = DropLevel); a count of 1 broadcasts like the reference does.
 * Scale bookkeeping (ct.Scale, the Rescale threshold loop) is metadata and stays with the caller: `nrescale` is the number of
 * ring.DivRoundByLastModulusNTT steps evaluator.Rescale(ct, params.Scale) performs (1 for two operands at params.Scale). */
/* relinearisation key cryptoParams.Rlk.Keys[0] (switching key s^2 -> s), same layout as a rotation key */
int sfg_ctx_set_relin_key(sfg_ctx *ctx, const uint64_t *key);
int sfg_ctx_set_relin_key_ptrs(sfg_ctx *ctx, const uint64_t *const *limbs /* beta*2*(nQ+nP) pointers */);
/* crypto.CMult / CMultScalar: out[k] = Rescale^nrescale(eval.MulRelinNew(x[k or 0], y[k or 0])), n = max(nx, ny) results at
 * level - nrescale, each [2][level+1-nrescale][N] */
int sfg_ct_mul_relin(sfg_ctx *ctx, int level, const uint64_t *x, int nx, int x_nl, const uint64_t *y, int ny, int y_nl, int nrescale,
                     uint64_t *out);
/* eval.MulRelinNew(plaintext, ct) + Rescale (crypto.MaskTrunc): pt k is [pt_nl][N], NTT domain, NOT Montgomery form */
int sfg_ct_mul_plain(sfg_ctx *ctx, int level, const uint64_t *pt, int npt, int pt_nl, const uint64_t *cts, int nct, int ct_nl, int nrescale,
                     uint64_t *out);
/* evaluator.Rescale: nrescale steps of ring.DivRoundByLastModulusNTT on cts [nct][2][level+1][N] -> [nct][2][level+1-nrescale][N] */
int sfg_ct_rescale(sfg_ctx *ctx, int level, const uint64_t *cts, int nct, int nrescale, uint64_t *out);
/* eval.Sub / eval.Add on operands of matching scale: out[k] = a[k or 0] -+ b[k or 0] at `level` ([2][level+1][N] each) */
int sfg_ct_sub(sfg_ctx *ctx, int level, const uint64_t *a, int na, int a_nl, const uint64_t *b, int nb, int b_nl, uint64_t *out);
int sfg_ct_add2(sfg_ctx *ctx, int level, const uint64_t *a, int na, int a_nl, const uint64_t *b, int nb, int b_nl, uint64_t *out);
/* crypto.InnerSumAll for nvec vectors of cnt ciphertexts ([nvec][cnt][2][level+1][N]): out [nvec][2][level+1][N]; needs the
 * power-of-two left-rotation keys 1, 2, 4, .. slots/2 (crypto/crypto.go:232-249) */
int sfg_inner_sum_all(sfg_ctx *ctx, int level, const uint64_t *cts, int nvec, int cnt, uint64_t *out);
/* EncodeNTT of an int8 slot vector v[slots] at params.Scale (the 0/1 mask of crypto.MaskTrunc), correctly rounded: out [level+1][N] */
int sfg_encode_slots_i8(sfg_ctx *ctx, const int8_t *v, int level, int mont, uint64_t *out);

/* ---- device-resident ciphertext vectors ("keeps Q on device across a power iteration") -------------------------------
 * The host-buffer entry points above move every operand over PCIe/C2C per call.  A sfg_cts is an array of n degree-1 ciphertexts
 * [n][2][nl][N] that STAYS in HBM between calls: the callers' algebra (CMult, CMultScalar, InnerSumAll, MaskTrunc, Sub) and the MatMult
 * itself chain on handles; only what the Go side really needs on the host (the ciphertexts that go into the network bootstrap, final
 * results) is downloaded.  Scale and level bookkeeping stays with the caller exactly as for the host-buffer variants. */
typedef struct sfg_cts sfg_cts;
int sfg_cts_upload(sfg_ctx *ctx, const uint64_t *host, int n, int nl, sfg_cts **out);
int sfg_cts_download(sfg_ctx *ctx, const sfg_cts *cts, uint64_t *host);
int sfg_cts_shape(const sfg_cts *cts, int *n, int *nl);
void sfg_cts_destroy(sfg_cts *cts);
/* copy of ciphertexts [first, first + count) (e.g. one row of a CipherMatrix stored row-major) */
int sfg_cts_slice(sfg_ctx *ctx, const sfg_cts *cts, int first, int count, sfg_cts **out);
/* MatMult4StreamCompute on handles: A holds s * num_block_rows ciphertexts (row-major), out s * m_ct at level max_level-1 */
int sfg_cts_matmult4_stream_compute(sfg_ctx *ctx, const sfg_cts *A, int s, int num_block_rows, int max_level, const sfg_cache *cache,
                                    sfg_cts **out);
/* crypto.CMult / CMultScalar, MaskTrunc (plaintext from the host), eval.Sub / Add, crypto.InnerSumAll -- see the host-buffer variants */
int sfg_cts_mul_relin(sfg_ctx *ctx, int level, const sfg_cts *x, const sfg_cts *y, int nrescale, sfg_cts **out);
int sfg_cts_mul_plain(sfg_ctx *ctx, int level, const uint64_t *pt, int npt, int pt_nl, const sfg_cts *cts, int nrescale, sfg_cts **out);
int sfg_cts_addsub(sfg_ctx *ctx, int level, const sfg_cts *a, const sfg_cts *b, int subtract, sfg_cts **out);
int sfg_cts_inner_sum_all(sfg_ctx *ctx, int level, const sfg_cts *cts, int nvec, int cnt, sfg_cts **out);

/* ---- local arithmetic of the collective bootstrap (mpc/mhe.go:262-341: CollectiveBootstrap / CollectiveBootstrapMat) ----
 * What every party computes per ciphertext around the two network aggregations (AggregateRefreshShare*, which stay in Go).  The random
 * draws stay in Go as well -- the mask (ring.RandInt, crypto/rand), the Gaussian noise, the common reference polynomial `crp` of the
 * shared PRG (crpGen.ReadNew()) -- and are handed over; the library does the exact arithmetic of dckks.RefreshProtocol
 * (un-vendored Lattigo fork: restated from the published v2.1 algorithm, oracle/refresh.py, [UNVERIFIED vs the fork]).
 * refProtocol.GenShares (mhe.go:303-311) for nct ciphertexts at `level`:
 *   c1 [nct][level+1][N]; sk_mont [nQ][N] = cps.Sk.Value.Coeffs (NTT + Montgomery form, as Lattigo stores it); crp [nct][nQ][N];
 *   mask as sign-magnitude big integers: mask_mag [nct][N][nwords] little-endian 64-bit words (big.Int.Bits()), mask_sign [nct][N] (big.Int.Sign());
 *   in_scale = ct.Scale(), out_scale = targetScale = params.Scale() (the recrypt share carries the mask scaled by their ratio);
 *   e0 / e1 [nct][N] the Gaussian noise coefficients  ->  share_decrypt [nct][level+1][N], share_recrypt [nct][nQ][N]. */
int sfg_refresh_gen_shares(sfg_ctx *ctx, int level, int nct, const uint64_t *c1, const uint64_t *sk_mont, const uint64_t *crp,
                           const uint64_t *mask_mag, const int8_t *mask_sign, int nwords, double in_scale, double out_scale, const int64_t *e0,
                           const int64_t *e1, uint64_t *share_decrypt, uint64_t *share_recrypt);
/* refProtocol.Decrypt + Recode + Recrypt (mhe.go:316-318): c0 [nct][c0_nl >= level+1][N] of the ciphertexts, the aggregated shares, the
 * same crp; in_scale = ct.Scale(), out_scale = params.Scale()  ->  out [nct][2][nQ][N]: refreshed ciphertexts at the top level, scale out_scale */
int sfg_refresh_finish(sfg_ctx *ctx, int level, int nct, const uint64_t *c0, int c0_nl, double in_scale, double out_scale,
                       const uint64_t *agg_decrypt, const uint64_t *agg_recrypt, const uint64_t *crp, uint64_t *out);

/* synchronise the context's stream (timing helper) */
int sfg_ctx_sync(sfg_ctx *ctx);
/* device-resident variant used by bench.py to time the kernels with inputs already in HBM:
 * d_A [s][nbr][2][level_a+1][N] and d_out [s][m_ct][2][max_level][N] are DEVICE pointers. */
int sfg_matmult4_stream_compute_dev(sfg_ctx *ctx, const uint64_t *d_A, int s, int num_block_rows, int level_a, int max_level,
                                    const sfg_cache *cache, uint64_t *d_out);
/* device-resident variants of sfg_ntt / sfg_rotate_right (d_* are DEVICE pointers; same layouts) for the NTT / key-switch throughput
 * sweep of BASELINE config 3 (profiles/sweep_ntt_ks.py) */
int sfg_ntt_dev(sfg_ctx *ctx, uint64_t *d_polys, int npoly, const int *limb_idx, int nsel, int inverse);
int sfg_rotate_right_dev(sfg_ctx *ctx, int level, const uint64_t *d_cts, int nct, int nrot, uint64_t *d_out);
/* last call's phase timings in milliseconds (CUDA events on the context stream):
 * {baby rotations, MAC phase, giant rotations, total, MAC kernel alone} */
int sfg_ctx_last_timings(const sfg_ctx *ctx, float out_ms[5]);
void *sfg_ctx_stream(const sfg_ctx *ctx); /* cudaStream_t of the context */

#ifdef __cplusplus
}
#endif
#endif
