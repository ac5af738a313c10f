// kernels.h -- launchers of the sm_100a kernels (kernels_*.cu). Host-callable, stream-ordered, no allocation inside.
#pragma once
#include "ctx.h"

namespace sfg {

struct LimbSel {
    int n;
    int idx[kMaxLimbs];
};

// Byte layout of one [nl][N] polynomial record with a per-limb element size: 8 = uint64 residues (the reference's
// layout; plaintext diagonals in Montgomery form), 4 = PACKED uint32 residues for moduli < 2^32 (plain, not Montgomery).
constexpr int kMaxLayoutLimbs = kMaxLimbs;  // full chains of the logN 15 / 16 sweep: 18 and 34 Q limbs
struct PolyLayout {
    int nl;
    int es[kMaxLayoutLimbs];
    long long off[kMaxLayoutLimbs];
    long long bytes;
};
PolyLayout make_layout(const Ctx *c, int nl, bool packed);

// ---- NTT (kernels_ntt.cu) ----
// Transforms npoly polynomials. poly p: group g = p / sel.n, k = p % sel.n; src = src_base + g*src_gstride + k*N,
// dst likewise; modulus index = sel.idx[k].  In-place allowed (src == dst with equal strides).
int launch_ntt(Ctx *c, const uint64_t *src, size_t src_gstride, uint64_t *dst, size_t dst_gstride, int npoly,
               const LimbSel &sel, bool inverse, cudaStream_t st);
// same with per-group source offsets (device array of element offsets, overrides g*src_gstride; may be null) and, for inverse
// transforms, inputs stored in TT order (ntt2.cuh)
int launch_ntt_gather(Ctx *c, const uint64_t *src, const long long *src_off, size_t src_gstride, uint64_t *dst, size_t dst_gstride,
                      int npoly, const LimbSel &sel, bool inverse, bool in_tt, cudaStream_t st);

// ---- lazy MAC primitives as stand-alone kernels (parity tests of K1/K2/K3) ----
int launch_mul_coeffs_and_add128(Ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *acc_hi_lo, size_t n, cudaStream_t st);
int launch_reduce_and_add128(Ctx *c, const uint64_t *acc_hi_lo, uint64_t *out, int limb, size_t n, cudaStream_t st);
int launch_mform(Ctx *c, uint64_t *p, int nlimbs, cudaStream_t st);  // [nlimbs][N], limbs 0..nlimbs-1

// ---- genotype preparation (kernels_geno.cu) ----
// in-place: missing (<0) -> 0, optional per-column sum / sum of squares (float64 atomics on exact small integers),
// optional squaring.  X is rows x ncols row-major.
int launch_geno_prep(Ctx *c, int8_t *X, size_t rows, size_t ncols, double *sum, double *sqsum, bool square, cudaStream_t st);

// count sketch + column sums of the PCA sketch (kernels_geno.cu; gwas/pca.go:152-162), see the kernel for the argument layout
int launch_count_sketch(Ctx *c, const int8_t *X, size_t nrows, size_t ncols, const int *rows_sorted, const int8_t *sgn_sorted,
                        const int *bucket_off, int kp, long long *sketch_i64, double *sketch_f64, unsigned long long *xsum,
                        unsigned long long *x2sum, int *bad, cudaStream_t st);

// ---- diagonal encoder (kernels_encode.cu) ----
struct EncJob {
    int row0;        // first matrix row of the block row (bi*slots)
    int col0;        // first matrix column of the block column (bj*slots)
    int r, cdim;     // block dimensions (<= slots)
    int shift;       // diagonal index in [0, slots)
    int nrot;        // right rotation applied before encoding (d*giant)
    long long out_off;  // BYTE offset of the output polynomial record
};
// lay: output record layout (limbs 0..lay.nl-1). 8-byte limbs are stored in Montgomery form iff mont; 4-byte limbs plain.
int launch_encode(Ctx *c, const int8_t *X, size_t ld, const EncJob *jobs_dev, int njobs, const PolyLayout &lay, bool mont, void *out,
                  long long *coeff_out /* optional [njobs][N] int64 coefficient-domain message, may be null */, cudaStream_t st);

// ---- tensor-core MAC (kernels_mactc.cu): byte-plane images + tcgen05 kind::i8 contraction ----
constexpr int kTcLimbs = 8;
struct TcGeomP {                 // geometry of the plaintext-diagonal image (fixed at preprocess time)
    int L, N, K, ncols;
    int ntiles;                  // column tiles of 128
    int SBN;                     // coefficient groups (of 4) per superblock
    int Kg, ngroups;             // K bytes per stage / K groups (one launch each)
    int nb[kTcLimbs];            // byte planes per residue of limb l
    long long pbase[kTcLimbs];   // byte offset of limb l inside one K-group image of `ntiles` tiles
    long long group_bytes;
};
struct TcGeomR {                 // geometry of the rotated-ciphertext image (per call: depends on the number of rows)
    int rows, RP;
    int cv_rows = 0, cv_row0 = 0;  // this launch's rows are rows [cv_row0, cv_row0 + rows) of a cv image with cv_rows rows per column (0 = rows)
    int npad[kTcLimbs];
    long long rbase[kTcLimbs];
    long long group_bytes;
    int tbuf_stride;
};
int tc_geom_p(Ctx *c, int L, int K, int ncols, TcGeomP *g);
int tc_geom_r(Ctx *c, const TcGeomP &gp, int rows, TcGeomR *g, int rp_min = 0);
// records -> image.  src_off_dev: device [Kg][128] (P) / [Kg][rows] (R) byte offsets of the source records, -1 = zero.
int launch_img_p(Ctx *c, const TcGeomP &g, const PolyLayout &lay, const void *records, const long long *src_off_dev, int img_ntiles,
                 int ct_in_img, void *img_group, cudaStream_t st);
int launch_img_r(Ctx *c, const TcGeomP &gp, const TcGeomR &gr, const PolyLayout &lay, const void *R, const long long *src_off_dev,
                 void *img_group, cudaStream_t st);
int launch_img_extract(Ctx *c, const TcGeomP &g, const void *img_group, int l, int col, int k_in_group, uint64_t *out_dev, cudaStream_t st);
// cv[col - col_lo][row][l][n] (+)= sum_k R*P mod q for the columns of tiles [tile_lo, tile_hi) that lie in [col_lo, col_hi), over `ngroups`
// consecutive K groups (images p_gstride / r_gstride bytes apart) accumulated in TMEM by ONE launch (<= tc_max_fused_groups)
int tc_max_fused_groups(const TcGeomP &gp);
int launch_mac_tc(Ctx *c, const TcGeomP &gp, const TcGeomR &gr, const void *Pimg_group, long long p_gstride, int img_ntiles, int img_tile0,
                  const void *Rimg_group, long long r_gstride, int ngroups, int tile_lo, int tile_hi, int col_lo, int col_hi, bool accumulate,
                  uint64_t *cv, cudaStream_t st, const void *Rimg_group2 = nullptr, int cv_row0_2 = 0, int rows2 = 0);
// Rimg_group2 != nullptr ("pair mode"): a second row part of the SAME image geometry (RP, npad) with rows2 rows starting at cv_row0_2 is computed by
// the same launch -- 2-CTA clusters, each CTA fetches half of every P stage and multicasts it to both, so P leaves HBM once for both parts

// ---- key-switch + automorphism (kernels_ks.cu) ----
// A batch is a list of ciphertexts, each rotated with its own Galois key.  All arrays are DEVICE arrays of nct entries.
struct KsBatch {
    int level;              // input level (nl = level+1 limbs)
    int nct;                // ciphertexts in the batch
    const uint64_t *in;     // ct k at in + in_off[k], layout [2][in_nl][N]
    const long long *in_off;
    int in_nl;              // limb count of the stored input cts (>= level+1)
    // INTT(c1) is computed once per DISTINCT input: slot j of c2 holds the transform of the ct at in + c2_src_off[j];
    // batch entry k reads slot c2_slot[k]
    int n_c2;
    const long long *c2_src_off;  // device [n_c2]
    const int *c2_slot;           // device [nct]
    const uint64_t *const *keys;  // device [nct]: converted Galois key (launch_key_convert) of entry k
    const uint32_t *const *perms; // device [nct]: PermuteNTTIndex table of entry k
    void *out;              // ct k at (char*)out + out_off[k] BYTES: two consecutive records of out_layout (c0 then c1);
    const long long *out_off;  //                                       only limbs < out_layout.nl are produced
    PolyLayout out_layout;
    bool accumulate;        // out += result (mod q) instead of out = result; entries of one batch must target distinct outputs
    // scratch (device): c2 [n_c2][nl][N], acc [nct][2][nl+nP][N]
    uint64_t *c2, *acc;
    uint64_t *dout = nullptr;  // internal: digit-transform output of the shared-decomposition path (kernels_ks.cu)
    const uint32_t *vq = nullptr;  // internal, nP > 1: quotient estimates of the digit base conversion, [n_c2][beta][N] (k_ks_bcprep)
    const uint64_t *extd = nullptr;  // internal, nP > 1: every digit extended to every target, [n_c2][beta][nt][N] (k_ks_extd)
    bool acc_dlog = false;     // internal: the Q limbs of acc are written in discrete-log order (giant-step sums)
    int acc_cap;            // ciphertexts the acc scratch holds; larger batches are processed in chunks
};
int launch_rotate(Ctx *c, const KsBatch &b, cudaStream_t st);
// Sum of rotations: entries a*nout + o (a < nct/nout), each with its own key, summed into output o with the mod-down hoisted out
// of the sum (exact; kernels_ks.cu).  ginv_dev: device [nct] galEl^-1 mod 2N of every entry.  S1 / C0 / E: device
// [nout][2][L][N] partial sums carried across calls (first = overwrite); launch_rotate_sum_final adds the result into out[o].
int launch_rotate_sum(Ctx *c, const KsBatch &b, int nout, const uint32_t *ginv_dev, uint64_t *S1, uint64_t *C0, uint64_t *E, bool first,
                      cudaStream_t st);
int launch_rotate_sum_final(Ctx *c, int level, int nout, const uint64_t *S1, const uint64_t *C0, const uint64_t *E, void *out,
                            const long long *out_off, const PolyLayout &olay, cudaStream_t st);
// out (+)= in, limb-wise mod q, same offset conventions (used for rotation by 0)
int launch_copy_add(Ctx *c, const KsBatch &b, cudaStream_t st);
// Lattigo SwitchingKey [beta][2][nQP][N] (NTT + Montgomery) -> device format of k_ks_inner2 (TT order; narrow moduli as Shoup pairs)
int launch_key_convert(Ctx *c, const uint64_t *in, uint64_t *out, cudaStream_t st);

// ---- ciphertext algebra of the callers around the path (kernels_ctalg.cu; gwas/matmult.go:27-116, crypto/basics.go) ----
// ct k of an operand lives at base + k*stride (stride 0 = broadcast), stored as [2][*_nl][N]; results are dense [nct][2][nl][N]
int launch_ct_tensor(Ctx *c, const uint64_t *a, long long a_stride, int a_nl, const uint64_t *b, long long b_stride, int b_nl, int nl, int nct,
                     uint64_t *tmp /* (d0, d2) */, uint64_t *out /* (0, d1) */, cudaStream_t st);
int launch_pt_mul(Ctx *c, const uint64_t *pt, long long pt_stride, const uint64_t *ct, long long ct_stride, int ct_nl, int nl, int nct,
                  uint64_t *out, cudaStream_t st);
int launch_ct_addsub(Ctx *c, const uint64_t *a, long long a_stride, int a_nl, const uint64_t *b, long long b_stride, int b_nl, int nl, int nct,
                     bool sub, uint64_t *out, cudaStream_t st);
int launch_ct_sum(Ctx *c, const uint64_t *in, int nvec, int cnt, int nl, uint64_t *out, cudaStream_t st);
// ring.DivRoundByLastModulusNTT: npoly polynomials [level+1][N] -> [level][N]; scratch T [npoly][N], U [npoly][level][N]
int launch_rescale(Ctx *c, int level, const uint64_t *in, int npoly, uint64_t *out, uint64_t *T, uint64_t *U, cudaStream_t st);

// ---- on-disk diagonal cache records (kernels_cachefile.cu; gwas/filestream.go:42-282) ----
int launch_bswap64(Ctx *c, uint64_t *x, size_t n, cudaStream_t st);  // little <-> big endian in place
int launch_file_to_rec(Ctx *c, const uint64_t *raw, const long long *dst_off_dev, int npoly, int file_nl, const PolyLayout &lay, void *out,
                       cudaStream_t st);

// modular canonicalisation of sums of residues: x[l][n] = x[l][n] mod q_l over npoly*[L][N] (multi-GPU reduce epilogue)
int launch_mod_reduce(Ctx *c, uint64_t *x, size_t npoly, int L, cudaStream_t st);

}  // namespace sfg
