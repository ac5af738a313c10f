// kernels.h -- launchers of the sm_100a kernels (kernels_*.cu). Host-callable, stream-ordered, no allocation inside.
#pragma once
#include "ctx.h"

namespace sfg {

struct LimbSel {
    int n;
    int idx[kMaxLimbs];
};

// ---- NTT (kernels_ntt.cu) ----
// Transforms npoly polynomials. poly p: group g = p / sel.n, k = p % sel.n; src = src_base + g*src_gstride + k*N,
// dst likewise; modulus index = sel.idx[k].  In-place allowed (src == dst with equal strides).
int launch_ntt(Ctx *c, const uint64_t *src, size_t src_gstride, uint64_t *dst, size_t dst_gstride, int npoly,
               const LimbSel &sel, bool inverse, cudaStream_t st);

// ---- lazy MAC primitives as stand-alone kernels (parity tests of K1/K2/K3) ----
int launch_mul_coeffs_and_add128(Ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *acc_hi_lo, size_t n, cudaStream_t st);
int launch_reduce_and_add128(Ctx *c, const uint64_t *acc_hi_lo, uint64_t *out, int limb, size_t n, cudaStream_t st);
int launch_mform(Ctx *c, uint64_t *p, int nlimbs, cudaStream_t st);  // [nlimbs][N], limbs 0..nlimbs-1

// ---- genotype preparation (kernels_geno.cu) ----
// in-place: missing (<0) -> 0, optional per-column sum / sum of squares (float64 atomics on exact small integers),
// optional squaring.  X is rows x ncols row-major.
int launch_geno_prep(Ctx *c, int8_t *X, size_t rows, size_t ncols, double *sum, double *sqsum, bool square, cudaStream_t st);

// ---- diagonal encoder (kernels_encode.cu) ----
struct EncJob {
    int row0;        // first matrix row of the block row (bi*slots)
    int col0;        // first matrix column of the block column (bj*slots)
    int r, cdim;     // block dimensions (<= slots)
    int shift;       // diagonal index in [0, slots)
    int nrot;        // right rotation applied before encoding (d*giant)
    long long out_off;  // element offset of the [nl][N] output polynomial
};
int launch_encode(Ctx *c, const int8_t *X, size_t ld, const EncJob *jobs_dev, int njobs, int nl, bool mont, uint64_t *out,
                  long long *coeff_out /* optional [njobs][N] int64 coefficient-domain message, may be null */, cudaStream_t st);

// ---- output-stationary MAC + Montgomery reduce (kernels_mac.cu) ----
// R: rotation cache [K][nrows][L][N]; P: plaintext diagonals, element offsets poff[col*K + k] (-1 = nil);
// cv: [ncols][nrows][L][N] canonical residues.
int launch_mac(Ctx *c, const uint64_t *R, const uint64_t *P, const long long *poff, int K, int nrows, int ncols, int L,
               uint64_t *cv, cudaStream_t st);

// ---- key-switch + automorphism (kernels_ks.cu) ----
struct KsBatch {
    int level;              // input level (nl = level+1 limbs)
    int nct;                // ciphertexts in the batch (all use the same Galois key)
    const uint64_t *in;     // ct k at in + in_off[k], layout [2][in_nl][N]
    const long long *in_off;   // device [nct]; must be the arithmetic progression in_first + k*in_stride
    long long in_first, in_stride;
    int in_nl;              // limb count of the stored input cts (>= level+1)
    uint64_t *out;          // ct k at out + out_off[k], layout [2][out_nl][N]; only limbs < out_limbs are produced
    const long long *out_off;  // device [nct]
    int out_nl;
    int out_limbs;
    bool accumulate;        // out += result (mod q) instead of out = result
    // scratch (device): c2 [nct][nl][N], acc [nct][2][nl+nP][N]
    uint64_t *c2, *acc;
};
int launch_rotate(Ctx *c, const KsBatch &b, const GaloisKey &key, cudaStream_t st);
// out (+)= in, limb-wise mod q, same offset conventions (used for rotation by 0)
int launch_copy_add(Ctx *c, const KsBatch &b, cudaStream_t st);

// modular canonicalisation of sums of residues: x[l][n] = x[l][n] mod q_l over npoly*[L][N] (multi-GPU reduce epilogue)
int launch_mod_reduce(Ctx *c, uint64_t *x, size_t npoly, int L, cudaStream_t st);

}  // namespace sfg
