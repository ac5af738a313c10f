// kernels_ctalg.cu -- element-wise kernels of the ciphertext algebra the callers of the MatMult path use around it
// (SURVEY 8 rows a4 / f2): crypto.CMult / CMultScalar / MaskTrunc / InnerSumAll / Sub in QXLazyNormStream and QXtLazyNormStream
// (gwas/matmult.go:27-116, crypto/basics.go:110-127,236-293,386-427,553-566).  Semantics: Lattigo v2.1 evaluator.mulRelin,
// Rescale -> ring.DivRoundByLastModulusNTT, Add / Sub.  All results are canonical residues, so only the mathematical value of each
// step matters; the relinearisation key-switch and the transforms reuse kernels_ks.cu / kernels_ntt.cu.
// These kernels are pure streaming work (HBM-bound): one thread per coefficient, unit-stride 8-byte accesses, grid = (N/256, limb, ct).
#include "kernels.h"

namespace sfg {

// tensor product of two degree-1 ciphertexts: d0 = a0*b0, d1 = a0*b1 + a1*b0, d2 = a1*b1  (mulRelin before relinearisation).
// tmp[ct] = (d0, d2) is the pseudo-ciphertext the key-switch reads (it returns (d0 + ks0(d2), ks1(d2))); out[ct] = (0, d1) is the
// accumulation target, so that out = (d0 + ks0, d1 + ks1) after the accumulate-mode mod-down.
__global__ void k_ct_tensor(const uint64_t *__restrict__ a, long long a_stride, int a_nl, const uint64_t *__restrict__ b, long long b_stride,
                            int b_nl, int nl, int N, const LimbConst *__restrict__ lcs, uint64_t *__restrict__ tmp, uint64_t *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y, ct = blockIdx.z;
    if (j >= N) return;
    const LimbConst lc = lcs[l];
    const uint64_t *A = a + (size_t)ct * a_stride, *B = b + (size_t)ct * b_stride;
    const uint64_t a0 = mform(A[(size_t)l * N + j], lc), a1 = mform(A[((size_t)a_nl + l) * N + j], lc);
    const uint64_t b0 = B[(size_t)l * N + j], b1 = B[((size_t)b_nl + l) * N + j];
    const size_t o0 = (((size_t)ct * 2 + 0) * nl + l) * N + j, o1 = (((size_t)ct * 2 + 1) * nl + l) * N + j;
    tmp[o0] = mred(a0, b0, lc);
    tmp[o1] = mred(a1, b1, lc);
    out[o0] = 0;
    out[o1] = add_mod(mred(a0, b1, lc), mred(a1, b0, lc), lc.q);
}
int launch_ct_tensor(Ctx *c, const uint64_t *a, long long a_stride, int a_nl, const uint64_t *b, long long b_stride, int b_nl, int nl, int nct,
                     uint64_t *tmp, uint64_t *out, cudaStream_t st) {
    if (nct <= 0) return 0;
    k_ct_tensor<<<dim3((c->N + 255) / 256, nl, nct), 256, 0, st>>>(a, a_stride, a_nl, b, b_stride, b_nl, nl, c->N, c->lc, tmp, out);
    SFG_LAUNCHED(c, "k_ct_tensor", st);
    return 0;
}

// plaintext x ciphertext (MulRelinNew(pt, ct), crypto/basics.go:121): both components times the NTT-domain plaintext
__global__ void k_pt_mul(const uint64_t *__restrict__ pt, long long pt_stride, const uint64_t *__restrict__ ct, long long ct_stride, int ct_nl,
                         int nl, int N, const LimbConst *__restrict__ lcs, uint64_t *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y % nl, comp = blockIdx.y / nl, t = blockIdx.z;
    if (j >= N) return;
    const LimbConst lc = lcs[l];
    const uint64_t p = mform(pt[(size_t)t * pt_stride + (size_t)l * N + j], lc);
    const uint64_t x = ct[(size_t)t * ct_stride + ((size_t)comp * ct_nl + l) * N + j];
    out[(((size_t)t * 2 + comp) * nl + l) * N + j] = mred(p, x, lc);
}
int launch_pt_mul(Ctx *c, const uint64_t *pt, long long pt_stride, const uint64_t *ct, long long ct_stride, int ct_nl, int nl, int nct,
                  uint64_t *out, cudaStream_t st) {
    if (nct <= 0) return 0;
    k_pt_mul<<<dim3((c->N + 255) / 256, 2 * nl, nct), 256, 0, st>>>(pt, pt_stride, ct, ct_stride, ct_nl, nl, c->N, c->lc, out);
    SFG_LAUNCHED(c, "k_pt_mul", st);
    return 0;
}

// out = a +- b limb-wise (evaluator.Add / Sub on operands of matching scale); operands may be stored with more limbs than nl
__global__ void k_ct_addsub(const uint64_t *__restrict__ a, long long a_stride, int a_nl, const uint64_t *__restrict__ b, long long b_stride,
                            int b_nl, int nl, int N, const LimbConst *__restrict__ lcs, int sub, uint64_t *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y % nl, comp = blockIdx.y / nl, t = blockIdx.z;
    if (j >= N) return;
    const uint64_t q = lcs[l].q;
    const uint64_t x = a[(size_t)t * a_stride + ((size_t)comp * a_nl + l) * N + j];
    const uint64_t y = b[(size_t)t * b_stride + ((size_t)comp * b_nl + l) * N + j];
    out[(((size_t)t * 2 + comp) * nl + l) * N + j] = sub ? sub_mod(x, y, q) : add_mod(x, y, q);
}
int launch_ct_addsub(Ctx *c, const uint64_t *a, long long a_stride, int a_nl, const uint64_t *b, long long b_stride, int b_nl, int nl, int nct,
                     bool sub, uint64_t *out, cudaStream_t st) {
    if (nct <= 0) return 0;
    k_ct_addsub<<<dim3((c->N + 255) / 256, 2 * nl, nct), 256, 0, st>>>(a, a_stride, a_nl, b, b_stride, b_nl, nl, c->N, c->lc, sub ? 1 : 0, out);
    SFG_LAUNCHED(c, "k_ct_addsub", st);
    return 0;
}

// out[v] = sum of the cnt ciphertexts of vector v (first step of crypto.InnerSumAll, crypto/basics.go:278-290)
__global__ void k_ct_sum(const uint64_t *__restrict__ in, int cnt, int nl, int N, const LimbConst *__restrict__ lcs, uint64_t *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y % nl, comp = blockIdx.y / nl, v = blockIdx.z;
    if (j >= N) return;
    const uint64_t q = lcs[l].q;
    const size_t ctsz = (size_t)2 * nl * N, o = ((size_t)comp * nl + l) * N + j;
    uint64_t s = 0;
    for (int t = 0; t < cnt; t++) s = add_mod(s, in[((size_t)v * cnt + t) * ctsz + o], q);
    out[(size_t)v * ctsz + o] = s;
}
int launch_ct_sum(Ctx *c, const uint64_t *in, int nvec, int cnt, int nl, uint64_t *out, cudaStream_t st) {
    if (nvec <= 0) return 0;
    k_ct_sum<<<dim3((c->N + 255) / 256, 2 * nl, nvec), 256, 0, st>>>(in, cnt, nl, c->N, c->lc, out);
    SFG_LAUNCHED(c, "k_ct_sum", st);
    return 0;
}

// ---- ring.DivRoundByLastModulusNTT (one step of evaluator.Rescale) -----------------------------------------------------------
//   t = InvNTT(x_L)                                  (launch_ntt, limb L of every polynomial)
//   u_l = ((t + (qL-1)/2) mod qL) - (qL-1)/2 mod q_l (k_rescale_prep: the centred representative of x mod qL, reduced mod q_l)
//   z_l = NTT_l(u_l)                                 (launch_ntt)
//   out_l = (x_l - z_l) * qL^-1 mod q_l              (k_rescale_fin)
struct RescaleInv {
    uint64_t inv[kMaxLimbs];  // qL^-1 mod q_l
};
__global__ void k_rescale_prep(const uint64_t *__restrict__ T, int level, int N, const LimbConst *__restrict__ lcs, uint64_t *__restrict__ U) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y, p = blockIdx.z;
    if (j >= N) return;
    const uint64_t qL = lcs[level].q, half = (qL - 1) >> 1;
    const LimbConst lc = lcs[l];
    uint64_t t = T[(size_t)p * N + j] + half;
    if (t >= qL) t -= qL;
    const uint64_t halfneg = lc.q - bred_add(half, lc);
    U[((size_t)p * level + l) * N + j] = bred_add(t + halfneg, lc);
}
__global__ void k_rescale_fin(const uint64_t *__restrict__ x, const uint64_t *__restrict__ U, int level, int N, const LimbConst *__restrict__ lcs,
                              RescaleInv ri, uint64_t *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y, p = blockIdx.z;
    if (j >= N) return;
    const LimbConst lc = lcs[l];
    const uint64_t d = sub_mod(x[((size_t)p * (level + 1) + l) * N + j], U[((size_t)p * level + l) * N + j], lc.q);
    out[((size_t)p * level + l) * N + j] = mul_mod(d, ri.inv[l], lc);
}
// in: npoly polynomials [level+1][N] (NTT domain), out: npoly polynomials [level][N]; T: scratch [npoly][N], U: scratch [npoly][level][N]
int launch_rescale(Ctx *c, int level, const uint64_t *in, int npoly, uint64_t *out, uint64_t *T, uint64_t *U, cudaStream_t st) {
    if (npoly <= 0) return 0;
    if (level < 1 || level >= c->nQ) SFG_FAIL(c, "rescale: level %d out of range [1, %d)", level, c->nQ);
    const int N = c->N, nl = level + 1;
    LimbSel last;
    last.n = 1;
    last.idx[0] = level;
    if (launch_ntt(c, in + (size_t)level * N, (size_t)nl * N, T, (size_t)N, npoly, last, true, st)) return -1;
    k_rescale_prep<<<dim3((N + 255) / 256, level, npoly), 256, 0, st>>>(T, level, N, c->lc, U);
    SFG_LAUNCHED(c, "k_rescale_prep", st);
    LimbSel low;
    low.n = level;
    for (int l = 0; l < level; l++) low.idx[l] = l;
    if (launch_ntt(c, U, (size_t)level * N, U, (size_t)level * N, npoly * level, low, false, st)) return -1;
    RescaleInv ri;
    for (int l = 0; l < level; l++) ri.inv[l] = h_invmod(c->mod[level] % c->mod[l], c->mod[l]);
    k_rescale_fin<<<dim3((N + 255) / 256, level, npoly), 256, 0, st>>>(in, U, level, N, c->lc, ri, out);
    SFG_LAUNCHED(c, "k_rescale_fin", st);
    return 0;
}

}  // namespace sfg
