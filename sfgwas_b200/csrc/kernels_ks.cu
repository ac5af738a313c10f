// kernels_ks.cu -- rotation = hybrid key-switch + NTT-domain automorphism (K6/K7 of SURVEY 2.2).
//
// Semantics: crypto.RotateRightWithEvaluator (crypto/basics.go:201-210) -> Lattigo v2.1 Evaluator.RotateNew ->
// permuteNTT: (d0, d1) = switchKeysInPlace(c1, rtk); (d0 + c0, d1) permuted with PermuteNTTIndex(galEl)
// ([UNVERIFIED vs the fork]).  All residues leave every step canonical, so the result depends
// only on the mathematical value of each step -- including the float64 quotient estimate `v` of Lattigo's fast exact
// base conversion, which is reproduced operation by operation (IEEE division and addition in index order).
//
// Batched over ciphertexts that share one Galois key (all rows / block columns of one baby or giant step):
//   1. INTT of c1                                                   (launch_ntt)
//   2. k_ks_inner : per (ct, target modulus t in Q_level U P): for every digit, base-convert the digit to t, NTT in
//                   shared memory (or reuse the NTT-domain input limb inside the digit), 128-bit lazy MAC with both
//                   key polynomials held in registers, Montgomery reduce -> acc[ct][2][t]
//   3. INTT of the P limbs of acc                                   (launch_ntt)
//   4. k_ks_moddown : per (ct, component, Q limb): base-convert P -> q_l, NTT, (acc - ext) * P^-1, + c0, automorphism
//                   gather from shared memory, store or accumulate (the giant-step sum of gwas/matmult.go:1223-1227).
#include "kernels.h"
#include "ntt.cuh"

namespace sfg {

// Lattigo fast exact base conversion for one coefficient: residues xs[k] (k < ns) -> target modulus t.
__device__ __forceinline__ uint64_t base_conv_coeff(const BaseConv &bc, const uint64_t *xs, const LimbConst *lcs, uint64_t t) {
    double vi = 0.0;
    uint64_t acc = 0;
#pragma unroll 1
    for (int k = 0; k < bc.ns; k++) {
        const uint64_t sk = lcs[bc.src_limb[k]].q;
        const uint64_t y = mul_shoup(xs[k], bc.sinv[k], bc.sinv_sh[k], sk);
        vi += __ddiv_rn((double)y, bc.sf[k]);
        acc = add_mod(acc, mul_shoup(y, bc.fac[k], bc.fac_sh[k], t), t);
    }
    const uint64_t v = (uint64_t)vi;
    return sub_mod(acc, mul_shoup(v, bc.smod, bc.smod_sh, t), t);
}

template <int NPER>
__global__ void __launch_bounds__(1024, 1)
k_ks_inner(const uint64_t *__restrict__ in, const long long *__restrict__ in_off, int in_nl, const uint64_t *__restrict__ c2,
           const uint64_t *__restrict__ key, const BaseConv *__restrict__ ks, int level, int nQ, int nP, int logN,
           const uint64_t *__restrict__ tw, const LimbConst *__restrict__ lcs, uint64_t *__restrict__ accout) {
    extern __shared__ __align__(16) uint64_t s[];
    const int N = 1 << logN, nl = level + 1, nt = nl + nP, nQP = nQ + nP;
    const int alpha = nP, beta = (nl + alpha - 1) / alpha;
    const int tt = blockIdx.x, ct = blockIdx.y;
    const int tgt = tt < nl ? tt : nQ + (tt - nl);
    const int T = blockDim.x, tid = threadIdx.x;
    const LimbConst lc = lcs[tgt];
    const NttTab tab = ntt_tab(tw, tgt, N);
    const uint64_t *c1 = in + in_off[ct] + (size_t)in_nl * N;  // second polynomial of the input ct
    const uint64_t *c2ct = c2 + (size_t)ct * nl * N;

    // canonical accumulators (2 registers each): one Montgomery reduction per product keeps the kernel at 1024 threads
    uint64_t a0[NPER], a1[NPER];
#pragma unroll
    for (int r = 0; r < NPER; r++) a0[r] = a1[r] = 0;

    for (int i = 0; i < beta; i++) {
        const BaseConv &bc = ks[(size_t)i * nt + tt];
        const uint64_t *k0 = key + ((size_t)(i * 2 + 0) * nQP + tgt) * N;
        const uint64_t *k1 = key + ((size_t)(i * 2 + 1) * nQP + tgt) * N;
        const int ns = bc.ns;
        if (ns == 0) {  // target lies inside the digit: reuse the NTT-domain input limb (decomposeAndSplitNTT)
#pragma unroll
            for (int r = 0; r < NPER; r++) {
                const int k = tid + r * T;
                const uint64_t v = c1[(size_t)tt * N + k];
                a0[r] = add_mod(a0[r], mred(v, k0[k], lc), lc.q);
                a1[r] = add_mod(a1[r], mred(v, k1[k], lc), lc.q);
            }
            continue;
        }
        if (ns == 1) {  // single-modulus digit: BRedAdd of the integer representative (DecomposeAndSplit)
            const uint64_t *x = c2ct + (size_t)bc.src_limb[0] * N;
#pragma unroll
            for (int r = 0; r < NPER; r++) {
                const int k = tid + r * T;
                s[k] = bred_add(x[k], lc);
            }
        } else {
#pragma unroll 1
            for (int r = 0; r < NPER; r++) {
                const int k = tid + r * T;
                uint64_t xs[kMaxAlpha];
                for (int j = 0; j < ns; j++) xs[j] = c2ct[(size_t)bc.src_limb[j] * N + k];
                s[k] = base_conv_coeff(bc, xs, lcs, lc.q);
            }
        }
        __syncthreads();
        ntt_fwd_smem(s, logN, 1, 0, tab, lc.q);
#pragma unroll
        for (int r = 0; r < NPER; r++) {
            const int k = tid + r * T;
            const uint64_t v = s[k];
            a0[r] = add_mod(a0[r], mred(v, k0[k], lc), lc.q);
            a1[r] = add_mod(a1[r], mred(v, k1[k], lc), lc.q);
        }
        __syncthreads();
    }
    uint64_t *o0 = accout + ((size_t)(ct * 2 + 0) * nt + tt) * N;
    uint64_t *o1 = accout + ((size_t)(ct * 2 + 1) * nt + tt) * N;
#pragma unroll
    for (int r = 0; r < NPER; r++) {
        const int k = tid + r * T;
        o0[k] = a0[r];
        o1[k] = a1[r];
    }
}

template <int NPER>
__global__ void __launch_bounds__(1024, 1)
k_ks_moddown(const uint64_t *__restrict__ in, const long long *__restrict__ in_off, int in_nl, const uint64_t *__restrict__ acc,
             const BaseConv *__restrict__ md, const uint64_t *__restrict__ pinv, const uint32_t *__restrict__ perm, int level,
             int nQ, int nP, int logN, const uint64_t *__restrict__ tw, const LimbConst *__restrict__ lcs,
             unsigned char *__restrict__ out, const long long *__restrict__ out_off, PolyLayout olay, int accumulate) {
    extern __shared__ __align__(16) uint64_t s[];
    const int N = 1 << logN, nl = level + 1, nt = nl + nP;
    const int l = blockIdx.x, comp = blockIdx.y, ct = blockIdx.z;
    const int T = blockDim.x, tid = threadIdx.x;
    const LimbConst lc = lcs[l];
    const NttTab tab = ntt_tab(tw, l, N);
    const BaseConv &bc = md[l];
    const uint64_t *accP = acc + ((size_t)(ct * 2 + comp) * nt + nl) * N;  // P limbs, coefficient domain
    const uint64_t *accQ = acc + ((size_t)(ct * 2 + comp) * nt + l) * N;
    const uint64_t pi = pinv[2 * l], pish = pinv[2 * l + 1];

#pragma unroll 1
    for (int r = 0; r < NPER; r++) {
        const int k = tid + r * T;
        uint64_t xs[kMaxAlpha];
        for (int j = 0; j < nP; j++) xs[j] = accP[(size_t)j * N + k];
        s[k] = base_conv_coeff(bc, xs, lcs, lc.q);
    }
    __syncthreads();
    ntt_fwd_smem(s, logN, 1, 0, tab, lc.q);
    const uint64_t *c0 = in + in_off[ct] + (size_t)l * N;
#pragma unroll
    for (int r = 0; r < NPER; r++) {
        const int k = tid + r * T;
        uint64_t v = mul_shoup(sub_mod(accQ[k], s[k], lc.q), pi, pish, lc.q);
        if (comp == 0) v = add_mod(v, c0[k], lc.q);
        s[k] = v;
    }
    __syncthreads();
    unsigned char *ob = out + out_off[ct] + (size_t)comp * olay.bytes + olay.off[l];
    if (olay.es[l] == 4) {
        uint32_t *o = reinterpret_cast<uint32_t *>(ob);
#pragma unroll
        for (int r = 0; r < NPER; r++) {
            const int k = tid + r * T;
            uint64_t v = s[perm[k]];
            if (accumulate) v = add_mod(v, (uint64_t)o[k], lc.q);
            o[k] = (uint32_t)v;
        }
    } else {
        uint64_t *o = reinterpret_cast<uint64_t *>(ob);
#pragma unroll
        for (int r = 0; r < NPER; r++) {
            const int k = tid + r * T;
            uint64_t v = s[perm[k]];  // PermuteNTTWithIndexLvl: out[k] = in[index[k]]
            if (accumulate) v = add_mod(v, o[k], lc.q);
            o[k] = v;
        }
    }
}

__global__ void k_copy_add(const uint64_t *__restrict__ in, const long long *__restrict__ in_off, int in_nl,
                           unsigned char *__restrict__ out, const long long *__restrict__ out_off, PolyLayout olay, int N,
                           const LimbConst *__restrict__ lcs, int accumulate) {
    const int l = blockIdx.x % olay.nl, comp = blockIdx.x / olay.nl, ct = blockIdx.y;
    const uint64_t q = lcs[l].q;
    const uint64_t *src = in + in_off[ct] + ((size_t)comp * in_nl + l) * N;
    unsigned char *db = out + out_off[ct] + (size_t)comp * olay.bytes + olay.off[l];
    if (olay.es[l] == 4) {
        uint32_t *dst = reinterpret_cast<uint32_t *>(db);
        for (int k = threadIdx.x; k < N; k += blockDim.x) dst[k] = (uint32_t)(accumulate ? add_mod((uint64_t)dst[k], src[k], q) : src[k]);
    } else {
        uint64_t *dst = reinterpret_cast<uint64_t *>(db);
        for (int k = threadIdx.x; k < N; k += blockDim.x) dst[k] = accumulate ? add_mod(dst[k], src[k], q) : src[k];
    }
}

int launch_copy_add(Ctx *c, const KsBatch &b, cudaStream_t st) {
    if (b.nct <= 0) return 0;
    dim3 g(2 * b.out_layout.nl, b.nct);
    k_copy_add<<<g, 256, 0, st>>>(b.in, b.in_off, b.in_nl, (unsigned char *)b.out, b.out_off, b.out_layout, c->N, c->lc, b.accumulate ? 1 : 0);
    SFG_LAUNCHED(c, "k_copy_add", st);
    return 0;
}

template <int NPER>
static int rotate_impl(Ctx *c, const KsBatch &b, const GaloisKey &key, cudaStream_t st) {
    const long long in_first = b.in_first, in_stride = b.in_stride;
    const int N = c->N, nl = b.level + 1, nt = nl + c->nP;
    const int T = N / NPER;
    BaseConv *ks, *md;
    uint64_t *pinv;
    if (ctx_get_ks_tables(c, b.level, &ks, &md, &pinv)) return -1;
    // 1. c2 = INTT(c1)
    LimbSel sel;
    sel.n = nl;
    for (int i = 0; i < nl; i++) sel.idx[i] = i;
    if (launch_ntt(c, b.in + in_first + (size_t)b.in_nl * N, (size_t)in_stride, b.c2, (size_t)nl * N, b.nct * nl, sel, true, st)) return -1;
    // 2. inner products with the switching key
    const size_t smem = (size_t)N * sizeof(uint64_t);
    SFG_CUDA(c, cudaFuncSetAttribute(k_ks_inner<NPER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SFG_CUDA(c, cudaFuncSetAttribute(k_ks_moddown<NPER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        dim3 g(nt, b.nct);
        k_ks_inner<NPER><<<g, T, smem, st>>>(b.in, b.in_off, b.in_nl, b.c2, key.key, ks, b.level, c->nQ, c->nP, c->logN, c->tw, c->lc, b.acc);
        SFG_LAUNCHED(c, "k_ks_inner", st);
    }
    // 3. INTT of the P limbs of acc: groups = (ct, comp), per-group limbs nQ..nQ+nP-1 located after the nl Q limbs
    LimbSel selp;
    selp.n = c->nP;
    for (int i = 0; i < c->nP; i++) selp.idx[i] = c->nQ + i;
    if (launch_ntt(c, b.acc + (size_t)nl * N, (size_t)nt * N, b.acc + (size_t)nl * N, (size_t)nt * N, b.nct * 2 * c->nP, selp, true, st)) return -1;
    // 4. mod-down, + c0, automorphism, store / accumulate
    {
        dim3 g(b.out_layout.nl, 2, b.nct);
        k_ks_moddown<NPER><<<g, T, smem, st>>>(b.in, b.in_off, b.in_nl, b.acc, md, pinv, key.perm, b.level, c->nQ, c->nP, c->logN, c->tw,
                                              c->lc, (unsigned char *)b.out, b.out_off, b.out_layout, b.accumulate ? 1 : 0);
        SFG_LAUNCHED(c, "k_ks_moddown", st);
    }
    return 0;
}

int launch_rotate(Ctx *c, const KsBatch &b, const GaloisKey &key, cudaStream_t st) {
    if (b.nct <= 0) return 0;
    if (c->logN > 14) SFG_FAIL(c, "fused key-switch kernels support logN <= 14 (got %d)", c->logN);
    if (c->logN < 6) SFG_FAIL(c, "logN >= 6 required");
    if (c->logN == 14) return rotate_impl<16>(c, b, key, st);
    if (c->logN >= 9) return rotate_impl<8>(c, b, key, st);
    return rotate_impl<2>(c, b, key, st);
}

}  // namespace sfg
