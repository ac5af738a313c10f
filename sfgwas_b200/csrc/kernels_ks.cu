// kernels_ks.cu -- rotation = hybrid key-switch + NTT-domain automorphism (K6/K7 of SURVEY 2.2).
//
// Semantics: crypto.RotateRightWithEvaluator (crypto/basics.go:201-210) -> Lattigo v2.1 Evaluator.RotateNew ->
// permuteNTT: (d0, d1) = switchKeysInPlace(c1, rtk); (d0 + c0, d1) permuted with PermuteNTTIndex(galEl)
// ([UNVERIFIED vs the fork]).  All residues leave every step canonical, so the result depends
// only on the mathematical value of each step -- including the float64 quotient estimate `v` of Lattigo's fast exact
// base conversion, which is reproduced operation by operation (IEEE division and addition in index order).
//
// A batch is any list of ciphertexts, each with its own Galois key (so all baby steps -- or many giant steps -- of a MatMult
// call go through ONE sequence of launches and fill the 148 SMs):
//   1. k_ntt2_inv    : c2 = INTT(c1)                                                         (kernels_ntt.cu)
//   2. k_ks_inner2   : per (ct, target modulus t in Q_level U P [, slice]): for every digit, reduce / base-convert the digit
//                      to t, forward NTT (register-tiled passes, ntt2.cuh), multiply with both key polynomials, accumulate in
//                      shared memory -> acc[ct][2][t] in TT order (the order the next kernels read with unit stride)
//   3. k_ntt2_inv    : INTT of the P limbs of acc (TT in, natural out)
//   4. k_ks_moddown2 : per (ct, component, Q limb): base-convert P -> q_l, NTT, (acc - ext) * P^-1, + c0, automorphism
//                      gather from shared memory, store or accumulate (the giant-step sum of gwas/matmult.go:1223-1227).
// Kernels are instantiated per arithmetic class of the target modulus (ntt2.cuh) and launched once per class present.
#include "kernels.h"
#include "ntt2.cuh"

namespace sfg {

// Lattigo fast exact base conversion for one coefficient: residues xs[k] (k < ns) -> target modulus t (canonical).
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

// The target-independent half of that conversion, once per (ciphertext, digit, coefficient) instead of once per target modulus (and
// per half-ring CTA):  y_k = x_k * (S/s_k)^-1 mod s_k  (written over c2)  and  v = floor(sum_k y_k / s_k)  (float64, the same operations
// in the same order).  The per-target half is  sum_k y_k * (S/s_k mod t) - v * (S mod t)  mod t  (base_conv_target).
__global__ void k_ks_bcprep(uint64_t *__restrict__ c2, const BaseConv *__restrict__ ks, int nl, int nt, int beta, int N,
                            const LimbConst *__restrict__ lcs, uint32_t *__restrict__ vq) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, slot = blockIdx.z;
    if (j >= N) return;
    const BaseConv &bc = ks[(size_t)i * nt + nl];  // any target outside the digit carries the source-side constants: the first P modulus
    uint64_t *x = c2 + (size_t)slot * nl * N + j;
    double vi = 0.0;
#pragma unroll 1
    for (int k = 0; k < bc.ns; k++) {
        const uint64_t sk = lcs[bc.src_limb[k]].q;
        const uint64_t y = mul_shoup(x[(size_t)bc.src_limb[k] * N], bc.sinv[k], bc.sinv_sh[k], sk);
        vi += __ddiv_rn((double)y, bc.sf[k]);
        x[(size_t)bc.src_limb[k] * N] = y;
    }
    vq[((size_t)slot * beta + i) * N + j] = (uint32_t)(uint64_t)vi;
}
// (deliberately NOT inlined: a radix-8 pass evaluates 16 conversions per thread, and inlined they are interleaved into ~2 KB of
// spills per thread; a call keeps one conversion's registers live at a time)
__device__ __noinline__ uint64_t base_conv_target(const BaseConv *bc, const uint64_t *y, size_t ystride, uint64_t v, uint64_t t) {
    uint64_t acc = 0;
#pragma unroll 1
    for (int k = 0; k < bc->ns; k++) acc = add_mod(acc, mul_shoup(y[(size_t)bc->src_limb[k] * ystride], bc->fac[k], bc->fac_sh[k], t), t);
    return sub_mod(acc, mul_shoup(v, bc->smod, bc->smod_sh, t), t);
}

// nP > 1: all digits of a ciphertext extended to all target moduli in one coalesced pass (coefficient domain, canonical), so that
// k_ks_inner2 -- where every (target, half-ring) CTA used to redo the conversion for both of its stage-0 inputs -- only loads them.
__global__ void k_ks_extd(const uint64_t *__restrict__ c2, const uint32_t *__restrict__ vq, const BaseConv *__restrict__ ks, int nl, int nt,
                          int nQ, int beta, int N, const LimbConst *__restrict__ lcs, uint64_t *__restrict__ extd) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y, slot = blockIdx.z;
    if (j >= N) return;
    const uint64_t *y = c2 + (size_t)slot * nl * N + j;
    const uint64_t v = vq[((size_t)slot * beta + i) * N + j];
    for (int tt = 0; tt < nt; tt++) {
        const BaseConv &bc = ks[(size_t)i * nt + tt];
        if (bc.ns == 0) continue;  // target inside the digit: the NTT-domain input limb is reused
        const uint64_t t = lcs[tt < nl ? tt : nQ + (tt - nl)].q;
        uint64_t a = 0;
#pragma unroll 1
        for (int k = 0; k < bc.ns; k++) a = add_mod(a, mul_shoup(y[(size_t)bc.src_limb[k] * N], bc.fac[k], bc.fac_sh[k], t), t);
        extd[(((size_t)slot * beta + i) * nt + tt) * N + j] = sub_mod(a, mul_shoup(v, bc.smod, bc.smod_sh, t), t);
    }
}

struct TgtSel {  // targets (or limbs) of one arithmetic class
    int n;
    int tt[kMaxLimbs];
};

// rings larger than one CTA's shared memory take the unfused key-switch (rotate_big); SFG_KS_UNFUSED=1 forces it for any ring so
// that the tests can run both implementations against each other on the same inputs
static bool ks_unfused(const Ctx *c) {
    static const bool forced = [] { const char *e = getenv("SFG_KS_UNFUSED"); return e && *e == '1'; }();
    return c->logN > 14 || forced;
}

// ---- Galois key conversion (once per key upload) ----------------------------------------------------------------------
// in : Lattigo SwitchingKey [beta][2][nQP][N] u64, NTT + Montgomery form.
// out: same shape, every polynomial in TT order; wide moduli keep the Montgomery u64, narrow moduli (q < 2^31) become
//      (k, floor(k 2^32 / q)) pairs with k the plain residue, for the 32-bit Shoup multiplication; FP64-class moduli the plain
//      residue as a double.
__global__ void k_key_convert(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, int N, int nQP, const LimbConst *__restrict__ lcs) {
    const int poly = blockIdx.y, limb = poly % nQP;
    const LimbConst lc = lcs[limb];
    const int kind = arith_kind(lc.q);
    const uint64_t *src = in + (size_t)poly * N;
    uint64_t *dst = out + (size_t)poly * N;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) {
        const uint64_t km = src[j];
        uint64_t o = km;
        if (kind == kArD) {
            o = (uint64_t)__double_as_longlong((double)(long long)mred(km, 1, lc));  // plain residue as an FP64 integer
        } else if (kind != kArW) {
            // 32-bit Montgomery form k * 2^32 mod q in the FIRST HALF of the polynomial's slot (4 bytes per coefficient: half the
            // L2 -> SM traffic of a (k, floor(k 2^32 / q)) Shoup pair; the product costs one more IMAD)
            const uint64_t k = mred(km, 1, lc);  // InvMForm
            reinterpret_cast<uint32_t *>(dst)[tt_index(j, N)] = (uint32_t)((k << 32) % lc.q);
            continue;
        }
        dst[tt_index(j, N)] = o;
    }
}
int launch_key_convert(Ctx *c, const uint64_t *in, uint64_t *out, cudaStream_t st) {
    if (ks_unfused(c)) {  // the unfused key-switch of large rings reads Lattigo's layout directly
        SFG_CUDA(c, cudaMemcpyAsync(out, in, (size_t)c->beta * 2 * c->nQP * c->N * 8, cudaMemcpyDefault, st));
        return 0;
    }
    dim3 g(std::max(1, c->N / 256), c->beta * 2 * c->nQP);
    k_key_convert<<<g, 256, 0, st>>>(in, out, c->N, c->nQP, c->lc);
    SFG_LAUNCHED(c, "k_key_convert", st);
    return 0;
}

// ---- 2. inner products with the switching key ---------------------------------------------------------------------------
// One CTA per (ciphertext, target modulus [, slice]); thread p owns the 16 consecutive NTT-domain coefficients 16p .. 16p+15 of
// both accumulators in REGISTERS for the whole digit loop (the last forward pass delivers exactly those).  A1: nP == 1 (every
// digit is a single modulus: plain reduction of the representative, Lattigo DecomposeAndSplit).
// LOGN > 0: the ring size is a compile-time constant (every index, stride and pass shape folds into immediates: ~30 % of the executed
// instructions of the generic version are address arithmetic); LOGN == 0: generic, ring size and pass plan from the arguments.
template <class A, int CS, bool A1, int LOGN, bool DM>
__global__ void __launch_bounds__(512, A::kMinBlocks)
k_ks_inner2(const uint64_t *__restrict__ in, const long long *__restrict__ in_off, int in_nl, const uint64_t *__restrict__ c2,
            const int *__restrict__ c2_slot, const uint64_t *const *__restrict__ keys, const BaseConv *__restrict__ ks, int level, int nQ,
            int nP, int logN_arg, PassPlan plan_arg, const TwTab *__restrict__ tabs, const LimbConst *__restrict__ lcs,
            uint64_t *__restrict__ accout, TgtSel sel, uint64_t *__restrict__ dout, const uint32_t *__restrict__ dlog_src,
            const uint32_t *__restrict__ vq /* nP > 1: c2 holds y_k and vq the quotient estimates (k_ks_bcprep) */,
            const uint64_t *__restrict__ extd /* nP > 1: the digits already extended to every target (k_ks_extd), or null */) {
    // DM (dout != nullptr): DIGIT mode.  The transforms NTT_t(Ext(digit_i(c1))) do not depend on the switching key, so rotations of the SAME
    // ciphertext by different amounts (the d-1 baby steps of every A[i][bi]) share them: the kernel then runs once per distinct input
    // (ct = slot, c2_slot == keys == nullptr) and stores the canonical transforms D[slot][i][tt] in TT order instead of multiplying
    // them with a key; k_ks_macd does the key products per rotation.  Bit-identical: the same values enter the same modular sums.
    using T = typename A::T;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int logN = LOGN ? LOGN : logN_arg;
    const PassPlan plan = LOGN ? make_pass_plan(LOGN - CS - kLastR) : plan_arg;
    const int N = 1 << logN, logS = logN - CS, S = 1 << logS, nl = level + 1, nt = nl + nP, nQP = nQ + nP;
    const int alpha = nP, beta = (nl + alpha - 1) / alpha;
    const int sl = blockIdx.x & ((1 << CS) - 1), tt = sel.tt[blockIdx.x >> CS], ct = blockIdx.y;
    const int tgt = tt < nl ? tt : nQ + (tt - nl);
    T *s = reinterpret_cast<T *>(smraw);
    const LimbConst lc = lcs[tgt];
    const typename A::C c = A::make(lc);
    const TwTab tab = tabs[tgt];
    const uint64_t *c1 = in + in_off[ct] + (size_t)in_nl * N;  // second polynomial of the input ct (NTT domain)
    const uint64_t *c2ct = c2 + (size_t)(c2_slot ? c2_slot[ct] : ct) * nl * N;  // its INTT, coefficient domain
    const uint64_t *key = keys ? keys[ct] : nullptr;
    const int gbase = sl << logS;                               // global index of local coefficient 0
    const int P = (gbase >> kLastR) + threadIdx.x;              // global index of this thread's group of 16
    const bool active = threadIdx.x < (S >> kLastR);
    const int NP = N >> kLastR;

    // accumulators: registers for the 32-bit classes; thread-private shared-memory slots for the 64-bit class (register budget)
    constexpr bool ACC_SMEM = true;  // thread-private shared-memory slots: keeps every class within its register budget
    constexpr int NREG = ACC_SMEM ? 1 : kLastE;
    T a0[NREG], a1[NREG];
    T *s0a = s + ntt_smem_elems(S), *s1a = s0a + ntt_smem_elems(S);
    // twiddles of the passes before the last one: staged once per CTA (one target modulus), read by every digit's transform
    typename A::TW *tw_s = reinterpret_cast<typename A::TW *>(s1a + ntt_smem_elems(S));
    ntt_stage_twiddles<A>(tw_s, tab, logN);
    const int abase = kLastE * threadIdx.x;
#pragma unroll
    for (int k = 0; k < kLastE; k++) {
        if constexpr (ACC_SMEM) {
            if (active) s0a[sidx<sizeof(T)>(abase + k)] = s1a[sidx<sizeof(T)>(abase + k)] = 0;
        } else {
            a0[k] = a1[k] = 0;
        }
    }
    __syncthreads();  // the staged twiddles are visible to every warp before the first transform

    for (int i = 0; i < beta; i++) {
        const BaseConv &bc = ks[(size_t)i * nt + tt];
        const uint64_t *k0 = key + ((size_t)(i * 2 + 0) * nQP + tgt) * N + P;
        const uint64_t *k1 = key + ((size_t)(i * 2 + 1) * nQP + tgt) * N + P;
        // multiply coefficient 16 P + k with the two key polynomials (TT order: unit stride across lanes) and accumulate
        auto mac = [&](int, T v, int k) {
            if constexpr (DM) {
                dout[(((size_t)ct * beta + i) * nt + tt) * N + P + (size_t)k * NP] = (uint64_t)A::canon(v, c);
                return;
            }
            if constexpr (A::kKind == kArD) {  // lazy FP64 accumulation: every product is in (-q, q)
                const int sj = sidx<sizeof(T)>(abase + k);
                s0a[sj] += A::mul_lazy(v, __ldg(reinterpret_cast<const double *>(k0) + (size_t)k * NP), c);
                s1a[sj] += A::mul_lazy(v, __ldg(reinterpret_cast<const double *>(k1) + (size_t)k * NP), c);
            } else if constexpr (A::kKind == kArW) {
                const int sj = sidx<sizeof(T)>(abase + k);
                s0a[sj] = add_mod(s0a[sj], mred(v, __ldg(k0 + (size_t)k * NP), lc), lc.q);
                s1a[sj] = add_mod(s1a[sj], mred(v, __ldg(k1 + (size_t)k * NP), lc), lc.q);
            } else {
                const int sj = sidx<sizeof(T)>(abase + k);
                const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t *>(k0 - P) + P + (size_t)k * NP);
                const uint32_t w1 = __ldg(reinterpret_cast<const uint32_t *>(k1 - P) + P + (size_t)k * NP);
                uint32_t p0 = ArN30::mul_mont(v, w0, c), p1 = ArN30::mul_mont(v, w1, c);
                p0 = min(p0, p0 - c.q) + s0a[sj];
                p1 = min(p1, p1 - c.q) + s1a[sj];
                s0a[sj] = min(p0, p0 - c.q);
                s1a[sj] = min(p1, p1 - c.q);
            }
        };
        const int ns = bc.ns;
        if (ns == 0) {  // target lies inside the digit: reuse the NTT-domain input limb (decomposeAndSplitNTT)
            const uint64_t *x = c1 + (size_t)tt * N + gbase;
            for (int j0 = threadIdx.x; j0 < S; j0 += 8 * blockDim.x) {  // 8 independent loads in flight per thread
                uint64_t xv[8];
#pragma unroll
                for (int u = 0; u < 8; u++) xv[u] = (j0 + u * (int)blockDim.x < S) ? __ldg(x + j0 + u * blockDim.x) : 0;
#pragma unroll
                for (int u = 0; u < 8; u++)
                    if (j0 + u * (int)blockDim.x < S) s[sidx<sizeof(T)>(j0 + u * blockDim.x)] = A::from_canon(xv[u], c);
            }
            __syncthreads();
            if (active) {
                if constexpr (A::kKind == kArN30 || A::kKind == kArN31) {  // 64-register classes: bound the loads in flight
#pragma unroll 4
                    for (int k = 0; k < kLastE; k++) mac(0, s[sidx<sizeof(T)>(kLastE * threadIdx.x + k)], k);
                } else {
#pragma unroll
                    for (int k = 0; k < kLastE; k++) mac(0, s[sidx<sizeof(T)>(kLastE * threadIdx.x + k)], k);
                }
            }
        } else {
            // coefficient-domain value of global coefficient g of this digit, reduced / base-converted to the target modulus
            auto digit = [&](int g) -> T {
                if constexpr (A1) {
                    return A::load_u64(c2ct[(size_t)bc.src_limb[0] * N + g], c);  // DecomposeAndSplit, single modulus
                } else {
                    const size_t slot = c2_slot ? (size_t)c2_slot[ct] : (size_t)ct;
                    if (extd) return A::load_u64(extd[((slot * beta + i) * nt + tt) * N + g], c);
                    return A::load_u64(base_conv_target(&bc, c2ct + g, (size_t)N, vq[(slot * beta + i) * N + g], lc.q), c);
                }
            };
            auto ld0 = [&](int j, int) -> T {
                if constexpr (CS == 0) {
                    return digit(j);
                } else {  // the ring is cut into 2 slices: stage 0 is evaluated by both CTAs, each keeps its half
                    T x = digit(j), y = digit(j + S);
                    A::fwd(x, y, __ldg(reinterpret_cast<const typename A::TW *>(tab.fwd) + 1), c);
                    return sl ? y : x;
                }
            };
            ntt_forward<A>(s, logN, logS, sl, plan, tab, c, ld0, mac, tw_s);
        }
        __syncthreads();
    }
    if (!DM && dlog_src && tt < nl) {
        // giant-step sums: Q limbs leave in discrete-log order (position q holds the coefficient dlog_src[q]; a half-ring slice is
        // exactly one sign class, i.e. a contiguous half), so that k_md_accum can add rotated polynomials as shifted contiguous reads
        // (the digit loop ended with a barrier: every accumulator slot is final and visible)
        uint64_t *o0 = accout + ((size_t)(ct * 2 + 0) * nt + tt) * N + gbase, *o1 = accout + ((size_t)(ct * 2 + 1) * nt + tt) * N + gbase;
        for (int j = threadIdx.x; j < S; j += blockDim.x) {
            const int sj = sidx<sizeof(T)>((int)__ldg(dlog_src + gbase + j) - gbase);
            if constexpr (A::kKind == kArD) {
                o0[j] = (uint64_t)A::canon(s0a[sj], c);
                o1[j] = (uint64_t)A::canon(s1a[sj], c);
            } else {
                o0[j] = s0a[sj];
                o1[j] = s1a[sj];
            }
        }
    } else if (active && !DM) {
        uint64_t *o0 = accout + ((size_t)(ct * 2 + 0) * nt + tt) * N + P;
        uint64_t *o1 = accout + ((size_t)(ct * 2 + 1) * nt + tt) * N + P;
#pragma unroll
        for (int k = 0; k < kLastE; k++) {
            if constexpr (A::kKind == kArD) {
                o0[(size_t)k * NP] = (uint64_t)A::canon(s0a[sidx<sizeof(T)>(abase + k)], c);
                o1[(size_t)k * NP] = (uint64_t)A::canon(s1a[sidx<sizeof(T)>(abase + k)], c);
            } else if constexpr (ACC_SMEM) {
                o0[(size_t)k * NP] = s0a[sidx<sizeof(T)>(abase + k)];
                o1[(size_t)k * NP] = s1a[sidx<sizeof(T)>(abase + k)];
            } else {
                o0[(size_t)k * NP] = a0[k];
                o1[(size_t)k * NP] = a1[k];
            }
        }
    }
}

// ---- 4. mod-down, + c0, automorphism, store / accumulate ---------------------------------------------------------------------
template <class A, bool A1, int LOGN>
__global__ void __launch_bounds__(512, A::kMinBlocks)
k_ks_moddown2(const uint64_t *__restrict__ in, const long long *__restrict__ in_off, int in_nl, const uint64_t *__restrict__ acc,
              const BaseConv *__restrict__ md, const uint64_t *__restrict__ pinv, const uint32_t *const *__restrict__ perms, int level,
              int nQ, int nP, int logN_arg, PassPlan plan_arg, const TwTab *__restrict__ tabs, const LimbConst *__restrict__ lcs,
              unsigned char *__restrict__ out, const long long *__restrict__ out_off, PolyLayout olay, int accumulate, TgtSel sel,
              const uint64_t *__restrict__ extP /* nP > 1: Ext(accP) per (ct, comp, limb < olay.nl) from k_md_extp, else null */) {
    using T = typename A::T;
    extern __shared__ __align__(16) unsigned char smraw[];
    T *s = reinterpret_cast<T *>(smraw);
    const int logN = LOGN ? LOGN : logN_arg;
    const PassPlan plan = LOGN ? make_pass_plan(LOGN - kLastR) : plan_arg;
    const int N = 1 << logN, nl = level + 1, nt = nl + nP;
    const int l = sel.tt[blockIdx.x], comp = blockIdx.y, ct = blockIdx.z;
    const LimbConst lc = lcs[l];
    const typename A::C c = A::make(lc);
    const TwTab tab = tabs[l];
    const BaseConv &bc = md[l];
    const uint64_t *accP = acc + ((size_t)(ct * 2 + comp) * nt + nl) * N;  // P limbs, coefficient domain, natural order
    const uint64_t *accQ = acc + ((size_t)(ct * 2 + comp) * nt + l) * N;   // Q limb, NTT domain, TT order
    const uint64_t pi = pinv[2 * l], pish = pinv[2 * l + 1];
    const uint32_t *perm = perms[ct];

    auto ld0 = [&](int j, int) -> T {
        if constexpr (A1) {
            return A::load_u64(accP[j], c);
        } else {
            if (extP) return A::load_u64(extP[((size_t)(ct * 2 + comp) * olay.nl + l) * N + j], c);
            uint64_t xs[kMaxAlpha];
            for (int k = 0; k < nP; k++) xs[k] = accP[(size_t)k * N + j];
            return A::load_u64(base_conv_coeff(bc, xs, lcs, lc.q), c);
        }
    };
    auto fin = [&](int j, T v, int) {
        const uint64_t e = A::canon(v, c);
        const uint64_t r = mul_shoup(sub_mod(accQ[tt_index(j, N)], e, lc.q), pi, pish, lc.q);
        s[sidx<sizeof(T)>(j)] = (T)r;
    };
    ntt_forward<A>(s, logN, logN, 0, plan, tab, c, ld0, fin);
    __syncthreads();
    if (comp == 0) {
        const uint64_t *c0 = in + in_off[ct] + (size_t)l * N;
        for (int j = threadIdx.x; j < N; j += blockDim.x) {
            const int sj = sidx<sizeof(T)>(j);
            s[sj] = (T)add_mod((uint64_t)s[sj], c0[j], lc.q);
        }
        __syncthreads();
    }
    unsigned char *ob = out + out_off[ct] + (size_t)comp * olay.bytes + olay.off[l];
    if (olay.es[l] == 4) {
        uint32_t *o = reinterpret_cast<uint32_t *>(ob);
        for (int k = threadIdx.x; k < N; k += blockDim.x) {
            uint64_t v = s[sidx<sizeof(T)>(perm[k])];  // PermuteNTTWithIndexLvl: out[k] = in[index[k]]
            if (accumulate) v = add_mod(v, (uint64_t)o[k], lc.q);
            o[k] = (uint32_t)v;
        }
    } else {
        uint64_t *o = reinterpret_cast<uint64_t *>(ob);
        for (int k = threadIdx.x; k < N; k += blockDim.x) {
            uint64_t v = s[sidx<sizeof(T)>(perm[k])];
            if (accumulate) v = add_mod(v, o[k], lc.q);
            o[k] = v;
        }
    }
}

// ---- 4'. giant-step sum: mod-down hoisted out of the sum over the giant steps ------------------------------------------------
// out[o] = sum_a perm_a( (accQ_a - NTT(Ext(accP_a))) * P^-1 + c0_a )   over the entries a*nout + o of one output ciphertext.
// Every operation is exact arithmetic mod q_l, the NTT is linear, and the NTT-domain permutation perm_a is the automorphism
// X -> X^galEl_a, so   sum_a perm_a(NTT(e_a)) = NTT( sum_a sigma_a(e_a) ):  ONE forward transform per output polynomial instead
// of one per entry.  k_md_accum gathers  S1 = sum perm(accQ),  C0 = sum perm(c0)  (NTT domain) and  E = sum sigma(Ext(accP))
// (coefficient domain: E[j'] += +-e[u mod N], u = j' * galEl^-1 mod 2N, minus iff u >= N);  k_md_final adds
// (S1 - NTT(E)) * P^-1 + C0 into the output.  The result is bit-identical to per-entry mod-down (canonical residues).
// Sources are staged whole (N x 8 bytes) by cp.async.bulk into a ring of three shared-memory buffers, prefetched two stages
// ahead of the gathers, so the kernel streams acc / cv at memory speed instead of paying one load latency per source.
__device__ __forceinline__ uint32_t md_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void md_bulk(void *smem, const void *gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(md_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(md_smem_u32(smem)), "l"(gmem),
                 "r"(bytes), "r"(md_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void md_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "MD_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MD_DONE;\n\t"
        "bra MD_WAIT;\n\t"
        "MD_DONE:\n\t}\n" ::"r"(md_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// nP > 1: the exact base conversion of the P limbs to every Q modulus, once per (entry, component, coefficient) with unit stride --
// k_md_accum then stages ext[l] like a single P limb.  (Evaluated inside k_md_accum it ran per limb CTA, at permuted -- uncoalesced --
// positions, with its float64 divisions in the innermost loop: 52 of the 285 ms of a logN = 14 step.)
__global__ void k_md_extp(const uint64_t *__restrict__ acc, const BaseConv *__restrict__ md, int nl, int nt, int nP, int L, int N,
                          const LimbConst *__restrict__ lcs, uint64_t *__restrict__ ext) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;  // p = ct * 2 + comp
    if (j >= N) return;
    const uint64_t *accP = acc + ((size_t)p * nt + nl) * N + j;
    const BaseConv &b0 = md[0];  // the source-side constants are the same for every target
    uint64_t ys[kMaxAlpha];
    double vi = 0.0;
#pragma unroll 1
    for (int k = 0; k < nP; k++) {
        const uint64_t sk = lcs[b0.src_limb[k]].q;
        ys[k] = mul_shoup(accP[(size_t)k * N], b0.sinv[k], b0.sinv_sh[k], sk);
        vi += __ddiv_rn((double)ys[k], b0.sf[k]);
    }
    const uint64_t v = (uint64_t)vi;
    for (int l = 0; l < L; l++) {
        const BaseConv &bc = md[l];
        const uint64_t t = lcs[l].q;
        uint64_t a = 0;
#pragma unroll 1
        for (int k = 0; k < nP; k++) a = add_mod(a, mul_shoup(ys[k], bc.fac[k], bc.fac_sh[k], t), t);
        ext[((size_t)p * L + l) * N + j] = sub_mod(a, mul_shoup(v, bc.smod, bc.smod_sh, t), t);
    }
}

template <typename T, int LOGN>  // accumulator type: uint32_t for q < 2^31, uint64_t otherwise; LOGN > 0: compile-time ring size
__global__ void __launch_bounds__(512, 1)
k_md_accum(const uint64_t *__restrict__ in, const long long *__restrict__ in_off, int in_nl, const uint64_t *__restrict__ acc,
           const BaseConv *__restrict__ md, const uint32_t *const *__restrict__ perms, const uint32_t *__restrict__ ginv, int level, int nQ,
           int nP, int logN_arg, const LimbConst *__restrict__ lcs, int nout, int nacc, uint64_t *__restrict__ S1o, uint64_t *__restrict__ C0o,
           uint64_t *__restrict__ Eo, int L, int first, TgtSel sel, int NBUF /* ring depth: 3, or what fits (2^14-coefficient rings: 1) */,
           const uint64_t *__restrict__ extP /* nP > 1: Ext(accP) per (entry, comp, limb), k_md_extp */) {
    extern __shared__ __align__(128) uint64_t sst[];
    const int logN = LOGN ? LOGN : logN_arg;
    const int N = 1 << logN, nl = level + 1, nt = nl + nP;
    uint64_t *bars = sst + (size_t)NBUF * N;
    // rings larger than 2^13 are cut into coefficient ranges of 2^13 owned by different CTAs (the sources are staged whole)
    const int nsplit = N > 8192 ? N / 8192 : 1;
    const int l = sel.tt[blockIdx.x / nsplit], comp = blockIdx.y, o = blockIdx.z, tid = threadIdx.x;
    const LimbConst lc = lcs[l];
    const T q = (T)lc.q;
    constexpr int E16 = 16;
    const int per = (N / nsplit) / E16;  // threads that own coefficients (<= blockDim.x)
    T s1[E16], c0s[E16], es[E16];
    const int kbase = (blockIdx.x % nsplit) * (N / nsplit);
    const size_t obase = (((size_t)o * 2 + comp) * L + l) * N + kbase;
    auto addm = [&](T a, T b2) { const T r = a + b2; return r >= q ? r - q : r; };
    auto subm = [&](T a, T b2) { return a >= b2 ? a - b2 : a + q - b2; };
#pragma unroll
    for (int m = 0; m < E16; m++) {
        const int k = tid + per * m;
        const bool ld = !first && tid < per;
        s1[m] = ld ? (T)S1o[obase + k] : 0;
        c0s[m] = (ld && comp == 0) ? (T)C0o[obase + k] : 0;
        es[m] = ld ? (T)Eo[obase + k] : 0;
    }
    // accQ is in discrete-log order (k_ks_inner2): perm_a is a cyclic shift inside each half, read straight from global memory.
    // stage sequence of the ring: per entry  [c0 if comp == 0] [accP if nP == 1]
    const bool stageP = nP == 1 || extP != nullptr;
    const int spe = (comp == 0 ? 1 : 0) + (stageP ? 1 : 0);
    const int nstage = nacc * spe;
    const uint32_t bytes = (uint32_t)N * 8;
    auto src_of = [&](int st) -> const uint64_t * {
        const int a = st / spe, w = st % spe, ct = a * nout + o;
        if (w == 0 && comp == 0) return in + in_off[ct] + (size_t)l * N;               // c0: NTT domain, natural order
        if (extP) return extP + ((size_t)(ct * 2 + comp) * L + l) * N;                 // Ext(accP) mod q_l: coefficient domain, natural order
        return acc + ((size_t)(ct * 2 + comp) * nt + nl) * N;                          // accP: coefficient domain, natural order
    };
    if (tid == 0) {
        for (int i = 0; i < NBUF; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(md_smem_u32(&bars[i])), "r"(1));
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        for (int st = 0; st < NBUF && st < nstage; st++) md_bulk(sst + (size_t)st * N, src_of(st), bytes, &bars[st]);
    }
    __syncthreads();
    int st = 0;
    auto next_stage = [&]() -> const uint64_t * {  // wait for stage st; returns its buffer
        md_wait(&bars[st % NBUF], (uint32_t)(st / NBUF) & 1u);
        return sst + (size_t)(st % NBUF) * N;
    };
    auto release_stage = [&]() {  // all threads are done with stage st: refill its buffer with stage st + NBUF
        __syncthreads();
        if (tid == 0 && st + NBUF < nstage) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            md_bulk(sst + (size_t)(st % NBUF) * N, src_of(st + NBUF), bytes, &bars[st % NBUF]);
        }
        st++;
    };
    for (int a = 0; a < nacc; a++) {
        const int ct = a * nout + o;
        const uint32_t *perm = perms[ct];
        uint32_t pk[E16];
#pragma unroll
        for (int m = 0; m < E16; m++) pk[m] = tid < per ? __ldg(perm + kbase + tid + per * m) : 0;
        const uint32_t gw = ginv[ct];  // galEl^-1 mod 2N in the low 18 bits, the rotation amount r (galEl = 5^r) above
        if (tid < per) {
            const uint64_t *aq = acc + ((size_t)(ct * 2 + comp) * nt + l) * N;
            const uint32_t r = gw >> 18, n2m = (uint32_t)(N >> 1) - 1;
            T q1[E16];
#pragma unroll
            for (int m = 0; m < E16; m++) {
                const uint32_t k = (uint32_t)(kbase + tid + per * m);
                q1[m] = (T)__ldg(aq + ((k & ~n2m) | ((k + r) & n2m)));
            }
#pragma unroll
            for (int m = 0; m < E16; m++) s1[m] = addm(s1[m], q1[m]);
        }
        if (comp == 0) {
            const uint64_t *sc = next_stage();
            if (tid < per) {
#pragma unroll
                for (int m = 0; m < E16; m++) c0s[m] = addm(c0s[m], (T)sc[pk[m]]);
            }
            release_stage();
        }
        const uint32_t gi = gw & 0x3FFFFu;
        if (stageP) {
            const uint64_t *sp = next_stage();
            if (tid < per) {
#pragma unroll
                for (int m = 0; m < E16; m++) {
                    const uint32_t u = ((uint32_t)(kbase + tid + per * m) * gi) & (uint32_t)(2 * N - 1);
                    const T e = (T)bred_add(sp[u & (N - 1)], lc);
                    es[m] = u < (uint32_t)N ? addm(es[m], e) : subm(es[m], e);
                }
            }
            release_stage();
        } else if (tid < per) {
            const uint64_t *accP = acc + ((size_t)(ct * 2 + comp) * nt + nl) * N;
            const BaseConv &bc = md[l];
            for (int m = 0; m < E16; m++) {
                const uint32_t u = ((uint32_t)(kbase + tid + per * m) * gi) & (uint32_t)(2 * N - 1);
                uint64_t xs[kMaxAlpha];
                for (int k = 0; k < nP; k++) xs[k] = accP[(size_t)k * N + (u & (N - 1))];
                const T e = (T)base_conv_coeff(bc, xs, lcs, lc.q);
                es[m] = u < (uint32_t)N ? addm(es[m], e) : subm(es[m], e);
            }
        }
    }
    if (tid < per) {
#pragma unroll
        for (int m = 0; m < E16; m++) {
            const int k = tid + per * m;
            S1o[obase + k] = s1[m];
            if (comp == 0) C0o[obase + k] = c0s[m];
            Eo[obase + k] = es[m];
        }
    }
}

template <class A>
__global__ void __launch_bounds__(512, A::kMinBlocks)
k_md_final(const uint64_t *__restrict__ S1, const uint64_t *__restrict__ C0, const uint64_t *__restrict__ E, const uint64_t *__restrict__ pinv,
           int logN, PassPlan plan, const TwTab *__restrict__ tabs, const LimbConst *__restrict__ lcs, unsigned char *__restrict__ out,
           const long long *__restrict__ out_off, PolyLayout olay, int L, TgtSel sel, const uint32_t *__restrict__ dlog_pos) {
    using T = typename A::T;
    extern __shared__ __align__(16) unsigned char smraw[];
    T *s = reinterpret_cast<T *>(smraw);
    const int N = 1 << logN;
    const int l = sel.tt[blockIdx.x], comp = blockIdx.y, o = blockIdx.z;
    const LimbConst lc = lcs[l];
    const typename A::C c = A::make(lc);
    const size_t obase = (((size_t)o * 2 + comp) * L + l) * N;
    const uint64_t pi = pinv[2 * l], pish = pinv[2 * l + 1];
    uint64_t *ob = reinterpret_cast<uint64_t *>(out + out_off[o] + (size_t)comp * olay.bytes + olay.off[l]);  // u64 output layout
    auto ld0 = [&](int j, int) -> T { return A::load_u64(E[obase + j], c); };
    auto fin = [&](int j, T v, int) {
        const uint64_t e = A::canon(v, c);
        uint64_t r = mul_shoup(sub_mod(S1[obase + __ldg(dlog_pos + j)], e, lc.q), pi, pish, lc.q);  // S1 is kept in discrete-log order
        if (comp == 0) r = add_mod(r, C0[obase + j], lc.q);
        ob[j] = add_mod(ob[j], r, lc.q);
    };
    ntt_forward<A>(s, logN, logN, 0, plan, tabs[l], c, ld0, fin);
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

// ---- 2'. key products over shared digit transforms: acc[ct][comp][tt] = sum_i D[slot(ct)][i][tt] * key_ct[i][comp][tt]  (TT order) -------
// Pure streaming (unit stride in the TT position): per output pair beta transforms and 2 beta key words are read once.  One launch per
// arithmetic class of the target modulus: 32-bit Montgomery products for the narrow classes, exact FP64 two-products for ArD, u64
// Montgomery for ArW (a single generic u64 kernel spent 600 instructions per thread, most of them 64-bit mul.hi sequences).
template <int KIND>
__global__ void __launch_bounds__(256)
k_ks_macd(const uint64_t *__restrict__ D, const int *__restrict__ c2_slot, const uint64_t *const *__restrict__ keys, int level, int nQ,
          int nP, int N, const LimbConst *__restrict__ lcs, uint64_t *__restrict__ accout, TgtSel sel) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x, tt = sel.tt[blockIdx.y], ct = blockIdx.z;
    if (p >= N) return;
    const int nl = level + 1, nt = nl + nP, nQP = nQ + nP, beta = (nl + nP - 1) / nP;
    const int tgt = tt < nl ? tt : nQ + (tt - nl);
    const LimbConst lc = lcs[tgt];
    const uint64_t *d = D + ((size_t)c2_slot[ct] * beta * nt + tt) * N + p;
    const uint64_t *key = keys[ct] + (size_t)tgt * N;
    uint64_t *o0 = accout + ((size_t)(ct * 2 + 0) * nt + tt) * N + p, *o1 = accout + ((size_t)(ct * 2 + 1) * nt + tt) * N + p;
    if constexpr (KIND == kArN30 || KIND == kArN31) {
        const ArN30::C c = ArN30::make(lc);
        uint32_t a0 = 0, a1 = 0;
        for (int i = 0; i < beta; i++) {
            const uint32_t v = (uint32_t)__ldg(d + (size_t)i * nt * N);
            const uint32_t k0 = __ldg(reinterpret_cast<const uint32_t *>(key + (size_t)(i * 2 + 0) * nQP * N) + p);  // 32-bit Montgomery form,
            const uint32_t k1 = __ldg(reinterpret_cast<const uint32_t *>(key + (size_t)(i * 2 + 1) * nQP * N) + p);  // first half of the slot
            uint32_t r0 = ArN30::mul_mont(v, k0, c), r1 = ArN30::mul_mont(v, k1, c);
            r0 = min(r0, r0 - c.q) + a0;
            r1 = min(r1, r1 - c.q) + a1;
            a0 = min(r0, r0 - c.q);
            a1 = min(r1, r1 - c.q);
        }
        *o0 = a0;
        *o1 = a1;
    } else if constexpr (KIND == kArD) {
        const ArD::C c = ArD::make(lc);
        double a0 = 0.0, a1 = 0.0;  // every product is in (-q, q): the sums stay exact integers far below 2^53
        for (int i = 0; i < beta; i++) {
            const double v = (double)(long long)__ldg(d + (size_t)i * nt * N);
            a0 += ArD::mul_lazy(v, __ldg(reinterpret_cast<const double *>(key + (size_t)(i * 2 + 0) * nQP * N) + p), c);  // plain residue as FP64
            a1 += ArD::mul_lazy(v, __ldg(reinterpret_cast<const double *>(key + (size_t)(i * 2 + 1) * nQP * N) + p), c);
        }
        *o0 = (uint64_t)ArD::canon(a0, c);
        *o1 = (uint64_t)ArD::canon(a1, c);
    } else {
        uint64_t a0 = 0, a1 = 0;
        for (int i = 0; i < beta; i++) {
            const uint64_t v = __ldg(d + (size_t)i * nt * N);
            a0 = add_mod(a0, mred(v, __ldg(key + (size_t)(i * 2 + 0) * nQP * N + p), lc), lc.q);  // Montgomery form, as uploaded
            a1 = add_mod(a1, mred(v, __ldg(key + (size_t)(i * 2 + 1) * nQP * N + p), lc), lc.q);
        }
        *o0 = a0;
        *o1 = a1;
    }
}

static int ntt_threads(int S) { return std::min(512, std::max(32, S >> kLastR)); }

template <class A, int CS, bool A1>
static int inner_launch2(Ctx *c, const KsBatch &b, BaseConv *ks, const TgtSel &sel, cudaStream_t st) {
    const int logN = c->logN, logS = logN - CS, S = 1 << logS;
    const PassPlan plan = make_pass_plan(logS - kLastR);
    const size_t smem = 3 * ntt_smem_elems(S) * sizeof(typename A::T) + (size_t)ntt_mid_twiddles(logN) * sizeof(typename A::TW);
    dim3 g(sel.n << CS, b.nct);
    auto go = [&](auto kern) -> int {
        SFG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<g, ntt_threads(S), smem, st>>>(b.in, b.in_off, b.in_nl, b.c2, b.c2_slot, b.keys, ks, b.level, c->nQ, c->nP, logN, plan, c->tw2, c->lc,
                                              b.acc, sel, b.dout, b.acc_dlog ? c->dlog_src : nullptr, b.vq, b.extd);
        SFG_LAUNCHED(c, "k_ks_inner2", st);
        return 0;
    };
    // rings of the reference's parameter sets are compiled with constant shapes; anything else takes the generic kernel
    if (b.dout) {  // digit mode (shared decomposition of the baby steps): the generic-shape kernel is enough for logN != 13
        if constexpr (CS == 0) {
            if (logN == 13) return go(k_ks_inner2<A, CS, A1, 13, true>);
        }
        return go(k_ks_inner2<A, CS, A1, 0, true>);
    }
    if constexpr (CS == 0) {
        if (logN == 13) return go(k_ks_inner2<A, CS, A1, 13, false>);
    } else {
        if (logN == 14) return go(k_ks_inner2<A, CS, A1, 14, false>);
    }
    return go(k_ks_inner2<A, CS, A1, 0, false>);
}
template <class A>
static int inner_launch(Ctx *c, const KsBatch &b, BaseConv *ks, const TgtSel &sel, cudaStream_t st) {
    if (sel.n == 0) return 0;
    const bool a1 = c->nP == 1;
    if (c->logN > 13) return a1 ? inner_launch2<A, 1, true>(c, b, ks, sel, st) : inner_launch2<A, 1, false>(c, b, ks, sel, st);
    return a1 ? inner_launch2<A, 0, true>(c, b, ks, sel, st) : inner_launch2<A, 0, false>(c, b, ks, sel, st);
}

template <class A, bool A1>
static int moddown_launch2(Ctx *c, const KsBatch &b, BaseConv *md, uint64_t *pinv, const TgtSel &sel, const uint64_t *extP, cudaStream_t st) {
    const int logN = c->logN, N = c->N;
    const PassPlan plan = make_pass_plan(logN - kLastR);
    const size_t smem = ntt_smem_elems(N) * sizeof(typename A::T);
    dim3 g(sel.n, 2, b.nct);
    auto go = [&](auto kern) -> int {
        SFG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<g, ntt_threads(N), smem, st>>>(b.in, b.in_off, b.in_nl, b.acc, md, pinv, b.perms, b.level, c->nQ, c->nP, logN, plan, c->tw2, c->lc,
                                              (unsigned char *)b.out, b.out_off, b.out_layout, b.accumulate ? 1 : 0, sel, extP);
        SFG_LAUNCHED(c, "k_ks_moddown2", st);
        return 0;
    };
    if (logN == 13) return go(k_ks_moddown2<A, A1, 13>);
    if (logN == 14) return go(k_ks_moddown2<A, A1, 14>);
    return go(k_ks_moddown2<A, A1, 0>);
}
template <class A>
static int moddown_launch(Ctx *c, const KsBatch &b, BaseConv *md, uint64_t *pinv, const TgtSel &sel, const uint64_t *extP, cudaStream_t st) {
    if (sel.n == 0) return 0;
    return c->nP == 1 ? moddown_launch2<A, true>(c, b, md, pinv, sel, extP, st) : moddown_launch2<A, false>(c, b, md, pinv, sel, extP, st);
}

static int inner_all(Ctx *c, const KsBatch &b, BaseConv *ks, cudaStream_t st) {
    const int nl = b.level + 1, nt = nl + c->nP;
    TgtSel ts[kNumArith] = {{0, {}}, {0, {}}, {0, {}}, {0, {}}};
    for (int tt = 0; tt < nt; tt++) {
        const int tgt = tt < nl ? tt : c->nQ + (tt - nl);
        TgtSel &t = ts[arith_kind(c->mod[tgt])];
        t.tt[t.n++] = tt;
    }
    if (inner_launch<ArW>(c, b, ks, ts[kArW], st) || inner_launch<ArD>(c, b, ks, ts[kArD], st) || inner_launch<ArN30>(c, b, ks, ts[kArN30], st) || inner_launch<ArN31>(c, b, ks, ts[kArN31], st))
        return -1;
    return 0;
}

// D != nullptr: the digit transforms of the batch's distinct inputs are already in D (shared-decomposition path)
static int rotate_chunk(Ctx *c, const KsBatch &b, BaseConv *ks, BaseConv *md, uint64_t *pinv, cudaStream_t st, bool moddown = true,
                        const uint64_t *D = nullptr) {
    const int N = c->N, nl = b.level + 1, nt = nl + c->nP;
    // 2. inner products with the switching keys, one launch per arithmetic class of the target modulus
    if (D) {
        TgtSel ts[kNumArith] = {{0, {}}, {0, {}}, {0, {}}, {0, {}}};
        for (int tt = 0; tt < nt; tt++) {
            TgtSel &t = ts[arith_kind(c->mod[tt < nl ? tt : c->nQ + (tt - nl)])];
            t.tt[t.n++] = tt;
        }
        const unsigned gx = (unsigned)((N + 255) / 256);
        if (ts[kArW].n) k_ks_macd<kArW><<<dim3(gx, ts[kArW].n, b.nct), 256, 0, st>>>(D, b.c2_slot, b.keys, b.level, c->nQ, c->nP, N, c->lc, b.acc, ts[kArW]);
        if (ts[kArD].n) k_ks_macd<kArD><<<dim3(gx, ts[kArD].n, b.nct), 256, 0, st>>>(D, b.c2_slot, b.keys, b.level, c->nQ, c->nP, N, c->lc, b.acc, ts[kArD]);
        if (ts[kArN30].n)
            k_ks_macd<kArN30><<<dim3(gx, ts[kArN30].n, b.nct), 256, 0, st>>>(D, b.c2_slot, b.keys, b.level, c->nQ, c->nP, N, c->lc, b.acc, ts[kArN30]);
        if (ts[kArN31].n)
            k_ks_macd<kArN31><<<dim3(gx, ts[kArN31].n, b.nct), 256, 0, st>>>(D, b.c2_slot, b.keys, b.level, c->nQ, c->nP, N, c->lc, b.acc, ts[kArN31]);
        c->launches += (ts[kArW].n > 0) + (ts[kArD].n > 0) + (ts[kArN30].n > 0) + (ts[kArN31].n > 0) - 1;
        SFG_LAUNCHED(c, "k_ks_macd", st);
    } else if (inner_all(c, b, ks, st)) {
        return -1;
    }
    // 3. INTT of the P limbs of acc (TT order in, natural order out, in place): groups = (ct, comp)
    LimbSel selp;
    selp.n = c->nP;
    for (int i = 0; i < c->nP; i++) selp.idx[i] = c->nQ + i;
    if (launch_ntt_gather(c, b.acc + (size_t)nl * N, nullptr, (size_t)nt * N, b.acc + (size_t)nl * N, (size_t)nt * N, b.nct * 2 * c->nP, selp, true,
                          true, st))
        return -1;
    if (!moddown) return 0;
    // 4. mod-down, + c0, automorphism, store / accumulate
    TgtSel ls[kNumArith] = {{0, {}}, {0, {}}, {0, {}}, {0, {}}};
    for (int l = 0; l < b.out_layout.nl; l++) {
        TgtSel &t = ls[arith_kind(c->mod[l])];
        t.tt[t.n++] = l;
    }
    const uint64_t *extP = nullptr;
    if (c->nP > 1) {  // the base conversion of the P limbs once per coefficient, not once per Q-limb CTA (k_md_extp)
        void *pe;
        const int Lo = b.out_layout.nl;
        if (ws_get(c, WS_MD, (size_t)b.nct * 2 * Lo * N * 8, &pe)) return -1;
        k_md_extp<<<dim3((N + 255) / 256, b.nct * 2), 256, 0, st>>>(b.acc, md, nl, nt, c->nP, Lo, N, c->lc, (uint64_t *)pe);
        SFG_LAUNCHED(c, "k_md_extp", st);
        extP = (const uint64_t *)pe;
    }
    if (moddown_launch<ArW>(c, b, md, pinv, ls[kArW], extP, st) || moddown_launch<ArD>(c, b, md, pinv, ls[kArD], extP, st) ||
        moddown_launch<ArN30>(c, b, md, pinv, ls[kArN30], extP, st) || moddown_launch<ArN31>(c, b, md, pinv, ls[kArN31], extP, st))
        return -1;
    return 0;
}

// ---- rings larger than one CTA's shared memory (logN 15, 16): the same algorithm, unfused ----------------------------------------
// The BASELINE sweep runs NTT / rotation / key-switch up to logN = 16 with the full modulus chain (18 + 3 and 34 + 4 moduli).  A ring of
// 2^15 64-bit residues does not fit one CTA, so the steps of switchKeysInPlace run as separate streaming kernels around the
// global-memory transforms of kernels_ntt.cu; keys stay in Lattigo's layout ([beta][2][nQP][N], NTT + Montgomery form).
__global__ void k_ksb_gather_c1(const uint64_t *__restrict__ in, const long long *__restrict__ src_off, int in_nl, int nl, int N, uint64_t *__restrict__ c2) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y, t = blockIdx.z;
    if (j < N) c2[((size_t)t * nl + l) * N + j] = in[src_off[t] + ((size_t)in_nl + l) * N + j];
}
// digit i of every ciphertext, reduced / base-converted to every target modulus (coefficient domain); a target inside the digit
// takes the NTT-domain input limb as it is (decomposeAndSplitNTT)
__global__ void k_ksb_extend(const uint64_t *__restrict__ in, const long long *__restrict__ in_off, int in_nl, const uint64_t *__restrict__ c2,
                             const int *__restrict__ c2_slot, const BaseConv *__restrict__ ks, int digit, int nl, int nt, int nQ, int N,
                             const LimbConst *__restrict__ lcs, uint64_t *__restrict__ ext) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, tt = blockIdx.y, ct = blockIdx.z;
    if (j >= N) return;
    const BaseConv &bc = ks[(size_t)digit * nt + tt];
    uint64_t v;
    if (bc.ns == 0) {
        v = in[in_off[ct] + ((size_t)in_nl + tt) * N + j];
    } else {
        const int tgt = tt < nl ? tt : nQ + (tt - nl);
        const uint64_t *c2ct = c2 + (size_t)c2_slot[ct] * nl * N;
        uint64_t xs[kMaxAlpha];
        for (int k = 0; k < bc.ns; k++) xs[k] = c2ct[(size_t)bc.src_limb[k] * N + j];
        v = base_conv_coeff(bc, xs, lcs, lcs[tgt].q);
    }
    ext[((size_t)ct * nt + tt) * N + j] = v;
}
__global__ void k_ksb_mac(const uint64_t *__restrict__ ext, const uint64_t *const *__restrict__ keys, int digit, int nl, int nt, int nQ, int nQP, int N,
                          const LimbConst *__restrict__ lcs, int first, uint64_t *__restrict__ acc) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, tt = blockIdx.y, ct = blockIdx.z;
    if (j >= N) return;
    const int tgt = tt < nl ? tt : nQ + (tt - nl);
    const LimbConst lc = lcs[tgt];
    const uint64_t d = ext[((size_t)ct * nt + tt) * N + j];
    const uint64_t *key = keys[ct];
#pragma unroll
    for (int comp = 0; comp < 2; comp++) {
        const uint64_t p = mred(d, key[((size_t)(digit * 2 + comp) * nQP + tgt) * N + j], lc);
        uint64_t *a = acc + ((size_t)(ct * 2 + comp) * nt + tt) * N + j;
        *a = first ? p : add_mod(*a, p, lc.q);
    }
}
__global__ void k_ksb_mdext(const uint64_t *__restrict__ acc, const BaseConv *__restrict__ md, int nl, int nt, int nP, int L, int N,
                            const LimbConst *__restrict__ lcs, uint64_t *__restrict__ E) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y, p = blockIdx.z;  // p = ct * 2 + comp
    if (j >= N) return;
    const uint64_t *accP = acc + ((size_t)p * nt + nl) * N;
    uint64_t xs[kMaxAlpha];
    for (int k = 0; k < nP; k++) xs[k] = accP[(size_t)k * N + j];
    E[((size_t)p * L + l) * N + j] = base_conv_coeff(md[l], xs, lcs, lcs[l].q);
}
__global__ void k_ksb_final(const uint64_t *__restrict__ in, const long long *__restrict__ in_off, const uint64_t *__restrict__ acc,
                            const uint64_t *__restrict__ E, const uint64_t *__restrict__ pinv, const uint32_t *const *__restrict__ perms, int nt, int L,
                            int N, const LimbConst *__restrict__ lcs, unsigned char *__restrict__ out, const long long *__restrict__ out_off,
                            PolyLayout olay, int accumulate) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y, p = blockIdx.z, ct = p >> 1, comp = p & 1;
    if (k >= N) return;
    const uint64_t q = lcs[l].q;
    const int j = (int)perms[ct][k];  // PermuteNTTWithIndexLvl: out[k] = in[index[k]]
    uint64_t v = mul_shoup(sub_mod(acc[((size_t)p * nt + l) * N + j], E[((size_t)p * L + l) * N + j], q), pinv[2 * l], pinv[2 * l + 1], q);
    if (comp == 0) v = add_mod(v, in[in_off[ct] + (size_t)l * N + j], q);
    uint64_t *o = reinterpret_cast<uint64_t *>(out + out_off[ct] + (size_t)comp * olay.bytes + olay.off[l]) + k;
    *o = accumulate ? add_mod(*o, v, q) : v;
}

static int rotate_big(Ctx *c, const KsBatch &b, BaseConv *ks, BaseConv *md, uint64_t *pinv, cudaStream_t st) {
    const int N = c->N, nl = b.level + 1, nP = c->nP, nt = nl + nP, nQ = c->nQ, L = b.out_layout.nl, alpha = nP, beta = (nl + alpha - 1) / alpha;
    for (int l = 0; l < L; l++)
        if (b.out_layout.es[l] != 8) SFG_FAIL(c, "key-switch for logN > 14 writes the u64 layout only");
    const unsigned gx = (unsigned)((N + 255) / 256);
    k_ksb_gather_c1<<<dim3(gx, nl, b.n_c2), 256, 0, st>>>(b.in, b.c2_src_off, b.in_nl, nl, N, b.c2);
    SFG_LAUNCHED(c, "k_ksb_gather_c1", st);
    LimbSel selq;
    selq.n = nl;
    for (int i = 0; i < nl; i++) selq.idx[i] = i;
    if (launch_ntt(c, b.c2, (size_t)nl * N, b.c2, (size_t)nl * N, b.n_c2 * nl, selq, true, st)) return -1;
    for (int k0 = 0; k0 < b.nct; k0 += b.acc_cap) {
        const int n = std::min(b.acc_cap, b.nct - k0);
        void *pe;
        if (ws_get(c, WS_KSB, (size_t)n * std::max(nt, 2 * L) * N * 8, &pe)) return -1;
        uint64_t *ext = (uint64_t *)pe;
        for (int i = 0; i < beta; i++) {
            k_ksb_extend<<<dim3(gx, nt, n), 256, 0, st>>>(b.in, b.in_off + k0, b.in_nl, b.c2, b.c2_slot + k0, ks, i, nl, nt, nQ, N, c->lc, ext);
            SFG_LAUNCHED(c, "k_ksb_extend", st);
            // forward NTT of every target outside the digit: the ranges [0, lo) and [hi, nt)
            const int lo = i * alpha, hi = std::min((i + 1) * alpha, nl);
            for (int part = 0; part < 2; part++) {
                const int t0 = part ? hi : 0, t1 = part ? nt : lo;
                if (t1 <= t0) continue;
                LimbSel sel;
                sel.n = t1 - t0;
                for (int tt = t0; tt < t1; tt++) sel.idx[tt - t0] = tt < nl ? tt : nQ + (tt - nl);
                if (launch_ntt(c, ext + (size_t)t0 * N, (size_t)nt * N, ext + (size_t)t0 * N, (size_t)nt * N, n * sel.n, sel, false, st)) return -1;
            }
            k_ksb_mac<<<dim3(gx, nt, n), 256, 0, st>>>(ext, b.keys + k0, i, nl, nt, nQ, c->nQP, N, c->lc, i == 0 ? 1 : 0, b.acc);
            SFG_LAUNCHED(c, "k_ksb_mac", st);
        }
        LimbSel selp;
        selp.n = nP;
        for (int i = 0; i < nP; i++) selp.idx[i] = nQ + i;
        if (launch_ntt(c, b.acc + (size_t)nl * N, (size_t)nt * N, b.acc + (size_t)nl * N, (size_t)nt * N, n * 2 * nP, selp, true, st)) return -1;
        k_ksb_mdext<<<dim3(gx, L, n * 2), 256, 0, st>>>(b.acc, md, nl, nt, nP, L, N, c->lc, ext);
        SFG_LAUNCHED(c, "k_ksb_mdext", st);
        LimbSel sell;
        sell.n = L;
        for (int i = 0; i < L; i++) sell.idx[i] = i;
        if (launch_ntt(c, ext, (size_t)L * N, ext, (size_t)L * N, n * 2 * L, sell, false, st)) return -1;
        k_ksb_final<<<dim3(gx, L, n * 2), 256, 0, st>>>(b.in, b.in_off + k0, b.acc, ext, pinv, b.perms + k0, nt, L, N, c->lc, (unsigned char *)b.out,
                                                         b.out_off + k0, b.out_layout, b.accumulate ? 1 : 0);
        SFG_LAUNCHED(c, "k_ksb_final", st);
    }
    return 0;
}

static int bc_prep(Ctx *c, KsBatch &b, BaseConv *ks, cudaStream_t st) {
    if (c->nP == 1) return 0;  // single-modulus digits: plain reduction of the representative, nothing to hoist
    const int N = c->N, nl = b.level + 1, nt = nl + c->nP, beta = (nl + c->nP - 1) / c->nP;
    void *pv;
    if (ws_get(c, WS_VQ, (size_t)b.n_c2 * beta * N * 4, &pv)) return -1;
    k_ks_bcprep<<<dim3((N + 255) / 256, beta, b.n_c2), 256, 0, st>>>(b.c2, ks, nl, nt, beta, N, c->lc, (uint32_t *)pv);
    SFG_LAUNCHED(c, "k_ks_bcprep", st);
    b.vq = (const uint32_t *)pv;
    // all digits x all targets, when the buffer is affordable (<= 16 GiB); otherwise k_ks_inner2 converts on the fly
    const size_t eb = (size_t)b.n_c2 * beta * nt * N * 8;
    if (eb <= ((size_t)16 << 30) && getenv("SFG_KS_NOEXTD") == nullptr) {
        void *pe;
        if (ws_get(c, WS_EXTD, eb, &pe)) return -1;
        k_ks_extd<<<dim3((N + 255) / 256, beta, b.n_c2), 256, 0, st>>>(b.c2, b.vq, ks, nl, nt, c->nQ, beta, N, c->lc, (uint64_t *)pe);
        SFG_LAUNCHED(c, "k_ks_extd", st);
        b.extd = (const uint64_t *)pe;
    }
    return 0;
}

int launch_rotate(Ctx *c, const KsBatch &b, cudaStream_t st) {
    if (b.nct <= 0) return 0;
    if (c->logN < 6) SFG_FAIL(c, "logN >= 6 required");
    if (b.acc_cap < 1) SFG_FAIL(c, "key-switch scratch is empty");
    const int N = c->N, nl = b.level + 1;
    BaseConv *ks, *md;
    uint64_t *pinv;
    if (ctx_get_ks_tables(c, b.level, &ks, &md, &pinv)) return -1;
    if (ks_unfused(c)) return rotate_big(c, b, ks, md, pinv, st);
    // 1. c2 = INTT(c1), once per distinct input ciphertext
    LimbSel sel;
    sel.n = nl;
    for (int i = 0; i < nl; i++) sel.idx[i] = i;
    if (launch_ntt_gather(c, b.in + (size_t)b.in_nl * N, b.c2_src_off, 0, b.c2, (size_t)nl * N, b.n_c2 * nl, sel, true, false, st)) return -1;
    KsBatch bq = b;
    if (bc_prep(c, bq, ks, st)) return -1;
    // many rotations of few ciphertexts (the baby steps: d-1 rotations of every A[i][bi]): transform the digits once per ciphertext
    const uint64_t *D = nullptr;
    if (b.n_c2 * 4 <= b.nct && b.nct <= 65535) {
        const int nt = nl + c->nP, beta = (nl + c->nP - 1) / c->nP;
        void *pd;
        if (ws_get(c, WS_KSB, (size_t)b.n_c2 * beta * nt * N * 8, &pd)) return -1;
        KsBatch bd = bq;
        bd.nct = b.n_c2;
        bd.in_off = b.c2_src_off;
        bd.c2_slot = nullptr;
        bd.keys = nullptr;
        bd.dout = (uint64_t *)pd;
        if (inner_all(c, bd, ks, st)) return -1;
        D = (const uint64_t *)pd;
    }
    for (int k0 = 0; k0 < b.nct; k0 += b.acc_cap) {
        KsBatch ch = bq;
        ch.nct = std::min(b.acc_cap, b.nct - k0);
        ch.in_off += k0;
        ch.c2_slot += k0;
        ch.keys += k0;
        ch.perms += k0;
        ch.out_off += k0;
        if (rotate_chunk(c, ch, ks, md, pinv, st, true, D)) return -1;
    }
    return 0;
}

// Giant-step sums (gwas/matmult.go:1203-1227): entries a*nout + o (a < nacc) are rotated with their own keys and summed into
// output o.  Steps 1-3 run over all entries at once; the mod-down is hoisted out of the sum (k_md_accum / k_md_final).
int launch_rotate_sum(Ctx *c, const KsBatch &b, int nout, const uint32_t *ginv_dev, uint64_t *S1, uint64_t *C0, uint64_t *E, bool first,
                      cudaStream_t st) {
    if (b.nct <= 0) return 0;
    if (c->logN > 14 || c->logN < 6) SFG_FAIL(c, "fused key-switch kernels support 6 <= logN <= 14 (got %d)", c->logN);
    if (b.nct % nout) SFG_FAIL(c, "rotate_sum: %d entries do not divide into %d outputs", b.nct, nout);
    const int N = c->N, nl = b.level + 1, L = b.out_layout.nl;
    const int cap = (b.acc_cap / nout) * nout;
    if (cap < nout) SFG_FAIL(c, "key-switch scratch holds %d ciphertexts, one giant step needs %d", b.acc_cap, nout);
    BaseConv *ks, *md;
    uint64_t *pinv;
    if (ctx_get_ks_tables(c, b.level, &ks, &md, &pinv)) return -1;
    LimbSel sel;
    sel.n = nl;
    for (int i = 0; i < nl; i++) sel.idx[i] = i;
    if (launch_ntt_gather(c, b.in + (size_t)b.in_nl * N, b.c2_src_off, 0, b.c2, (size_t)nl * N, b.n_c2 * nl, sel, true, false, st)) return -1;
    KsBatch bq = b;
    if (bc_prep(c, bq, ks, st)) return -1;
    // whole source polynomials are staged (the permutation reaches anywhere): as many ring buffers as fit one CTA, at most 3
    const int nbuf = (int)std::max<size_t>(1, std::min<size_t>(3, ((size_t)220 << 10) / ((size_t)N * 8)));
    const size_t smem = (size_t)nbuf * N * 8 + 64;
    if (smem > ((size_t)227 << 10)) SFG_FAIL(c, "giant-step sums: a ring of 2^%d coefficients does not fit one CTA's shared memory", c->logN);
    const uint64_t *extP = nullptr;
    auto accum = [&](auto kern, const TgtSel &ts, const KsBatch &ch, int k0, int nsplit, int thr, int fst) -> int {
        SFG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<dim3(ts.n * nsplit, 2, nout), thr, smem, st>>>(ch.in, ch.in_off, ch.in_nl, ch.acc, md, ch.perms, ginv_dev + k0, ch.level, c->nQ, c->nP,
                                                              c->logN, c->lc, nout, ch.nct / nout, S1, C0, E, L, fst, ts, nbuf, extP);
        SFG_LAUNCHED(c, "k_md_accum", st);
        return 0;
    };
    TgtSel nar{0, {}}, wid{0, {}};
    for (int l = 0; l < L; l++) {
        TgtSel &t = c->mod[l] < (1ULL << 31) ? nar : wid;
        t.tt[t.n++] = l;
    }
    for (int k0 = 0; k0 < b.nct; k0 += cap) {
        KsBatch ch = bq;
        ch.nct = std::min(cap, b.nct - k0);
        ch.in_off += k0;
        ch.c2_slot += k0;
        ch.keys += k0;
        ch.perms += k0;
        ch.out_off += k0;
        ch.acc_dlog = true;
        if (rotate_chunk(c, ch, ks, md, pinv, st, false)) return -1;
        if (c->nP > 1) {
            void *pe;
            if (ws_get(c, WS_KSB, (size_t)ch.nct * 2 * L * N * 8, &pe)) return -1;
            k_md_extp<<<dim3((N + 255) / 256, ch.nct * 2), 256, 0, st>>>(ch.acc, md, nl, nl + c->nP, c->nP, L, N, c->lc, (uint64_t *)pe);
            SFG_LAUNCHED(c, "k_md_extp", st);
            extP = (const uint64_t *)pe;
        }
        const int nsplit = N > 8192 ? N / 8192 : 1;
        const int thr = std::min(512, std::max(32, N / nsplit / 16));
        const int fst = (first && k0 == 0) ? 1 : 0;
        if (wid.n && (c->logN == 13 ? accum(k_md_accum<uint64_t, 13>, wid, ch, k0, nsplit, thr, fst) : accum(k_md_accum<uint64_t, 0>, wid, ch, k0, nsplit, thr, fst)))
            return -1;
        if (nar.n && (c->logN == 13 ? accum(k_md_accum<uint32_t, 13>, nar, ch, k0, nsplit, thr, fst) : accum(k_md_accum<uint32_t, 0>, nar, ch, k0, nsplit, thr, fst)))
            return -1;
    }
    return 0;
}

template <class A>
static int md_final_launch(Ctx *c, int level, int nout, const uint64_t *S1, const uint64_t *C0, const uint64_t *E, uint64_t *pinv, void *out,
                           const long long *out_off, const PolyLayout &olay, const TgtSel &sel, cudaStream_t st) {
    if (sel.n == 0) return 0;
    const int logN = c->logN, N = c->N;
    const PassPlan plan = make_pass_plan(logN - kLastR);
    const size_t smem = ntt_smem_elems(N) * sizeof(typename A::T);
    SFG_CUDA(c, cudaFuncSetAttribute(k_md_final<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 g(sel.n, 2, nout);
    k_md_final<A><<<g, ntt_threads(N), smem, st>>>(S1, C0, E, pinv, logN, plan, c->tw2, c->lc, (unsigned char *)out, out_off, olay, olay.nl, sel,
                                                   c->dlog_pos);
    SFG_LAUNCHED(c, "k_md_final", st);
    return 0;
}

// out[o] += (S1 - NTT(E)) * P^-1 + C0   (out in the plain u64 layout)
int launch_rotate_sum_final(Ctx *c, int level, int nout, const uint64_t *S1, const uint64_t *C0, const uint64_t *E, void *out,
                            const long long *out_off, const PolyLayout &olay, cudaStream_t st) {
    BaseConv *ks, *md;
    uint64_t *pinv;
    if (ctx_get_ks_tables(c, level, &ks, &md, &pinv)) return -1;
    for (int l = 0; l < olay.nl; l++)
        if (olay.es[l] != 8) SFG_FAIL(c, "rotate_sum_final writes the u64 layout only");
    TgtSel ls[kNumArith] = {{0, {}}, {0, {}}, {0, {}}, {0, {}}};
    for (int l = 0; l < olay.nl; l++) {
        TgtSel &t = ls[arith_kind(c->mod[l])];
        t.tt[t.n++] = l;
    }
    if (md_final_launch<ArW>(c, level, nout, S1, C0, E, pinv, out, out_off, olay, ls[kArW], st) ||
        md_final_launch<ArD>(c, level, nout, S1, C0, E, pinv, out, out_off, olay, ls[kArD], st) ||
        md_final_launch<ArN30>(c, level, nout, S1, C0, E, pinv, out, out_off, olay, ls[kArN30], st) ||
        md_final_launch<ArN31>(c, level, nout, S1, C0, E, pinv, out, out_off, olay, ls[kArN31], st))
        return -1;
    return 0;
}

}  // namespace sfg
