// kernels_ntt.cu -- batched negacyclic NTT/INTT over RNS limbs, plus the stand-alone K1/K2/K3 kernels and the
// genotype preparation scan.  sm_100a only.
//
// NTT layout: one CTA owns one polynomial-limb (or a 2^14-coefficient sub-block of it for logN > 14) in shared memory;
// stages run over smem; the first (forward) / last (inverse) log2(N/S) stages of logN > 14 run over global memory.
#include <cstdlib>

#include "kernels.h"
#include "ntt.cuh"
#include "ntt2.cuh"

namespace sfg {

constexpr int kMaxLogS = 14;  // 2^14 * 8 B = 128 KB of shared memory per CTA

template <bool INV>
__global__ void __launch_bounds__(1024, 1)
k_ntt_smem(const uint64_t *__restrict__ src, size_t src_gstride, uint64_t *__restrict__ dst, size_t dst_gstride, LimbSel sel,
           int logN, int logS, const uint64_t *__restrict__ tw, const LimbConst *__restrict__ lcs) {
    extern __shared__ __align__(16) uint64_t s[];
    const int N = 1 << logN, S = 1 << logS, nblk = N >> logS;
    const int p = blockIdx.x / nblk, blk = blockIdx.x % nblk;
    const int g = p / sel.n, k = p % sel.n, limb = sel.idx[k];
    const uint64_t *in = src + (size_t)g * src_gstride + (size_t)k * N + (size_t)blk * S;
    uint64_t *out = dst + (size_t)g * dst_gstride + (size_t)k * N + (size_t)blk * S;
    const LimbConst lc = lcs[limb];
    const NttTab tab = ntt_tab(tw, limb, N);
    for (int i = threadIdx.x; i < S; i += blockDim.x) s[i] = in[i];
    __syncthreads();
    if (!INV)
        ntt_fwd_smem(s, logS, nblk, blk, tab, lc.q);
    else
        ntt_inv_smem(s, logS, N, blk, tab, lc, nblk == 1);
    for (int i = threadIdx.x; i < S; i += blockDim.x) out[i] = s[i];
}

// One global-memory radix-2 stage (only used for logN > 14).  Forward: stage with m groups (m < N/S).
// Inverse: stage with h groups (h < N/S ... 1); `last` multiplies by N^-1.
template <bool INV>
__global__ void k_ntt_gstage(uint64_t *__restrict__ data, size_t gstride, LimbSel sel, int logN, int m, int last,
                             const uint64_t *__restrict__ tw, const LimbConst *__restrict__ lcs) {
    const int N = 1 << logN;
    const int p = blockIdx.y;
    const int g = p / sel.n, k = p % sel.n, limb = sel.idx[k];
    uint64_t *a = data + (size_t)g * gstride + (size_t)k * N;
    const LimbConst lc = lcs[limb];
    const NttTab tab = ntt_tab(tw, limb, N);
    const int t = N / (2 * m);
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < N / 2; b += gridDim.x * blockDim.x) {
        const int i = b / t, j = b % t, idx = 2 * i * t + j;
        if (!INV) {
            const uint64_t U = a[idx], V = mul_shoup(a[idx + t], tab.w[m + i], tab.wsh[m + i], lc.q);
            a[idx] = add_mod(U, V, lc.q);
            a[idx + t] = sub_mod(U, V, lc.q);
        } else {
            const uint64_t U = a[idx], V = a[idx + t];
            uint64_t x = add_mod(U, V, lc.q), y = mul_shoup(sub_mod(U, V, lc.q), tab.wi[m + i], tab.wish[m + i], lc.q);
            if (last) {
                x = mul_shoup(x, lc.ninv, lc.ninv_sh, lc.q);
                y = mul_shoup(y, lc.ninv, lc.ninv_sh, lc.q);
            }
            a[idx] = x;
            a[idx + t] = y;
        }
    }
}

__global__ void k_copy_polys(const uint64_t *__restrict__ src, size_t src_gstride, uint64_t *__restrict__ dst, size_t dst_gstride,
                             int n_per_group, int N) {
    const int p = blockIdx.y, g = p / n_per_group, k = p % n_per_group;
    const uint64_t *in = src + (size_t)g * src_gstride + (size_t)k * N;
    uint64_t *out = dst + (size_t)g * dst_gstride + (size_t)k * N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) out[i] = in[i];
}

// ---------------------------------------------------------------------------------------------------------------
// register-tiled transforms (ntt2.cuh) for rings that fit one CTA's shared memory (logN <= 14), one launch per arithmetic class
// ---------------------------------------------------------------------------------------------------------------
struct SubSel {  // the polynomials of a group that belong to one arithmetic class
    int n;
    int pos[kMaxLimbs];   // position k inside the group (polynomial at group base + k*N)
    int limb[kMaxLimbs];  // modulus index
};

// LOGN > 0: ring size known at compile time (indices and the pass plan fold into immediates); LOGN == 0: generic
template <class A, int LOGN>
__global__ void __launch_bounds__(512, A::kMinBlocks)
k_ntt2_fwd(const uint64_t *__restrict__ src, const long long *__restrict__ src_off, size_t src_gstride, uint64_t *__restrict__ dst,
           size_t dst_gstride, SubSel sub, int logN_arg, PassPlan plan_arg, const TwTab *__restrict__ tabs, const LimbConst *__restrict__ lcs) {
    using T = typename A::T;
    extern __shared__ __align__(16) unsigned char smraw[];
    T *s = reinterpret_cast<T *>(smraw);
    const int logN = LOGN ? LOGN : logN_arg;
    const PassPlan plan = LOGN ? make_pass_plan(LOGN - kLastR) : plan_arg;
    const int N = 1 << logN, g = blockIdx.x / sub.n, kk = blockIdx.x % sub.n, k = sub.pos[kk], limb = sub.limb[kk];
    const uint64_t *in = src + (src_off ? (size_t)src_off[g] : (size_t)g * src_gstride) + (size_t)k * N;
    uint64_t *out = dst + (size_t)g * dst_gstride + (size_t)k * N;
    const typename A::C c = A::make(lcs[limb]);
    ntt_forward<A>(s, logN, logN, 0, plan, tabs[limb], c, [&](int j, int) { return A::from_canon(in[j], c); },
                   [&](int j, T v, int) { s[sidx<sizeof(T)>(j)] = A::canon(v, c); });
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += blockDim.x) out[j] = s[sidx<sizeof(T)>(j)];
}

template <class A, int LOGN>
__global__ void __launch_bounds__(512, A::kMinBlocks)
k_ntt2_inv(const uint64_t *__restrict__ src, const long long *__restrict__ src_off, size_t src_gstride, uint64_t *__restrict__ dst,
           size_t dst_gstride, SubSel sub, int logN_arg, PassPlan plan_arg, const TwTab *__restrict__ tabs, const LimbConst *__restrict__ lcs,
           int in_tt) {
    using T = typename A::T;
    extern __shared__ __align__(16) unsigned char smraw[];
    T *s = reinterpret_cast<T *>(smraw);
    const int logN = LOGN ? LOGN : logN_arg;
    const PassPlan plan = LOGN ? make_pass_plan(LOGN - kLastR) : plan_arg;
    const int N = 1 << logN, g = blockIdx.x / sub.n, kk = blockIdx.x % sub.n, k = sub.pos[kk], limb = sub.limb[kk];
    const uint64_t *in = src + (src_off ? (size_t)src_off[g] : (size_t)g * src_gstride) + (size_t)k * N;
    uint64_t *out = dst + (size_t)g * dst_gstride + (size_t)k * N;
    const typename A::C c = A::make(lcs[limb]);
    auto fin = [&](int j, T v, int) { out[j] = v; };
    if (in_tt) {
        ntt_inverse<A>(s, logN, plan, tabs[limb], c, [&](int j, int) { return A::from_canon(in[tt_index(j, N)], c); }, fin);
    } else {
        for (int j0 = threadIdx.x; j0 < N; j0 += 8 * blockDim.x) {  // 8 independent loads in flight per thread
            uint64_t xv[8];
#pragma unroll
            for (int u = 0; u < 8; u++) xv[u] = (j0 + u * (int)blockDim.x < N) ? __ldg(in + j0 + u * blockDim.x) : 0;
#pragma unroll
            for (int u = 0; u < 8; u++)
                if (j0 + u * (int)blockDim.x < N) s[sidx<sizeof(T)>(j0 + u * blockDim.x)] = A::from_canon(xv[u], c);
        }
        __syncthreads();
        ntt_inverse<A>(s, logN, plan, tabs[limb], c, [&](int j, int) { return s[sidx<sizeof(T)>(j)]; }, fin);
    }
}

template <class A>
static int ntt2_launch(Ctx *c, const uint64_t *src, const long long *src_off, size_t sgs, uint64_t *dst, size_t dgs, int ngroups,
                       const SubSel &sub, bool inverse, bool in_tt, cudaStream_t st) {
    if (sub.n == 0 || ngroups == 0) return 0;
    const int logN = c->logN, N = c->N;
    const PassPlan plan = make_pass_plan(logN - kLastR);
    const size_t smem = ntt_smem_elems(N) * sizeof(typename A::T);
    const int threads = std::min(512, std::max(32, N >> kLastR));
    auto inv = [&](auto kern) -> int {
        SFG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<ngroups * sub.n, threads, smem, st>>>(src, src_off, sgs, dst, dgs, sub, logN, plan, c->tw2, c->lc, in_tt ? 1 : 0);
        SFG_LAUNCHED(c, "k_ntt2_inv", st);
        return 0;
    };
    auto fwd = [&](auto kern) -> int {
        SFG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<ngroups * sub.n, threads, smem, st>>>(src, src_off, sgs, dst, dgs, sub, logN, plan, c->tw2, c->lc);
        SFG_LAUNCHED(c, "k_ntt2_fwd", st);
        return 0;
    };
    if (inverse) {
        if (logN == 13) return inv(k_ntt2_inv<A, 13>);
        if (logN == 14) return inv(k_ntt2_inv<A, 14>);
        return inv(k_ntt2_inv<A, 0>);
    }
    if (logN == 13) return fwd(k_ntt2_fwd<A, 13>);
    if (logN == 14) return fwd(k_ntt2_fwd<A, 14>);
    return fwd(k_ntt2_fwd<A, 0>);
    return 0;
}

static int launch_ntt_old(Ctx *c, const uint64_t *src, size_t src_gstride, uint64_t *dst, size_t dst_gstride, int npoly, const LimbSel &sel,
                          bool inverse, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------------------
// Four-step transforms for rings larger than one CTA's shared memory (logN 15, 16: the NTT / key-switch sweep of BASELINE config 3).
// forward : ONE global-memory pass does the first CS = logN - 13 stages on 2^CS coefficients N / 2^CS apart per thread (registers,
//           coalesced across threads), then every slice of 2^13 consecutive coefficients is an independent transform with the
//           register-tiled passes of ntt2.cuh (one CTA per (polynomial, slice)).
// inverse : the slices first, then the last CS Gentleman-Sande stages and N^-1 in one global pass.
// 2 x (read + write) of the data instead of CS + 1 radix-2 global passes around a radix-2 shared-memory kernel.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSliceLog = 13;

template <class A, int R, bool INV>
__global__ void __launch_bounds__(256)
k_ntt_gpass(const uint64_t *__restrict__ src, size_t src_gstride, uint64_t *__restrict__ dst, size_t dst_gstride, SubSel sub, int logN,
            const TwTab *__restrict__ tabs, const LimbConst *__restrict__ lcs) {
    using T = typename A::T;
    using TW = typename A::TW;
    constexpr int E = 1 << R;
    const int N = 1 << logN, lobits = logN - R;
    const int g = blockIdx.y / sub.n, kk = blockIdx.y % sub.n, k = sub.pos[kk], limb = sub.limb[kk];
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (N >> R)) return;
    const uint64_t *in = src + (size_t)g * src_gstride + (size_t)k * N;
    uint64_t *out = dst + (size_t)g * dst_gstride + (size_t)k * N;
    const typename A::C c = A::make(lcs[limb]);
    const TW *tw = reinterpret_cast<const TW *>(INV ? tabs[limb].inv : tabs[limb].fwd);
    T v[E];
#pragma unroll
    for (int i = 0; i < E; i++) v[i] = A::from_canon(in[p + ((size_t)i << lobits)], c);
    if (!INV) {
#pragma unroll
        for (int r = 0; r < R; r++) {  // global stage r: 2^r groups, twiddle NttPsi[2^r + group]
            const int half = E >> (r + 1);
#pragma unroll
            for (int gg = 0; gg < (1 << r); gg++) {
                const TW w = __ldg(tw + (1 << r) + gg);
#pragma unroll
                for (int j = 0; j < half; j++) A::fwd(v[gg * 2 * half + j], v[gg * 2 * half + j + half], w, c);
            }
        }
#pragma unroll
        for (int i = 0; i < E; i++) out[p + ((size_t)i << lobits)] = (uint64_t)A::canon(v[i], c);
    } else {
#pragma unroll
        for (int r = R - 1; r >= 0; r--) {
            const int half = E >> (r + 1);
#pragma unroll
            for (int gg = 0; gg < (1 << r); gg++) {
                const TW w = __ldg(tw + (1 << r) + gg);
#pragma unroll
                for (int j = 0; j < half; j++) A::inv(v[gg * 2 * half + j], v[gg * 2 * half + j + half], w, c);
            }
        }
#pragma unroll
        for (int i = 0; i < E; i++) out[p + ((size_t)i << lobits)] = (uint64_t)A::inv_final(v[i], c);
    }
}

template <class A, bool INV>
__global__ void __launch_bounds__(512, 1)
k_ntt2_slice(const uint64_t *__restrict__ src, size_t src_gstride, uint64_t *__restrict__ dst, size_t dst_gstride, SubSel sub, int logN,
             PassPlan plan, const TwTab *__restrict__ tabs, const LimbConst *__restrict__ lcs) {
    using T = typename A::T;
    extern __shared__ __align__(16) unsigned char smraw[];
    T *s = reinterpret_cast<T *>(smraw);
    const int N = 1 << logN, CS = logN - kSliceLog, S = 1 << kSliceLog;
    const int sl = blockIdx.x & ((1 << CS) - 1), pk = blockIdx.x >> CS;
    const int g = pk / sub.n, kk = pk % sub.n, k = sub.pos[kk], limb = sub.limb[kk];
    const uint64_t *in = src + (size_t)g * src_gstride + (size_t)k * N + (size_t)sl * S;
    uint64_t *out = dst + (size_t)g * dst_gstride + (size_t)k * N + (size_t)sl * S;
    const typename A::C c = A::make(lcs[limb]);
    if (!INV) {
        ntt_forward<A>(s, logN, kSliceLog, sl, plan, tabs[limb], c, [&](int j, int) { return A::from_canon(in[j], c); },
                       [&](int j, T v, int) { s[sidx<sizeof(T)>(j)] = A::canon(v, c); });
        __syncthreads();
        for (int j = threadIdx.x; j < S; j += blockDim.x) out[j] = (uint64_t)s[sidx<sizeof(T)>(j)];
    } else {
        for (int j = threadIdx.x; j < S; j += blockDim.x) s[sidx<sizeof(T)>(j)] = A::from_canon(in[j], c);
        __syncthreads();
        ntt_inverse_slice<A>(s, logN, kSliceLog, sl, plan, tabs[limb], c, [&](int j, int) { return s[sidx<sizeof(T)>(j)]; },
                             [&](int j, T v, int) { out[j] = (uint64_t)A::canon(v, c); });
    }
}

template <class A>
static int ntt4_launch(Ctx *c, const uint64_t *src, size_t sgs, uint64_t *dst, size_t dgs, int ngroups, const SubSel &sub, bool inverse,
                       cudaStream_t st) {
    if (sub.n == 0 || ngroups == 0) return 0;
    const int logN = c->logN, N = c->N, CS = logN - kSliceLog;
    const PassPlan plan = make_pass_plan(kSliceLog - kLastR);
    const size_t smem = ntt_smem_elems(1 << kSliceLog) * sizeof(typename A::T);
    const dim3 gp((unsigned)(((N >> CS) + 255) / 256), (unsigned)(ngroups * sub.n));
    const unsigned nslice = (unsigned)(ngroups * sub.n) << CS;
    auto gpass = [&](bool inv, const uint64_t *a, size_t as, uint64_t *b, size_t bs) -> int {
        if (CS == 1) {
            if (inv) k_ntt_gpass<A, 1, true><<<gp, 256, 0, st>>>(a, as, b, bs, sub, logN, c->tw2, c->lc);
            else k_ntt_gpass<A, 1, false><<<gp, 256, 0, st>>>(a, as, b, bs, sub, logN, c->tw2, c->lc);
        } else if (CS == 2) {
            if (inv) k_ntt_gpass<A, 2, true><<<gp, 256, 0, st>>>(a, as, b, bs, sub, logN, c->tw2, c->lc);
            else k_ntt_gpass<A, 2, false><<<gp, 256, 0, st>>>(a, as, b, bs, sub, logN, c->tw2, c->lc);
        } else {
            if (inv) k_ntt_gpass<A, 3, true><<<gp, 256, 0, st>>>(a, as, b, bs, sub, logN, c->tw2, c->lc);
            else k_ntt_gpass<A, 3, false><<<gp, 256, 0, st>>>(a, as, b, bs, sub, logN, c->tw2, c->lc);
        }
        SFG_LAUNCHED(c, "k_ntt_gpass", st);
        return 0;
    };
    if (!inverse) {
        if (gpass(false, src, sgs, dst, dgs)) return -1;
        SFG_CUDA(c, cudaFuncSetAttribute(k_ntt2_slice<A, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_ntt2_slice<A, false><<<nslice, 512, smem, st>>>(dst, dgs, dst, dgs, sub, logN, plan, c->tw2, c->lc);
        SFG_LAUNCHED(c, "k_ntt2_slice", st);
    } else {
        SFG_CUDA(c, cudaFuncSetAttribute(k_ntt2_slice<A, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_ntt2_slice<A, true><<<nslice, 512, smem, st>>>(src, sgs, dst, dgs, sub, logN, plan, c->tw2, c->lc);
        SFG_LAUNCHED(c, "k_ntt2_slice", st);
        if (gpass(true, dst, dgs, dst, dgs)) return -1;
    }
    return 0;
}

// src_off (device, optional): element offset of every group's first polynomial (overrides g*src_gstride).
// in_tt: the inputs are in TT order (inverse transforms only).
int launch_ntt_gather(Ctx *c, const uint64_t *src, const long long *src_off, size_t src_gstride, uint64_t *dst, size_t dst_gstride, int npoly,
                      const LimbSel &sel, bool inverse, bool in_tt, cudaStream_t st) {
    if (npoly <= 0) return 0;
    static const bool old_big = [] { const char *e = getenv("SFG_NTT_OLDBIG"); return e && *e == '1'; }();
    if (c->logN > 14 && (src_off || in_tt)) SFG_FAIL(c, "gathered / TT transforms support logN <= 14 (got %d)", c->logN);
    if (c->logN > 16 || (c->logN > 14 && old_big)) return launch_ntt_old(c, src, src_gstride, dst, dst_gstride, npoly, sel, inverse, st);
    SubSel sub[kNumArith] = {{0, {}, {}}, {0, {}, {}}, {0, {}, {}}, {0, {}, {}}};
    for (int k = 0; k < sel.n; k++) {
        SubSel &s = sub[arith_kind(c->mod[sel.idx[k]])];
        s.pos[s.n] = k;
        s.limb[s.n++] = sel.idx[k];
    }
    const int ngroups = npoly / sel.n;
    if (c->logN > 14) {  // four-step: one global pass + register-tiled slices of 2^13 coefficients
        if (ntt4_launch<ArW>(c, src, src_gstride, dst, dst_gstride, ngroups, sub[kArW], inverse, st)) return -1;
        if (ntt4_launch<ArD>(c, src, src_gstride, dst, dst_gstride, ngroups, sub[kArD], inverse, st)) return -1;
        if (ntt4_launch<ArN30>(c, src, src_gstride, dst, dst_gstride, ngroups, sub[kArN30], inverse, st)) return -1;
        if (ntt4_launch<ArN31>(c, src, src_gstride, dst, dst_gstride, ngroups, sub[kArN31], inverse, st)) return -1;
        return 0;
    }
    if (ntt2_launch<ArW>(c, src, src_off, src_gstride, dst, dst_gstride, ngroups, sub[kArW], inverse, in_tt, st)) return -1;
    if (ntt2_launch<ArD>(c, src, src_off, src_gstride, dst, dst_gstride, ngroups, sub[kArD], inverse, in_tt, st)) return -1;
    if (ntt2_launch<ArN30>(c, src, src_off, src_gstride, dst, dst_gstride, ngroups, sub[kArN30], inverse, in_tt, st)) return -1;
    if (ntt2_launch<ArN31>(c, src, src_off, src_gstride, dst, dst_gstride, ngroups, sub[kArN31], inverse, in_tt, st)) return -1;
    return 0;
}
int launch_ntt(Ctx *c, const uint64_t *src, size_t src_gstride, uint64_t *dst, size_t dst_gstride, int npoly, const LimbSel &sel,
               bool inverse, cudaStream_t st) {
    return launch_ntt_gather(c, src, nullptr, src_gstride, dst, dst_gstride, npoly, sel, inverse, false, st);
}

static int launch_ntt_old(Ctx *c, const uint64_t *src, size_t src_gstride, uint64_t *dst, size_t dst_gstride, int npoly, const LimbSel &sel,
               bool inverse, cudaStream_t st) {
    if (npoly <= 0) return 0;
    const int logN = c->logN, N = c->N;
    const int logS = logN < kMaxLogS ? logN : kMaxLogS, S = 1 << logS, nblk = N >> logS;
    const int threads = S / 2 < 1024 ? S / 2 : 1024;
    const size_t smem = (size_t)S * sizeof(uint64_t);
    static bool attr_set = false;
    if (!attr_set) {
        SFG_CUDA(c, cudaFuncSetAttribute(k_ntt_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << kMaxLogS) * 8));
        SFG_CUDA(c, cudaFuncSetAttribute(k_ntt_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << kMaxLogS) * 8));
        attr_set = true;
    }
    if (nblk == 1) {
        if (!inverse)
            k_ntt_smem<false><<<npoly, threads, smem, st>>>(src, src_gstride, dst, dst_gstride, sel, logN, logS, c->tw, c->lc);
        else
            k_ntt_smem<true><<<npoly, threads, smem, st>>>(src, src_gstride, dst, dst_gstride, sel, logN, logS, c->tw, c->lc);
        SFG_LAUNCHED(c, "k_ntt_smem", st);
    } else {
        // large rings: global stages work in place on dst
        dim3 gg(64, npoly);
        if (!inverse) {
            if (src != dst) {
                k_copy_polys<<<gg, 256, 0, st>>>(src, src_gstride, dst, dst_gstride, sel.n, N);
                SFG_LAUNCHED(c, "k_copy_polys", st);
            }
            for (int m = 1; m < nblk; m <<= 1) {
                k_ntt_gstage<false><<<gg, 256, 0, st>>>(dst, dst_gstride, sel, logN, m, 0, c->tw, c->lc);
                SFG_LAUNCHED(c, "k_ntt_gstage", st);
            }
            k_ntt_smem<false><<<npoly * nblk, threads, smem, st>>>(dst, dst_gstride, dst, dst_gstride, sel, logN, logS, c->tw, c->lc);
            SFG_LAUNCHED(c, "k_ntt_smem", st);
        } else {
            k_ntt_smem<true><<<npoly * nblk, threads, smem, st>>>(src, src_gstride, dst, dst_gstride, sel, logN, logS, c->tw, c->lc);
            SFG_LAUNCHED(c, "k_ntt_smem", st);
            for (int h = nblk >> 1; h >= 1; h >>= 1) {
                k_ntt_gstage<true><<<gg, 256, 0, st>>>(dst, dst_gstride, sel, logN, h, h == 1, c->tw, c->lc);
                SFG_LAUNCHED(c, "k_ntt_gstage", st);
            }
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// K1 / K2 / K3 as stand-alone kernels: the exact restatement of gwas/matmult.go:247-324,411-440, used by the parity
// tests; the production MAC fuses K1+K2 (kernels_mac.cu).  Accumulators are (hi, lo) pairs like the reference's uint128.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_mul_coeffs_and_add128(const uint64_t *__restrict__ a, const uint64_t *__restrict__ b, ulonglong2 *__restrict__ acc,
                                        size_t n) {
    for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        ulonglong2 z = acc[j];  // x = hi, y = lo (gwas/matmult.go:196-199)
        u128 t{z.y, z.x};
        mac128(t, a[j], b[j]);
        acc[j] = make_ulonglong2(t.hi, t.lo);
    }
}
__global__ void k_reduce_and_add128(const ulonglong2 *__restrict__ acc, uint64_t *__restrict__ out, LimbConst lc, size_t n) {
    for (size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        const ulonglong2 z = acc[j];
        const uint64_t hhi = __umul64hi(z.y * lc.qinv, lc.q);
        out[j] += z.x - hhi + lc.q;  // gwas/matmult.go:300-301 (not range-reduced)
    }
}
__global__ void k_mform(uint64_t *__restrict__ p, int N, const LimbConst *__restrict__ lcs) {
    const LimbConst lc = lcs[blockIdx.y];
    uint64_t *a = p + (size_t)blockIdx.y * N;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) a[j] = mform(a[j], lc);
}
__global__ void k_mod_reduce(uint64_t *__restrict__ x, int L, int N, const LimbConst *__restrict__ lcs) {
    const size_t poly = blockIdx.y;
    const LimbConst lc = lcs[poly % L];
    uint64_t *a = x + poly * (size_t)N;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) a[j] = bred_add(a[j], lc);
}

int launch_mul_coeffs_and_add128(Ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *acc, size_t n, cudaStream_t st) {
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    k_mul_coeffs_and_add128<<<blocks, 256, 0, st>>>(a, b, (ulonglong2 *)acc, n);
    SFG_LAUNCHED(c, "k_mul_coeffs_and_add128", st);
    return 0;
}
int launch_reduce_and_add128(Ctx *c, const uint64_t *acc, uint64_t *out, int limb, size_t n, cudaStream_t st) {
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 8);
    k_reduce_and_add128<<<blocks, 256, 0, st>>>((const ulonglong2 *)acc, out, c->lc_h[limb], n);
    SFG_LAUNCHED(c, "k_reduce_and_add128", st);
    return 0;
}
int launch_mform(Ctx *c, uint64_t *p, int nlimbs, cudaStream_t st) {
    dim3 g((c->N + 255) / 256, nlimbs);
    k_mform<<<g, 256, 0, st>>>(p, c->N, c->lc);
    SFG_LAUNCHED(c, "k_mform", st);
    return 0;
}
int launch_mod_reduce(Ctx *c, uint64_t *x, size_t npoly, int L, cudaStream_t st) {
    const size_t maxy = 65535;
    for (size_t off = 0; off < npoly; off += maxy - (maxy % L)) {
        const size_t cnt = std::min(npoly - off, maxy - (maxy % L));
        dim3 g((c->N + 255) / 256, (unsigned)cnt);
        k_mod_reduce<<<g, 256, 0, st>>>(x + off * c->N, L, c->N, c->lc);
        SFG_LAUNCHED(c, "k_mod_reduce", st);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Genotype preparation: gwas/matmult.go:1289-1304 (missing -> 0, sum / sqSum before squaring, optional square).
// One thread per column (k_geno_prep) or per 16 consecutive columns (k_geno_prep16) of a row slab; partial sums are reduced over the slab in
// registers and flushed with one atomicAdd per (slab, column).  The sums are exact: every partial is a small integer.
// ---------------------------------------------------------------------------------------------------------------
__global__ void k_geno_prep(int8_t *__restrict__ X, size_t rows, size_t ncols, double *__restrict__ sum, double *__restrict__ sqsum,
                            int square, int rows_per_slab) {
    const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    const size_t r0 = (size_t)blockIdx.y * rows_per_slab;
    const size_t r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
    long long s1 = 0, s2 = 0;
    for (size_t r = r0; r < r1; r++) {
        int8_t v = X[r * ncols + col];
        if (v < 0) v = 0;
        const int8_t v2 = (int8_t)(v * v);  // int8 arithmetic like the reference (row[rj]*row[rj])
        s1 += v;
        s2 += v2;
        X[r * ncols + col] = square ? v2 : v;
    }
    if (sum) atomicAdd(&sum[col], (double)s1);
    if (sqsum) atomicAdd(&sqsum[col], (double)s2);
}

// The same scan with 16-byte accesses (ncols % 16 == 0): a thread owns 16 consecutive columns of a slab of rows; a row segment is
// written back only when it changes (a missing value, or squaring).  HBM-bound: rows * ncols bytes read once.
__global__ void __launch_bounds__(256)
k_geno_prep16(int8_t *__restrict__ X, size_t rows, size_t ncols, double *__restrict__ sum, double *__restrict__ sqsum, int square,
              int rows_per_slab) {
    const size_t col0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (col0 >= ncols) return;
    const size_t r0 = (size_t)blockIdx.y * rows_per_slab;
    const size_t r1 = r0 + rows_per_slab < rows ? r0 + rows_per_slab : rows;
    int s1[16], s2[16];
#pragma unroll
    for (int k = 0; k < 16; k++) s1[k] = s2[k] = 0;
    for (size_t r = r0; r < r1; r++) {
        uint4 *p = reinterpret_cast<uint4 *>(X + r * ncols + col0);
        const uint4 in = *p;
        uint32_t w[4] = {in.x, in.y, in.z, in.w}, o[4];
        bool changed = false;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint32_t ow = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) {
                int8_t v = (int8_t)(w[q] >> (8 * b));
                if (v < 0) v = 0;
                const int8_t v2 = (int8_t)(v * v);  // int8 arithmetic like the reference (row[rj]*row[rj])
                s1[4 * q + b] += v;
                s2[4 * q + b] += v2;
                ow |= (uint32_t)(uint8_t)(square ? v2 : v) << (8 * b);
            }
            o[q] = ow;
            changed |= ow != w[q];
        }
        if (changed) *p = make_uint4(o[0], o[1], o[2], o[3]);
    }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (sum && s1[k]) atomicAdd(&sum[col0 + k], (double)s1[k]);
        if (sqsum && s2[k]) atomicAdd(&sqsum[col0 + k], (double)s2[k]);
    }
}

int launch_geno_prep(Ctx *c, int8_t *X, size_t rows, size_t ncols, double *sum, double *sqsum, bool square, cudaStream_t st) {
    if (rows == 0 || ncols == 0) return 0;
    if (ncols % 16 == 0 && ((uintptr_t)X % 16) == 0) {
        const int rows_per_slab = 64;
        dim3 g((unsigned)((ncols / 16 + 255) / 256), (unsigned)((rows + rows_per_slab - 1) / rows_per_slab));
        k_geno_prep16<<<g, 256, 0, st>>>(X, rows, ncols, sum, sqsum, square ? 1 : 0, rows_per_slab);
        SFG_LAUNCHED(c, "k_geno_prep16", st);
        return 0;
    }
    const int rows_per_slab = 256;
    dim3 g((unsigned)((ncols + 255) / 256), (unsigned)((rows + rows_per_slab - 1) / rows_per_slab));
    k_geno_prep<<<g, 256, 0, st>>>(X, rows, ncols, sum, sqsum, square ? 1 : 0, rows_per_slab);
    SFG_LAUNCHED(c, "k_geno_prep", st);
    return 0;
}

}  // namespace sfg
