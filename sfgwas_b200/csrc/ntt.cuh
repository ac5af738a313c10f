// ntt.cuh -- negacyclic RNS NTT / INTT building blocks that run on one CTA over shared memory.
//
// Semantics follow Lattigo ring.NTT / ring.InvNTT ([UNVERIFIED vs fork]):
//   forward : Cooley-Tukey, input natural order, output position i holds a(psi^(2*brv(i)+1)), canonical [0,q)
//   inverse : Gentleman-Sande, then multiplication by N^-1, canonical [0,q)
// Twiddles: tw[0][m+i] = psi^brv-ordered powers (Lattigo NttPsi[m+i], but kept in plain form),
//           tw[1] = their Shoup companions floor(w*2^64/q); tw[2], tw[3] the same for psi^-1 (NttPsiInv).
// Because outputs are canonical any exact algorithm with the same psi is bit-identical to Lattigo's.
#pragma once
#include "modarith.cuh"

namespace sfg {

// Per-limb twiddle tables: base + (limb*4 + k)*N, k = 0: w, 1: w_shoup, 2: winv, 3: winv_shoup.
struct NttTab {
    const uint64_t *w, *wsh, *wi, *wish;
};
__device__ __forceinline__ NttTab ntt_tab(const uint64_t *base, int limb, int N) {
    const uint64_t *p = base + (size_t)limb * 4 * N;
    return NttTab{p, p + N, p + 2 * N, p + 3 * N};
}

// Forward stages m = m0, 2*m0, ... , S/2 * m0 ... on a sub-block of S = 2^logS coefficients held in smem.
// `blk` is the index of this sub-block among the N/S sub-blocks (0 when S == N), m0 = N/S.
// All threads of the CTA must call; ends with __syncthreads().
__device__ __forceinline__ void ntt_fwd_smem(uint64_t *s, int logS, int m0, int blk, const NttTab &tab, uint64_t q) {
    const int S = 1 << logS;
    int t = S;
    int logt = logS;
    for (int ml = 1; ml < S; ml <<= 1) {  // ml = local number of groups
        t >>= 1;
        logt -= 1;
        const int mg = ml * m0;           // global m
        const int ibase = blk * ml;       // global index of local group 0
        for (int b = threadIdx.x; b < (S >> 1); b += blockDim.x) {
            const int i = b >> logt;
            const int j = b & (t - 1);
            const int idx = (i << (logt + 1)) + j;
            const uint64_t w = __ldg(tab.w + mg + ibase + i), wsh = __ldg(tab.wsh + mg + ibase + i);
            const uint64_t U = s[idx];
            const uint64_t V = mul_shoup(s[idx + t], w, wsh, q);
            s[idx] = add_mod(U, V, q);
            s[idx + t] = sub_mod(U, V, q);
        }
        __syncthreads();
    }
}

// Inverse stages with t = 1, 2, ..., S/2 on a sub-block of S coefficients (Gentleman-Sande).
// If `scale` the result is multiplied by N^-1 (only valid when S == N).
__device__ __forceinline__ void ntt_inv_smem(uint64_t *s, int logS, int N, int blk, const NttTab &tab, const LimbConst &lc,
                                             bool scale) {
    const int S = 1 << logS;
    const uint64_t q = lc.q;
    int t = 1, logt = 0;
    int hg = N >> 1;  // global number of groups at this stage
    for (int hl = S >> 1; hl >= 1; hl >>= 1) {  // local number of groups
        const int ibase = blk * hl;
        for (int b = threadIdx.x; b < (S >> 1); b += blockDim.x) {
            const int i = b >> logt;
            const int j = b & (t - 1);
            const int idx = (i << (logt + 1)) + j;
            const uint64_t w = __ldg(tab.wi + hg + ibase + i), wsh = __ldg(tab.wish + hg + ibase + i);
            const uint64_t U = s[idx], V = s[idx + t];
            s[idx] = add_mod(U, V, q);
            s[idx + t] = mul_shoup(sub_mod(U, V, q), w, wsh, q);
        }
        __syncthreads();
        t <<= 1;
        logt += 1;
        hg >>= 1;
    }
    if (scale) {
        for (int k = threadIdx.x; k < S; k += blockDim.x) s[k] = mul_shoup(s[k], lc.ninv, lc.ninv_sh, q);
        __syncthreads();
    }
}

}  // namespace sfg
