// ntt2.cuh -- register-tiled negacyclic NTT / INTT over one CTA (sm_100a), specialised by limb width.
//
// Semantics are Lattigo's ring.NTT / ring.InvNTT: forward Cooley-Tukey from natural order to the
// bit-reversed evaluation order a(psi^(2 brv(i)+1)), inverse Gentleman-Sande times N^-1; stage `s` has m = 2^s groups and uses
// twiddle NttPsi[m + i] (NttPsiInv[m + i]).  Outputs are canonical, so any exact evaluation order is bit-identical.
//
// Organisation: a transform of 2^logS coefficients held in (padded) shared memory is a sequence of PASSES; a pass performs R <= 4
// consecutive stages on 2^R coefficients held in registers (radix-2^R), so a 2^13-point transform is 4 passes (3 + 3 + 3 + 4 stages)
// with 3 CTA barriers instead of 13.  The last forward pass (first inverse pass) works on 16 CONSECUTIVE coefficients per thread
// and takes its 15 twiddles from a table transposed for coalesced access; polynomials that only feed such a pass can be kept in
// global memory in the matching "TT" order (coefficient 16 P + k stored at k * N/16 + P).
//
// Arithmetic policies (the integer pipe is the bound, B200: IMAD 64 lanes/clk/SM, IMAD.WIDE 32):
//   ArW   q < 2^62 : u64, Shoup multiplication with 64-bit companions, lazy forward butterflies (no conditional subtraction:
//                    values grow by 2q per stage and fit 64 bits for q < 2^57; larger q fall back to Harvey's [0, 4q) form),
//                    Harvey [0, 2q) inverse butterflies.                                  ~16 FMA-pipe slots per butterfly
//   ArD   q < 2^46 : residues kept as exact integers in FP64 (the 64 lanes/clk FP64 pipe is otherwise idle): product by a twiddle
//                    = DMUL + DFMA (exact two-product), quotient by the magic-number round, remainder by DFMA; values stay in
//                    (-q, q) after a multiplication and grow by q per forward stage, so no conditional corrections at all.
//                                                                                        8 FP64 ops per forward butterfly
//   ArN30 q < 2^30 : u32, 32-bit Shoup, Harvey lazy butterflies in [0, 4q) / [0, 2q).      3 IMAD + 4 ALU per butterfly
//   ArN31 q < 2^31 : u32, canonical butterflies (4q does not fit 32 bits).                 3 IMAD + 8 ALU per butterfly
#pragma once
#include "modarith.cuh"

namespace sfg {

enum ArithKind : int { kArW = 0, kArN30 = 1, kArN31 = 2, kArD = 3, kNumArith = 4 };
__host__ __device__ inline int arith_kind(uint64_t q) {
    // ArD up to 2^46 (round 2; was 2^45): PN14QP438's q0 = 2^45 + 2^15 + 1 and the 45-46-bit moduli of PN16QP1761 run on the FP64 pipe
    // (3x the u64 butterfly rate).  Exactness bound, with Y = the largest magnitude a transform value can reach:
    //   mul_lazy(y, w): h = RN(y w), l = y w - h exactly (FMA); k = rint(h * RN(1/q)) in ONE rounding (FMA onto 1.5 * 2^52), so
    //   |k - h/q| <= 1/2 + |h/q| 2^-53 <= 1/2 + Y 2^-53;  h - k q and l are integers below 2^53, so r = fma(-k, q, h) + l is exact,
    //   r = y w (mod q) and |r| <= q (1/2 + Y 2^-53) + Y q 2^-53 = q (1/2 + Y 2^-52).  Needs Y < 2^51 (the magic-number round) and
    //   gives |r| < q as soon as Y <= 2^50.6.  Y <= 2^49 (load_u64 reduces anything larger) + 16 stages of growth by q < 2^46
    //   = 1.5 * 2^50: |r| <= 0.875 q.  Lazy key-product sums add at most 2^3 values below q: far below 2^53.
    return q < (1ULL << 30) ? kArN30 : (q < (1ULL << 31) ? kArN31 : (q < (1ULL << 46) ? kArD : kArW));
}

// Stages per pass: at most kLastR = 4 (16 coefficients in registers per thread).
constexpr int kLastR = 4;
constexpr int kLastE = 1 << kLastR;

// Padded shared-memory index: one pad element per 32 (4-byte classes) / per 16 (8-byte classes).  Unit stride across lanes and
// the stride-16 pattern of the last pass are conflict-free; the address costs 2 instructions (an XOR swizzle that also made the
// 16 x 2 pattern of the lobits == 4 pass conflict-free cost 6 and measured 7 % slower end to end).
template <int BYTES>
__device__ __forceinline__ int sidx(int i) {
    if (BYTES == 8) return i + (i >> 4);
    return i + (i >> 5);
}
__host__ __device__ inline size_t ntt_smem_elems(int S) { return (size_t)S + ((size_t)S >> 4) + 32; }
// "TT" global order of a polynomial of N coefficients: coefficient 16 P + k stored at k * N/16 + P
__device__ __forceinline__ int tt_index(int idx, int N) { return (idx & (kLastE - 1)) * (N >> kLastR) + (idx >> kLastR); }

struct PassPlan {  // stage counts of the passes before the last pass (forward order); sum + kLastR == logS
    int n;
    int R[4];
};
__host__ __device__ inline PassPlan make_pass_plan(int nstages /* = logS - kLastR >= 0 */) {
    PassPlan p{0, {0, 0, 0, 0}};
    const int np = (nstages + kLastR - 1) / kLastR;
    for (int i = 0; i < np; i++) p.R[i] = nstages / np + (i < nstages % np ? 1 : 0);
    p.n = np;
    return p;
}

// ---------------------------------------------------------------------------------------------------------------
// arithmetic policies
// ---------------------------------------------------------------------------------------------------------------
struct ArW {
    static constexpr int kMinBlocks = 1;
    using T = uint64_t;
    using TW = ulonglong2;  // (w, floor(w 2^64 / q))
    static constexpr int kKind = kArW;
    struct C {
        uint64_t q, q2, bred_hi, ninv, ninv_sh;
        int big;  // q >= 2^57: keep forward values in [0, 4q)
    };
    __device__ static __forceinline__ C make(const LimbConst &lc) { return C{lc.q, 2 * lc.q, lc.bred_hi, lc.ninv, lc.ninv_sh, lc.q >= (1ULL << 57)}; }
    __device__ static __forceinline__ T mul_lazy(T y, TW w, const C &c) { return y * w.x - __umul64hi(y, w.y) * c.q; }  // [0, 2q), any y
    // any 64-bit value -> something the forward butterflies accept
    __device__ static __forceinline__ T load_u64(uint64_t x, const C &c) { return c.big ? (x - __umul64hi(x, c.bred_hi) * c.q) : x; }
    __device__ static __forceinline__ void fwd(T &x, T &y, TW w, const C &c) {
        T x0 = x;
        if (c.big) x0 = x0 >= c.q2 ? x0 - c.q2 : x0;
        const T v = mul_lazy(y, w, c);
        x = x0 + v;
        y = x0 - v + c.q2;
    }
    __device__ static __forceinline__ void inv(T &x, T &y, TW w, const C &c) {  // in / out [0, 2q)
        T u = x + y;
        u = u >= c.q2 ? u - c.q2 : u;
        const T d = x - y + c.q2;
        y = mul_lazy(d, w, c);
        x = u;
    }
    __device__ static __forceinline__ T canon(T x, const C &c) {  // any value -> [0, q)
        const T r = x - __umul64hi(x, c.bred_hi) * c.q;
        return r >= c.q ? r - c.q : r;
    }
    __device__ static __forceinline__ T inv_final(T x, const C &c) {  // x * N^-1 mod q, canonical
        const T r = x * c.ninv - __umul64hi(x, c.ninv_sh) * c.q;
        return r >= c.q ? r - c.q : r;
    }
    __device__ static __forceinline__ T from_canon(uint64_t x, const C &) { return x; }  // residue of THIS modulus
    __host__ static TW make_tw(uint64_t w, uint64_t q) { return make_ulonglong2(w, h_shoup(w, q)); }
};

struct ArD {
    static constexpr int kMinBlocks = 1;
    using T = double;
    using TW = double;  // w, plain residue
    static constexpr int kKind = kArD;
    struct C {
        double q, qinv, ninv;
        uint64_t q64, bred_hi;
        int red_inv;  // inverse butterflies: reduce the sum at every stage (q * N would leave the exact-integer range otherwise)
    };
    __device__ static __forceinline__ C make(const LimbConst &lc) {
        return C{(double)lc.q, 1.0 / (double)lc.q, (double)lc.ninv, lc.q, lc.bred_hi, lc.q >= (1ULL << 34)};
    }
    // x mod q in (-q, q) (|x| < 2^50): nearest-integer quotient by the 1.5 * 2^52 magic constant, exact remainder by FMA
    __device__ static __forceinline__ T red(T x, const C &c) {
        const double k = __fma_rn(x, c.qinv, 6755399441055744.0) - 6755399441055744.0;
        return __fma_rn(-k, c.q, x);
    }
    // y * w mod q in (-q, q) for integer-valued |y| < 2^50, 0 <= w < q: (h, l) is the exact product
    __device__ static __forceinline__ T mul_lazy(T y, TW w, const C &c) {
        const double h = y * w;
        const double l = __fma_rn(y, w, -h);
        const double k = __fma_rn(h, c.qinv, 6755399441055744.0) - 6755399441055744.0;
        return __fma_rn(-k, c.q, h) + l;
    }
    __device__ static __forceinline__ T load_u64(uint64_t x, const C &c) {  // any 64-bit value
        if (x >= (1ULL << 49)) x -= __umul64hi(x, c.bred_hi) * c.q64;
        return (double)(long long)x;
    }
    __device__ static __forceinline__ void fwd(T &x, T &y, TW w, const C &c) {
        const T v = mul_lazy(y, w, c);
        y = x - v;
        x = x + v;
    }
    __device__ static __forceinline__ void inv(T &x, T &y, TW w, const C &c) {
        T u = x + y;
        const T d = x - y;
        if (c.red_inv) u = red(u, c);
        y = mul_lazy(d, w, c);
        x = u;
    }
    __device__ static __forceinline__ T canon(T x, const C &c) {  // -> [0, q)
        const T r = red(x, c);
        return r < 0.0 ? r + c.q : r;
    }
    __device__ static __forceinline__ T inv_final(T x, const C &c) {
        const T r = mul_lazy(x, c.ninv, c);
        return r < 0.0 ? r + c.q : r;
    }
    __device__ static __forceinline__ T from_canon(uint64_t x, const C &) { return (double)(long long)x; }
    __host__ static TW make_tw(uint64_t w, uint64_t) { return (double)w; }
};

struct ArN30 {
    static constexpr int kMinBlocks = 2;  // 32-bit classes: 64 registers per thread -> two 512-thread CTAs per SM
    using T = uint32_t;
    using TW = uint2;  // (w, floor(w 2^32 / q))
    static constexpr int kKind = kArN30;
    struct C {
        uint32_t q, q2, ninv, ninv_sh, qinv;
        uint64_t q64, bred_hi;
    };
    __device__ static __forceinline__ C make(const LimbConst &lc) {
        return C{(uint32_t)lc.q, (uint32_t)(2 * lc.q), (uint32_t)lc.ninv, (uint32_t)(lc.ninv_sh >> 32), (uint32_t)lc.qinv, lc.q, lc.bred_hi};
    }
    // y * k mod q in [0, 2q) for any 32-bit y, km = k * 2^32 mod q (32-bit Montgomery form), qinv = q^-1 mod 2^32
    __device__ static __forceinline__ T mul_mont(T y, uint32_t km, const C &c) {
        const uint32_t lo = y * km, hi = __umulhi(y, km);
        uint32_t m;  // opaque to the optimiser: it would otherwise precompute km * qinv for every key word ahead of time and spill them
        asm("mul.lo.u32 %0, %1, %2;" : "=r"(m) : "r"(lo), "r"(c.qinv));
        return hi - __umulhi(m, c.q) + c.q;
    }
    __device__ static __forceinline__ T mul_lazy(T y, TW w, const C &c) { return y * w.x - __umulhi(y, w.y) * c.q; }  // [0, 2q), any y
    __device__ static __forceinline__ T load_u64(uint64_t x, const C &c) {  // any 64-bit value -> [0, 4q)
        return x < 4 * c.q64 ? (uint32_t)x : (uint32_t)(x - __umul64hi(x, c.bred_hi) * c.q64);  // BRedAdd without the final subtraction
    }
    __device__ static __forceinline__ void fwd(T &x, T &y, TW w, const C &c) {  // in / out [0, 4q)
        const T x0 = min(x, x - c.q2);
        const T v = mul_lazy(y, w, c);
        x = x0 + v;
        y = x0 - v + c.q2;
    }
    __device__ static __forceinline__ void inv(T &x, T &y, TW w, const C &c) {  // in / out [0, 2q)
        T u = x + y;
        u = min(u, u - c.q2);
        const T d = x - y + c.q2;
        y = mul_lazy(d, w, c);
        x = u;
    }
    __device__ static __forceinline__ T canon(T x, const C &c) {  // [0, 4q) -> [0, q)
        x = min(x, x - c.q2);
        return min(x, x - c.q);
    }
    __device__ static __forceinline__ T inv_final(T x, const C &c) {
        const T r = x * c.ninv - __umulhi(x, c.ninv_sh) * c.q;
        return min(r, r - c.q);
    }
    __device__ static __forceinline__ T from_canon(uint64_t x, const C &) { return (uint32_t)x; }
    __host__ static TW make_tw(uint64_t w, uint64_t q) { return make_uint2((uint32_t)w, (uint32_t)((w << 32) / q)); }
};

struct ArN31 {
    static constexpr int kMinBlocks = 2;
    using T = uint32_t;
    using TW = uint2;
    static constexpr int kKind = kArN31;
    using C = ArN30::C;
    __device__ static __forceinline__ C make(const LimbConst &lc) { return ArN30::make(lc); }
    __device__ static __forceinline__ T mul_lazy(T y, TW w, const C &c) { return y * w.x - __umulhi(y, w.y) * c.q; }
    __device__ static __forceinline__ T mul_canon(T y, TW w, const C &c) {
        const T r = mul_lazy(y, w, c);
        return min(r, r - c.q);
    }
    __device__ static __forceinline__ T load_u64(uint64_t x, const C &c) {  // -> [0, q)
        const uint32_t r = (uint32_t)(x - __umul64hi(x, c.bred_hi) * c.q64);  // [0, 2q)
        return min(r, r - c.q);
    }
    __device__ static __forceinline__ void fwd(T &x, T &y, TW w, const C &c) {  // canonical in / out
        const T v = mul_canon(y, w, c);
        const T s = x + v, d = x - v;
        x = min(s, s - c.q);
        y = min(d, d + c.q);
    }
    __device__ static __forceinline__ void inv(T &x, T &y, TW w, const C &c) {
        const T s = x + y;
        T d = x - y;
        d = min(d, d + c.q);
        x = min(s, s - c.q);
        y = mul_canon(d, w, c);
    }
    __device__ static __forceinline__ T canon(T x, const C &) { return x; }
    __device__ static __forceinline__ T inv_final(T x, const C &c) {
        const T r = x * c.ninv - __umulhi(x, c.ninv_sh) * c.q;
        return min(r, r - c.q);
    }
    __device__ static __forceinline__ T from_canon(uint64_t x, const C &) { return (uint32_t)x; }
    __host__ static TW make_tw(uint64_t w, uint64_t q) { return ArN30::make_tw(w, q); }
};

// Twiddle tables of one modulus (device pointers, element type A::TW):
//   fwd[m + i]  = NttPsi[m + i]           fwd_last[(2^r - 1 + g) * N/16 + P] = NttPsi[2^(logN-4+r) + (P << r) + g]
//   inv / inv_last: the same for NttPsiInv
struct TwTab {
    const void *fwd, *fwd_last, *inv, *inv_last;
};

// ---------------------------------------------------------------------------------------------------------------
// passes.  A transform works on the slice `sl` (2^logS consecutive coefficients) of a 2^logN ring; local index j is global
// index (sl << logS) + j.  ld(j) / st(j, v) move one coefficient (registers <-> wherever the caller keeps it).
// ---------------------------------------------------------------------------------------------------------------
// TWS: the twiddles of the passes before the last one (NttPsi[0 .. N/16), warp-uniform or nearly so) come from a SHARED-memory copy: with
// 105-209 KB of shared memory per CTA the L1 is ~20 KB, every __ldg of a twiddle is an L2 round trip, and ncu shows the first multiply of
// every butterfly group waiting for it (k_ks_inner2<ArD>: DMUL 6 % of the instructions, 13 % of the stall samples).
template <class A, int R, bool LAST, class LD, class ST, bool TWS = false>
__device__ __forceinline__ void fwd_pass(int s0, int logN, int logS, int sl, const typename A::TW *__restrict__ tw,
                                         const typename A::C &c, LD ld, ST st) {
    using T = typename A::T;
    using TW = typename A::TW;
    constexpr int E = 1 << R;
    const int lobits = LAST ? 0 : logN - s0 - R;
    const int nsub = 1 << (logS - R);
    for (int p = threadIdx.x; p < nsub; p += blockDim.x) {
        const int hl = p >> lobits, lo = p & ((1 << lobits) - 1);
        const int base = (hl << (lobits + R)) | lo;
        const int hi = ((sl << logS) | base) >> (lobits + R);  // global group index at stage s0
        T v[E];
#pragma unroll
        for (int k = 0; k < E; k++) v[k] = ld(base + (k << lobits), k);
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int half = E >> (r + 1);
#pragma unroll
            for (int g = 0; g < (1 << r); g++) {
                TW w;
                if (LAST) w = __ldg(tw + (size_t)((1 << r) - 1 + g) * (size_t)(1 << (logN - kLastR)) + hi);
                else if (TWS) w = tw[(1 << (s0 + r)) + (hi << r) + g];
                else w = __ldg(tw + (1 << (s0 + r)) + (hi << r) + g);
#pragma unroll
                for (int j = 0; j < half; j++) A::fwd(v[g * 2 * half + j], v[g * 2 * half + j + half], w, c);
            }
        }
#pragma unroll
        for (int k = 0; k < E; k++) st(base + (k << lobits), v[k], k);
    }
}

template <class A, int R, bool LAST, class LD, class ST>
__device__ __forceinline__ void inv_pass(int s0, int logN, int logS, int sl, const typename A::TW *__restrict__ tw,
                                         const typename A::C &c, LD ld, ST st) {
    using T = typename A::T;
    using TW = typename A::TW;
    constexpr int E = 1 << R;
    const int lobits = LAST ? 0 : logN - s0 - R;
    const int nsub = 1 << (logS - R);
    for (int p = threadIdx.x; p < nsub; p += blockDim.x) {
        const int hl = p >> lobits, lo = p & ((1 << lobits) - 1);
        const int base = (hl << (lobits + R)) | lo;
        const int hi = ((sl << logS) | base) >> (lobits + R);
        T v[E];
#pragma unroll
        for (int k = 0; k < E; k++) v[k] = ld(base + (k << lobits), k);
#pragma unroll
        for (int r = R - 1; r >= 0; r--) {
            const int half = E >> (r + 1);
#pragma unroll
            for (int g = 0; g < (1 << r); g++) {
                TW w;
                if (LAST) w = __ldg(tw + (size_t)((1 << r) - 1 + g) * (size_t)(1 << (logN - kLastR)) + hi);
                else w = __ldg(tw + (1 << (s0 + r)) + (hi << r) + g);
#pragma unroll
                for (int j = 0; j < half; j++) A::inv(v[g * 2 * half + j], v[g * 2 * half + j + half], w, c);
            }
        }
#pragma unroll
        for (int k = 0; k < E; k++) st(base + (k << lobits), v[k], k);
    }
}

// run-time stage count -> compile-time radix
template <class A, bool INV, class LD, class ST, bool TWS = false>
__device__ __forceinline__ void mid_pass(int R, int s0, int logN, int logS, int sl, const typename A::TW *tw, const typename A::C &c, LD ld,
                                         ST st) {
    switch (R) {
        case 1: INV ? inv_pass<A, 1, false>(s0, logN, logS, sl, tw, c, ld, st) : fwd_pass<A, 1, false, LD, ST, TWS>(s0, logN, logS, sl, tw, c, ld, st); break;
        case 2: INV ? inv_pass<A, 2, false>(s0, logN, logS, sl, tw, c, ld, st) : fwd_pass<A, 2, false, LD, ST, TWS>(s0, logN, logS, sl, tw, c, ld, st); break;
        case 3: INV ? inv_pass<A, 3, false>(s0, logN, logS, sl, tw, c, ld, st) : fwd_pass<A, 3, false, LD, ST, TWS>(s0, logN, logS, sl, tw, c, ld, st); break;
        default: INV ? inv_pass<A, 4, false>(s0, logN, logS, sl, tw, c, ld, st) : fwd_pass<A, 4, false, LD, ST, TWS>(s0, logN, logS, sl, tw, c, ld, st); break;
    }
}

// Forward transform of slice `sl`.  ld0(j, k) supplies the input of local coefficient j in the policy's input range, AFTER the
// first (logN - logS) stages when the ring is sliced (see slice_input()).  The passes before the last one go through the padded
// shared-memory array `s`; the results of the last pass are handed to fin(j, value, k), k = j mod 16 the register index -- 16 consecutive j per thread, lazy range.
// number of twiddles the passes before the last one can touch: NttPsi[0 .. 2^(logN - kLastR))
__host__ __device__ inline int ntt_mid_twiddles(int logN) { return logN > kLastR ? 1 << (logN - kLastR) : 1; }
// cooperative copy of those twiddles into shared memory (call once per CTA and modulus, then __syncthreads())
template <class A>
__device__ __forceinline__ void ntt_stage_twiddles(typename A::TW *tw_s, const TwTab &tab, int logN) {
    const typename A::TW *g = reinterpret_cast<const typename A::TW *>(tab.fwd);
    for (int i = threadIdx.x; i < ntt_mid_twiddles(logN); i += blockDim.x) tw_s[i] = __ldg(g + i);
}

template <class A, class LD, class FIN>
__device__ __forceinline__ void ntt_forward(typename A::T *s, int logN, int logS, int sl, const PassPlan &plan, const TwTab &tab,
                                            const typename A::C &c, LD ld0, FIN fin, const typename A::TW *tw_s = nullptr) {
    using T = typename A::T;
    using TW = typename A::TW;
    const TW *tw = reinterpret_cast<const TW *>(tab.fwd);
    auto lds = [=](int j, int) { return s[sidx<sizeof(T)>(j)]; };
    auto sts = [=](int j, T v, int) { s[sidx<sizeof(T)>(j)] = v; };
    int s0 = logN - logS;
    if (plan.n == 0) {
        for (int j = threadIdx.x; j < (1 << logS); j += blockDim.x) sts(j, ld0(j, 0), 0);
    }
#pragma unroll
    for (int i = 0; i < plan.n; i++) {
        if (tw_s) {  // twiddles staged in shared memory (ntt_stage_twiddles)
            if (i == 0) mid_pass<A, false, LD, decltype(sts), true>(plan.R[i], s0, logN, logS, sl, tw_s, c, ld0, sts);
            else mid_pass<A, false, decltype(lds), decltype(sts), true>(plan.R[i], s0, logN, logS, sl, tw_s, c, lds, sts);
        } else {
            if (i == 0) mid_pass<A, false>(plan.R[i], s0, logN, logS, sl, tw, c, ld0, sts);
            else mid_pass<A, false>(plan.R[i], s0, logN, logS, sl, tw, c, lds, sts);
        }
        s0 += plan.R[i];
        __syncthreads();
    }
    if (plan.n == 0) __syncthreads();
    fwd_pass<A, kLastR, true>(s0, logN, logS, sl, reinterpret_cast<const TW *>(tab.fwd_last), c, lds, fin);
}

// Inverse transform of a whole ring (logS == logN).  ld_last(j, k) supplies the canonical input of coefficient j for the first
// (stride-16) pass; fin(j, v, k) receives v = canonical coefficient j of the result (times N^-1), unit-stride across lanes.
template <class A, class LD, class FIN>
__device__ __forceinline__ void ntt_inverse(typename A::T *s, int logN, const PassPlan &plan, const TwTab &tab, const typename A::C &c,
                                            LD ld_last, FIN fin) {
    using T = typename A::T;
    using TW = typename A::TW;
    const TW *tw = reinterpret_cast<const TW *>(tab.inv);
    auto lds = [=](int j, int) { return s[sidx<sizeof(T)>(j)]; };
    auto sts = [=](int j, T v, int) { s[sidx<sizeof(T)>(j)] = v; };
    int s0 = logN - kLastR;
    inv_pass<A, kLastR, true>(s0, logN, logN, 0, reinterpret_cast<const TW *>(tab.inv_last), c, ld_last, sts);
    __syncthreads();
#pragma unroll
    for (int i = plan.n - 1; i >= 0; i--) {
        s0 -= plan.R[i];
        if (i == 0) {
            auto fin2 = [&](int j, T v, int k) { fin(j, A::inv_final(v, c), k); };
            mid_pass<A, true>(plan.R[i], s0, logN, logN, 0, tw, c, lds, fin2);
        } else {
            mid_pass<A, true>(plan.R[i], s0, logN, logN, 0, tw, c, lds, sts);
            __syncthreads();
        }
    }
    if (plan.n == 0) {
        for (int j = threadIdx.x; j < (1 << logN); j += blockDim.x) fin(j, A::inv_final(lds(j, 0), c), 0);
    }
}

// Inverse transform of slice `sl` (2^logS consecutive coefficients) of a 2^logN ring: the Gentleman-Sande stages logN-1 .. logN-logS, i.e.
// everything that stays inside the slice; the remaining logN-logS stages pair coefficients of different slices and run as one
// global-memory pass (kernels_ntt.cu: four-step transforms of rings larger than one CTA).  fin(j, v, k) receives the lazy value of local
// coefficient j (NOT multiplied by N^-1).
template <class A, class LD, class FIN>
__device__ __forceinline__ void ntt_inverse_slice(typename A::T *s, int logN, int logS, int sl, const PassPlan &plan, const TwTab &tab,
                                                  const typename A::C &c, LD ld_last, FIN fin) {
    using T = typename A::T;
    using TW = typename A::TW;
    const TW *tw = reinterpret_cast<const TW *>(tab.inv);
    auto lds = [=](int j, int) { return s[sidx<sizeof(T)>(j)]; };
    auto sts = [=](int j, T v, int) { s[sidx<sizeof(T)>(j)] = v; };
    int s0 = logN - kLastR;
    inv_pass<A, kLastR, true>(s0, logN, logS, sl, reinterpret_cast<const TW *>(tab.inv_last), c, ld_last, sts);
    __syncthreads();
#pragma unroll
    for (int i = plan.n - 1; i >= 0; i--) {
        s0 -= plan.R[i];
        if (i == 0) {
            mid_pass<A, true>(plan.R[i], s0, logN, logS, sl, tw, c, lds, fin);
        } else {
            mid_pass<A, true>(plan.R[i], s0, logN, logS, sl, tw, c, lds, sts);
            __syncthreads();
        }
    }
    if (plan.n == 0) {
        for (int j = threadIdx.x; j < (1 << logS); j += blockDim.x) fin(j, lds(j, 0), 0);
    }
}

}  // namespace sfg
