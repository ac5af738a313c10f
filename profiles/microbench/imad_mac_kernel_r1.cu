// NOT BUILT: the round-1 IMAD.WIDE multiply-accumulate kernel that k_mac_tc (sfgwas_b200/csrc/kernels_mactc.cu) replaced; kept as the
// design record behind profiles/r1_baseline (89.8 ms on config 2 vs 13.4 ms on the tensor cores).
// kernels_mac.cu -- the dominant kernel: output-stationary lazy multiply-accumulate of rotated ciphertext residues with
// NTT-domain plaintext diagonals, fused with the modular reduce (K1 + K2 of SURVEY 2.2; gwas/matmult.go:247-324,
// 343-399, 1154-1168).
//
// Dense-contraction view: for every RNS limb l and coefficient n,
//     CV[col][row] = sum_k  R[k][row] * P[col][k]      (mod q_l)
// with row = (i, c) over the 2s ciphertext polynomials, k = (block row bi, baby step b), col = (giant g, block column bj).
// The reference keeps u128 accumulators in memory under a mutex and streams the diagonals once.  Here the accumulators
// live in registers (TR x TC per thread), the K loop is innermost, P is streamed from HBM exactly once, and the R / P
// tiles of KB consecutive K steps are staged in shared memory by producer warps (cp.async / LDGSTS completing on
// mbarriers); the consumer warps execute only LDS + MAC (no address math, no CTA-wide barrier).
//
// Only the canonical residue sum_k a_k*b_k mod q_l is observable, so the arithmetic is specialised by
// limb width -- measured on B200 (profiles/microbench/mac_rate.cu): IMAD.WIDE.U32 issues at 32 lanes/clk/SM, so
//   wide  limbs (q >= 2^32): u64 operands, 64x64->128 as 4 IMAD.WIDE into even/odd accumulators + 2 carry adds
//                            (exactly the reference's wrap-around u128 sum, Montgomery-reduced at the end)  ~7 MAC/clk/SM
//   narrow limbs (q < 2^32): u32 operands kept PACKED in HBM and shared memory, 32x32->64 as ONE IMAD.WIDE into a
//                            96-bit accumulator + 1 carry add, plain (non-Montgomery) residues               ~25-30 MAC/clk/SM
//
//   CTA tile : NB = 32 consecutive coefficients  x  CB = CG*TC columns  x  RG*TR rows, KB K-steps per pipeline stage
//   warp     : lanes = the 32 coefficients  -> every shared-memory read is one conflict-free row
#include "kernels.h"

namespace sfg {

// ---- mbarrier / bulk-copy (TMA) primitives ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 16-byte global -> shared async copy (LDGSTS) with zero-fill when src_bytes == 0
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem), "r"(src_bytes) : "memory");
}
// the executing thread arrives on `bar` once all its prior cp.async operations have completed (no pending-count increment)
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// ---- accumulators ----
// The 32-bit halves are kept in 64-bit PTX registers ("l" constraints) so that ptxas allocates aligned pairs and fuses
// mad.lo.cc / madc.hi.cc into one IMAD.WIDE.U32 with carry-out without register moves (checked with cuobjdump -sass).
struct AccW {  // e23:e01 collects a0*b0 and a1*b1 (128 bit); o2:o01 (weight 2^32) collects the two cross terms (96 bit)
    uint64_t e01, e23, o01;
    uint32_t o2;
};
__device__ __forceinline__ void mac_wide(AccW &A, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
    asm("{\n\t.reg .u32 x0, x1, x2, x3, y0, y1;\n\t"
        "mov.b64 {x0, x1}, %0;\n\tmov.b64 {x2, x3}, %1;\n\tmov.b64 {y0, y1}, %2;\n\t"
        "mad.lo.cc.u32 x0, %4, %6, x0;\n\tmadc.hi.cc.u32 x1, %4, %6, x1;\n\t"
        "madc.lo.cc.u32 x2, %5, %7, x2;\n\tmadc.hi.u32 x3, %5, %7, x3;\n\t"
        "mad.lo.cc.u32 y0, %4, %7, y0;\n\tmadc.hi.cc.u32 y1, %4, %7, y1;\n\taddc.u32 %3, %3, 0;\n\t"
        "mad.lo.cc.u32 y0, %5, %6, y0;\n\tmadc.hi.cc.u32 y1, %5, %6, y1;\n\taddc.u32 %3, %3, 0;\n\t"
        "mov.b64 %0, {x0, x1};\n\tmov.b64 %1, {x2, x3};\n\tmov.b64 %2, {y0, y1};\n\t}"
        : "+l"(A.e01), "+l"(A.e23), "+l"(A.o01), "+r"(A.o2)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}
// the reference's u128 accumulator value (mod 2^128), then ReduceAndAddUint128 + eval.Reduce
__device__ __forceinline__ uint64_t finish_wide(const AccW &A, const LimbConst &lc) {
    u128 t;
    t.lo = A.e01;
    t.hi = A.e23;
    const uint64_t add_lo = A.o01 << 32;
    const uint64_t add_hi = ((uint64_t)A.o2 << 32) | (A.o01 >> 32);
    t.lo += add_lo;
    t.hi += add_hi + (t.lo < add_lo);
    return mred128(t, lc);
}
struct AccN {
    uint64_t e01;
    uint32_t e2;
};
__device__ __forceinline__ void mac_narrow(AccN &A, uint32_t a, uint32_t b) {
    asm("{\n\t.reg .u32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\t"
        "mad.lo.cc.u32 lo, %2, %3, lo;\n\tmadc.hi.cc.u32 hi, %2, %3, hi;\n\taddc.u32 %1, %1, 0;\n\t"
        "mov.b64 %0, {lo, hi};\n\t}"
        : "+l"(A.e01), "+r"(A.e2)
        : "r"(a), "r"(b));
}
// (e2*2^64 + e01) mod q for q < 2^32, r32 = 2^32 mod q
__device__ __forceinline__ uint64_t finish_narrow(const AccN &A, const LimbConst &lc, uint64_t r32) {
    const uint64_t hi = bred_add(((uint64_t)A.e2 << 32) | (A.e01 >> 32), lc);  // < q < 2^32
    return bred_add(hi * r32 + (uint32_t)A.e01, lc);
}

constexpr int kNB = 32;     // coefficients per CTA
constexpr int kStages = 4;  // pipeline depth (each stage holds KB K-steps)

struct LimbList {
    int n;
    int idx[kMaxLayoutLimbs];
};

// Warp-specialised: warps 0 .. CG*RG-1 are consumers (LDS + MAC only); the last NPW warps are producers that fill the
// stage ring with 16-byte cp.async (LDGSTS, zero-fill for nil diagonals / padding) and signal full[s] through
// cp.async.mbarrier.arrive; consumers release a stage through empty[s].  No CTA-wide barrier inside the K loop.
// (A first version issued one cp.async.bulk (TMA) per 128-256 B row: 3.4x slower -- per-copy overhead dominates.)
template <int TR, int TC, int CG, int RG, int KB, int NPW, bool NARROW>
__global__ void __launch_bounds__(kNB *(CG *RG + NPW), 1)
k_mac(const char *__restrict__ R, const char *__restrict__ P, const int *__restrict__ pidx, int K, int nrows, int ncols, PolyLayout lay,
      LimbList limbs, int N, const LimbConst *__restrict__ lcs, uint64_t *__restrict__ cv, int Lcv) {
    using T = typename std::conditional<NARROW, uint32_t, uint64_t>::type;
    constexpr int CB = CG * TC, RB = RG * TR, NCW = CG * RG, NPT = 32 * NPW;
    constexpr int ROWS = RB + CB;                      // smem rows per K-step: R rows then P rows
    constexpr int ROWBYTES = kNB * (int)sizeof(T);
    constexpr int PARTS = ROWBYTES / 16;               // 16-byte chunks per row (8 or 16)
    constexpr int SROWS = KB * ROWS;                   // rows per stage
    constexpr int CHUNKS = SROWS * PARTS;
    extern __shared__ __align__(128) unsigned char sm_raw[];
    T *sm = reinterpret_cast<T *>(sm_raw);             // [kStages][KB][ROWS][kNB]
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(sm_raw + (size_t)kStages * SROWS * ROWBYTES);
    uint64_t *empty_bar = full_bar + kStages;
    uint64_t *rtab = empty_bar + kStages;              // [SROWS] stage-invariant row descriptors
    //   R row : (kk*nrows + row)*rec  (byte offset, < 2^62)           ; 0xFFFF... = padding row
    //   P row : (1<<63) | address of pidx[col*K + kk]                  ; 0xFFFF... = column outside the problem

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int col0 = blockIdx.x * CB;
    const int n0 = blockIdx.y * kNB;
    const int l = limbs.idx[blockIdx.z];
    const size_t rec = (size_t)lay.bytes;
    const int nstage = (K + KB - 1) / KB;

    if (tid == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(&full_bar[s], NPT);
            mbar_init(&empty_bar[s], NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int rr = tid; rr < SROWS; rr += blockDim.x) {
        const int kk = rr / ROWS, row = rr % ROWS;
        uint64_t e = ~0ULL;
        if (row < RB) {
            if (row < nrows) e = ((uint64_t)kk * nrows + row) * rec;
        } else {
            const int col = col0 + (row - RB);
            if (col < ncols) e = (1ULL << 63) | (uint64_t)(uintptr_t)(pidx + (size_t)col * K + kk);
        }
        rtab[rr] = e;
    }
    __syncthreads();

    if (wid >= NCW) {
        // ===================== producer warps =====================
        const int ptid = tid - NCW * 32;
        const int part = ptid % PARTS, rsub = ptid / PARTS;
        constexpr int RSTEP = NPT / PARTS;             // rows covered per pass of the producer threads
        const char *Rl = R + lay.off[l] + (size_t)n0 * sizeof(T) + part * 16;
        const char *Pl = P + lay.off[l] + (size_t)n0 * sizeof(T) + part * 16;
        const size_t rstride = (size_t)KB * nrows * rec;
        for (int sidx = 0; sidx < nstage; sidx++) {
            const int s = sidx % kStages;
            mbar_wait(&empty_bar[s], ((uint32_t)(sidx / kStages) & 1u) ^ 1u);  // slot free (immediate in the first round)
            unsigned char *stage = sm_raw + (size_t)s * SROWS * ROWBYTES + part * 16;
            const int k0 = sidx * KB;
            const char *Rs = Rl + (size_t)sidx * rstride;
            // phase 1: resolve every source address of this thread's rows (independent pidx loads overlap) ...
            constexpr int NIT = (SROWS + RSTEP - 1) / RSTEP;
            const char *src[NIT];
#pragma unroll
            for (int j = 0; j < NIT; j++) {
                const int rr = rsub + j * RSTEP;
                src[j] = nullptr;
                if (rr < SROWS) {
                    const uint64_t e = rtab[rr];
                    const int kk = rr / ROWS;
                    if (e != ~0ULL && k0 + kk < K) {
                        if (e >> 63) {
                            const int pi = __ldg(reinterpret_cast<const int *>(e & ~(1ULL << 63)) + k0);
                            if (pi >= 0) src[j] = Pl + (size_t)pi * rec;
                        } else {
                            src[j] = Rs + e;
                        }
                    }
                }
            }
            // ... phase 2: issue the copies (zero-fill when there is no source)
#pragma unroll
            for (int j = 0; j < NIT; j++) {
                const int rr = rsub + j * RSTEP;
                if (rr < SROWS) cp_async16(stage + (size_t)rr * ROWBYTES, src[j] ? src[j] : R, src[j] ? 16 : 0);
            }
            cp_async_mbar_arrive(&full_bar[s]);
        }
        return;
    }

    // ===================== consumer warps =====================
    const int cg = wid % CG, rg = wid / CG;
    const LimbConst lc = lcs[l];
    typename std::conditional<NARROW, AccN, AccW>::type acc[TR][TC];
#pragma unroll
    for (int r = 0; r < TR; r++)
#pragma unroll
        for (int c = 0; c < TC; c++) {
            if constexpr (NARROW) acc[r][c] = AccN{0, 0};
            else acc[r][c] = AccW{0, 0, 0, 0};
        }

    for (int sidx = 0; sidx < nstage; sidx++) {
        const int s = sidx % kStages;
        mbar_wait(&full_bar[s], (uint32_t)(sidx / kStages) & 1u);
        const T *st = sm + (size_t)s * SROWS * kNB;
#pragma unroll
        for (int kk = 0; kk < KB; kk++) {
            const T *sk = st + (size_t)kk * ROWS * kNB;
            T a[TR], b[TC];
#pragma unroll
            for (int r = 0; r < TR; r++) a[r] = sk[(rg * TR + r) * kNB + lane];
#pragma unroll
            for (int c = 0; c < TC; c++) b[c] = sk[(RB + cg * TC + c) * kNB + lane];
#pragma unroll
            for (int r = 0; r < TR; r++)
#pragma unroll
                for (int c = 0; c < TC; c++) {
                    if constexpr (NARROW) mac_narrow(acc[r][c], a[r], b[c]);
                    else mac_wide(acc[r][c], (uint32_t)a[r], (uint32_t)(a[r] >> 32), (uint32_t)b[c], (uint32_t)(b[c] >> 32));
                }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

    // epilogue: reduce and store canonical residues, cv[col][row][l][n] (u64)
    uint64_t r32 = 0;
    if constexpr (NARROW) r32 = bred_add(1ULL << 32, lc);
    const size_t LN = (size_t)Lcv * N;
#pragma unroll
    for (int c = 0; c < TC; c++) {
        const int col = col0 + cg * TC + c;
        if (col >= ncols) continue;
#pragma unroll
        for (int r = 0; r < TR; r++) {
            const int row = rg * TR + r;
            if (row >= nrows) continue;
            uint64_t v;
            if constexpr (NARROW) v = finish_narrow(acc[r][c], lc, r32);
            else v = finish_wide(acc[r][c], lc);
            cv[((size_t)col * nrows + row) * LN + (size_t)l * N + n0 + lane] = v;
        }
    }
}

template <int TR, int TC, int CG, int RG, int KB, int NPW, bool NARROW>
static int launch_cfg(Ctx *c, const char *R, const char *P, const int *pidx, int K, int nrows, int ncols, const PolyLayout &lay,
                      const LimbList &limbs, uint64_t *cv, int Lcv, cudaStream_t st) {
    constexpr int CB = CG * TC, RB = RG * TR, THREADS = kNB * (CG * RG + NPW);
    const size_t smem = (size_t)kStages * KB * (RB + CB) * kNB * (NARROW ? 4 : 8) + (2 * kStages + KB * (RB + CB)) * sizeof(uint64_t);
    SFG_CUDA(c, cudaFuncSetAttribute(k_mac<TR, TC, CG, RG, KB, NPW, NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((ncols + CB - 1) / CB, c->N / kNB, limbs.n);
    k_mac<TR, TC, CG, RG, KB, NPW, NARROW><<<grid, THREADS, smem, st>>>(R, P, pidx, K, nrows, ncols, lay, limbs, c->N, c->lc, cv, Lcv);
    SFG_LAUNCHED(c, "k_mac", st);
    return 0;
}

int launch_mac(Ctx *c, const void *R, const void *P, const int *pidx, int K, int nrows, int ncols, const PolyLayout &lay,
               uint64_t *cv, cudaStream_t st) {
    if (K <= 0 || nrows <= 0 || ncols <= 0) return 0;
    if (c->N % kNB) SFG_FAIL(c, "N must be a multiple of %d", kNB);
    if (nrows > 32) SFG_FAIL(c, "more than 16 ciphertext rows per MAC launch (nrows=%d): split the call", nrows);
    LimbList nar{0, {}}, wid{0, {}};
    for (int l = 0; l < lay.nl; l++) {
        if (lay.es[l] == 4) nar.idx[nar.n++] = l;
        else wid.idx[wid.n++] = l;
    }
    const char *Rc = (const char *)R, *Pc = (const char *)P;
    const int Lcv = lay.nl;
#define SFG_MAC(TR, TC, CG, RG, KB, NARROW, LIMBS) launch_cfg<TR, TC, CG, RG, KB, 4, NARROW>(c, Rc, Pc, pidx, K, nrows, ncols, lay, LIMBS, cv, Lcv, st)
    if (nar.n) {
        int rc;
        if (nrows <= 8) rc = SFG_MAC(8, 2, 16, 1, 4, true, nar);
        else if (nrows <= 10) rc = SFG_MAC(10, 2, 16, 1, 4, true, nar);
        else if (nrows <= 16) rc = SFG_MAC(8, 2, 8, 2, 4, true, nar);
        else if (nrows <= 20) rc = SFG_MAC(10, 2, 8, 2, 4, true, nar);
        else if (nrows <= 24) rc = SFG_MAC(8, 2, 5, 3, 4, true, nar);
        else if (nrows <= 30) rc = SFG_MAC(10, 2, 5, 3, 4, true, nar);
        else rc = SFG_MAC(8, 2, 4, 4, 4, true, nar);
        if (rc) return rc;
    }
    if (wid.n) {
        int rc;
        if (nrows <= 4) rc = SFG_MAC(4, 2, 16, 1, 4, false, wid);
        else if (nrows <= 8) rc = SFG_MAC(4, 2, 8, 2, 4, false, wid);
        else if (nrows <= 12) rc = SFG_MAC(4, 2, 5, 3, 4, false, wid);
        else if (nrows <= 16) rc = SFG_MAC(4, 2, 4, 4, 4, false, wid);
        else if (nrows <= 20) rc = SFG_MAC(4, 2, 3, 5, 4, false, wid);
        else if (nrows <= 24) rc = SFG_MAC(4, 2, 2, 6, 4, false, wid);
        else if (nrows <= 28) rc = SFG_MAC(4, 2, 2, 7, 4, false, wid);
        else rc = SFG_MAC(4, 2, 2, 8, 4, false, wid);
        if (rc) return rc;
    }
#undef SFG_MAC
    return 0;
}

}  // namespace sfg
