// kernels_mac.cu -- the dominant kernel: output-stationary lazy multiply-accumulate of rotated ciphertext residues with
// NTT-domain plaintext diagonals, fused with the modular reduce (K1 + K2 of SURVEY 2.2; gwas/matmult.go:247-324,
// 343-399, 1154-1168).
//
// Dense-contraction view: for every RNS limb l and coefficient n,
//     CV[col][row] = sum_k  R[k][row] * P[col][k]      (mod q_l)
// with row = (i, c) over the 2s ciphertext polynomials, k = (block row bi, baby step b), col = (giant g, block column bj).
// The reference keeps u128 accumulators in memory under a mutex and streams the diagonals once.  Here the accumulators
// live in registers (TR x TC per thread), the K loop is innermost, P is streamed from HBM exactly once, and the R / P
// tiles of KB consecutive K steps are staged in shared memory by cp.async (LDGSTS, zero-fill for nil diagonals).
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

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int NWAIT>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(NWAIT));
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
constexpr int kStages = 3;  // cp.async ring depth (each stage holds KB K-steps)

struct LimbList {
    int n;
    int idx[kMaxLayoutLimbs];
};

template <int TR, int TC, int CG, int RG, int KB, bool NARROW>
__global__ void __launch_bounds__(kNB *CG *RG, 1)
k_mac(const char *__restrict__ R, const char *__restrict__ P, const int *__restrict__ pidx, int K, int nrows, int ncols, PolyLayout lay,
      LimbList limbs, int N, const LimbConst *__restrict__ lcs, uint64_t *__restrict__ cv, int Lcv) {
    using T = typename std::conditional<NARROW, uint32_t, uint64_t>::type;
    constexpr int CB = CG * TC, RB = RG * TR, THREADS = kNB * CG * RG;
    constexpr int ROWS = RB + CB;                       // smem rows per K-step: R rows then P rows
    constexpr int PARTS = kNB * (int)sizeof(T) / 16;    // 16-byte chunks per row
    constexpr int CHUNKS = KB * ROWS * PARTS;
    constexpr int EPC = 16 / (int)sizeof(T);            // elements per chunk
    extern __shared__ __align__(16) unsigned char sm_raw[];
    T *sm = reinterpret_cast<T *>(sm_raw);              // [kStages][KB][ROWS][kNB]

    const int tid = threadIdx.x;
    const int lane = tid % kNB, cg = (tid / kNB) % CG, rg = tid / (kNB * CG);
    const int col0 = blockIdx.x * CB;
    const int n0 = blockIdx.y * kNB;
    const int l = limbs.idx[blockIdx.z];
    const LimbConst lc = lcs[l];
    const size_t rec = (size_t)lay.bytes;
    const char *Rl = R + lay.off[l] + (size_t)n0 * sizeof(T);
    const char *Pl = P + lay.off[l] + (size_t)n0 * sizeof(T);
    const int nstage = (K + KB - 1) / KB;

    // Per-thread cp.async descriptors, fixed across pipeline stages: chunk j of this thread copies 16 bytes of one row of
    // one K-step of the stage.  R rows advance by a constant stride per stage; P rows look up the record index.
    constexpr int NCH = (CHUNKS + THREADS - 1) / THREADS;
    const char *csrc[NCH];   // R chunks: source pointer for stage 0 ; P chunks: Pl + part*16
    const int *cpi[NCH];     // P chunks: &pidx[col*K + kk] for stage 0 (nullptr for R chunks)
    int cdst[NCH];           // element offset inside a stage buffer, -1 = no chunk
    int ckk[NCH];
    bool cok[NCH];           // row / column inside the problem
#pragma unroll
    for (int j = 0; j < NCH; j++) {
        const int ch = tid + j * THREADS;
        const int part = ch % PARTS, row = (ch / PARTS) % ROWS, kk = ch / (PARTS * ROWS);
        cdst[j] = ch < CHUNKS ? (kk * ROWS + row) * kNB + part * EPC : -1;
        ckk[j] = kk;
        cpi[j] = nullptr;
        if (row < RB) {
            cok[j] = row < nrows;
            csrc[j] = Rl + ((size_t)kk * nrows + (cok[j] ? row : 0)) * rec + part * 16;
        } else {
            const int col = col0 + (row - RB);
            cok[j] = col < ncols;
            csrc[j] = Pl + part * 16;
            cpi[j] = pidx + (size_t)(cok[j] ? col : 0) * K + kk;
        }
    }
    const size_t rstride = (size_t)KB * nrows * rec;

    auto issue = [&](int sidx) {
        T *dst = sm + (size_t)(sidx % kStages) * KB * ROWS * kNB;
        const int k0 = sidx * KB;
#pragma unroll
        for (int j = 0; j < NCH; j++) {
            if (cdst[j] < 0) continue;
            const char *src = R;
            int bytes = 0;
            if (cok[j] && k0 + ckk[j] < K) {
                if (cpi[j] == nullptr) {
                    src = csrc[j] + (size_t)sidx * rstride;
                    bytes = 16;
                } else {
                    const int pi = __ldg(cpi[j] + k0);
                    if (pi >= 0) {
                        src = csrc[j] + (size_t)pi * rec;
                        bytes = 16;
                    }
                }
            }
            cp_async16(dst + cdst[j], src, bytes);
        }
    };

    typename std::conditional<NARROW, AccN, AccW>::type acc[TR][TC];
#pragma unroll
    for (int r = 0; r < TR; r++)
#pragma unroll
        for (int c = 0; c < TC; c++) {
            if constexpr (NARROW) acc[r][c] = AccN{0, 0};
            else acc[r][c] = AccW{0, 0, 0, 0};
        }

#pragma unroll
    for (int st = 0; st < kStages - 1; st++) {
        if (st < nstage) issue(st);
        cp_async_commit();
    }
    for (int sidx = 0; sidx < nstage; sidx++) {
        cp_async_wait<kStages - 2>();
        __syncthreads();
        {   // prefetch stage sidx + kStages - 1 into the slot freed by stage sidx - 1
            const int sn = sidx + kStages - 1;
            if (sn < nstage) issue(sn);
            cp_async_commit();
        }
        const T *st = sm + (size_t)(sidx % kStages) * KB * ROWS * kNB;
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
    }
    cp_async_wait<0>();

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

template <int TR, int TC, int CG, int RG, int KB, bool NARROW>
static int launch_cfg(Ctx *c, const char *R, const char *P, const int *pidx, int K, int nrows, int ncols, const PolyLayout &lay,
                      const LimbList &limbs, uint64_t *cv, int Lcv, cudaStream_t st) {
    constexpr int CB = CG * TC, RB = RG * TR, THREADS = kNB * CG * RG;
    const size_t smem = (size_t)kStages * KB * (RB + CB) * kNB * (NARROW ? 4 : 8);
    SFG_CUDA(c, cudaFuncSetAttribute(k_mac<TR, TC, CG, RG, KB, NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((ncols + CB - 1) / CB, c->N / kNB, limbs.n);
    k_mac<TR, TC, CG, RG, KB, NARROW><<<grid, THREADS, smem, st>>>(R, P, pidx, K, nrows, ncols, lay, limbs, c->N, c->lc, cv, Lcv);
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
#define SFG_MAC(TR, TC, CG, RG, KB, NARROW, LIMBS) launch_cfg<TR, TC, CG, RG, KB, NARROW>(c, Rc, Pc, pidx, K, nrows, ncols, lay, LIMBS, cv, Lcv, st)
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
        if (nrows <= 4) rc = SFG_MAC(4, 2, 16, 1, 2, false, wid);
        else if (nrows <= 8) rc = SFG_MAC(4, 2, 8, 2, 2, false, wid);
        else if (nrows <= 12) rc = SFG_MAC(4, 2, 5, 3, 2, false, wid);
        else if (nrows <= 16) rc = SFG_MAC(4, 2, 4, 4, 2, false, wid);
        else if (nrows <= 20) rc = SFG_MAC(4, 2, 3, 5, 2, false, wid);
        else if (nrows <= 24) rc = SFG_MAC(4, 2, 2, 6, 2, false, wid);
        else if (nrows <= 28) rc = SFG_MAC(4, 2, 2, 7, 2, false, wid);
        else rc = SFG_MAC(4, 2, 2, 8, 2, false, wid);
        if (rc) return rc;
    }
#undef SFG_MAC
    return 0;
}

}  // namespace sfg
