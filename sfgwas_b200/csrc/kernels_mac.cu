// kernels_mac.cu -- the dominant kernel: output-stationary lazy multiply-accumulate of rotated ciphertext residues with
// NTT-domain plaintext diagonals, fused with the Montgomery reduce (K1 + K2 of SURVEY 2.2; gwas/matmult.go:247-324,
// 343-399, 1154-1168).
//
// Dense-contraction view: for every RNS limb l and coefficient n,
//     CV[col][row] = sum_k  R[k][row] * P[col][k]      (mod q_l, via 128-bit lazy accumulation)
// with row = (i, c) over the 2s ciphertext polynomials, k = (block row bi, baby step b) and col = (giant g, block column bj).
// The reference keeps the u128 accumulators in memory under a mutex and streams the diagonals once; here the
// accumulators live in registers (TR x TC per thread), the K loop is innermost, P is streamed from HBM exactly once and
// the R tile of every K step is staged in shared memory with cp.async and shared by all columns of the CTA.
//
//   CTA tile : NB = 32 consecutive coefficients  x  CB = CG*TC columns  x  RG*TR rows
//   thread   : 1 coefficient, TR rows, TC columns  (u128 accumulators: 4*TR*TC registers)
//   warp     : lanes = the 32 coefficients  -> every shared-memory read is a conflict-free 256 B row
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

constexpr int kNB = 32;      // coefficients per CTA
constexpr int kStages = 4;   // cp.async ring depth

template <int TR, int TC, int CG, int RG>
__global__ void __launch_bounds__(kNB *CG *RG, 1)
k_mac(const uint64_t *__restrict__ R, const uint64_t *__restrict__ P, const long long *__restrict__ poff, int K, int nrows,
      int ncols, int L, int N, const LimbConst *__restrict__ lcs, uint64_t *__restrict__ cv) {
    constexpr int CB = CG * TC, RB = RG * TR, THREADS = kNB * CG * RG;
    constexpr int ROWS = RB + CB;                    // smem rows per stage: R rows then P rows
    constexpr int CHUNKS = ROWS * (kNB * 8 / 16);    // 16-byte chunks per stage
    extern __shared__ __align__(16) uint64_t sm[];   // [kStages][ROWS][kNB]

    const int tid = threadIdx.x;
    const int lane = tid % kNB, cg = (tid / kNB) % CG, rg = tid / (kNB * CG);
    const int col0 = blockIdx.x * CB;
    const int n0 = blockIdx.y * kNB;
    const int l = blockIdx.z;
    const LimbConst lc = lcs[l];
    const size_t LN = (size_t)L * N;

    auto issue = [&](int k, int stage) {
        uint64_t *dst = sm + (size_t)stage * ROWS * kNB;
        for (int ch = tid; ch < CHUNKS; ch += THREADS) {
            const int row = ch >> 4, part = ch & 15;  // kNB*8/16 = 16 chunks per row
            const uint64_t *src = R;
            int bytes = 0;
            if (row < RB) {
                if (row < nrows) {
                    src = R + ((size_t)k * nrows + row) * LN + (size_t)l * N + n0 + part * 2;
                    bytes = 16;
                }
            } else {
                const int col = col0 + (row - RB);
                if (col < ncols) {
                    const long long po = poff[(size_t)col * K + k];
                    if (po >= 0) {
                        src = P + po + (size_t)l * N + n0 + part * 2;
                        bytes = 16;
                    }
                }
            }
            cp_async16(dst + row * kNB + part * 2, src, bytes);
        }
    };

    u128 acc[TR][TC];
#pragma unroll
    for (int r = 0; r < TR; r++)
#pragma unroll
        for (int c = 0; c < TC; c++) acc[r][c] = u128{0, 0};

#pragma unroll
    for (int st = 0; st < kStages - 1; st++) {
        if (st < K) issue(st, st);
        cp_async_commit();
    }
    for (int k = 0; k < K; k++) {
        cp_async_wait<kStages - 2>();
        __syncthreads();
        {   // prefetch step k + kStages - 1 into the slot freed by step k - 1
            const int kn = k + kStages - 1;
            if (kn < K) issue(kn, kn % kStages);
            cp_async_commit();
        }
        const uint64_t *st = sm + (size_t)(k % kStages) * ROWS * kNB;
        uint64_t a[TR], b[TC];
#pragma unroll
        for (int r = 0; r < TR; r++) a[r] = st[(rg * TR + r) * kNB + lane];
#pragma unroll
        for (int c = 0; c < TC; c++) b[c] = st[(RB + cg * TC + c) * kNB + lane];
#pragma unroll
        for (int r = 0; r < TR; r++)
#pragma unroll
            for (int c = 0; c < TC; c++) mac128(acc[r][c], a[r], b[c]);
    }
    cp_async_wait<0>();

    // epilogue: Montgomery reduce (ReduceAndAddUint128 + eval.Reduce) and store canonical residues
#pragma unroll
    for (int c = 0; c < TC; c++) {
        const int col = col0 + cg * TC + c;
        if (col >= ncols) continue;
#pragma unroll
        for (int r = 0; r < TR; r++) {
            const int row = rg * TR + r;
            if (row >= nrows) continue;
            cv[((size_t)col * nrows + row) * LN + (size_t)l * N + n0 + lane] = mred128(acc[r][c], lc);
        }
    }
}

template <int TR, int TC, int CG, int RG>
static int launch_mac_cfg(Ctx *c, const uint64_t *R, const uint64_t *P, const long long *poff, int K, int nrows, int ncols, int L,
                          uint64_t *cv, cudaStream_t st) {
    constexpr int CB = CG * TC, RB = RG * TR, THREADS = kNB * CG * RG;
    const size_t smem = (size_t)kStages * (RB + CB) * kNB * sizeof(uint64_t);
    SFG_CUDA(c, cudaFuncSetAttribute(k_mac<TR, TC, CG, RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((ncols + CB - 1) / CB, c->N / kNB, L);
    k_mac<TR, TC, CG, RG><<<grid, THREADS, smem, st>>>(R, P, poff, K, nrows, ncols, L, c->N, c->lc, cv);
    SFG_LAUNCHED(c, "k_mac", st);
    return 0;
}

int launch_mac(Ctx *c, const uint64_t *R, const uint64_t *P, const long long *poff, int K, int nrows, int ncols, int L,
               uint64_t *cv, cudaStream_t st) {
    if (K <= 0 || nrows <= 0 || ncols <= 0) return 0;
    if (c->N % kNB) SFG_FAIL(c, "N must be a multiple of %d", kNB);
    // row tiling: RG*TR >= nrows with the least padding; 512 threads per CTA
    if (nrows <= 8) return launch_mac_cfg<8, 2, 16, 1>(c, R, P, poff, K, nrows, ncols, L, cv, st);
    if (nrows <= 16) return launch_mac_cfg<8, 2, 8, 2>(c, R, P, poff, K, nrows, ncols, L, cv, st);
    if (nrows <= 20) return launch_mac_cfg<10, 2, 8, 2>(c, R, P, poff, K, nrows, ncols, L, cv, st);
    if (nrows <= 24) return launch_mac_cfg<8, 2, 5, 3>(c, R, P, poff, K, nrows, ncols, L, cv, st);
    if (nrows <= 30) return launch_mac_cfg<10, 2, 5, 3>(c, R, P, poff, K, nrows, ncols, L, cv, st);
    if (nrows <= 32) return launch_mac_cfg<8, 2, 4, 4>(c, R, P, poff, K, nrows, ncols, L, cv, st);
    SFG_FAIL(c, "more than 16 ciphertext rows per MAC launch (nrows=%d): split the call", nrows);
}

}  // namespace sfg
