// kernels_mactc.cu -- the dominant kernel on the 5th-generation tensor cores: output-stationary multiply-accumulate of the
// rotated ciphertext residues with the NTT-domain plaintext diagonals, fused with the modular reduce
// (K1 + K2 of SURVEY 2.2; gwas/matmult.go:247-324, 343-399, 1154-1168).
//
// Dense-contraction view: for every RNS limb l and coefficient n,
//     CV[col][row] = sum_k  R[k][row] * P[col][k]      (mod q_l)
// with row = (i, c) over the 2s ciphertext polynomials, k = (block row bi, baby step b), col = (giant g, block column bj):
// a real (2s) x K x (d*m_ct) integer GEMM per (l, n), batch L'*N.  Only the canonical residue is observable (App. A.4), so
// the residues are split into BYTE limbs and the contraction runs as u8 x u8 -> s32 tcgen05.mma (kind::i8):
//     P[col][k] = sum_j P_j 2^(8j),  R[k][row] = sum_i R_i 2^(8i)   =>   sum_k R*P = sum_s 2^(8s) * T_s,
//     T_s[col][row] = sum_{i+j=s} sum_k P_j[col][k] * R_i[k][row]
//   A operand (M = 128 lanes)  : one byte plane P_j of 128 columns, K-major              (streamed from HBM, exactly once)
//   B operand (N = nb*RP cols) : all byte planes of R stacked, row (i*RP + r), K-major    (L2-resident, 15-30 KB per (l, n))
//   D (TMEM)                   : plane j accumulates at column offset j*RP, so that products with equal i+j land in the
//                                same accumulator column: (2nb-1)*RP columns instead of nb*nb*RP, and the K-sum, the
//                                (i+j)-sum and the zero padding all happen inside the tensor core.
// The epilogue (4 warps, one TMEM lane = one column each) reads the 2nb-1 partial sums per output, recombines them with
// the constants 2^(8s) mod q, Barrett/Montgomery-reduces to the canonical residue and stages NGF consecutive coefficients
// in shared memory so that every global store is a full 32-byte sector of cv[col][row][l][n..n+3].
//
// Both operand images are laid out in HBM exactly as the UMMA shared-memory descriptors expect them (no-swizzle K-major
// core matrices: 8 rows x 16 bytes contiguous, SBO = 128 B between 8-row groups, LBO = rows*16 B between 16-byte K chunks),
// so one stage is ONE contiguous cp.async.bulk of 128*Kg bytes -- no tensor map, no swizzle, no address math on the SM.
//
// Warp roles (384 threads, 1 CTA/SM, persistent over (l, n-group, column tile) items):
//   warp 0 lane 0 : bulk-copy producer (A ring of SA stages, B ring of 2 slots; mbarrier expect_tx)
//   warp 1 lane 0 : tcgen05.mma issuer (zeroing MMA + nb*Kg/32 MMAs per tile), tcgen05.commit -> stage / TMEM barriers
//   warp 2        : TMEM allocation (512 columns = 2 accumulator buffers)
//   warps 4..11   : epilogue, two warps per TMEM lane quadrant (tcgen05.ld -> recombine -> reduce -> smem -> 32-byte stores)
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "kernels.h"

namespace sfg {

namespace {

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
// one contiguous global -> shared bulk copy (TMA engine, UBLKCP), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void *smem, const void *gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(smem)),
                 "l"(gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void *smem, const void *gmem, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
                     smem_u32(smem)),
                 "l"(gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}
// the same copy delivered to the same shared-memory offset (and mbarrier offset) of every CTA of the cluster named in cta_mask
__device__ __forceinline__ void bulk_g2s_mc_hint(void *smem, const void *gmem, uint32_t bytes, uint64_t *bar, uint16_t cta_mask, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint [%0], [%1], %2, [%3], %4, %5;\n" ::"r"(
            smem_u32(smem)),
        "l"(gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// the same arrival on the mbarrier at this offset in every CTA of cta_mask
__device__ __forceinline__ void tc_commit_mc(uint64_t *bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, u8 x u8 -> s32, M = 128, K = 32 per instruction
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE: start address, LBO (between the 16-byte K chunks of one MMA),
// SBO (between 8-row groups), version 1 (Blackwell)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ULL << 46);
}
// instruction descriptor: c = S32 (2 << 4), a/b = unsigned 8 bit (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t umma_idesc_u8(int n) { return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

constexpr int kTcMaxL = 8;     // limbs the tensor-core path carries per launch (maxLevel is 5 at every reference call site)
constexpr int kTcMaxS = 11;    // 2*nb - 1 partial sums, nb <= 6 byte planes (q < 2^48)
constexpr int kZeroBytes = 8192;

}  // namespace

struct MacTcParams {
    const uint8_t *P;     // P image of this K group
    const uint8_t *R;     // R image of this K group
    const uint8_t *R2;    // pair mode: the R image of the second row part (CTA rank 1 of every 2-CTA cluster), same geometry
    int pair, cv_row0_2, rows2;  // pair mode: both CTAs of a cluster walk the same items; each fetches HALF of every A stage and multicasts it
                          // to both, so the P image is read from HBM once for two row parts; cv_row0_2 / rows2 = first cv row / rows of the second part
    uint64_t *cv;         // [col - col_lo][rows][L][N] canonical residues
    const LimbConst *lc;
    int L, N, rows, RP, Kg;
    int ngroups;                // K groups accumulated in TMEM by this launch (the s32 partial sums hold nb * ngroups * Kg * 255^2)
    long long p_gstride, r_gstride;  // bytes between consecutive K-group images of P / of R
    int cv_rows, cv_row0;       // rows per column of the cv image and the first row this launch writes
    int img_ntiles, img_tile0;  // geometry of the P image: tiles it holds and the global index of its first tile
    int tile_lo, tile_hi;       // global column tiles processed by this launch
    int col_lo, col_hi;         // global columns written
    int accumulate;             // cv += (mod q) instead of cv =
    int SA, NGF, SBN;
    int ni[kTcMaxL];              // coefficients per item of limb l: 4, or 8 for the 32-bit classes (their results are staged as u32, so 8
                                  // consecutive coefficients leave as ONE 64-byte store instead of two scattered 32-byte sectors)
    long long item_base[kTcMaxL + 1];  // first item of limb l (items are limb-major)
    int tbuf_stride;            // TMEM column stride between the two accumulator buffers (256) or 0 = single buffer
    int bslot_bytes;
    int dbg;                    // debug knobs (SFG_TC_DBG): 1 = no global stores, 2 = no epilogue math
    int nb[kTcMaxL], npad[kTcMaxL], fast[kTcMaxL];
    long long pbase[kTcMaxL], rbase[kTcMaxL];
    uint64_t cs[kTcMaxL][kTcMaxS];  // fast: 2^(8s) mod q ; otherwise 2^(8s) * 2^64 mod q (Montgomery form)
};

namespace {

struct Item {
    int l, ct, NI, n0;   // limb, column tile, coefficients in this item, first coefficient
    long long tg0;       // index of coefficient n0's A-stage group inside the limb's part of the P image
};
__device__ __forceinline__ Item decode_item(const MacTcParams &p, long long item) {
    const int ntr = p.tile_hi - p.tile_lo;
    Item it;
    it.l = 0;
    while (it.l + 1 < p.L && item >= p.item_base[it.l + 1]) it.l++;
    it.NI = p.ni[it.l];
    const int rem = (int)(item - p.item_base[it.l]);
    const int SBC = p.SBN * 4;            // coefficients per superblock
    const int GS = SBC / it.NI;           // items per (superblock, tile)
    const int per_sb = ntr * GS;
    const int sb = rem / per_sb, rem2 = rem - sb * per_sb;
    it.ct = p.tile_lo + rem2 / GS;
    const int gl = rem2 % GS;
    it.n0 = sb * SBC + gl * it.NI;
    it.tg0 = ((long long)sb * p.img_ntiles + (it.ct - p.img_tile0)) * SBC + gl * it.NI;
    return it;
}

// recombination + reduction of the partial sums of one output:  sum_s T_s * 2^(8s) mod q, canonical.
//   FAST 2: q < 2^31.  cs[s] = 2^(8s) 2^32 mod q (u32); x = sum T_s cs[s] < q 2^32 (host-checked), one IMAD.WIDE.U32 per term,
//           then a 32-bit Montgomery reduction: (x + ((x qinv') mod 2^32) q) >> 32 with qinv' = -q^-1 mod 2^32.
//   FAST 1: cs[s] = 2^(8s) mod q (u64), x < 2^64 (host-checked), Barrett (Lattigo BRedAdd).
//   FAST 0: cs[s] = 2^(8s) 2^64 mod q, 128-bit accumulate, the reference's Montgomery reduce (gwas/matmult.go:291-324).
template <int NS, int FAST>
__device__ __forceinline__ uint64_t recombine(const uint32_t *T, const uint64_t *cs, const LimbConst &lc, uint32_t nqinv32) {
    if constexpr (FAST == 2) {
        uint64_t x = 0;
#pragma unroll
        for (int s = 0; s < NS; s++) x += (uint64_t)T[s] * (uint32_t)cs[s];
        const uint32_t q = (uint32_t)lc.q;
        const uint32_t m = (uint32_t)x * nqinv32;
        const uint32_t t = (uint32_t)((x + (uint64_t)m * q) >> 32);  // < 2q, exact: the low word cancels
        return min(t, t - q);
    } else if constexpr (FAST == 1) {
        uint64_t x = 0;
#pragma unroll
        for (int s = 0; s < NS; s++) x += (uint64_t)T[s] * cs[s];  // host guarantees the sum stays below 2^64
        return bred_add(x, lc);
    } else {
        u128 acc{0, 0};
#pragma unroll
        for (int s = 0; s < NS; s++) mac128(acc, (uint64_t)T[s], cs[s]);
        return mred128(acc, lc);
    }
}

// One accumulator tile: straight-line code per chunk of 4 rows (NB and the reduction class are compile-time, the four rows
// are independent dependency chains), results staged in shared memory as outb[row][slot][column].
template <int NB, int FAST>
__device__ __forceinline__ void epilogue_tile(const MacTcParams &p, int rows, int l, uint32_t taddr, int slot, uint64_t *outb, int t, int half,
                                              int ngf) {
    constexpr int NS = 2 * NB - 1;
    const int RP = p.RP;
    const LimbConst lc = p.lc[l];
    const uint32_t nqinv32 = 0u - (uint32_t)lc.qinv;  // -q^-1 mod 2^32
    uint64_t cs[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) cs[s] = p.cs[l][s];
    for (int q = half; q < RP / 4; q += 2) {  // the two epilogue warps of a TMEM lane quadrant take alternate row chunks
        uint32_t v[NS][4];
#pragma unroll
        for (int s = 0; s < NS; s++) tmem_ld4(taddr + s * RP + 4 * q, v[s]);
        tmem_wait_ld();
        if (p.dbg & 4) {  // TMEM reads only
            uint32_t x = 0;
#pragma unroll
            for (int s = 0; s < NS; s++) x ^= v[s][0] ^ v[s][1] ^ v[s][2] ^ v[s][3];
            if (x == 0xdeadbeef) outb[t] = x;
            continue;
        }
        uint64_t val[4];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            uint32_t T[NS];
#pragma unroll
            for (int s = 0; s < NS; s++) T[s] = v[s][r];
            val[r] = recombine<NS, FAST>(T, cs, lc, nqinv32);
        }
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const int row = 4 * q + r;
            if (row < rows && (!(p.dbg & 8) || val[r] == 0xdeadbeefULL)) {
                if (FAST == 2 && p.ni[l] == 8) reinterpret_cast<uint32_t *>(outb)[((size_t)row * ngf + slot) * 128 + t] = (uint32_t)val[r];  // < q < 2^31
                else outb[((size_t)row * ngf + slot) * 128 + t] = val[r];
            }
        }
    }
}

template <int NB>
__device__ __forceinline__ void epilogue_nb(const MacTcParams &p, int rows, int l, uint32_t taddr, int slot, uint64_t *outb, int t, int half,
                                            int ngf) {
    const int f = p.fast[l];
    if (f == 2) epilogue_tile<NB, 2>(p, rows, l, taddr, slot, outb, t, half, ngf);
    else if (f == 1) epilogue_tile<NB, 1>(p, rows, l, taddr, slot, outb, t, half, ngf);
    else epilogue_tile<NB, 0>(p, rows, l, taddr, slot, outb, t, half, ngf);
}

}  // namespace

__global__ void __launch_bounds__(384, 1) k_mac_tc(const __grid_constant__ MacTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int stage_bytes = 128 * p.Kg;
    uint8_t *a_ring = smem;
    uint8_t *b_ring = a_ring + (size_t)p.SA * stage_bytes;
    uint8_t *zblk = b_ring + 2 * (size_t)p.bslot_bytes;
    uint64_t *outb = reinterpret_cast<uint64_t *>(zblk + kZeroBytes);
    uint64_t *bars = outb + (size_t)p.RP * p.NGF * 128;
    uint64_t *a_full = bars, *a_empty = bars + p.SA;
    uint64_t *b_full = bars + 2 * p.SA, *b_empty = b_full + 2;
    uint64_t *t_full = b_empty + 2, *t_empty = t_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t_empty + 2);

    const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
    const int ntbuf = p.tbuf_stride ? 2 : 1;
    // pair mode: CTA `rank` of cluster `cta0` handles row part `rank`; items are strided over the clusters
    const uint32_t rank = p.pair ? cluster_ctarank() : 0u;
    const int cta0 = p.pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, ncta = p.pair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const uint8_t *Rimg = rank ? p.R2 : p.R;
    const int cv_row0 = rank ? p.cv_row0_2 : p.cv_row0, rows = rank ? p.rows2 : p.rows;

    if (tid == 0) {
        for (int s = 0; s < p.SA; s++) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], p.pair ? 2 : 1);  // pair mode: a stage is free when BOTH CTAs' MMAs have read their copy of it
        }
        for (int s = 0; s < 2; s++) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
            mbar_init(&t_full[s], 1);
            mbar_init(&t_empty[s], 256);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int i = tid; i < kZeroBytes / 16; i += blockDim.x) reinterpret_cast<uint4 *>(zblk)[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");  // generic-proxy zeros -> visible to the tensor core (async proxy)
    if (wid == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (p.pair) cluster_sync_all();  // the partner's mbarriers exist before anything is multicast to them
    const uint32_t tmem_base = *tmem_slot;

    const long long nitems = p.item_base[p.L];

    if (wid == 0) {
        // ===================== producer =====================
        if (lane == 0) {
            uint32_t as = 0, aph = 0, bs = 0, bph = 0;
            // the diagonals are read exactly once (evict first); the rotated ciphertext image is re-read by every column tile
            const uint64_t pol_p = policy_evict_first(), pol_r = policy_evict_last();
            for (long long item = cta0; item < nitems; item += ncta) {
                const Item it = decode_item(p, item);
                const int nb = p.nb[it.l];
                const uint32_t bbytes = (uint32_t)p.npad[it.l] * p.Kg;
                const long long tg0 = it.tg0;
                for (int i = 0; i < it.NI; i++) {
                    const int n = it.n0 + i;
                    for (int g = 0; g < p.ngroups; g++) {  // every K group of this coefficient accumulates into the same TMEM tile
                        mbar_wait(&b_empty[bs], bph ^ 1u);
                        mbar_arrive_expect_tx(&b_full[bs], bbytes);
                        bulk_g2s_hint(b_ring + (size_t)bs * p.bslot_bytes, Rimg + g * p.r_gstride + p.rbase[it.l] + (long long)n * bbytes, bbytes,
                                      &b_full[bs], pol_r);
                        bs ^= 1u;
                        if (bs == 0) bph ^= 1u;
                        const uint8_t *src = p.P + g * p.p_gstride + p.pbase[it.l] + (tg0 + i) * nb * (long long)stage_bytes;
                        for (int j = 0; j < nb; j++) {
                            mbar_wait(&a_empty[as], aph ^ 1u);
                            mbar_arrive_expect_tx(&a_full[as], (uint32_t)stage_bytes);
                            if (p.pair) {  // my half of the stage, to both CTAs (the partner sends the other half)
                                const uint32_t hb = (uint32_t)stage_bytes >> 1;
                                bulk_g2s_mc_hint(a_ring + (size_t)as * stage_bytes + rank * hb, src + (long long)j * stage_bytes + rank * hb, hb,
                                                 &a_full[as], (uint16_t)3, pol_p);
                            } else {
                                bulk_g2s_hint(a_ring + (size_t)as * stage_bytes, src + (long long)j * stage_bytes, (uint32_t)stage_bytes, &a_full[as], pol_p);
                            }
                            if (++as == (uint32_t)p.SA) {
                                as = 0;
                                aph ^= 1u;
                            }
                        }
                    }
                }
            }
        }
    } else if (wid == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t as = 0, aph = 0, bs = 0, bph = 0, tb = 0, tph = 0;
            const uint32_t zaddr = smem_u32(zblk);
            const int ksteps = p.Kg >> 5;
            for (long long item = cta0; item < nitems; item += ncta) {
                const Item it = decode_item(p, item);
                const int nb = p.nb[it.l], npad = p.npad[it.l];
                const int region = (nb - 1) * p.RP + npad;
                const uint32_t idesc = umma_idesc_u8(npad);
                for (int i = 0; i < it.NI; i++) {
                    mbar_wait(&t_empty[tb], tph ^ 1u);
                    tc_fence_after();
                    const uint32_t d_base = tmem_base + tb * (uint32_t)p.tbuf_stride;
                    // zero the accumulator region: the byte planes overlap at different column offsets, so no single
                    // MMA can carry the "overwrite" flag for all of it
                    for (int c0 = 0; c0 < region; c0 += 256) {
                        const int rem = region - c0;
                        const int w = (((rem < 256 ? rem : 256) + 15) >> 4) << 4;
                        umma_i8(d_base + c0, umma_desc(zaddr, 2048, 128), umma_desc(zaddr, 4096, 128), umma_idesc_u8(w), 0u);
                    }
                    for (int g = 0; g < p.ngroups; g++) {
                        mbar_wait(&b_full[bs], bph);
                        tc_fence_after();
                        const uint32_t baddr = smem_u32(b_ring + (size_t)bs * p.bslot_bytes);
                        for (int j = 0; j < nb; j++) {
                            mbar_wait(&a_full[as], aph);
                            tc_fence_after();
                            const uint32_t aaddr = smem_u32(a_ring + (size_t)as * stage_bytes);
                            for (int ks = 0; ks < ksteps; ks++) {
                                umma_i8(d_base + j * p.RP, umma_desc(aaddr + ks * 4096, 2048, 128),
                                        umma_desc(baddr + ks * npad * 32, npad * 16, 128), idesc, 1u);
                            }
                            if (p.pair) tc_commit_mc(&a_empty[as], (uint16_t)3);
                            else tc_commit(&a_empty[as]);
                            if (++as == (uint32_t)p.SA) {
                                as = 0;
                                aph ^= 1u;
                            }
                        }
                        tc_commit(&b_empty[bs]);
                        bs ^= 1u;
                        if (bs == 0) bph ^= 1u;
                    }
                    tc_commit(&t_full[tb]);
                    if (++tb == (uint32_t)ntbuf) {
                        tb = 0;
                        tph ^= 1u;
                    }
                }
            }
        }
    } else if (wid >= 4) {
        // ===================== epilogue =====================
        const int t = (wid & 3) * 32 + lane, half = (wid - 4) >> 2;  // TMEM lane = column of the tile; 2 warps per lane quadrant
        const uint32_t lane_base = (uint32_t)((wid & 3) * 32) << 16;
        uint32_t tb = 0, tph = 0;
        const size_t LN = (size_t)p.L * p.N;
        for (long long item = cta0; item < nitems; item += ncta) {
            const Item it = decode_item(p, item);
            const int l = it.l;
            const uint64_t q = p.lc[l].q;
            const int col = it.ct * 128 + t;
            const bool narrow8 = it.NI == 8;                 // u32 staging, 8 coefficients per flush
            const int ngf = narrow8 ? 8 : p.NGF;             // staged coefficients per flush (slots of 4 resp. 8 bytes: same bytes per row)
            for (int i = 0; i < it.NI; i++) {
                mbar_wait(&t_full[tb], tph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + lane_base + tb * (uint32_t)p.tbuf_stride;
                if (p.dbg & 2) {
                    uint32_t v[4];
                    tmem_ld4(taddr, v);
                    tmem_wait_ld();
                    outb[t] = v[0];
                } else {
                    switch (p.nb[l]) {  // uniform across the CTA
                        case 1: epilogue_nb<1>(p, rows, l, taddr, i % ngf, outb, t, half, ngf); break;
                        case 2: epilogue_nb<2>(p, rows, l, taddr, i % ngf, outb, t, half, ngf); break;
                        case 3: epilogue_nb<3>(p, rows, l, taddr, i % ngf, outb, t, half, ngf); break;
                        case 4: epilogue_nb<4>(p, rows, l, taddr, i % ngf, outb, t, half, ngf); break;
                        case 5: epilogue_nb<5>(p, rows, l, taddr, i % ngf, outb, t, half, ngf); break;
                        default: epilogue_nb<6>(p, rows, l, taddr, i % ngf, outb, t, half, ngf); break;
                    }
                }
                tc_fence_before();
                mbar_arrive(&t_empty[tb]);
                if (++tb == (uint32_t)ntbuf) {
                    tb = 0;
                    tph ^= 1u;
                }
                if ((i + 1) % ngf == 0) {
                    asm volatile("bar.sync 1, 256;\n" ::: "memory");
                    if (col >= p.col_lo && col < p.col_hi && !(p.dbg & 1)) {
                        const int n0 = it.n0 + (i + 1 - ngf);
                        uint64_t *dst = p.cv + ((size_t)(col - p.col_lo) * p.cv_rows + cv_row0 + half) * LN + (size_t)l * p.N + n0;
                        for (int row = half; row < rows; row += 2, dst += 2 * LN) {
                            if (narrow8) {
                                // 8 consecutive coefficients of a 32-bit-class limb: widened to u64 and written as one 64-byte run
                                const uint32_t *src = reinterpret_cast<const uint32_t *>(outb) + (size_t)row * 8 * 128 + t;
                                uint64_t v[8];
#pragma unroll
                                for (int x = 0; x < 8; x++) v[x] = src[x * 128];
                                if (p.accumulate) {
#pragma unroll
                                    for (int x = 0; x < 8; x++) v[x] = add_mod(v[x], dst[x], q);
                                }
                                asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};\n" ::"l"(dst), "l"(v[0]), "l"(v[1]), "l"(v[2]), "l"(v[3]) : "memory");
                                asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};\n" ::"l"(dst + 4), "l"(v[4]), "l"(v[5]), "l"(v[6]), "l"(v[7]) : "memory");
                                continue;
                            }
                            const uint64_t *src = outb + (size_t)row * ngf * 128 + t;
                            if (ngf == 4) {
                                uint64_t v0 = src[0], v1 = src[128], v2 = src[256], v3 = src[384];
                                if (p.accumulate) {
                                    uint64_t o0, o1, o2, o3;
                                    asm volatile("ld.global.v4.b64 {%0, %1, %2, %3}, [%4];\n" : "=l"(o0), "=l"(o1), "=l"(o2), "=l"(o3) : "l"(dst));
                                    v0 = add_mod(v0, o0, q);
                                    v1 = add_mod(v1, o1, q);
                                    v2 = add_mod(v2, o2, q);
                                    v3 = add_mod(v3, o3, q);
                                }
                                // one full 32-byte sector per store (STG.256)
                                asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};\n" ::"l"(dst), "l"(v0), "l"(v1), "l"(v2), "l"(v3) : "memory");
                            } else {
                                for (int x = 0; x < ngf; x++) {
                                    uint64_t v = src[x * 128];
                                    if (p.accumulate) v = add_mod(v, dst[x], q);
                                    dst[x] = v;
                                }
                            }
                        }
                    }
                    asm volatile("bar.sync 1, 256;\n" ::: "memory");
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.pair) cluster_sync_all();  // the partner's last commits arrive on this CTA's mbarriers: do not leave before it is done too
    if (wid == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// image builder: polynomial records -> K-major byte-plane image (P tiles at preprocess time, R once per call)
// ---------------------------------------------------------------------------------------------------------------
struct ImgParams {
    const uint8_t *src;        // record base
    const long long *src_off;  // device [Kg][NRS]: byte offset of the source record of (k, row), -1 = zero
    uint8_t *dst;              // image base (of this K group)
    int NRS;                   // source rows per k (128 columns of a P tile; `rows` ciphertext polynomials for R)
    int Kg;
    int mode;                  // 0 = P image, 1 = R image
    int RP;                    // R: row pitch between byte planes
    int img_ntiles, ct, SBN;   // P: tile position inside the image
    int nlimbs;
    int limb[kTcMaxL], nb[kTcMaxL], npad[kTcMaxL];
    long long limb_off[kTcMaxL], base[kTcMaxL];
};

template <typename T>
__global__ void __launch_bounds__(512, 1) k_img_build(const __grid_constant__ ImgParams p) {
    constexpr int NCH = 32 / (int)sizeof(T);  // coefficients per CTA: one 32-byte sector of every source record
    extern __shared__ __align__(16) uint8_t sm_raw[];
    T *vals = reinterpret_cast<T *>(sm_raw);  // [Kg][NCH][16]
    const int li = blockIdx.z;
    const int nb = p.nb[li];
    const int n0 = blockIdx.x * NCH, r0 = blockIdx.y * 16;
    const int tid = threadIdx.x;
    const uint8_t *src = p.src + p.limb_off[li] + (size_t)n0 * sizeof(T);
    // (one 32-byte sector per source record and thread.  Requesting 2-4 sectors per thread before the first use was measured SLOWER,
    //  profiles/r2/ab_img_build_load_batching.txt: the phase is not bound by the bytes in flight)
    for (int idx = tid; idx < p.Kg * 16; idx += blockDim.x) {
        const int c = idx & 15, k = idx >> 4;
        const int row = r0 + c;
        long long off = -1;
        if (row < p.NRS) off = p.src_off[(size_t)k * p.NRS + row];
        uint4 a = make_uint4(0, 0, 0, 0), b = a;
        if (off >= 0) {
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src + off);
            a = __ldg(s4);
            b = __ldg(s4 + 1);
        }
        T *dstv = vals + (size_t)k * NCH * 16 + c;
        if constexpr (sizeof(T) == 4) {
            dstv[0 * 16] = a.x; dstv[1 * 16] = a.y; dstv[2 * 16] = a.z; dstv[3 * 16] = a.w;
            dstv[4 * 16] = b.x; dstv[5 * 16] = b.y; dstv[6 * 16] = b.z; dstv[7 * 16] = b.w;
        } else {
            dstv[0 * 16] = ((uint64_t)a.y << 32) | a.x;
            dstv[1 * 16] = ((uint64_t)a.w << 32) | a.z;
            dstv[2 * 16] = ((uint64_t)b.y << 32) | b.x;
            dstv[3 * 16] = ((uint64_t)b.w << 32) | b.z;
        }
    }
    __syncthreads();
    const int nkc = p.Kg >> 4;
    const int total = NCH * nb * nkc * 16;
    const int NR = p.mode == 0 ? 128 : p.npad[li];
    for (int idx = tid; idx < total; idx += blockDim.x) {
        const int c = idx & 15;
        int rest = idx >> 4;
        const int kc = rest % nkc;
        rest /= nkc;
        const int j = rest % nb, nn = rest / nb;
        const int row = r0 + c;
        if (row >= p.NRS) continue;
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            const T v = vals[((size_t)(kc * 16 + kk) * NCH + nn) * 16 + c];
            w[kk >> 2] |= (uint32_t)((v >> (8 * j)) & 0xff) << (8 * (kk & 3));
        }
        const int n = n0 + nn;
        long long o;
        int orow;
        if (p.mode == 0) {
            const int n4 = n >> 2, sb = n4 / p.SBN, n4l = n4 % p.SBN;
            const long long tg = (((long long)sb * p.img_ntiles + p.ct) * p.SBN + n4l) * 4 + (n & 3);
            o = p.base[li] + (tg * nb + j) * (128LL * p.Kg);
            orow = row;
        } else {
            o = p.base[li] + (long long)n * NR * p.Kg;
            orow = j * p.RP + row;
        }
        o += (long long)kc * NR * 16 + (orow >> 3) * 128 + (orow & 7) * 16;
        *reinterpret_cast<uint4 *>(p.dst + o) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// one polynomial record back out of the P image (sfg_cache_get_diag: parity tests of the cached diagonals)
__global__ void k_img_extract(const uint8_t *__restrict__ img, long long base, int nb, int Kg, int img_ntiles, int ct, int SBN, int c,
                              int k, int N, uint64_t *__restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const int n4 = n >> 2, sb = n4 / SBN, n4l = n4 % SBN;
    const long long tg = (((long long)sb * img_ntiles + ct) * SBN + n4l) * 4 + (n & 3);
    uint64_t v = 0;
    for (int j = 0; j < nb; j++) {
        const long long o = base + (tg * nb + j) * (128LL * Kg) + (long long)(k >> 4) * 2048 + (c >> 3) * 128 + (c & 7) * 16 + (k & 15);
        v |= (uint64_t)img[o] << (8 * j);
    }
    out[n] = v;
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int bytes_of(uint64_t q) {
    int b = 0;
    while (q) {
        b++;
        q >>= 8;
    }
    return b;  // residues are < q
}

int tc_geom_p(Ctx *c, int L, int K, int ncols, TcGeomP *g) {
    if (L > kTcMaxL) SFG_FAIL(c, "tensor-core MAC carries at most %d limbs per call (maxLevel = %d)", kTcMaxL, L);
    if (c->N < 16) SFG_FAIL(c, "ring degree %d too small for the tensor-core MAC", c->N);
    memset(g, 0, sizeof *g);
    g->L = L;
    g->N = c->N;
    g->K = K;
    g->ncols = ncols;
    g->ntiles = (ncols + 127) / 128;
    g->SBN = std::min(64, c->N / 4);
    // K groups: Kg bytes of K per stage (multiple of 32, <= 256), as few padded K steps as possible
    int best_ng = 0, best_kg = 0;
    long long best = -1;
    const int ng0 = (K + 255) / 256;
    for (int ng = ng0; ng <= ng0 + 8; ng++) {
        const int kg = (((K + ng - 1) / ng) + 31) / 32 * 32;
        if (kg > 256) continue;
        const long long tot = (long long)ng * kg + ng * 8;  // small penalty per extra group (one more launch, one more epilogue)
        if (best < 0 || tot < best) best = tot, best_ng = ng, best_kg = kg;
    }
    g->ngroups = best_ng;
    g->Kg = best_kg;
    long long off = 0;
    for (int l = 0; l < L; l++) {
        g->nb[l] = bytes_of(c->mod[l] - 1);
        if (g->nb[l] > 6) SFG_FAIL(c, "modulus %d has more than 48 bits: not supported by the tensor-core MAC", l);
        g->pbase[l] = off;
        off += (long long)c->N * g->ntiles * g->nb[l] * 128 * g->Kg;
    }
    g->group_bytes = off;
    return 0;
}

int tc_geom_r(Ctx *c, const TcGeomP &gp, int rows, TcGeomR *g, int rp_min) {
    memset(g, 0, sizeof *g);
    g->rows = rows;
    g->RP = std::max((rows + 3) / 4 * 4, rp_min);  // rp_min: the row pitch of the part this one is paired with (same image geometry)
    long long off = 0;
    int maxreg = 0;
    for (int l = 0; l < gp.L; l++) {
        const int nb = gp.nb[l];
        g->npad[l] = (nb * g->RP + 15) / 16 * 16;
        if (g->npad[l] > 256) SFG_FAIL(c, "%d ciphertext polynomials x %d byte planes exceed one MMA (N <= 256): split the call", rows, nb);
        const int region = ((nb - 1) * g->RP + g->npad[l] + 15) / 16 * 16;
        maxreg = std::max(maxreg, region);
        g->rbase[l] = off;
        off += (long long)c->N * g->npad[l] * gp.Kg;
    }
    if (maxreg > 512) SFG_FAIL(c, "accumulator region of %d TMEM columns exceeds 512: split the call", maxreg);
    g->tbuf_stride = maxreg <= 256 ? 256 : 0;
    g->group_bytes = off;
    return 0;
}

static void split_limbs(const PolyLayout &lay, int L, int want_es, ImgParams &ip, const int *nb, const int *npad, const long long *base) {
    ip.nlimbs = 0;
    for (int l = 0; l < L; l++)
        if (lay.es[l] == want_es) {
            const int i = ip.nlimbs++;
            ip.limb[i] = l;
            ip.nb[i] = nb[l];
            ip.npad[i] = npad ? npad[l] : 0;
            ip.limb_off[i] = lay.off[l];
            ip.base[i] = base[l];
        }
}

static int launch_img(Ctx *c, ImgParams &ip, const PolyLayout &lay, int L, const int *nb, const int *npad, const long long *base,
                      cudaStream_t st) {
    for (int es : {4, 8}) {
        split_limbs(lay, L, es, ip, nb, npad, base);
        if (!ip.nlimbs) continue;
        const int NCH = 32 / es;
        const size_t smem = (size_t)ip.Kg * 512;
        dim3 grid(c->N / NCH, (ip.NRS + 15) / 16, ip.nlimbs);
        if (es == 4) {
            SFG_CUDA(c, cudaFuncSetAttribute(k_img_build<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_img_build<uint32_t><<<grid, 512, smem, st>>>(ip);
        } else {
            SFG_CUDA(c, cudaFuncSetAttribute(k_img_build<uint64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_img_build<uint64_t><<<grid, 512, smem, st>>>(ip);
        }
        SFG_LAUNCHED(c, "k_img_build", st);
    }
    return 0;
}

int launch_img_p(Ctx *c, const TcGeomP &g, const PolyLayout &lay, const void *records, const long long *src_off_dev, int img_ntiles,
                 int ct_in_img, void *img_group, cudaStream_t st) {
    if (c->N % 8) SFG_FAIL(c, "N must be a multiple of 8");
    ImgParams ip{};
    ip.src = (const uint8_t *)records;
    ip.src_off = src_off_dev;
    ip.dst = (uint8_t *)img_group;
    ip.NRS = 128;
    ip.Kg = g.Kg;
    ip.mode = 0;
    ip.img_ntiles = img_ntiles;
    ip.ct = ct_in_img;
    ip.SBN = g.SBN;
    // limb bases scale with the number of tiles the image holds
    long long base[kTcMaxL], off = 0;
    for (int l = 0; l < g.L; l++) {
        base[l] = off;
        off += (long long)c->N * img_ntiles * g.nb[l] * 128 * g.Kg;
    }
    return launch_img(c, ip, lay, g.L, g.nb, nullptr, base, st);
}

int launch_img_r(Ctx *c, const TcGeomP &gp, const TcGeomR &gr, const PolyLayout &lay, const void *R, const long long *src_off_dev,
                 void *img_group, cudaStream_t st) {
    ImgParams ip{};
    ip.src = (const uint8_t *)R;
    ip.src_off = src_off_dev;
    ip.dst = (uint8_t *)img_group;
    ip.NRS = gr.rows;
    ip.Kg = gp.Kg;
    ip.mode = 1;
    ip.RP = gr.RP;
    SFG_CUDA(c, cudaMemsetAsync(img_group, 0, (size_t)gr.group_bytes, st));  // padding rows / planes / K steps are zero
    return launch_img(c, ip, lay, gp.L, gp.nb, gr.npad, gr.rbase, st);
}

int launch_img_extract(Ctx *c, const TcGeomP &g, const void *img_group, int l, int col, int k_in_group, uint64_t *out_dev, cudaStream_t st) {
    const int ct = col / 128, cc = col % 128;
    k_img_extract<<<(c->N + 255) / 256, 256, 0, st>>>((const uint8_t *)img_group, g.pbase[l], g.nb[l], g.Kg, g.ntiles, ct, g.SBN, cc,
                                                      k_in_group, c->N, out_dev);
    SFG_LAUNCHED(c, "k_img_extract", st);
    return 0;
}

// K groups one launch may accumulate in TMEM: the s32 partial sums hold nb * (groups * Kg) products of two bytes
int tc_max_fused_groups(const TcGeomP &gp) {
    int nbmax = 1;
    for (int l = 0; l < gp.L; l++) nbmax = std::max(nbmax, gp.nb[l]);
    long long g = std::max<long long>(1, 2147483647LL / ((long long)nbmax * gp.Kg * 65025LL));
    if (const char *e = getenv("SFG_TC_MAXGROUPS")) g = std::max<long long>(1, std::min<long long>(g, atoll(e)));  // tests: force the multi-launch path
    return (int)g;
}

int launch_mac_tc(Ctx *c, const TcGeomP &gp, const TcGeomR &gr, const void *Pimg_group, long long p_gstride, int img_ntiles, int img_tile0,
                  const void *Rimg_group, long long r_gstride, int ngroups, int tile_lo, int tile_hi, int col_lo, int col_hi, bool accumulate,
                  uint64_t *cv, cudaStream_t st, const void *Rimg_group2, int cv_row0_2, int rows2) {
    if (tile_hi <= tile_lo || col_hi <= col_lo || ngroups < 1) return 0;
    if (ngroups > tc_max_fused_groups(gp)) SFG_FAIL(c, "tensor-core MAC: %d K groups exceed the s32 accumulator range", ngroups);
    const int Ktot = ngroups * gp.Kg;  // products summed per accumulator by this launch
    MacTcParams p{};
    p.P = (const uint8_t *)Pimg_group;
    p.R = (const uint8_t *)Rimg_group;
    p.R2 = (const uint8_t *)Rimg_group2;
    p.pair = Rimg_group2 ? 1 : 0;
    p.cv_row0_2 = cv_row0_2;
    p.rows2 = rows2;
    p.ngroups = ngroups;
    p.p_gstride = p_gstride;
    p.r_gstride = r_gstride;
    p.cv = cv;
    p.lc = c->lc;
    p.L = gp.L;
    p.N = c->N;
    p.rows = gr.rows;
    p.cv_rows = gr.cv_rows ? gr.cv_rows : gr.rows;
    p.cv_row0 = gr.cv_row0;
    p.RP = gr.RP;
    p.Kg = gp.Kg;
    p.img_ntiles = img_ntiles;
    p.img_tile0 = img_tile0;
    p.tile_lo = tile_lo;
    p.tile_hi = tile_hi;
    p.col_lo = col_lo;
    p.col_hi = col_hi;
    p.accumulate = accumulate ? 1 : 0;
    p.SBN = gp.SBN;
    p.tbuf_stride = gr.tbuf_stride;
    int maxnpad = 0;
    long long off = 0;
    for (int l = 0; l < gp.L; l++) {
        const uint64_t q = c->mod[l];
        const int nb = gp.nb[l];
        p.nb[l] = nb;
        p.npad[l] = gr.npad[l];
        p.pbase[l] = off;
        off += (long long)c->N * img_ntiles * nb * 128 * gp.Kg;
        p.rbase[l] = gr.rbase[l];
        maxnpad = std::max(maxnpad, gr.npad[l]);
        // u64 recombination is exact iff (2nb-1) * [nb * Kg * 255^2] * q < 2^64
        const long double bound = (long double)(2 * nb - 1) * nb * Ktot * 65025.0L * (long double)q;
        p.fast[l] = bound < 18446744073709551616.0L ? 1 : 0;
        // 32-bit Montgomery: terms T_s * c_s with c_s < q < 2^31 and the sum below q * 2^32
        if (q < (1ULL << 31) && (long double)(2 * nb - 1) * nb * Ktot * 65025.0L < 4294967296.0L) p.fast[l] = 2;
        if ((long double)nb * Ktot * 65025.0L >= 2147483648.0L) SFG_FAIL(c, "s32 accumulator overflow (K = %d)", Ktot);
        for (int s = 0; s < 2 * nb - 1; s++) {
            uint64_t v = h_powmod(2, 8 * s, q);
            if (p.fast[l] == 0) v = h_mulmod(v, c->lc_h[l].r64, q);
            if (p.fast[l] == 2) v = h_mulmod(v, (1ULL << 32) % q, q);
            p.cs[l][s] = v;
        }
    }
    p.bslot_bytes = maxnpad * gp.Kg;
    const int stage = 128 * gp.Kg;
    const size_t fixed = 2 * (size_t)p.bslot_bytes + kZeroBytes + 512;
    const size_t cap = 232448;  // 227 KB
    int NGF = 4, SA = 0;
    for (; NGF >= 1; NGF >>= 1) {
        const size_t outb = (size_t)gr.RP * NGF * 128 * 8;
        if (fixed + outb + 3 * (size_t)stage <= cap) {
            SA = (int)std::min<size_t>(8, (cap - fixed - outb) / stage);
            break;
        }
    }
    if (SA < 3) SFG_FAIL(c, "tensor-core MAC does not fit shared memory (rows = %d, Kg = %d)", gr.rows, gp.Kg);
    p.NGF = NGF;
    // items: 4 coefficients of one (limb, column tile); 8 for the 32-bit classes when the staging buffer holds 8 of their 4-byte results
    // per row (NGF == 4) -- their outputs then leave as 64-byte runs (SFG_TC_NI4=1 keeps 4 everywhere for A/B runs)
    static const bool ni4 = [] { const char *e = getenv("SFG_TC_NI4"); return e && *e == '1'; }();
    p.item_base[0] = 0;
    for (int l = 0; l < gp.L; l++) {
        p.ni[l] = (!ni4 && p.fast[l] == 2 && NGF == 4 && (gp.SBN * 4) % 8 == 0) ? 8 : 4;
        p.item_base[l + 1] = p.item_base[l] + (long long)(c->N / p.ni[l]) * (tile_hi - tile_lo);
    }
    if (const char *e = getenv("SFG_TC_DBG")) p.dbg = atoi(e);
    if (const char *e = getenv("SFG_TC_SA")) SA = std::min(SA, atoi(e));
    p.SA = SA;
    const size_t smem = (size_t)SA * stage + fixed + (size_t)gr.RP * NGF * 128 * 8;
    SFG_CUDA(c, cudaFuncSetAttribute(k_mac_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int nsm = 0;
    SFG_CUDA(c, cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device));
    const long long nitems = p.item_base[gp.L];
    if (p.pair) {
        // 2-CTA clusters, one CTA per SM: as many clusters as the device can hold at once (the kernel is persistent)
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(nsm & ~1), 1, 1);
        cfg.blockDim = dim3(384, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int nclusters = 0;
        SFG_CUDA(c, cudaOccupancyMaxActiveClusters(&nclusters, k_mac_tc, &cfg));
        if (nclusters < 1) SFG_FAIL(c, "tensor-core MAC: no 2-CTA cluster fits the device");
        const int grid = 2 * (int)std::min<long long>(std::min(nclusters, nsm / 2), nitems);
        cfg.gridDim = dim3((unsigned)grid, 1, 1);
        SFG_CUDA(c, cudaLaunchKernelEx(&cfg, k_mac_tc, p));
    } else {
        const int grid = (int)std::min<long long>(nsm, nitems);
        k_mac_tc<<<grid, 384, smem, st>>>(p);
    }
    SFG_LAUNCHED(c, "k_mac_tc", st);
    return 0;
}

}  // namespace sfg
