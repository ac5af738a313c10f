// kernels_geno.cu -- genotype scan of the randomized-PCA sketch (gwas/pca.go:152-162):
//     localSketch[randIndex[i]][j] += sgn[i] * float64(row[j]);  xsum[j] += uint64(row[j]);  x2sum[j] += uint64(row[j] * row[j])
// over the int8 rows the GenoFileStream delivers (dosages 0/1/2, missing already replaced by 0: gwas/filestream.go:349-351).
// Pure int8 streaming, HBM-bound: the matrix (nrows x ncols bytes) is read exactly once with 16-byte loads.  Rows are grouped by
// bucket on the host (counting sort of randIndex), a CTA = (512-column tile, bucket, row split); each thread owns 16 columns and adds
// whole 32-bit words of four packed dosages (no carry between bytes for up to 63 rows: 63 * 4 < 256), split by the row's sign, then
// spills the byte lanes into 32-bit accumulators.  All sums are exact integers; the float64 sketch is produced from them at the end.
#include "kernels.h"

namespace sfg {

template <bool ALIGNED>
__device__ __forceinline__ uint4 load16(const int8_t *__restrict__ p, int nvalid) {
    if constexpr (ALIGNED) {
        return __ldg(reinterpret_cast<const uint4 *>(p));
    } else {
        uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 16; k++)
            if (k < nvalid) w[k >> 2] |= (uint32_t)(uint8_t)__ldg(p + k) << (8 * (k & 3));
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256)
k_count_sketch(const int8_t *__restrict__ X, size_t ncols, const int *__restrict__ rows_sorted, const int8_t *__restrict__ sgn_sorted,
               const int *__restrict__ bucket_off, int nsplit, long long *__restrict__ sketch, unsigned long long *__restrict__ xsum,
               unsigned long long *__restrict__ x2sum, int *__restrict__ bad) {
    // CTA = (512-column tile, bucket, row split): the 8 warps read the SAME 512 columns (lane -> 16 columns) of different rows, so
    // their partial sums are combined in shared memory and only one warp's worth of atomics leaves the CTA
    __shared__ uint32_t red[8][24][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t col0 = ((size_t)blockIdx.x * 32 + lane) * 16;
    const bool live = col0 < ncols;
    const int nvalid = live ? (int)min((size_t)16, ncols - col0) : 0;
    const int b = blockIdx.y, sp = blockIdx.z;
    const int lo = bucket_off[b], hi = bucket_off[b + 1];
    // byte lanes (<= 63 rows) -> 16-bit lanes (even / odd bytes of each word; the launcher bounds the rows per warp so they cannot overflow)
    uint32_t P[4] = {0, 0, 0, 0}, Ng[4] = {0, 0, 0, 0}, Sq[4] = {0, 0, 0, 0}, badacc = 0;
    uint32_t P16[8], N16[8], S16[8];
#pragma unroll
    for (int k = 0; k < 8; k++) P16[k] = N16[k] = S16[k] = 0;
    int cnt = 0;
    auto spill = [&]() {
#pragma unroll
        for (int w = 0; w < 4; w++) {
            P16[2 * w] += P[w] & 0x00FF00FFu;
            P16[2 * w + 1] += (P[w] >> 8) & 0x00FF00FFu;
            N16[2 * w] += Ng[w] & 0x00FF00FFu;
            N16[2 * w + 1] += (Ng[w] >> 8) & 0x00FF00FFu;
            S16[2 * w] += Sq[w] & 0x00FF00FFu;
            S16[2 * w + 1] += (Sq[w] >> 8) & 0x00FF00FFu;
            P[w] = Ng[w] = Sq[w] = 0;
        }
        cnt = 0;
    };
    constexpr int U = 8;  // rows in flight per thread
    const int stride = 8 * nsplit;
    if (live) {
        for (int i0 = lo + sp * 8 + warp; i0 < hi; i0 += U * stride) {
            uint4 v[U];
            int sg[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + u * stride;
                if (i < hi) {
                    v[u] = load16<ALIGNED>(X + (size_t)rows_sorted[i] * ncols + col0, nvalid);
                    sg[u] = sgn_sorted[i];
                } else {
                    v[u] = make_uint4(0, 0, 0, 0);
                    sg[u] = 1;
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    badacc |= (w[k] & 0xFCFCFCFCu) | (w[k] & (w[k] >> 1) & 0x01010101u);  // a byte outside {0, 1, 2}
                    if (sg[u] > 0) P[k] += w[k]; else Ng[k] += w[k];
                    Sq[k] += w[k] + (w[k] & 0x02020202u);  // v*v for v in {0,1,2}: 0, 1, 4
                }
            }
            cnt += U;
            if (cnt + U > 63) spill();
        }
        spill();
    }
    if (badacc) atomicOr(bad, 1);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        red[warp][k][lane] = P16[k];
        red[warp][8 + k][lane] = N16[k];
        red[warp][16 + k][lane] = S16[k];
    }
    __syncthreads();
    if (!live) return;
    // warp m combines word m of P / N / S over the 8 warps: two columns each (16-bit lanes h = 0, 1 <-> byte j = 2h + (m & 1) of word m >> 1)
    const int m = warp;
    int p[2] = {0, 0}, n[2] = {0, 0}, q[2] = {0, 0};
#pragma unroll
    for (int ww = 0; ww < 8; ww++) {
        const uint32_t pv = red[ww][m][lane], nv = red[ww][8 + m][lane], sv = red[ww][16 + m][lane];
        p[0] += pv & 0xffff; p[1] += pv >> 16;
        n[0] += nv & 0xffff; n[1] += nv >> 16;
        q[0] += sv & 0xffff; q[1] += sv >> 16;
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int k = 4 * (m >> 1) + 2 * h + (m & 1);
        if (k < nvalid) {
            if (p[h] != n[h]) atomicAdd(reinterpret_cast<unsigned long long *>(sketch) + (size_t)b * ncols + col0 + k, (unsigned long long)(long long)(p[h] - n[h]));
            if (p[h] + n[h]) atomicAdd(xsum + col0 + k, (unsigned long long)(p[h] + n[h]));
            if (q[h]) atomicAdd(x2sum + col0 + k, (unsigned long long)q[h]);
        }
    }
}

__global__ void k_i64_to_f64(const long long *__restrict__ in, double *__restrict__ out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = (double)in[i];
}

// rows_sorted / sgn_sorted: rows grouped by bucket (device [nrows]), bucket_off device [kp+1]; sketch_i64 [kp][ncols], xsum, x2sum
// [ncols] zero-initialised by the caller; sketch_f64 [kp][ncols]; bad: device int, set to 1 when a dosage outside {0,1,2} was seen
int launch_count_sketch(Ctx *c, const int8_t *X, size_t nrows, size_t ncols, const int *rows_sorted, const int8_t *sgn_sorted,
                        const int *bucket_off, int kp, long long *sketch_i64, double *sketch_f64, unsigned long long *xsum,
                        unsigned long long *x2sum, int *bad, cudaStream_t st) {
    if (nrows == 0 || ncols == 0 || kp <= 0) return 0;
    const unsigned tiles = (unsigned)((ncols + 511) / 512);
    // enough CTAs to fill the 148 SMs several times over; and at most 252 * 256 rows per warp (16-bit lanes: 65535 / 4 per row... the
    // squares add up to 4 per row, so 16 000 rows per warp keep every lane below 2^16)
    size_t nsplit = std::max<size_t>(1, (size_t)(148 * 8 + tiles * kp - 1) / ((size_t)tiles * kp));
    nsplit = std::max(nsplit, (nrows + 8 * 16000 - 1) / (8 * 16000));
    nsplit = std::min<size_t>(nsplit, 65535);
    dim3 g(tiles, kp, (unsigned)nsplit);
    const bool aligned = ncols % 16 == 0 && ((uintptr_t)X % 16) == 0;
    if (aligned)
        k_count_sketch<true><<<g, 256, 0, st>>>(X, ncols, rows_sorted, sgn_sorted, bucket_off, (int)nsplit, sketch_i64, xsum, x2sum, bad);
    else
        k_count_sketch<false><<<g, 256, 0, st>>>(X, ncols, rows_sorted, sgn_sorted, bucket_off, (int)nsplit, sketch_i64, xsum, x2sum, bad);
    SFG_LAUNCHED(c, "k_count_sketch", st);
    const size_t n = (size_t)kp * ncols;
    k_i64_to_f64<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(sketch_i64, sketch_f64, n);
    SFG_LAUNCHED(c, "k_i64_to_f64", st);
    return 0;
}

}  // namespace sfg
