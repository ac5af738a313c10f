// kernels_encode.cu -- on-device generation of the NTT-domain plaintext diagonals (K5 + K4 + K3 of SURVEY 2.2):
//   GetDiag (gwas/matmult.go:636-664)  ->  convertToComplex128WithRot (:666-672)  ->  EncoderBig.EncodeNTT (:711-731,
//   Lattigo App. B.6)  ->  ToMontgomeryForm (:401-440).
//
// One CTA per diagonal polynomial.  The CKKS canonical-embedding inverse ("special" inverse FFT over the rotation
// group 5^j) runs in FP64 in shared memory.  The reference rounds scale*w_k computed with 256-bit floats, i.e. the
// correctly rounded value.  FP64 alone is not bit-exact, so every coefficient whose scaled value lies within
// `delta` of a half-integer is recomputed exactly-enough by a direct double-double sum of the n terms
// v_j * cos/sin(2 pi 5^j k / 2N) (error ~1e-25), which resolves the rounding.  Coefficients that are still within
// 1e-13 of a tie after that are counted in stats[1] (never observed; a genuine tie needs an exact half-integer).
#include <cstdlib>

#include "kernels.h"
#include "ntt.cuh"
#include "ntt2.cuh"

namespace sfg {

// ---- double-double helpers ----
struct dd {
    double hi, lo;
};
__device__ __forceinline__ dd two_sum(double a, double b) {
    double s = a + b, bb = s - a;
    return dd{s, (a - (s - bb)) + (b - bb)};
}
__device__ __forceinline__ dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    dd t = two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return two_sum(s.hi, s.lo);
}
__device__ __forceinline__ dd dd_mul_d(dd a, double b) {
    double p = a.hi * b;
    double e = __fma_rn(a.hi, b, -p);
    e = __fma_rn(a.lo, b, e);
    return two_sum(p, e);
}

constexpr int kMaxFlag = 512;  // near-tie re-checks per polynomial (observed: ~0.1 at PN13QP218, ~2 at PN14QP438); more are counted as unresolved

// Residue of the integer message coefficient v (exact in FP64, |v| < 2^50) in the input range of the class's forward butterflies.
template <class A>
__device__ __forceinline__ typename A::T msg_residue(double v, const typename A::C &c, const LimbConst &lc) {
    if constexpr (A::kKind == kArD) {
        return ArD::red(v, c);  // (-q, q): the FP64 class works on signed representatives
    } else {
        uint64_t r = bred_add((uint64_t)fabs(v), lc);
        if (v < 0.0 && r != 0) r = lc.q - r;
        return A::from_canon(r, c);
    }
}

// NTT of the message over one limb with the register-tiled, class-specialised transform of ntt2.cuh (radix-8/16 passes: 3 CTA barriers
// for 2^13 coefficients instead of 13), canonical residues out.  m: the integer message [N] in shared memory; s: transform scratch.
// CS = 1: the ring is cut in two halves transformed one after the other (stage 0 folded into the first pass), which keeps the scratch at
// 2^13 coefficients for logN = 14.  The result is staged in `s` and leaves with unit-stride stores.
// LOGN > 0: the ring size is a compile-time constant (13 -> whole ring, 14 -> two halves), so the pass plan, strides and indices fold
// into immediates (a third of the executed instructions of the generic version are address arithmetic, as in the key-switch kernels).
template <class A, int LOGN>
__device__ __forceinline__ void encode_limb(const double *__restrict__ m, void *sraw, int logN_arg, int CS_arg, const TwTab &tab,
                                            const LimbConst &lc, unsigned char *__restrict__ o, int es, int mont) {
    using T = typename A::T;
    using TW = typename A::TW;
    T *s = reinterpret_cast<T *>(sraw);
    const typename A::C c = A::make(lc);
    const int logN = LOGN ? LOGN : logN_arg;
    const int CS = LOGN ? (LOGN > 13 ? 1 : 0) : CS_arg;
    const int logS = logN - CS, S = 1 << logS;
    const PassPlan plan = make_pass_plan(logS - kLastR);
    for (int sl = 0; sl < (1 << CS); sl++) {
        auto ld0 = [&](int j, int) -> T {
            if (CS == 0) return msg_residue<A>(m[j], c, lc);
            T x = msg_residue<A>(m[j], c, lc), y = msg_residue<A>(m[j + S], c, lc);
            A::fwd(x, y, __ldg(reinterpret_cast<const TW *>(tab.fwd) + 1), c);
            return sl ? y : x;
        };
        auto fin = [&](int j, T v, int) { s[sidx<sizeof(T)>(j)] = (T)A::canon(v, c); };
        ntt_forward<A>(s, logN, logS, sl, plan, tab, c, ld0, fin);
        __syncthreads();
        // (tried and measured slower, profiles/r2/ab_encode_variants.txt and DESIGN.md: 16-byte stores straight from the last pass's
        //  registers -- 128-byte stride between lanes, +9 % at logN 14 --, and twiddles staged in shared memory -- no gain at logN 13,
        //  +12 % at logN 14, where the kernel sits at the 64-register cap)
        if (es == 4) {
            uint32_t *o32 = reinterpret_cast<uint32_t *>(o) + (size_t)sl * S;
            for (int k = threadIdx.x; k < S; k += blockDim.x) o32[k] = (uint32_t)s[sidx<sizeof(T)>(k)];
        } else {
            uint64_t *o64 = reinterpret_cast<uint64_t *>(o) + (size_t)sl * S;
            for (int k = threadIdx.x; k < S; k += blockDim.x) {
                const uint64_t x = (uint64_t)s[sidx<sizeof(T)>(k)];
                o64[k] = mont ? mform(x, lc) : x;
            }
        }
        __syncthreads();
    }
}

// index of element i of the FFT buffers: one pad element per 512 makes the bit-reversed read of step 3 conflict-free (consecutive
// threads read positions 512 apart: all in one bank otherwise) and leaves the unit-stride accesses of the FFT passes alone
__device__ __forceinline__ int fidx(int i) { return i + (i >> 9); }
__host__ __device__ inline size_t fft_buf_bytes(int N) { return ((size_t)N + 2 * ((size_t)N >> 10)) * 8; }  // re + im, padded

// The front part (steps 1-4) of the rings the reference's parameter sets use (logN 13, 14) is compiled with the ring shape as a constant,
// keeps all loads of the gather in flight at once and pads the FFT buffers ("NEW"); the generic kernel keeps the plain loops.
// Measured per diagonal incl. image build (profiles/microbench/ab_encode.py, profiles/r2/ab_encode_*.txt): logN 13 0.551 -> 0.521 us,
// logN 14 1.453 -> 1.416 us.  At logN 14 the kernel sits at the 64-register cap of 1 024 threads (NPER = 16: 32 registers of rounded
// message): with the front part INLINED each of these changes alone cost 3-7 % and all together 20 % (1.751 us) through the register
// allocation of the limb transforms behind it; as a function of its own they gain.  SFG_ENC_NEW14 / SFG_ENC_FRONT_ATTR: A/B builds.
#ifndef SFG_ENC_NEW14
#define SFG_ENC_NEW14 1
#endif
#ifndef SFG_ENC_FRONT_ATTR
#define SFG_ENC_FRONT_ATTR __noinline__
#endif
__host__ __device__ constexpr bool enc_new_front(int LOGN) { return LOGN == 13 || (SFG_ENC_NEW14 && LOGN == 14); }

// Steps 1-4: gather, special inverse FFT, rounding with exact re-check of near ties; leaves the integer message [N] as exact FP64
// integers at the start of the CTA's shared memory.  Its own function (not inlined): its register allocation and that of the limb
// transforms after it do not disturb each other under the 64-register cap.
//
// NEW: compile-time ring shape, all loads of the gather in flight at once, padded FFT buffers.
template <int NPER, int LOGN, bool NEW>
__device__ SFG_ENC_FRONT_ATTR void encode_front(unsigned char *smem_raw, const int8_t *__restrict__ X, size_t ld, const EncJob &job, int logN_arg,
                                                double sc, double delta, const int *__restrict__ rot5, const double2 *__restrict__ ddcos,
                                                long long *__restrict__ coeff_out, unsigned long long *__restrict__ stats,
                                                const double2 *__restrict__ fft_tw) {
    const int logN = NEW ? LOGN : logN_arg;
    const int N = 1 << logN, n = N >> 1, M = N << 1, logn = logN - 1;
    double *re = reinterpret_cast<double *>(smem_raw);
    double *im = re + n + (NEW ? n >> 9 : 0);
    int8_t *vals = reinterpret_cast<int8_t *>(smem_raw + (NEW ? fft_buf_bytes(N) : (size_t)N * 8));
    int *flag_idx = reinterpret_cast<int *>(vals + n);
    long long *flag_val = reinterpret_cast<long long *>(flag_idx + kMaxFlag);
    dd *red = reinterpret_cast<dd *>(flag_val + kMaxFlag);  // [32] warp partials
    __shared__ int nflag;

    const int T = blockDim.x, tid = threadIdx.x;  // T = N / NPER: every loop below has a compile-time trip count
    if (tid == 0) nflag = 0;

    if constexpr (NEW) {
        // 1. gather the generalized diagonal, right-rotated by nrot:  v[(j+nrot) mod n] = X[(shift+j) mod n][j].  One byte per 32-byte sector
        //    of X: all loads of a thread are issued before the first use.
        {
            constexpr int G = NPER / 2;
            int8_t gv[G];
    #pragma unroll
            for (int it = 0; it < G; it++) {
                const int j = tid + it * T;
                int row = job.shift + j;
                if (row >= n) row -= n;
                gv[it] = (row < job.r && j < job.cdim) ? X[(size_t)(job.row0 + row) * ld + job.col0 + j] : (int8_t)0;
            }
    #pragma unroll
            for (int it = 0; it < G; it++) {
                int dst = tid + it * T + job.nrot;
                if (dst >= n) dst -= n;
                vals[dst] = gv[it];
                re[fidx(dst)] = (double)gv[it];
                im[fidx(dst)] = 0.0;
            }
        }
        __syncthreads();

        // 2. special inverse FFT (Lattigo invfft, App. B.6), decimation in frequency, result in bit-reversed order.  Two stages per CTA
        //    barrier (radix 4 in registers), twiddles from the per-stage table in butterfly order (unit stride: no 5^j / root gathers).
        {
            auto bfly = [](double &ar, double &ai, double &br, double &bi, const double2 w) {
                const double ur = ar + br, ui = ai + bi, vr = ar - br, vi = ai - bi;
                ar = ur;
                ai = ui;
                br = vr * w.x - vi * w.y;
                bi = vr * w.y + vi * w.x;
            };
            int len = n, loglen = logn;
    #pragma unroll
            for (; len >= 4; len >>= 2, loglen -= 2) {
                const int lenq4 = len >> 2, lenh = len >> 1;
    #pragma unroll
                for (int it = 0; it < NPER / 8; it++) {
                    const int b = tid + it * T;
                    const int grp = b >> (loglen - 2), j = b & (lenq4 - 1);
                    const int i0 = (grp << loglen) + j;
                    const int a0 = fidx(i0), a1 = fidx(i0 + lenq4), a2 = fidx(i0 + lenh), a3 = fidx(i0 + lenh + lenq4);
                    double r0 = re[a0], m0 = im[a0], r1 = re[a1], m1 = im[a1], r2 = re[a2], m2 = im[a2], r3 = re[a3], m3 = im[a3];
                    const double2 wa = __ldg(fft_tw + lenh + j), wb = __ldg(fft_tw + lenh + j + lenq4), wc = __ldg(fft_tw + lenq4 + j);
                    bfly(r0, m0, r2, m2, wa);  // stage of length len: pairs (i, i + len/2)
                    bfly(r1, m1, r3, m3, wb);
                    bfly(r0, m0, r1, m1, wc);  // stage of length len/2 inside both halves: pairs (i, i + len/4), same twiddle index j
                    bfly(r2, m2, r3, m3, wc);
                    re[a0] = r0; im[a0] = m0; re[a1] = r1; im[a1] = m1; re[a2] = r2; im[a2] = m2; re[a3] = r3; im[a3] = m3;
                }
                __syncthreads();
            }
            if (len == 2) {  // odd number of stages: the last one alone
                const double2 w1 = __ldg(fft_tw + 1);
    #pragma unroll
                for (int it = 0; it < NPER / 4; it++) {
                    const int b = tid + it * T;
                    const int a0 = fidx(b << 1), a1 = fidx((b << 1) + 1);
                    double r0 = re[a0], m0 = im[a0], r1 = re[a1], m1 = im[a1];
                    bfly(r0, m0, r1, m1, w1);
                    re[a0] = r0; im[a0] = m0; re[a1] = r1; im[a1] = m1;
                }
                __syncthreads();
            }
        }
    } else {
        // 1. gather the generalized diagonal, right-rotated by nrot:  v[(j+nrot) mod n] = X[(shift+j) mod n][j]
        for (int j = tid; j < n; j += T) {
            int row = job.shift + j;
            if (row >= n) row -= n;
            int8_t v = 0;
            if (row < job.r && j < job.cdim) v = X[(size_t)(job.row0 + row) * ld + job.col0 + j];
            int dst = j + job.nrot;
            if (dst >= n) dst -= n;
            vals[dst] = v;
            re[dst] = (double)v;
            im[dst] = 0.0;
        }
        __syncthreads();

        // 2. special inverse FFT (Lattigo invfft, App. B.6), decimation in frequency, result in bit-reversed order.  Two stages per CTA
        //    barrier (radix 4 in registers), twiddles from the per-stage table in butterfly order (unit stride: no 5^j / root gathers).
        {
            auto bfly = [](double &ar, double &ai, double &br, double &bi, const double2 w) {
                const double ur = ar + br, ui = ai + bi, vr = ar - br, vi = ai - bi;
                ar = ur;
                ai = ui;
                br = vr * w.x - vi * w.y;
                bi = vr * w.y + vi * w.x;
            };
            int len = n, loglen = logn;
            for (; len >= 4; len >>= 2, loglen -= 2) {
                const int lenq4 = len >> 2, lenh = len >> 1;
                for (int b = tid; b < (n >> 2); b += T) {
                    const int grp = b >> (loglen - 2), j = b & (lenq4 - 1);
                    const int i0 = (grp << loglen) + j, i1 = i0 + lenq4, i2 = i0 + lenh, i3 = i2 + lenq4;
                    double r0 = re[i0], m0 = im[i0], r1 = re[i1], m1 = im[i1], r2 = re[i2], m2 = im[i2], r3 = re[i3], m3 = im[i3];
                    const double2 wa = fft_tw[lenh + j], wb = fft_tw[lenh + j + lenq4], wc = fft_tw[lenq4 + j];
                    bfly(r0, m0, r2, m2, wa);  // stage of length len: pairs (i, i + len/2)
                    bfly(r1, m1, r3, m3, wb);
                    bfly(r0, m0, r1, m1, wc);  // stage of length len/2 inside both halves: pairs (i, i + len/4), same twiddle index j
                    bfly(r2, m2, r3, m3, wc);
                    re[i0] = r0; im[i0] = m0; re[i1] = r1; im[i1] = m1; re[i2] = r2; im[i2] = m2; re[i3] = r3; im[i3] = m3;
                }
                __syncthreads();
            }
            if (len == 2) {  // odd number of stages: the last one alone
                for (int b = tid; b < (n >> 1); b += T) {
                    const int i0 = b << 1, i1 = i0 + 1;
                    double r0 = re[i0], m0 = im[i0], r1 = re[i1], m1 = im[i1];
                    bfly(r0, m0, r1, m1, fft_tw[1]);
                    re[i0] = r0; im[i0] = m0; re[i1] = r1; im[i1] = m1;
                }
                __syncthreads();
            }
        }
    }

    // 3. scale, round half away from zero, flag near-ties
    long long m[NPER];
#pragma unroll
    for (int r = 0; r < NPER; r++) {
        const int k = tid + r * T;
        const int kk = k < n ? k : k - n;
        const int pos = __brev((unsigned)kk) >> (32 - logn);
        const int pp = NEW ? fidx(pos) : pos;
        const double x = (k < n ? re[pp] : im[pp]) * sc;
        const double ax = fabs(x);
        const double fl = floor(ax);
        const double fr = ax - fl;
        long long v = (long long)fl + (fr >= 0.5 ? 1 : 0);
        m[r] = x < 0 ? -v : v;
        if (fabs(fr - 0.5) < delta) {
            const int f = atomicAdd(&nflag, 1);
            if (f < kMaxFlag) flag_idx[f] = k;
        }
    }
    __syncthreads();

    // 4. exact re-evaluation of the flagged coefficients in double-double
    const int nf = nflag < kMaxFlag ? nflag : kMaxFlag;
    if (nflag > 0) {
        for (int f = 0; f < nf; f++) {
            const int k = flag_idx[f];
            const int kk = k < n ? k : k - n;
            const int shiftq = k < n ? 0 : (M >> 2);  // imaginary part: -sin(t) = -cos(t - pi/2)
            dd acc{0.0, 0.0};
            for (int j = tid; j < n; j += T) {
                const int v = vals[j];
                if (v != 0) {
                    const int t = (int)(((long long)rot5[j] * kk - shiftq) & (M - 1));
                    const double2 cs = ddcos[t];
                    acc = dd_add(acc, dd_mul_d(dd{cs.x, cs.y}, (double)v));
                }
            }
            // block reduction
            for (int o = 16; o > 0; o >>= 1) {
                dd other{__shfl_down_sync(0xffffffffu, acc.hi, o), __shfl_down_sync(0xffffffffu, acc.lo, o)};
                acc = dd_add(acc, other);
            }
            if ((tid & 31) == 0) red[tid >> 5] = acc;
            __syncthreads();
            if (tid < 32) {  // warp 0 combines the warp partials with a shuffle tree (a serial sum by one thread was ~5 000 cycles of dependent FP64)
                dd tot = tid < (T + 31) / 32 ? red[tid] : dd{0.0, 0.0};
                for (int o = 16; o > 0; o >>= 1) {
                    dd other{__shfl_down_sync(0xffffffffu, tot.hi, o), __shfl_down_sync(0xffffffffu, tot.lo, o)};
                    tot = dd_add(tot, other);
                }
              if (tid == 0) {
                if (k >= n) { tot.hi = -tot.hi; tot.lo = -tot.lo; }
                dd x = dd_mul_d(tot, sc);
                const bool neg = x.hi < 0;
                if (neg) { x.hi = -x.hi; x.lo = -x.lo; }
                double ip = floor(x.hi);
                double fr = (x.hi - ip) + x.lo;
                if (fr < 0) { ip -= 1.0; fr += 1.0; }
                if (fr >= 1.0) { ip += 1.0; fr -= 1.0; }
                long long v = (long long)ip + (fr >= 0.5 ? 1 : 0);
                flag_val[f] = neg ? -v : v;
                atomicAdd(&stats[0], 1ULL);
                if (fabs(fr - 0.5) < 1e-13) atomicAdd(&stats[1], 1ULL);
              }
            }
            __syncthreads();
        }
        if (nflag > kMaxFlag && tid == 0) atomicAdd(&stats[1], (unsigned long long)(nflag - kMaxFlag));
#pragma unroll
        for (int r = 0; r < NPER; r++) {
            const int k = tid + r * T;
            for (int f = 0; f < nf; f++)
                if (flag_idx[f] == k) m[r] = flag_val[f];
        }
        __syncthreads();
    }

    if (coeff_out) {
#pragma unroll
        for (int r = 0; r < NPER; r++) coeff_out[(size_t)blockIdx.x * N + tid + r * T] = m[r];
    }

    // the message goes to shared memory as exact FP64 integers (over the FFT buffers: every thread has read its values)
    __syncthreads();
    double *mm = reinterpret_cast<double *>(smem_raw);
#pragma unroll
    for (int r = 0; r < NPER; r++) mm[tid + r * T] = (double)m[r];
    __syncthreads();
}

template <int NPER, int LOGN>
__global__ void __launch_bounds__(1024, 1)
k_encode(const int8_t *__restrict__ X, size_t ld, const EncJob *__restrict__ jobs, int logN_arg, PolyLayout lay, int mont, double sc,
         double delta, const double2 *__restrict__ roots, const int *__restrict__ rot5, const double2 *__restrict__ ddcos,
         const uint64_t *__restrict__ tw, const LimbConst *__restrict__ lcs, unsigned char *__restrict__ out,
         long long *__restrict__ coeff_out, unsigned long long *__restrict__ stats, const TwTab *__restrict__ tabs2, int CS,
         const double2 *__restrict__ fft_tw) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool NEW = enc_new_front(LOGN);
    const int logN = LOGN ? LOGN : logN_arg;
    const int N = 1 << logN, n = N >> 1;
    const int T = blockDim.x, tid = threadIdx.x;
    const EncJob job = jobs[blockIdx.x];
    encode_front<NPER, LOGN, NEW>(smem_raw, X, ld, job, logN_arg, sc, delta, rot5, ddcos, coeff_out, stats, fft_tw);
    double *mm = reinterpret_cast<double *>(smem_raw);  // the integer message
    // scratch of the register-tiled limb transforms, behind the buffers of encode_front
    void *s2 = smem_raw + (NEW ? fft_buf_bytes(N) : (size_t)N * 8) + n + kMaxFlag * (sizeof(int) + sizeof(long long)) + 32 * sizeof(dd);

    // 5. RNS reduce, NTT per limb, Montgomery form, store
    if (tabs2) {
        // one class-specialised register-tiled transform per limb
        for (int l = 0; l < lay.nl; l++) {
            const LimbConst lc = lcs[l];
            unsigned char *o = out + job.out_off + lay.off[l];
            switch (arith_kind(lc.q)) {  // uniform across the CTA
                case kArN30: encode_limb<ArN30, LOGN>(mm, s2, logN, CS, tabs2[l], lc, o, lay.es[l], mont); break;
                case kArN31: encode_limb<ArN31, LOGN>(mm, s2, logN, CS, tabs2[l], lc, o, lay.es[l], mont); break;
                case kArD: encode_limb<ArD, LOGN>(mm, s2, logN, CS, tabs2[l], lc, o, lay.es[l], mont); break;
                default: encode_limb<ArW, LOGN>(mm, s2, logN, CS, tabs2[l], lc, o, lay.es[l], mont); break;
            }
        }
        return;
    }
    uint64_t *s = reinterpret_cast<uint64_t *>(smem_raw);  // radix-2 shared-memory transform in place of the message
    long long m[NPER];
#pragma unroll
    for (int r = 0; r < NPER; r++) m[r] = (long long)mm[tid + r * T];
    __syncthreads();
    for (int l = 0; l < lay.nl; l++) {
        const LimbConst lc = lcs[l];
        const NttTab tab = ntt_tab(tw, l, N);
#pragma unroll
        for (int r = 0; r < NPER; r++) {
            const long long v = m[r];
            uint64_t x = bred_add((uint64_t)(v < 0 ? -v : v), lc);
            if (v < 0 && x != 0) x = lc.q - x;
            s[tid + r * T] = x;
        }
        __syncthreads();
        ntt_fwd_smem(s, logN, 1, 0, tab, lc.q);
        unsigned char *o = out + job.out_off + lay.off[l];
        if (lay.es[l] == 4) {  // packed narrow limb: plain residue
#pragma unroll
            for (int r = 0; r < NPER; r++) reinterpret_cast<uint32_t *>(o)[tid + r * T] = (uint32_t)s[tid + r * T];
        } else {
#pragma unroll
            for (int r = 0; r < NPER; r++) {
                const uint64_t x = s[tid + r * T];
                reinterpret_cast<uint64_t *>(o)[tid + r * T] = mont ? mform(x, lc) : x;
            }
        }
        __syncthreads();
    }
}

int launch_encode(Ctx *c, const int8_t *X, size_t ld, const EncJob *jobs_dev, int njobs, const PolyLayout &lay, bool mont, void *out,
                  long long *coeff_out, cudaStream_t st) {
    if (njobs <= 0) return 0;
    const int logN = c->logN, N = c->N, n = c->slots;
    if (logN < 8 || logN > 14) SFG_FAIL(c, "on-device diagonal encoder supports 8 <= logN <= 14 (got %d)", logN);
    const int NPER = logN == 14 ? 16 : 8;
    const int T = N / NPER;
    // limb transforms: register-tiled passes of ntt2.cuh (SFG_ENC_OLDNTT=1 keeps the radix-2 shared-memory transform for A/B runs);
    // logN = 14 transforms the ring as two halves so that message + scratch fit one CTA's shared memory
    static const bool old_ntt = [] { const char *e = getenv("SFG_ENC_OLDNTT"); return e && *e == '1'; }();
    const TwTab *tabs2 = old_ntt ? nullptr : c->tw2;
    const int CS = logN > 13 ? 1 : 0;
    const size_t scratch = tabs2 ? ntt_smem_elems(N >> CS) * 8 + 16 : 0;
    const bool newk = logN == 14 ? enc_new_front(14) : (logN == 13 && tabs2);  // the kernel launched below pads its FFT buffers
    const size_t smem = (newk ? fft_buf_bytes(N) : (size_t)N * 8) + n + kMaxFlag * (sizeof(int) + sizeof(long long)) + 32 * sizeof(dd) + 64 + scratch;
    const double sc = c->scale / (double)n;
    auto go = [&](auto kern) -> int {
        SFG_CUDA(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<njobs, T, smem, st>>>(X, ld, jobs_dev, logN, lay, mont ? 1 : 0, sc, c->enc_delta, c->roots, c->rot5, c->ddcos, c->tw, c->lc,
                                     (unsigned char *)out, coeff_out, c->enc_stats, tabs2, CS, c->fft_tw);
        return 0;
    };
    // the rings of the reference's parameter sets are compiled with constant shapes; anything else takes the generic kernel
    if (logN == 14 ? go(k_encode<16, 14>) : (logN == 13 && tabs2 ? go(k_encode<8, 13>) : go(k_encode<8, 0>))) return -1;
    SFG_LAUNCHED(c, "k_encode", st);
    return 0;
}

}  // namespace sfg
