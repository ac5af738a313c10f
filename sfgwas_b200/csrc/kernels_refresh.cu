// kernels_refresh.cu -- SURVEY 8f row 4: the LOCAL arithmetic of the collective bootstrap (mpc/mhe.go:262-341), i.e. what every party
// computes per ciphertext around the two network aggregations:
//     refProtocol.GenShares(skShard, levelStart, nParties-1, ct, scale, crp, share1, share2)            mhe.go:303-311
//     refProtocol.Decrypt(ct, agg1); refProtocol.Recode(ct, scale); refProtocol.Recrypt(ct, crp, agg2)  mhe.go:316-318
// The protocol is Lattigo's dckks.RefreshProtocol (un-vendored fork; restated in oracle/refresh.py from the published v2.1 algorithm,
// [UNVERIFIED vs the fork]).  Everything random -- the masks (crypto/rand), the Gaussian noise, the common reference polynomials of the
// shared PRG -- is drawn by the Go side and handed over; what runs here is exact integer arithmetic, embarrassingly parallel over the
// s * m_ct ciphertexts of a MatMult output (BootstrapMatAll right after every MatMult4StreamCompute, gwas/matmult.go:48,93):
//   GenShares : h0 = NTT(mask mod q_l) + sk * c1 + NTT(e0)   (l <= level),   h1 = -(NTT(mask mod q_l) + sk * a + NTT(e1))   (all nQ limbs)
//   finish    : c0 += agg(h0); INTT; centred CRT reconstruction mod Q_level (Garner mixed radix + multi-word Horner);
//               x -> Quo(x * floor(targetScale), floor(ct.Scale)) (big.Int.Quo: truncation toward zero; bit-serial long division);
//               reduce into ALL nQ limbs (big.Int.Mod: non-negative); NTT; + agg(h1); c1 = a.
#include <algorithm>
#include <vector>

#include "matmult.h"

namespace sfg {

constexpr int kBigWords = 32;  // multi-word integers of the recode: 2048 bits (full PN16QP1761 chain x scale)

// r = (|x| as nw little-endian words) mod q, negated mod q when neg
__device__ __forceinline__ uint64_t words_mod(const uint64_t *w, int nw, bool neg, const LimbConst &lc) {
    uint64_t r = 0;
    for (int i = nw - 1; i >= 0; i--) r = add_mod(mul_shoup(r, lc.r64, lc.r64_sh, lc.q), bred_add(w[i], lc), lc.q);  // r * 2^64 + w_i
    return (neg && r != 0) ? lc.q - r : r;
}

// X (nw little-endian words, magnitude) <- floor(X * so / si): big.Int Mul + Quo on the magnitude (Quo truncates toward zero, so the sign
// is handled by the caller).  Bit-serial restoring division by the (up to) 128-bit divisor; X * so must fit nw words.
__device__ __forceinline__ void big_scale_quo(uint64_t *X, int nw, uint64_t so, uint64_t d0, uint64_t d1) {
    uint64_t carry = 0;
    for (int w = 0; w < nw; w++) {
        const uint64_t lo = X[w] * so, hi = __umul64hi(X[w], so);
        const uint64_t s = lo + carry;
        carry = hi + (s < lo);
        X[w] = s;
    }
    uint64_t r0 = 0, r1 = 0, r2 = 0;  // running remainder (< 2 * si < 2^129)
    for (int w = nw - 1; w >= 0; w--) {
        const uint64_t yw = X[w];
        uint64_t qw = 0;
        for (int b = 63; b >= 0; b--) {
            r2 = (r2 << 1) | (r1 >> 63);
            r1 = (r1 << 1) | (r0 >> 63);
            r0 = (r0 << 1) | ((yw >> b) & 1);
            const bool big = r2 != 0 || r1 > d1 || (r1 == d1 && r0 >= d0);
            if (big) {
                const uint64_t br = r0 < d0;
                const uint64_t br2 = (r1 < d1) || (r1 == d1 && br);
                r0 -= d0;
                r1 = r1 - d1 - br;
                r2 -= br2;
                qw |= 1ULL << b;
            }
        }
        X[w] = qw;
    }
}

// the same integers scaled first: x -> Quo(x * so, si) ("scales the mask by the ratio between the two scales" of GenShares), then reduced
template <int W>
__global__ void k_bigint_scaled_to_rns(const uint64_t *__restrict__ mag, const int8_t *__restrict__ sign, int nw_in, int nw, uint64_t so, uint64_t d0,
                                       uint64_t d1, int nl, int N, const LimbConst *__restrict__ lcs, uint64_t *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (j >= N) return;
    uint64_t X[W];
#pragma unroll
    for (int i = 0; i < W; i++) X[i] = 0;
    for (int i = 0; i < nw_in; i++) X[i] = mag[((size_t)p * N + j) * nw_in + i];
    big_scale_quo(X, nw, so, d0, d1);
    const bool neg = sign[(size_t)p * N + j] < 0;
    for (int l = 0; l < nl; l++) out[((size_t)p * nl + l) * N + j] = words_mod(X, nw, neg, lcs[l]);
}

// ring.SetCoefficientsBigintLvl: sign-magnitude multi-word integers [npoly][N][nw] -> residues [npoly][nl][N] (big.Int.Mod, non-negative)
__global__ void k_bigint_to_rns(const uint64_t *__restrict__ mag, const int8_t *__restrict__ sign, int nw, int nl, int N,
                                const LimbConst *__restrict__ lcs, uint64_t *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (j >= N) return;
    const uint64_t *w = mag + ((size_t)p * N + j) * nw;
    const bool neg = sign[(size_t)p * N + j] < 0;
    for (int l = 0; l < nl; l++) out[((size_t)p * nl + l) * N + j] = words_mod(w, nw, neg, lcs[l]);
}
// small signed coefficients (the Gaussian noise polynomials) -> residues [npoly][nl][N]
__global__ void k_i64_to_rns(const long long *__restrict__ e, int nl, int N, const LimbConst *__restrict__ lcs, uint64_t *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (j >= N) return;
    const long long v = e[(size_t)p * N + j];
    const uint64_t a = (uint64_t)(v < 0 ? -v : v);
    for (int l = 0; l < nl; l++) {
        const uint64_t r = bred_add(a, lcs[l]);
        out[((size_t)p * nl + l) * N + j] = (v < 0 && r != 0) ? lcs[l].q - r : r;
    }
}
// h = mask + sk * x + e (mod q), optionally negated.  sk is Lattigo's NTT + Montgomery form: MRed(sk, x) is the plain product
// (ring.MulCoeffsMontgomeryAndAdd).  x: [npoly][x_nl][N] (c1 of the ciphertext, or the common reference polynomial)
__global__ void k_refresh_share(const uint64_t *__restrict__ mask, const uint64_t *__restrict__ sk, const uint64_t *__restrict__ x, int x_nl,
                                const uint64_t *__restrict__ e, int nl, int N, const LimbConst *__restrict__ lcs, int negate,
                                uint64_t *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y, p = blockIdx.z;
    if (j >= N) return;
    const LimbConst lc = lcs[l];
    const size_t o = ((size_t)p * nl + l) * N + j;
    uint64_t h = add_mod(mask[o], mred(sk[(size_t)l * N + j], x[((size_t)p * x_nl + l) * N + j], lc), lc.q);
    h = add_mod(h, e[o], lc.q);
    out[o] = (negate && h != 0) ? lc.q - h : h;
}
// out = (a + b) mod q over [npoly][nl][N]; a may have more stored limbs per polynomial (a_nl >= nl)
__global__ void k_poly_add(const uint64_t *__restrict__ a, int a_nl, const uint64_t *__restrict__ b, int nl, int N,
                           const LimbConst *__restrict__ lcs, uint64_t *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y, p = blockIdx.z;
    if (j >= N) return;
    const size_t o = ((size_t)p * nl + l) * N + j;
    out[o] = add_mod(a[((size_t)p * a_nl + l) * N + j], b[o], lcs[l].q);
}

// Tables of the recode at one level (device): Garner inverses inv[i][k] = q_k^-1 mod q_i (k < i) with Shoup companions, and the words of
// Q_level and floor(Q_level / 2)
struct RecodeTab {
    int nl, nw;                      // limbs of Q_level, words of the big integers (enough for Q_level * outScale)
    uint64_t so;                     // floor(targetScale)
    uint64_t si[2];                  // floor(ct.Scale) as 128 bits
    uint64_t Q[kBigWords], Qhalf[kBigWords];
};

// One thread per coefficient: residues x_l of INTT(c0) (l < nl) -> centred integer -> * so / si (toward zero) -> residues mod all nQ limbs
template <int W>
__global__ void k_recode(const uint64_t *__restrict__ coeff /* [npoly][nl][N] */, const uint64_t *__restrict__ ginv /* [nl][nl][2] */,
                         const RecodeTab tab, int nQ, int N, const LimbConst *__restrict__ lcs, uint64_t *__restrict__ out /* [npoly][nQ][N] */) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (j >= N) return;
    const int nl = tab.nl, nw = tab.nw;
    // Garner: mixed-radix digits v_i < q_i with x = v_0 + v_1 q_0 + v_2 q_0 q_1 + ...
    uint64_t v[kMaxLimbs];
    for (int i = 0; i < nl; i++) {
        const LimbConst lc = lcs[i];
        uint64_t t = coeff[((size_t)p * nl + i) * N + j];
        for (int k = 0; k < i; k++) {
            const uint64_t vk = v[k] >= lc.q ? bred_add(v[k], lc) : v[k];
            t = mul_shoup(sub_mod(t, vk, lc.q), ginv[((size_t)i * nl + k) * 2], ginv[((size_t)i * nl + k) * 2 + 1], lc.q);
        }
        v[i] = t;
    }
    // Horner: X = (..(v_{nl-1} q_{nl-2} + v_{nl-2}) q_{nl-3} + ..) q_0 + v_0   (nw words, exact: X < Q_level)
    uint64_t X[W];
#pragma unroll
    for (int i = 0; i < W; i++) X[i] = 0;
    X[0] = v[nl - 1];
    for (int i = nl - 2; i >= 0; i--) {
        const uint64_t q = lcs[i].q;
        uint64_t carry = v[i];
        for (int w = 0; w < nw; w++) {
            const uint64_t lo = X[w] * q, hi = __umul64hi(X[w], q);
            const uint64_t s = lo + carry;
            carry = hi + (s < lo);
            X[w] = s;
        }
    }
    // centre: x >= floor(Q/2)  ->  x - Q  (magnitude Q - x, negative)
    bool ge = true;
    for (int w = nw - 1; w >= 0; w--)
        if (X[w] != tab.Qhalf[w]) {
            ge = X[w] > tab.Qhalf[w];
            break;
        }
    const bool neg = ge;
    if (neg) {
        uint64_t borrow = 0;
        for (int w = 0; w < nw; w++) {
            const uint64_t a = tab.Q[w], b = X[w];
            const uint64_t d = a - b - borrow;
            borrow = (a < b) || (a == b && borrow) ? 1 : 0;
            X[w] = d;
        }
    }
    // |x| -> floor(|x| * so / si)   (Mul + Quo: truncation toward zero)
    big_scale_quo(X, nw, tab.so, tab.si[0], tab.si[1]);
    for (int l = 0; l < nQ; l++) out[((size_t)p * nQ + l) * N + j] = words_mod(X, nw, neg, lcs[l]);
}

// ---- host side ---------------------------------------------------------------------------------------------------------------
static void big_mul_word(std::vector<uint64_t> &x, uint64_t m) {
    unsigned __int128 carry = 0;
    for (auto &w : x) {
        const unsigned __int128 t = (unsigned __int128)w * m + carry;
        w = (uint64_t)t;
        carry = t >> 64;
    }
}

int refresh_gen_shares_dev(Ctx *c, int level, int nct, const uint64_t *d_c1, const uint64_t *d_sk, const uint64_t *d_crp, const uint64_t *d_mask,
                           const int8_t *d_sign, int nwords, double in_scale, double out_scale, const long long *d_e0, const long long *d_e1,
                           uint64_t *d_h0, uint64_t *d_h1) {
    SFG_CUDA(c, cudaSetDevice(c->device));
    const int N = c->N, nl = level + 1, nQ = c->nQ;
    if (level < 0 || level >= nQ || nct < 1 || nwords < 1) SFG_FAIL(c, "refresh_gen_shares: bad level / counts");
    Buf m0, m1, n0, n1;
    if (m0.alloc(c, (size_t)nct * nl * N * 8) || m1.alloc(c, (size_t)nct * nQ * N * 8) || n0.alloc(c, (size_t)nct * nl * N * 8) ||
        n1.alloc(c, (size_t)nct * nQ * N * 8))
        return -1;
    const dim3 g((N + 255) / 256, nct);
    cudaStream_t st = c->stream;
    if (!(in_scale >= 1.0) || !(out_scale >= 1.0) || in_scale >= 3.0e38 || out_scale >= 9.2e18) SFG_FAIL(c, "refresh_gen_shares: scales out of range");
    const uint64_t so = (uint64_t)out_scale;  // big.Float.Int(): truncation
    const unsigned __int128 si = (unsigned __int128)in_scale;
    const int nw = nwords + 1;  // mask * so
    if (nw > kBigWords) SFG_FAIL(c, "refresh_gen_shares: masks of %d words are not supported (> %d)", nwords, kBigWords - 1);
    k_bigint_to_rns<<<g, 256, 0, st>>>(d_mask, d_sign, nwords, nl, N, c->lc, m0.as<uint64_t>());
    const dim3 g2((N + 127) / 128, nct);
    if (nw <= 8) k_bigint_scaled_to_rns<8><<<g2, 128, 0, st>>>(d_mask, d_sign, nwords, nw, so, (uint64_t)si, (uint64_t)(si >> 64), nQ, N, c->lc, m1.as<uint64_t>());
    else if (nw <= 16) k_bigint_scaled_to_rns<16><<<g2, 128, 0, st>>>(d_mask, d_sign, nwords, nw, so, (uint64_t)si, (uint64_t)(si >> 64), nQ, N, c->lc, m1.as<uint64_t>());
    else k_bigint_scaled_to_rns<kBigWords><<<g2, 128, 0, st>>>(d_mask, d_sign, nwords, nw, so, (uint64_t)si, (uint64_t)(si >> 64), nQ, N, c->lc, m1.as<uint64_t>());
    k_i64_to_rns<<<g, 256, 0, st>>>(d_e0, nl, N, c->lc, n0.as<uint64_t>());
    k_i64_to_rns<<<g, 256, 0, st>>>(d_e1, nQ, N, c->lc, n1.as<uint64_t>());
    c->launches += 3;
    SFG_LAUNCHED(c, "k_bigint_to_rns", st);
    LimbSel s0, s1;
    s0.n = nl;
    s1.n = nQ;
    for (int i = 0; i < nQ; i++) s1.idx[i] = i;
    for (int i = 0; i < nl; i++) s0.idx[i] = i;
    if (launch_ntt(c, m0.as<uint64_t>(), (size_t)nl * N, m0.as<uint64_t>(), (size_t)nl * N, nct * nl, s0, false, st) ||
        launch_ntt(c, n0.as<uint64_t>(), (size_t)nl * N, n0.as<uint64_t>(), (size_t)nl * N, nct * nl, s0, false, st) ||
        launch_ntt(c, m1.as<uint64_t>(), (size_t)nQ * N, m1.as<uint64_t>(), (size_t)nQ * N, nct * nQ, s1, false, st) ||
        launch_ntt(c, n1.as<uint64_t>(), (size_t)nQ * N, n1.as<uint64_t>(), (size_t)nQ * N, nct * nQ, s1, false, st))
        return -1;
    k_refresh_share<<<dim3((N + 255) / 256, nl, nct), 256, 0, st>>>(m0.as<uint64_t>(), d_sk, d_c1, nl, n0.as<uint64_t>(), nl, N, c->lc, 0, d_h0);
    k_refresh_share<<<dim3((N + 255) / 256, nQ, nct), 256, 0, st>>>(m1.as<uint64_t>(), d_sk, d_crp, nQ, n1.as<uint64_t>(), nQ, N, c->lc, 1, d_h1);
    c->launches += 1;
    SFG_LAUNCHED(c, "k_refresh_share", st);
    SFG_CUDA(c, cudaStreamSynchronize(st));
    return 0;
}

int refresh_finish_dev(Ctx *c, int level, int nct, const uint64_t *d_c0, int c0_nl, double in_scale, double out_scale, const uint64_t *d_agg0,
                       const uint64_t *d_agg1, const uint64_t *d_crp, uint64_t *d_out) {
    SFG_CUDA(c, cudaSetDevice(c->device));
    const int N = c->N, nl = level + 1, nQ = c->nQ;
    if (level < 0 || level >= nQ || nct < 1 || c0_nl < nl) SFG_FAIL(c, "refresh_finish: bad level / counts");
    if (!(in_scale >= 1.0) || !(out_scale >= 1.0) || in_scale >= 3.0e38 || out_scale >= 9.2e18) SFG_FAIL(c, "refresh_finish: scales out of range");
    // tables: Q_level, floor(Q_level / 2), Garner inverses
    RecodeTab tab{};
    tab.nl = nl;
    tab.so = (uint64_t)out_scale;  // big.Float.Int(): truncation
    const unsigned __int128 si = (unsigned __int128)in_scale;
    tab.si[0] = (uint64_t)si;
    tab.si[1] = (uint64_t)(si >> 64);
    int bits = 64 + 2;  // * so
    for (int l = 0; l < nl; l++) bits += 64 - __builtin_clzll(c->mod[l]);
    tab.nw = (bits + 63) / 64;
    if (tab.nw > kBigWords) SFG_FAIL(c, "refresh_finish: Q_level * scale needs %d words (> %d)", tab.nw, kBigWords);
    std::vector<uint64_t> Q(tab.nw, 0);
    Q[0] = 1;
    for (int l = 0; l < nl; l++) big_mul_word(Q, c->mod[l]);
    for (int w = 0; w < tab.nw; w++) {
        tab.Q[w] = Q[w];
        tab.Qhalf[w] = (Q[w] >> 1) | (w + 1 < tab.nw ? Q[w + 1] << 63 : 0);
    }
    std::vector<uint64_t> ginv((size_t)nl * nl * 2, 0);
    for (int i = 0; i < nl; i++)
        for (int k = 0; k < i; k++) {
            const uint64_t qi = c->mod[i], inv = h_invmod(c->mod[k] % qi, qi);
            ginv[((size_t)i * nl + k) * 2] = inv;
            ginv[((size_t)i * nl + k) * 2 + 1] = h_shoup(inv, qi);
        }
    Buf dg, t0, t1;
    if (dg.alloc(c, ginv.size() * 8) || t0.alloc(c, (size_t)nct * nl * N * 8) || t1.alloc(c, (size_t)nct * nQ * N * 8)) return -1;
    if (upload(c, dg.p, ginv.data(), ginv.size() * 8)) return -1;
    cudaStream_t st = c->stream;
    // Decrypt: c0 + agg(h0), then to the coefficient domain
    k_poly_add<<<dim3((N + 255) / 256, nl, nct), 256, 0, st>>>(d_c0, c0_nl, d_agg0, nl, N, c->lc, t0.as<uint64_t>());
    SFG_LAUNCHED(c, "k_poly_add", st);
    LimbSel s0, s1;
    s0.n = nl;
    s1.n = nQ;
    for (int i = 0; i < nQ; i++) s1.idx[i] = i;
    for (int i = 0; i < nl; i++) s0.idx[i] = i;
    if (launch_ntt(c, t0.as<uint64_t>(), (size_t)nl * N, t0.as<uint64_t>(), (size_t)nl * N, nct * nl, s0, true, st)) return -1;
    // Recode
    const dim3 g((N + 127) / 128, nct);
    if (tab.nw <= 8) k_recode<8><<<g, 128, 0, st>>>(t0.as<uint64_t>(), dg.as<uint64_t>(), tab, nQ, N, c->lc, t1.as<uint64_t>());
    else if (tab.nw <= 16) k_recode<16><<<g, 128, 0, st>>>(t0.as<uint64_t>(), dg.as<uint64_t>(), tab, nQ, N, c->lc, t1.as<uint64_t>());
    else k_recode<kBigWords><<<g, 128, 0, st>>>(t0.as<uint64_t>(), dg.as<uint64_t>(), tab, nQ, N, c->lc, t1.as<uint64_t>());
    SFG_LAUNCHED(c, "k_recode", st);
    if (launch_ntt(c, t1.as<uint64_t>(), (size_t)nQ * N, t1.as<uint64_t>(), (size_t)nQ * N, nct * nQ, s1, false, st)) return -1;
    // Recrypt: c0' = recoded + agg(h1) -> out[ct][0]; c1' = a -> out[ct][1]
    for (int t = 0; t < nct; t++) {
        uint64_t *o = d_out + (size_t)t * 2 * nQ * N;
        k_poly_add<<<dim3((N + 255) / 256, nQ, 1), 256, 0, st>>>(t1.as<uint64_t>() + (size_t)t * nQ * N, nQ, d_agg1 + (size_t)t * nQ * N, nQ, N, c->lc, o);
        SFG_CUDA(c, cudaMemcpyAsync(o + (size_t)nQ * N, d_crp + (size_t)t * nQ * N, (size_t)nQ * N * 8, cudaMemcpyDefault, st));
    }
    SFG_LAUNCHED(c, "k_poly_add", st);
    SFG_CUDA(c, cudaStreamSynchronize(st));
    return 0;
}

}  // namespace sfg
