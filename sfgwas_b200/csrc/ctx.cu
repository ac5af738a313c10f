// ctx.cu -- context construction: RNS constants, NTT twiddles, encoder tables, key-switch base-conversion tables.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "kernels.h"

namespace sfg {

int launch_check(Ctx *c, const char *what, cudaStream_t st) {
    static const bool debug = getenv("SFG_DEBUG") != nullptr;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && debug) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "kernel %s: %s", what, cudaGetErrorString(e));
        c->err = buf;
        return -1;
    }
    return 0;
}

bool poison_enabled() {
    static const bool on = [] { const char *e = getenv("SFG_POISON"); return e && *e && *e != '0'; }();
    return on;
}
void poison_fill(Ctx *c, void *p, size_t bytes) {
    if (poison_enabled() && p && bytes) cudaMemsetAsync(p, 0xA5, bytes, c->stream);
}
int upload(Ctx *c, void *dst, const void *src, size_t bytes) {
    if (bytes == 0) return 0;
    SFG_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int dev_alloc(Ctx *c, void **p, size_t bytes, const char *what) {
    *p = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *p = nullptr;
        SFG_FAIL(c, "cudaMalloc of %zu bytes (%s) failed: %s", bytes, what, cudaGetErrorString(e));
    }
    poison_fill(c, *p, bytes);
    return 0;
}

int ws_get(Ctx *c, int slot, size_t bytes, void **out) {
    Ctx::WsBuf &b = c->ws[slot];
    if (bytes == 0) bytes = 16;
    if (b.bytes < bytes) {
        if (b.p) {
            SFG_CUDA(c, cudaStreamSynchronize(c->stream));
            cudaFree(b.p);
            b.p = nullptr;
            b.bytes = 0;
        }
        cudaError_t e = cudaMalloc(&b.p, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            b.p = nullptr;
            SFG_FAIL(c, "workspace allocation of %zu bytes (slot %d) failed: %s", bytes, slot, cudaGetErrorString(e));
        }
        b.bytes = bytes;
    }
    *out = b.p;
    poison_fill(c, b.p, bytes);  // contents are undefined by contract: under SFG_POISON=1 every hand-out is re-poisoned (stream-ordered)
    return 0;
}
void ws_release(Ctx *c) {
    for (auto &b : c->ws) {
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.bytes = 0;
    }
}

int ctx_build_tables(Ctx *c, const uint64_t *psi_opt) {
    const int N = c->N, logN = c->logN, nQP = c->nQP;
    SFG_CUDA(c, cudaSetDevice(c->device));
    c->lc_h.resize(nQP);
    c->psi.resize(nQP);
    std::vector<uint64_t> tw((size_t)nQP * 4 * N);
    for (int i = 0; i < nQP; i++) {
        const uint64_t q = c->mod[i];
        if ((q & 1) == 0 || q >= (1ULL << 62) || (q - 1) % (2ULL * N) != 0) SFG_FAIL(c, "modulus %d (%llu) is not an NTT-friendly prime < 2^62", i, (unsigned long long)q);
        c->lc_h[i] = h_limb_const(q, N);
        // Lattigo ring.genNTTParams: psi = g^((q-1)/2N), psiInv = g^(q-1-(q-1)/2N)   
        uint64_t psi = psi_opt ? psi_opt[i] : h_powmod(h_primitive_root(q), (q - 1) / (2ULL * N), q);
        if (h_powmod(psi, N, q) != q - 1) SFG_FAIL(c, "psi for modulus %d is not a primitive 2N-th root of unity", i);
        c->psi[i] = psi;
        const uint64_t psiInv = h_invmod(psi, q);
        uint64_t *w = &tw[(size_t)i * 4 * N], *wsh = w + N, *wi = w + 2 * N, *wish = w + 3 * N;
        uint64_t p = 1, pi = 1;
        for (int j = 0; j < N; j++) {  // NttPsi[brv(j)] = psi^j, NttPsiInv[brv(j)] = psi^-j
            const uint64_t r = h_bitrev((uint64_t)j, logN);
            w[r] = p;
            wsh[r] = h_shoup(p, q);
            wi[r] = pi;
            wish[r] = h_shoup(pi, q);
            p = h_mulmod(p, psi, q);
            pi = h_mulmod(pi, psiInv, q);
        }
    }
    if (dev_alloc(c, (void **)&c->lc, sizeof(LimbConst) * nQP, "table") || upload(c, c->lc, c->lc_h.data(), sizeof(LimbConst) * nQP)) return -1;
    if (dev_alloc(c, (void **)&c->tw, sizeof(uint64_t) * tw.size(), "table") || upload(c, c->tw, tw.data(), sizeof(uint64_t) * tw.size())) return -1;
    // per-class tables of the register-tiled transforms (ntt2.cuh): [N] in Lattigo's order + the last-pass table in TT order
    if (logN <= 16) {  // logN 15 / 16: the four-step transforms of kernels_ntt.cu use the same tables
        std::vector<TwTab> tabs(nQP);
        const int P = N >> kLastR, s0 = logN - kLastR, NL = kLastE - 1;  // last-pass table: NL twiddles per group of 16
        for (int i = 0; i < nQP; i++) {
            const uint64_t q = c->mod[i];
            const uint64_t *w = &tw[(size_t)i * 4 * N], *wi = w + 2 * N;
            const int kind = arith_kind(q);
            const size_t es = kind == kArW ? sizeof(ulonglong2) : 8;
            std::vector<unsigned char> h((size_t)(2 * N + 2 * NL * P) * es);
            auto put = [&](size_t pos, uint64_t x) {
                if (kind == kArW) reinterpret_cast<ulonglong2 *>(h.data())[pos] = ArW::make_tw(x, q);
                else if (kind == kArD) reinterpret_cast<double *>(h.data())[pos] = ArD::make_tw(x, q);
                else reinterpret_cast<uint2 *>(h.data())[pos] = ArN30::make_tw(x, q);
            };
            for (int j = 0; j < N; j++) {
                put(j, w[j]);
                put((size_t)N + NL * P + j, wi[j]);
            }
            for (int r = 0; r < kLastR; r++)
                for (int g = 0; g < (1 << r); g++)
                    for (int p = 0; p < P; p++) {
                        const size_t idx = ((size_t)1 << (s0 + r)) + ((size_t)p << r) + g;
                        put((size_t)N + (size_t)((1 << r) - 1 + g) * P + p, w[idx]);
                        put((size_t)2 * N + NL * P + (size_t)((1 << r) - 1 + g) * P + p, wi[idx]);
                    }
            unsigned char *d = nullptr;
            if (dev_alloc(c, (void **)&d, h.size(), "table") || upload(c, d, h.data(), h.size())) return -1;
            c->tw2_bufs.push_back(d);
            tabs[i] = TwTab{d, d + (size_t)N * es, d + (size_t)(N + NL * P) * es, d + (size_t)(2 * N + NL * P) * es};
        }
        if (dev_alloc(c, (void **)&c->tw2, sizeof(TwTab) * nQP, "table") || upload(c, c->tw2, tabs.data(), sizeof(TwTab) * nQP)) return -1;
    }

    // encoder tables (Lattigo ckks encoder: m = 2N, rotGroup[j] = 5^j mod m, roots[k] = exp(2 pi i k / m); App. B.6)
    const int M = 2 * N;
    std::vector<double> roots(2 * (size_t)(M + 1)), dd(2 * (size_t)M);
    h_trig_tables(M, roots.data(), dd.data());
    std::vector<int> rot5(c->slots);
    uint64_t five = 1;
    for (int j = 0; j < c->slots; j++) {
        rot5[j] = (int)five;
        five = five * 5 % (uint64_t)M;
    }
    if (dev_alloc(c, (void **)&c->roots, sizeof(double) * roots.size(), "table") || upload(c, c->roots, roots.data(), sizeof(double) * roots.size())) return -1;
    if (dev_alloc(c, (void **)&c->ddcos, sizeof(double) * dd.size(), "table") || upload(c, c->ddcos, dd.data(), sizeof(double) * dd.size())) return -1;
    if (dev_alloc(c, (void **)&c->rot5, sizeof(int) * rot5.size(), "table") || upload(c, c->rot5, rot5.data(), sizeof(int) * rot5.size())) return -1;
    {
        // per-stage twiddles of the special inverse FFT in the order the butterflies read them (unit stride in j): the stage of length
        // len = 2h uses roots[(4 len - (5^j mod 4 len)) * (M / (4 len))] for j < h (Lattigo invfft); stored at [h + j]
        const int n = c->slots;
        std::vector<double> ft(2 * (size_t)n, 0.0);
        for (int len = 2; len <= n; len <<= 1) {
            const int lenh = len >> 1, lenq = len << 2, gap = M / lenq;
            for (int j = 0; j < lenh; j++) {
                const int idx = (lenq - (rot5[j] & (lenq - 1))) * gap;
                ft[2 * (size_t)(lenh + j)] = roots[2 * (size_t)idx];
                ft[2 * (size_t)(lenh + j) + 1] = roots[2 * (size_t)idx + 1];
            }
        }
        if (dev_alloc(c, (void **)&c->fft_tw, sizeof(double) * ft.size(), "table") || upload(c, c->fft_tw, ft.data(), sizeof(double) * ft.size())) return -1;
    }
    {
        const uint64_t twoN = 2 * (uint64_t)N, n2 = (uint64_t)N / 2;
        std::vector<uint32_t> of_exp(twoN, 0), pos(N), src(N);
        uint64_t p5 = 1;
        for (uint64_t t = 0; t < n2; t++) {
            of_exp[p5] = (uint32_t)t;
            of_exp[twoN - p5] = (uint32_t)(n2 + t);
            p5 = p5 * 5 % twoN;
        }
        for (int i = 0; i < N; i++) {
            pos[i] = of_exp[2 * h_bitrev((uint64_t)i, c->logN) + 1];
            src[pos[i]] = (uint32_t)i;
        }
        if (dev_alloc(c, (void **)&c->dlog_pos, sizeof(uint32_t) * N, "table") || upload(c, c->dlog_pos, pos.data(), sizeof(uint32_t) * N)) return -1;
        if (dev_alloc(c, (void **)&c->dlog_src, sizeof(uint32_t) * N, "table") || upload(c, c->dlog_src, src.data(), sizeof(uint32_t) * N)) return -1;
    }
    if (dev_alloc(c, (void **)&c->enc_stats, sizeof(unsigned long long) * 2, "encoder statistics")) return -1;
    SFG_CUDA(c, cudaMemsetAsync(c->enc_stats, 0, sizeof(unsigned long long) * 2, c->stream));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    // FP64 special-FFT error model for slot values |v| <= 4: scale * eps * log2(n) * 4 / sqrt(n); the window is 64x that
    // (DESIGN.md "encoder exactness"; validated against the quad-precision oracle in tests/test_encode.py).
    const double n = (double)c->slots;
    const double err = c->scale * 1.1102230246251565e-16 * std::log2(n) * 4.0 / std::sqrt(n);
    c->enc_delta = std::min(0.2, std::max(1e-9, 64.0 * err));
    return 0;
}

PolyLayout make_layout(const Ctx *c, int nl, bool packed) {
    PolyLayout lay{};
    lay.nl = nl;
    long long off = 0;
    // wide limbs first keeps every limb 16-byte aligned regardless of N
    for (int l = 0; l < nl; l++) {
        lay.es[l] = (packed && c->mod[l] < (1ULL << 32)) ? 4 : 8;
        lay.off[l] = off;
        off += (long long)c->N * lay.es[l];
    }
    lay.bytes = off;
    return lay;
}

static void fill_bc(const Ctx *c, const int *src, int ns, int tgt, BaseConv &b) {
    memset(&b, 0, sizeof b);
    b.ns = ns;
    const uint64_t t = c->mod[tgt];
    uint64_t smod = 1;
    for (int k = 0; k < ns; k++) {
        const uint64_t sk = c->mod[src[k]];
        uint64_t prod = 1, prodT = 1;
        for (int j = 0; j < ns; j++)
            if (j != k) {
                prod = h_mulmod(prod, c->mod[src[j]] % sk, sk);
                prodT = h_mulmod(prodT, c->mod[src[j]] % t, t);
            }
        b.src_limb[k] = src[k];
        b.sinv[k] = h_invmod(prod, sk);
        b.sinv_sh[k] = h_shoup(b.sinv[k], sk);
        b.fac[k] = prodT;
        b.fac_sh[k] = h_shoup(prodT, t);
        b.sf[k] = (double)sk;
        smod = h_mulmod(smod, sk % t, t);
    }
    b.smod = smod;
    b.smod_sh = h_shoup(smod, t);
}

// Tables for a key-switch at `level`: digit i covers Q limbs [i*alpha, min((i+1)*alpha, level+1)).
//   ks: [beta_l][level+1+nP]  (target tt < level+1 -> Q limb tt ; else P limb tt-(level+1)); ns = 0 when the target
//       lies inside the digit (the NTT-domain input limb is reused, Lattigo decomposeAndSplitNTT).
//   md: [level+1]  P -> q_l ;  pinv: [level+1][2]  P^-1 mod q_l with Shoup companion.
int ctx_get_ks_tables(Ctx *c, int level, BaseConv **ks, BaseConv **md, uint64_t **pinv) {
    std::lock_guard<std::mutex> g(c->mu);
    if (level < 0 || level >= c->nQ) SFG_FAIL(c, "key-switch level %d out of range", level);
    if (!c->bc_ks.count(level)) {
        const int nl = level + 1, alpha = c->nP, beta = (nl + alpha - 1) / alpha, nt = nl + c->nP;
        if (alpha > kMaxAlpha) SFG_FAIL(c, "more than %d special primes are not supported", kMaxAlpha);
        std::vector<BaseConv> h((size_t)beta * nt);
        for (int i = 0; i < beta; i++) {
            int src[kMaxAlpha], ns = 0;
            for (int k = i * alpha; k < std::min((i + 1) * alpha, nl); k++) src[ns++] = k;
            for (int tt = 0; tt < nt; tt++) {
                const int tgt = tt < nl ? tt : c->nQ + (tt - nl);
                BaseConv &b = h[(size_t)i * nt + tt];
                if (tt >= i * alpha && tt < i * alpha + ns) {
                    memset(&b, 0, sizeof b);
                } else {
                    fill_bc(c, src, ns, tgt, b);
                }
            }
        }
        std::vector<BaseConv> hm(nl);
        std::vector<uint64_t> hp((size_t)nl * 2);
        int psrc[kMaxAlpha];
        for (int p = 0; p < c->nP; p++) psrc[p] = c->nQ + p;
        for (int l = 0; l < nl; l++) {
            fill_bc(c, psrc, c->nP, l, hm[l]);
            const uint64_t q = c->mod[l];
            uint64_t Pm = 1;
            for (int p = 0; p < c->nP; p++) Pm = h_mulmod(Pm, c->mod[c->nQ + p] % q, q);
            hp[2 * l] = h_invmod(Pm, q);
            hp[2 * l + 1] = h_shoup(hp[2 * l], q);
        }
        BaseConv *dks = nullptr, *dmd = nullptr;
        uint64_t *dp = nullptr;
        if (dev_alloc(c, (void **)&dks, sizeof(BaseConv) * h.size(), "table") || upload(c, dks, h.data(), sizeof(BaseConv) * h.size())) return -1;
        if (dev_alloc(c, (void **)&dmd, sizeof(BaseConv) * hm.size(), "table") || upload(c, dmd, hm.data(), sizeof(BaseConv) * hm.size())) return -1;
        if (dev_alloc(c, (void **)&dp, sizeof(uint64_t) * hp.size(), "table") || upload(c, dp, hp.data(), sizeof(uint64_t) * hp.size())) return -1;
        c->bc_ks[level] = dks;
        c->bc_md[level] = dmd;
        c->pinv[level] = dp;
    }
    *ks = c->bc_ks[level];
    *md = c->bc_md[level];
    *pinv = c->pinv[level];
    return 0;
}

}  // namespace sfg
