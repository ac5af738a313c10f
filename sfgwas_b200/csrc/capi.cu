// capi.cu -- extern "C" boundary of libsfgwas_b200.so (include/sfgwas_b200.h).
#include <cmath>
#include <cstring>

#include "../../include/sfgwas_b200.h"
#include "matmult.h"

using namespace sfg;

struct sfg_ctx {
    Ctx c;
};
struct sfg_geno {
    Geno *g;
};
struct sfg_cache {
    Cache *ca;
};
struct sfg_cts {  // n ciphertexts [n][2][nl][N] resident in HBM
    int device;
    int n, nl;
    uint64_t *d;
};

static std::string g_create_err;

namespace sfg {  // cmfile.cpp
int cm_save(const char *filename, int logN, const uint64_t *cts, const double *scales, int nrows, int ncols, int nl, std::string &err);
int cm_info(const char *filename, int *nrows, int *ncols, int *nl, int *logN, std::string &err);
int cm_load(const char *filename, int logN, uint64_t *cts, double *scales, int nrows, int ncols, int nl, std::string &err);
}  // namespace sfg

extern "C" {

int sfg_version(void) { return 1; }

int sfg_ctx_create(int device, int logN, const uint64_t *qi, int nQ, const uint64_t *pi, int nP, double scale, const uint64_t *psi,
                   sfg_ctx **out) {
    *out = nullptr;
    sfg_ctx *h = new sfg_ctx();
    Ctx *c = &h->c;
    auto fail = [&](const std::string &m) {
        g_create_err = m;
        delete h;
        return -1;
    };
    if (logN < 6 || logN > 17) return fail("logN out of range [6, 17]");
    if (nQ < 1 || nP < 1 || nQ + nP > kMaxLimbs) return fail("unsupported number of moduli");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return fail(std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libsfgwas_b200 has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail("device index out of range");
    c->device = device;
    c->logN = logN;
    c->N = 1 << logN;
    c->slots = c->N >> 1;
    c->d = (int)std::ceil(std::sqrt((double)c->slots));  // gwas/matmult.go:918,1047
    c->nQ = nQ;
    c->nP = nP;
    c->nQP = nQ + nP;
    c->beta = (nQ + nP - 1) / nP;
    c->scale = scale;
    c->mod.assign(qi, qi + nQ);
    c->mod.insert(c->mod.end(), pi, pi + nP);
    if (cudaSetDevice(device) != cudaSuccess) return fail("cudaSetDevice failed");
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) return fail("cudaStreamCreate failed");
    if (ctx_build_tables(c, psi)) {
        std::string m = c->err;
        cudaStreamDestroy(c->stream);
        c->stream = nullptr;
        return fail(m);
    }
    *out = h;
    return 0;
}

void sfg_ctx_destroy(sfg_ctx *h) {
    if (!h) return;
    Ctx *c = &h->c;
    cudaSetDevice(c->device);
    cudaFree(c->lc);
    cudaFree(c->tw);
    cudaFree(c->tw2);
    ws_release(c);
    pinned_release(c);
    for (void *p : c->tw2_bufs) cudaFree(p);
    cudaFree(c->roots);
    cudaFree(c->ddcos);
    cudaFree(c->rot5);
    cudaFree(c->fft_tw);
    cudaFree(c->dlog_pos);
    cudaFree(c->dlog_src);
    cudaFree(c->enc_stats);
    for (auto &kv : c->bc_ks) cudaFree(kv.second);
    for (auto &kv : c->bc_md) cudaFree(kv.second);
    for (auto &kv : c->pinv) cudaFree(kv.second);
    for (auto &kv : c->keys) {
        cudaFree(kv.second.key);
        cudaFree(kv.second.perm);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete h;
}

const char *sfg_last_error(const sfg_ctx *h) { return h ? h->c.err.c_str() : g_create_err.c_str(); }

int sfg_ctx_set_cache_budget(sfg_ctx *h, size_t bytes) {
    h->c.cache_budget = bytes;
    return 0;
}
unsigned long long sfg_ctx_launch_count(const sfg_ctx *h) { return h->c.launches; }
int sfg_ctx_encoder_stats(sfg_ctx *h, unsigned long long out[2]) {
    Ctx *c = &h->c;
    SFG_CUDA(c, cudaSetDevice(c->device));
    SFG_CUDA(c, cudaMemcpy(out, c->enc_stats, 2 * sizeof(unsigned long long), cudaMemcpyDefault));
    return 0;
}
int sfg_ctx_psi(const sfg_ctx *h, uint64_t *psi_out) {
    for (int i = 0; i < h->c.nQP; i++) psi_out[i] = h->c.psi[i];
    return 0;
}
int sfg_ctx_sync(sfg_ctx *h) {
    Ctx *c = &h->c;
    SFG_CUDA(c, cudaSetDevice(c->device));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
void *sfg_ctx_stream(const sfg_ctx *h) { return (void *)h->c.stream; }
int sfg_ctx_last_timings(const sfg_ctx *, float out_ms[5]) {
    for (int i = 0; i < 5; i++) out_ms[i] = g_last_ms[i];
    return 0;
}

static int set_key_dev(Ctx *c, int rot_left, uint64_t *draw) {
    // convert to the device format of the key-switch kernels (TT order, Shoup pairs for narrow moduli)
    const size_t n = (size_t)c->beta * 2 * c->nQP * c->N;
    uint64_t *dkey = nullptr;
    if (dev_alloc(c, (void **)&dkey, n * 8, "Galois key")) {
        cudaFree(draw);
        return -1;
    }
    if (launch_key_convert(c, draw, dkey, c->stream)) return -1;
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(draw);
    const uint64_t galEl = h_galois_element(c->logN, rot_left);
    std::vector<uint32_t> idx(c->N);
    h_permute_ntt_index(c->logN, galEl, idx.data());
    uint32_t *dperm = nullptr;
    if (dev_alloc(c, (void **)&dperm, sizeof(uint32_t) * c->N, "permutation table") || upload(c, dperm, idx.data(), sizeof(uint32_t) * c->N)) return -1;
    std::lock_guard<std::mutex> g(c->mu);
    auto it = c->keys.find(galEl);
    if (it != c->keys.end()) {
        cudaFree(it->second.key);
        cudaFree(it->second.perm);
    }
    c->keys[galEl] = GaloisKey{galEl, dkey, dperm};
    return 0;
}

int sfg_ctx_set_rotation_key(sfg_ctx *h, int rot_left, const uint64_t *key) {
    Ctx *c = &h->c;
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t n = (size_t)c->beta * 2 * c->nQP * c->N;
    uint64_t *d = nullptr;
    if (dev_alloc(c, (void **)&d, n * 8, "Galois key staging")) return -1;
    if (upload(c, d, key, n * 8)) {  // ordered on the context stream: k_key_convert runs there (see ctx.h: upload)
        cudaFree(d);
        return -1;
    }
    return set_key_dev(c, rot_left, d);
}
int sfg_ctx_set_rotation_key_ptrs(sfg_ctx *h, int rot_left, const uint64_t *const *limbs) {
    Ctx *c = &h->c;
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t np = (size_t)c->beta * 2 * c->nQP;
    uint64_t *d = nullptr;
    if (dev_alloc(c, (void **)&d, np * c->N * 8, "Galois key staging")) return -1;
    // one Go slice per limb (cgo): gathered through the pinned bounce buffer, ordered on the context stream
    if (gather_limbs_to_device(c, limbs, np, d)) {
        cudaFree(d);
        return -1;
    }
    return set_key_dev(c, rot_left, d);
}
int sfg_ctx_has_rotation_key(const sfg_ctx *h, int rot_left) {
    return h->c.keys.count(h_galois_element(h->c.logN, rot_left)) ? 1 : 0;
}

// ---- primitives ----
int sfg_ntt(sfg_ctx *h, uint64_t *polys, int npoly, const int *limb_idx, int nsel, int inverse) {
    Ctx *c = &h->c;
    if (nsel < 1 || nsel > kMaxLimbs || npoly % nsel) SFG_FAIL(c, "sfg_ntt: npoly must be a multiple of nsel (<= %d)", kMaxLimbs);
    SFG_CUDA(c, cudaSetDevice(c->device));
    LimbSel sel;
    sel.n = nsel;
    for (int i = 0; i < nsel; i++) {
        if (limb_idx[i] < 0 || limb_idx[i] >= c->nQP) SFG_FAIL(c, "sfg_ntt: modulus index %d out of range", limb_idx[i]);
        sel.idx[i] = limb_idx[i];
    }
    Buf d;
    const size_t bytes = (size_t)npoly * c->N * 8;
    if (d.alloc(c, bytes)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(d.p, polys, bytes, cudaMemcpyDefault, c->stream));
    if (launch_ntt(c, d.as<uint64_t>(), (size_t)nsel * c->N, d.as<uint64_t>(), (size_t)nsel * c->N, npoly, sel, inverse != 0, c->stream)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(polys, d.p, bytes, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int sfg_mul_coeffs_and_add128(sfg_ctx *h, const uint64_t *a, const uint64_t *b, uint64_t *acc, size_t n) {
    Ctx *c = &h->c;
    SFG_CUDA(c, cudaSetDevice(c->device));
    Buf da, db, dacc;
    if (da.alloc(c, n * 8) || db.alloc(c, n * 8) || dacc.alloc(c, n * 16)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(da.p, a, n * 8, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaMemcpyAsync(db.p, b, n * 8, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaMemcpyAsync(dacc.p, acc, n * 16, cudaMemcpyDefault, c->stream));
    if (launch_mul_coeffs_and_add128(c, da.as<uint64_t>(), db.as<uint64_t>(), dacc.as<uint64_t>(), n, c->stream)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(acc, dacc.p, n * 16, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int sfg_reduce_and_add_uint128(sfg_ctx *h, const uint64_t *acc, uint64_t *out, int limb, size_t n) {
    Ctx *c = &h->c;
    if (limb < 0 || limb >= c->nQP) SFG_FAIL(c, "modulus index out of range");
    SFG_CUDA(c, cudaSetDevice(c->device));
    Buf dacc, dout;
    if (dacc.alloc(c, n * 16) || dout.alloc(c, n * 8)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(dacc.p, acc, n * 16, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaMemcpyAsync(dout.p, out, n * 8, cudaMemcpyDefault, c->stream));
    if (launch_reduce_and_add128(c, dacc.as<uint64_t>(), dout.as<uint64_t>(), limb, n, c->stream)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(out, dout.p, n * 8, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int sfg_mform_lvl(sfg_ctx *h, int level, uint64_t *p) {
    Ctx *c = &h->c;
    if (level < 0 || level >= c->nQ) SFG_FAIL(c, "level out of range");
    SFG_CUDA(c, cudaSetDevice(c->device));
    Buf d;
    const size_t bytes = (size_t)(level + 1) * c->N * 8;
    if (d.alloc(c, bytes)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(d.p, p, bytes, cudaMemcpyDefault, c->stream));
    if (launch_mform(c, d.as<uint64_t>(), level + 1, c->stream)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(p, d.p, bytes, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int sfg_rotate_right(sfg_ctx *h, int level, const uint64_t *cts, int nct, int nrot, uint64_t *out) {
    Ctx *c = &h->c;
    if (level < 0 || level >= c->nQ || nct < 1) SFG_FAIL(c, "sfg_rotate_right: bad level / count");
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t bytes = (size_t)nct * 2 * (level + 1) * c->N * 8;
    Buf din, dout;
    if (din.alloc(c, bytes) || dout.alloc(c, bytes)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(din.p, cts, bytes, cudaMemcpyDefault, c->stream));
    if (rotate_right_dev(c, level, din.as<uint64_t>(), nct, nrot, dout.as<uint64_t>())) return -1;
    SFG_CUDA(c, cudaMemcpy(out, dout.p, bytes, cudaMemcpyDefault));
    return 0;
}

// ---- genotype ----
int sfg_geno_create(sfg_ctx *h, size_t nrows, size_t ncols, sfg_geno **out) {
    *out = nullptr;
    Geno *g = nullptr;
    if (geno_create(&h->c, nrows, ncols, &g)) return -1;
    *out = new sfg_geno{g};
    return 0;
}
int sfg_geno_push_rows(sfg_geno *g, const int8_t *rows, size_t n) { return geno_push(g->g, rows, n); }
void sfg_geno_destroy(sfg_geno *g) {
    if (!g) return;
    geno_release(g->g);
    delete g;
}
int sfg_encode_diag(sfg_ctx *h, const sfg_geno *g, int block_row, int shift, int nrot, int level, int mont, uint64_t *out,
                    uint8_t *present, int64_t *coeffs_out) {
    return encode_diag_host(&h->c, g->g, block_row, shift, nrot, level, mont != 0, out, present, coeffs_out);
}

// ---- stream entry points ----
int sfg_matmult4_stream_preprocess(sfg_ctx *h, const sfg_geno *g, int max_level, sfg_cache **out) {
    *out = nullptr;
    Cache *ca = nullptr;
    if (cache_build(&h->c, g->g, max_level, &ca)) return -1;
    *out = new sfg_cache{ca};
    return 0;
}
int sfg_matmult4_stream_preprocess_rows(sfg_ctx *h, const sfg_geno *g, int max_level, int bi_lo, int bi_hi, sfg_cache **out) {
    *out = nullptr;
    Cache *ca = nullptr;
    if (cache_build(&h->c, g->g, max_level, &ca, bi_lo, bi_hi)) return -1;
    *out = new sfg_cache{ca};
    return 0;
}
int sfg_matmult4_stream_preprocess_giants(sfg_ctx *h, const sfg_geno *g, int max_level, int part, int nparts, sfg_cache **out) {
    *out = nullptr;
    Cache *ca = nullptr;
    if (cache_build(&h->c, g->g, max_level, &ca, 0, -1, part, nparts)) return -1;
    *out = new sfg_cache{ca};
    return 0;
}
void sfg_cache_destroy(sfg_cache *cache) {
    if (!cache) return;
    cache_destroy(cache->ca);
    delete cache;
}
int sfg_cache_info(const sfg_cache *cache, size_t *num_polys, size_t *bytes, int *materialised, int *m_ct, int *nbr) {
    const Cache *ca = cache->ca;
    if (num_polys) *num_polys = ca->npoly;
    if (bytes) *bytes = ca->materialised ? ca->img_bytes : 0;
    if (materialised) *materialised = ca->materialised ? 1 : 0;
    if (m_ct) *m_ct = ca->m_ct;
    if (nbr) *nbr = ca->nbr;
    return 0;
}
int sfg_cache_get_diag(sfg_ctx *h, const sfg_cache *cache, int bi, int shift, int bj, uint64_t *out, int *present) {
    Ctx *c = &h->c;
    const Cache *ca = cache->ca;
    if (bi < 0 || bi >= ca->nbr || shift < 0 || shift >= ca->slots || bj < 0 || bj >= ca->m_ct) SFG_FAIL(c, "cache_get_diag: index out of range");
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t N = c->N;
    Buf tmp;
    if (tmp.alloc(c, (size_t)ca->L * N * 8)) return -1;
    if (cache_get_diag_dev(c, ca, bi, shift, bj, tmp.as<uint64_t>(), present)) return -1;
    if (!*present) return 0;
    SFG_CUDA(c, cudaMemcpyAsync(out, tmp.p, (size_t)ca->L * N * 8, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    // the image holds plain residues; present them in the reference's cache form b*2^64 mod q (gwas/matmult.go:401-440)
    for (int l = 0; l < ca->L; l++) {
        const uint64_t q = c->mod[l], r64 = c->lc_h[l].r64;
        for (size_t k = 0; k < N; k++) out[l * N + k] = h_mulmod(out[l * N + k], r64, q);
    }
    return 0;
}

int sfg_matmult4_stream_compute_dev(sfg_ctx *h, const uint64_t *d_A, int s, int nbr, int level_a, int max_level,
                                    const sfg_cache *cache, uint64_t *d_out) {
    return mm_compute_dev(&h->c, d_A, s, nbr, level_a, max_level, cache->ca, d_out);
}

int sfg_matmult4_stream_compute(sfg_ctx *h, const uint64_t *A, int s, int nbr, int level_a, int max_level, const sfg_cache *cache,
                                uint64_t *out) {
    Ctx *c = &h->c;
    if (s < 1 || nbr < 1 || level_a < 0) SFG_FAIL(c, "bad A dimensions");
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t abytes = (size_t)s * nbr * 2 * (level_a + 1) * c->N * 8;
    const size_t obytes = (size_t)s * cache->ca->m_ct * 2 * max_level * c->N * 8;
    void *dA, *dO;  // grow-only workspace: no cudaMalloc / cudaFree in the steady state
    if (ws_get(c, WS_A, abytes, &dA) || ws_get(c, WS_OUT, obytes, &dO)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(dA, A, abytes, cudaMemcpyDefault, c->stream));
    // device -> host copies of the output rows overlap the giant-step rotations of the following rows (matmult.cu: HostSink)
    return mm_compute_dev(c, (const uint64_t *)dA, s, nbr, level_a, max_level, cache->ca, (uint64_t *)dO, out);
}

int sfg_matmult4_stream_compute_ptrs(sfg_ctx *h, const uint64_t *const *A_limbs, int s, int nbr, int level_a, int max_level,
                                     const sfg_cache *cache, uint64_t *const *out_limbs) {
    Ctx *c = &h->c;
    if (s < 1 || nbr < 1 || level_a < 0) SFG_FAIL(c, "bad A dimensions");
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t N = c->N;
    const size_t na = (size_t)s * nbr * 2 * (level_a + 1), no = (size_t)s * cache->ca->m_ct * 2 * max_level;
    void *dA, *dO;  // grow-only workspace, like the flat entry point
    if (ws_get(c, WS_A, na * N * 8, &dA) || ws_get(c, WS_OUT, no * N * 8, &dO)) return -1;
    // one pageable Go slice per limb: gathered into pinned staging by a few host threads, ONE transfer each way; the result rows are
    // scattered back to the limb pointers while the giant-step sums of the following rows run (matmult.cu: HostSink)
    if (gather_limbs_to_device(c, A_limbs, na, (uint64_t *)dA)) return -1;
    return mm_compute_dev(c, (const uint64_t *)dA, s, nbr, level_a, max_level, cache->ca, (uint64_t *)dO, nullptr, out_limbs);
}

int sfg_matmult4_stream(sfg_ctx *h, const uint64_t *A, int s, int level_a, const sfg_geno *g, int max_level, int compute_squared_sum,
                        int square, uint64_t *out, double *sum, double *sq_sum) {
    Ctx *c = &h->c;
    const Geno *src = g->g;
    if (src->filled != src->nrows) SFG_FAIL(c, "genotype matrix incomplete");
    if (compute_squared_sum && (!sum || !sq_sum)) SFG_FAIL(c, "sum / sq_sum buffers required when compute_squared_sum is set");
    SFG_CUDA(c, cudaSetDevice(c->device));
    // working copy: missing -> 0, (sum, sqSum), optional squaring  (gwas/matmult.go:1289-1304)
    Geno *w = nullptr;
    if (geno_create(c, src->nrows, src->ncols, &w)) return -1;
    w->filled = src->nrows;
    int rc = -1;
    Cache *ca = nullptr;
    Buf dsum, dsq;
    do {
        if (cudaMemcpyAsync(w->d, src->d, src->nrows * src->ncols, cudaMemcpyDeviceToDevice, c->stream) != cudaSuccess) { c->err = "D2D copy failed"; break; }
        if (compute_squared_sum) {
            if (dsum.alloc(c, src->ncols * 8) || dsq.alloc(c, src->ncols * 8)) break;
            cudaMemsetAsync(dsum.p, 0, src->ncols * 8, c->stream);
            cudaMemsetAsync(dsq.p, 0, src->ncols * 8, c->stream);
        }
        if (launch_geno_prep(c, w->d, src->nrows, src->ncols, compute_squared_sum ? dsum.as<double>() : nullptr,
                             compute_squared_sum ? dsq.as<double>() : nullptr, square != 0, c->stream)) break;
        if (cache_build(c, w, max_level, &ca)) break;
        const int nbr = ca->nbr;
        const size_t abytes = (size_t)s * nbr * 2 * (level_a + 1) * c->N * 8;
        const size_t obytes = (size_t)s * ca->m_ct * 2 * max_level * c->N * 8;
        Buf dA, dO;
        if (dA.alloc(c, abytes) || dO.alloc(c, obytes)) break;
        if (cudaMemcpyAsync(dA.p, A, abytes, cudaMemcpyDefault, c->stream) != cudaSuccess) { c->err = "H2D copy of A failed"; break; }
        if (mm_compute_dev(c, dA.as<uint64_t>(), s, nbr, level_a, max_level, ca, dO.as<uint64_t>())) break;
        if (cudaMemcpyAsync(out, dO.p, obytes, cudaMemcpyDefault, c->stream) != cudaSuccess) { c->err = "D2H copy failed"; break; }
        if (compute_squared_sum) {
            cudaMemcpyAsync(sum, dsum.p, src->ncols * 8, cudaMemcpyDefault, c->stream);
            cudaMemcpyAsync(sq_sum, dsq.p, src->ncols * 8, cudaMemcpyDefault, c->stream);
        }
        if (cudaStreamSynchronize(c->stream) != cudaSuccess) { c->err = "stream sync failed"; break; }
        rc = 0;
    } while (0);
    if (ca) cache_destroy(ca);
    geno_release(w);
    return rc;
}

// ---- multi-GPU pieces ----
size_t sfg_cv_elems(const sfg_ctx *h, const sfg_cache *cache, int s, int max_level) {
    const Cache *ca = cache->ca;
    (void)max_level;
    return ca->gact.size() * (size_t)ca->m_ct * 2 * s * ca->L * h->c.N;
}
int sfg_matmult4_partial(sfg_ctx *h, const uint64_t *A, int s, int nbr, int level_a, int max_level, const sfg_cache *cache, int bi_lo,
                         int bi_hi, uint64_t *d_cv) {
    Ctx *c = &h->c;
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t abytes = (size_t)s * nbr * 2 * (level_a + 1) * c->N * 8;
    Buf dA;
    if (dA.alloc(c, abytes)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(dA.p, A, abytes, cudaMemcpyDefault, c->stream));
    return mm_partial_dev(c, dA.as<uint64_t>(), s, nbr, level_a, max_level, cache->ca, bi_lo, bi_hi, d_cv);
}
int sfg_cv_mod_reduce(sfg_ctx *h, const sfg_cache *cache, int s, int max_level, uint64_t *d_cv, size_t first_elem, size_t num_elems) {
    Ctx *c = &h->c;
    (void)s;
    (void)max_level;
    const size_t N = c->N, L = cache->ca->L;
    if (first_elem % (L * N) || num_elems % (L * N)) SFG_FAIL(c, "cv_mod_reduce: range must cover whole [L][N] polynomials");
    SFG_CUDA(c, cudaSetDevice(c->device));
    if (launch_mod_reduce(c, d_cv + first_elem, num_elems / N, (int)L, c->stream)) return -1;
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
size_t sfg_matmult4_baby_chunk_bytes(const sfg_ctx *, const sfg_cache *cache, int s, int nparts) {
    return nparts < 1 ? 0 : mm_baby_chunk_bytes(cache->ca, s, nparts);
}
int sfg_matmult4_baby_dev(sfg_ctx *h, const uint64_t *d_A, int s, int nbr, int level_a, int max_level, const sfg_cache *cache, int part, int nparts,
                          void *d_R) {
    return mm_baby_dev(&h->c, d_A, s, nbr, level_a, max_level, cache->ca, part, nparts, d_R);
}
int sfg_matmult4_stream_compute_r_dev(sfg_ctx *h, const void *d_R, int s, int max_level, const sfg_cache *cache, uint64_t *d_out) {
    return mm_compute_r_dev(&h->c, d_R, s, max_level, cache->ca, d_out);
}
int sfg_ct_mod_reduce(sfg_ctx *h, uint64_t *d_polys, size_t npoly, int nl) {
    Ctx *c = &h->c;
    if (nl < 1 || nl > c->nQ) SFG_FAIL(c, "ct_mod_reduce: bad limb count %d", nl);
    SFG_CUDA(c, cudaSetDevice(c->device));
    if (launch_mod_reduce(c, d_polys, npoly * (size_t)nl, nl, c->stream)) return -1;
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int sfg_matmult4_finish(sfg_ctx *h, const sfg_cache *cache, int s, int max_level, const uint64_t *d_cv, int g_lo, int g_hi, uint64_t *out) {
    Ctx *c = &h->c;
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t obytes = (size_t)s * cache->ca->m_ct * 2 * max_level * c->N * 8;
    Buf dO;
    if (dO.alloc(c, obytes)) return -1;
    if (mm_finish_dev(c, cache->ca, s, max_level, d_cv, g_lo, g_hi, dO.as<uint64_t>())) return -1;
    SFG_CUDA(c, cudaMemcpy(out, dO.p, obytes, cudaMemcpyDefault));
    return 0;
}
int sfg_matmult4_finish_dev(sfg_ctx *h, const sfg_cache *cache, int s, int max_level, const uint64_t *d_cv_share, int g_lo, int g_hi,
                            uint64_t *d_out) {
    Ctx *c = &h->c;
    const Cache *ca = cache->ca;
    // mm_finish_dev addresses giant gi at d_cv + gi * per_g: rebase the share so that its first giant lands at g_lo
    const size_t per_g = (size_t)ca->m_ct * 2 * s * ca->L * c->N;
    return mm_finish_dev(c, cache->ca, s, max_level, d_cv_share - (size_t)g_lo * per_g, g_lo, g_hi, d_out);
}
int sfg_ct_add(sfg_ctx *h, const uint64_t *a, const uint64_t *b, int ncts, int nl, uint64_t *out) {
    Ctx *c = &h->c;
    if (nl < 1 || nl > c->nQ) SFG_FAIL(c, "ct_add: bad limb count");
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t n = (size_t)ncts * 2 * nl * c->N;
    Buf da, db;
    if (da.alloc(c, n * 8) || db.alloc(c, n * 8)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(da.p, a, n * 8, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaMemcpyAsync(db.p, b, n * 8, cudaMemcpyDefault, c->stream));
    std::vector<long long> offs(ncts), offs_b(ncts);
    for (int t = 0; t < ncts; t++) { offs[t] = (long long)t * 2 * nl * c->N; offs_b[t] = offs[t] * 8; }
    Buf doffs, doffs_b;
    if (doffs.alloc(c, std::max(1, ncts) * sizeof(long long)) || doffs_b.alloc(c, std::max(1, ncts) * sizeof(long long))) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(doffs.p, offs.data(), ncts * sizeof(long long), cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaMemcpyAsync(doffs_b.p, offs_b.data(), ncts * sizeof(long long), cudaMemcpyDefault, c->stream));
    KsBatch kb{};
    kb.nct = ncts;
    kb.in = da.as<uint64_t>();
    kb.in_off = doffs.as<long long>();
    kb.in_nl = nl;
    kb.out = db.p;
    kb.out_off = doffs_b.as<long long>();
    kb.out_layout = make_layout(c, nl, false);
    kb.accumulate = true;
    if (launch_copy_add(c, kb, c->stream)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(out, db.p, n * 8, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int sfg_cache_write_files(sfg_ctx *h, const sfg_cache *cache, const char *prefix) { return cache_write_files(&h->c, cache->ca, prefix); }
int sfg_cache_load_files(sfg_ctx *h, const char *prefix, size_t nrows, size_t ncols, int max_level, sfg_cache **out) {
    *out = nullptr;
    Cache *ca = nullptr;
    if (cache_load_files(&h->c, prefix, nrows, ncols, max_level, &ca)) return -1;
    *out = new sfg_cache{ca};
    return 0;
}

// crypto.SaveCipherMatrixToFile / LoadCipherMatrixFromFile: host-only (no device, no context needed); errors via sfg_last_error(NULL)
int sfg_cipher_matrix_save(const char *filename, int logN, const uint64_t *cts, const double *scales, int nrows, int ncols, int level) {
    return cm_save(filename, logN, cts, scales, nrows, ncols, level + 1, g_create_err);
}
int sfg_cipher_matrix_info(const char *filename, int *nrows, int *ncols, int *level, int *logN) {
    int nl = 0;
    if (cm_info(filename, nrows, ncols, &nl, logN, g_create_err)) return -1;
    *level = nl - 1;
    return 0;
}
int sfg_cipher_matrix_load(const char *filename, int logN, uint64_t *cts, double *scales, int nrows, int ncols, int level) {
    return cm_load(filename, logN, cts, scales, nrows, ncols, level + 1, g_create_err);
}

int sfg_geno_count_sketch(sfg_ctx *h, const sfg_geno *g, const int32_t *rand_index, const int8_t *sgn, int kp, double *sketch, uint64_t *xsum,
                          uint64_t *x2sum, float *scan_ms) {
    return geno_count_sketch(&h->c, g->g, rand_index, sgn, kp, sketch, xsum, x2sum, scan_ms);
}

int sfg_ntt_dev(sfg_ctx *h, uint64_t *d_polys, int npoly, const int *limb_idx, int nsel, int inverse) {
    Ctx *c = &h->c;
    if (nsel < 1 || nsel > kMaxLimbs || npoly % nsel) SFG_FAIL(c, "sfg_ntt_dev: npoly must be a multiple of nsel (<= %d)", kMaxLimbs);
    SFG_CUDA(c, cudaSetDevice(c->device));
    LimbSel sel;
    sel.n = nsel;
    for (int i = 0; i < nsel; i++) {
        if (limb_idx[i] < 0 || limb_idx[i] >= c->nQP) SFG_FAIL(c, "sfg_ntt_dev: modulus index %d out of range", limb_idx[i]);
        sel.idx[i] = limb_idx[i];
    }
    return launch_ntt(c, d_polys, (size_t)nsel * c->N, d_polys, (size_t)nsel * c->N, npoly, sel, inverse != 0, c->stream);
}
int sfg_rotate_right_dev(sfg_ctx *h, int level, const uint64_t *d_cts, int nct, int nrot, uint64_t *d_out) {
    Ctx *c = &h->c;
    if (level < 0 || level >= c->nQ || nct < 1) SFG_FAIL(c, "sfg_rotate_right_dev: bad level / count");
    return rotate_right_dev(c, level, d_cts, nct, nrot, d_out);
}

// ---- ciphertext algebra of the callers (gwas/matmult.go:27-116) ----
int sfg_ctx_set_relin_key(sfg_ctx *h, const uint64_t *key) {
    // stored as the "rotation by 0" key: galEl = 1, whose NTT permutation is the identity (rotation by 0 itself never key-switches)
    return sfg_ctx_set_rotation_key(h, 0, key);
}
int sfg_ctx_set_relin_key_ptrs(sfg_ctx *h, const uint64_t *const *limbs) { return sfg_ctx_set_rotation_key_ptrs(h, 0, limbs); }

namespace {
struct DevIO {  // host operand -> device, device result -> host
    Ctx *c;
    Buf in[2], out;
    int up(int k, const void *src, size_t bytes) {
        if (in[k].alloc(c, bytes)) return -1;
        SFG_CUDA(c, cudaMemcpyAsync(in[k].p, src, bytes, cudaMemcpyDefault, c->stream));
        return 0;
    }
    int down(void *dst, size_t bytes) {
        SFG_CUDA(c, cudaMemcpyAsync(dst, out.p, bytes, cudaMemcpyDefault, c->stream));
        SFG_CUDA(c, cudaStreamSynchronize(c->stream));
        return 0;
    }
};
}  // namespace

int sfg_ct_mul_relin(sfg_ctx *h, int level, const uint64_t *x, int nx, int x_nl, const uint64_t *y, int ny, int y_nl, int nrescale, uint64_t *out) {
    Ctx *c = &h->c;
    if (nx < 1 || ny < 1 || x_nl < 1 || y_nl < 1 || nrescale < 0 || nrescale > level) SFG_FAIL(c, "sfg_ct_mul_relin: bad counts");
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t N = c->N, n = std::max(nx, ny), ob = n * 2 * (size_t)(level + 1 - nrescale) * N * 8;
    DevIO io{c};
    if (io.up(0, x, (size_t)nx * 2 * x_nl * N * 8) || io.up(1, y, (size_t)ny * 2 * y_nl * N * 8) || io.out.alloc(c, n * 2 * (size_t)(level + 1) * N * 8)) return -1;
    if (mul_relin_dev(c, level, io.in[0].as<uint64_t>(), nx, x_nl, io.in[1].as<uint64_t>(), ny, y_nl, nrescale, io.out.as<uint64_t>())) return -1;
    return io.down(out, ob);
}
int sfg_ct_mul_plain(sfg_ctx *h, int level, const uint64_t *pt, int npt, int pt_nl, const uint64_t *cts, int nct, int ct_nl, int nrescale, uint64_t *out) {
    Ctx *c = &h->c;
    if (npt < 1 || nct < 1 || pt_nl < 1 || ct_nl < 1 || nrescale < 0 || nrescale > level) SFG_FAIL(c, "sfg_ct_mul_plain: bad counts");
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t N = c->N, ob = (size_t)nct * 2 * (size_t)(level + 1 - nrescale) * N * 8;
    DevIO io{c};
    if (io.up(0, pt, (size_t)npt * pt_nl * N * 8) || io.up(1, cts, (size_t)nct * 2 * ct_nl * N * 8) || io.out.alloc(c, (size_t)nct * 2 * (level + 1) * N * 8)) return -1;
    if (mul_plain_dev(c, level, io.in[0].as<uint64_t>(), npt, pt_nl, io.in[1].as<uint64_t>(), nct, ct_nl, nrescale, io.out.as<uint64_t>())) return -1;
    return io.down(out, ob);
}
int sfg_ct_rescale(sfg_ctx *h, int level, const uint64_t *cts, int nct, int nrescale, uint64_t *out) {
    Ctx *c = &h->c;
    if (nct < 1 || level < 0 || level >= c->nQ || nrescale < 0 || nrescale > level) SFG_FAIL(c, "sfg_ct_rescale: bad level / counts");
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t N = c->N;
    DevIO io{c};
    if (io.up(0, cts, (size_t)nct * 2 * (level + 1) * N * 8) || io.out.alloc(c, (size_t)nct * 2 * (level + 1) * N * 8)) return -1;
    if (rescale_dev(c, level, io.in[0].as<uint64_t>(), nct, nrescale, io.out.as<uint64_t>())) return -1;
    return io.down(out, (size_t)nct * 2 * (level + 1 - nrescale) * N * 8);
}
static int ct_addsub_host(sfg_ctx *h, int level, const uint64_t *a, int na, int a_nl, const uint64_t *b, int nb, int b_nl, bool sub, uint64_t *out) {
    Ctx *c = &h->c;
    if (na < 1 || nb < 1 || a_nl < 1 || b_nl < 1) SFG_FAIL(c, "ct add/sub: bad counts");
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t N = c->N, n = std::max(na, nb), ob = n * 2 * (size_t)(level + 1) * N * 8;
    DevIO io{c};
    if (io.up(0, a, (size_t)na * 2 * a_nl * N * 8) || io.up(1, b, (size_t)nb * 2 * b_nl * N * 8) || io.out.alloc(c, ob)) return -1;
    if (addsub_dev(c, level, io.in[0].as<uint64_t>(), na, a_nl, io.in[1].as<uint64_t>(), nb, b_nl, sub, io.out.as<uint64_t>())) return -1;
    return io.down(out, ob);
}
int sfg_ct_sub(sfg_ctx *h, int level, const uint64_t *a, int na, int a_nl, const uint64_t *b, int nb, int b_nl, uint64_t *out) {
    return ct_addsub_host(h, level, a, na, a_nl, b, nb, b_nl, true, out);
}
int sfg_ct_add2(sfg_ctx *h, int level, const uint64_t *a, int na, int a_nl, const uint64_t *b, int nb, int b_nl, uint64_t *out) {
    return ct_addsub_host(h, level, a, na, a_nl, b, nb, b_nl, false, out);
}
int sfg_inner_sum_all(sfg_ctx *h, int level, const uint64_t *cts, int nvec, int cnt, uint64_t *out) {
    Ctx *c = &h->c;
    if (nvec < 1 || cnt < 1 || level < 0 || level >= c->nQ) SFG_FAIL(c, "sfg_inner_sum_all: bad level / counts");
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t N = c->N, ct = (size_t)2 * (level + 1) * N * 8;
    DevIO io{c};
    if (io.up(0, cts, (size_t)nvec * cnt * ct) || io.out.alloc(c, (size_t)nvec * ct)) return -1;
    if (inner_sum_all_dev(c, level, io.in[0].as<uint64_t>(), nvec, cnt, io.out.as<uint64_t>())) return -1;
    return io.down(out, (size_t)nvec * ct);
}
// ---- device-resident ciphertext vectors: the *_dev orchestration of matmult.cu behind handles ----
static int cts_new(Ctx *c, int n, int nl, sfg_cts **out) {
    *out = nullptr;
    if (n < 1 || nl < 1 || nl > c->nQ) SFG_FAIL(c, "sfg_cts: bad shape (%d ciphertexts, %d limbs)", n, nl);
    SFG_CUDA(c, cudaSetDevice(c->device));
    uint64_t *d = nullptr;
    if (dev_alloc(c, (void **)&d, (size_t)n * 2 * nl * c->N * 8, "ciphertext vector")) return -1;
    *out = new sfg_cts{c->device, n, nl, d};
    return 0;
}
int sfg_cts_upload(sfg_ctx *h, const uint64_t *host, int n, int nl, sfg_cts **out) {
    Ctx *c = &h->c;
    if (cts_new(c, n, nl, out)) return -1;
    if (upload(c, (*out)->d, host, (size_t)n * 2 * nl * c->N * 8)) {
        sfg_cts_destroy(*out);
        *out = nullptr;
        return -1;
    }
    return 0;
}
int sfg_cts_download(sfg_ctx *h, const sfg_cts *v, uint64_t *host) {
    Ctx *c = &h->c;
    SFG_CUDA(c, cudaSetDevice(c->device));
    SFG_CUDA(c, cudaMemcpyAsync(host, v->d, (size_t)v->n * 2 * v->nl * c->N * 8, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int sfg_cts_shape(const sfg_cts *v, int *n, int *nl) {
    if (n) *n = v->n;
    if (nl) *nl = v->nl;
    return 0;
}
void sfg_cts_destroy(sfg_cts *v) {
    if (!v) return;
    cudaSetDevice(v->device);
    cudaFree(v->d);
    delete v;
}
int sfg_cts_slice(sfg_ctx *h, const sfg_cts *v, int first, int count, sfg_cts **out) {
    Ctx *c = &h->c;
    if (first < 0 || count < 1 || first + count > v->n) SFG_FAIL(c, "sfg_cts_slice: [%d, %d) out of %d ciphertexts", first, first + count, v->n);
    if (cts_new(c, count, v->nl, out)) return -1;
    const size_t ct = (size_t)2 * v->nl * c->N;
    SFG_CUDA(c, cudaMemcpyAsync((*out)->d, v->d + first * ct, count * ct * 8, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int sfg_cts_matmult4_stream_compute(sfg_ctx *h, const sfg_cts *A, int s, int nbr, int max_level, const sfg_cache *cache, sfg_cts **out) {
    Ctx *c = &h->c;
    *out = nullptr;
    if (s < 1 || nbr < 1 || s * nbr != A->n) SFG_FAIL(c, "sfg_cts_matmult4_stream_compute: A holds %d ciphertexts, not %d x %d", A->n, s, nbr);
    sfg_cts *o = nullptr;
    if (cts_new(c, s * cache->ca->m_ct, max_level, &o)) return -1;
    if (mm_compute_dev(c, A->d, s, nbr, A->nl - 1, max_level, cache->ca, o->d)) {
        sfg_cts_destroy(o);
        return -1;
    }
    *out = o;
    return 0;
}
int sfg_cts_mul_relin(sfg_ctx *h, int level, const sfg_cts *x, const sfg_cts *y, int nrescale, sfg_cts **out) {
    Ctx *c = &h->c;
    *out = nullptr;
    if (nrescale < 0 || nrescale > level) SFG_FAIL(c, "sfg_cts_mul_relin: bad rescale count");
    const int n = std::max(x->n, y->n);
    sfg_cts *full = nullptr, *o = nullptr;
    if (cts_new(c, n, level + 1, &full)) return -1;  // mul_relin_dev needs the un-rescaled product as scratch when nrescale > 0
    int rc = mul_relin_dev(c, level, x->d, x->n, x->nl, y->d, y->n, y->nl, nrescale, full->d);
    if (!rc && nrescale > 0) {  // results are dense [n][2][level+1-nrescale][N] at the front of the buffer: move into a right-sized handle
        rc = cts_new(c, n, level + 1 - nrescale, &o);
        if (!rc) {
            cudaMemcpyAsync(o->d, full->d, (size_t)n * 2 * (level + 1 - nrescale) * c->N * 8, cudaMemcpyDefault, c->stream);
            cudaStreamSynchronize(c->stream);
        }
        sfg_cts_destroy(full);
        full = o;
    }
    if (rc) {
        sfg_cts_destroy(full);
        return -1;
    }
    *out = full;
    return 0;
}
int sfg_cts_mul_plain(sfg_ctx *h, int level, const uint64_t *pt, int npt, int pt_nl, const sfg_cts *v, int nrescale, sfg_cts **out) {
    Ctx *c = &h->c;
    *out = nullptr;
    if (npt < 1 || pt_nl < level + 1 || nrescale < 0 || nrescale > level) SFG_FAIL(c, "sfg_cts_mul_plain: bad counts");
    SFG_CUDA(c, cudaSetDevice(c->device));
    Buf dpt;
    if (dpt.alloc(c, (size_t)npt * pt_nl * c->N * 8) || upload(c, dpt.p, pt, (size_t)npt * pt_nl * c->N * 8)) return -1;
    sfg_cts *full = nullptr, *o = nullptr;
    if (cts_new(c, v->n, level + 1, &full)) return -1;
    int rc = mul_plain_dev(c, level, dpt.as<uint64_t>(), npt, pt_nl, v->d, v->n, v->nl, nrescale, full->d);
    if (!rc && nrescale > 0) {
        rc = cts_new(c, v->n, level + 1 - nrescale, &o);
        if (!rc) {
            cudaMemcpyAsync(o->d, full->d, (size_t)v->n * 2 * (level + 1 - nrescale) * c->N * 8, cudaMemcpyDefault, c->stream);
            cudaStreamSynchronize(c->stream);
        }
        sfg_cts_destroy(full);
        full = o;
    }
    if (rc) {
        sfg_cts_destroy(full);
        return -1;
    }
    *out = full;
    return 0;
}
int sfg_cts_addsub(sfg_ctx *h, int level, const sfg_cts *a, const sfg_cts *b, int subtract, sfg_cts **out) {
    Ctx *c = &h->c;
    *out = nullptr;
    sfg_cts *o = nullptr;
    if (cts_new(c, std::max(a->n, b->n), level + 1, &o)) return -1;
    if (addsub_dev(c, level, a->d, a->n, a->nl, b->d, b->n, b->nl, subtract != 0, o->d)) {
        sfg_cts_destroy(o);
        return -1;
    }
    *out = o;
    return 0;
}
int sfg_cts_inner_sum_all(sfg_ctx *h, int level, const sfg_cts *v, int nvec, int cnt, sfg_cts **out) {
    Ctx *c = &h->c;
    *out = nullptr;
    if (nvec < 1 || cnt < 1 || nvec * cnt != v->n || v->nl != level + 1)
        SFG_FAIL(c, "sfg_cts_inner_sum_all: %d ciphertexts of %d limbs are not %d vectors of %d at level %d", v->n, v->nl, nvec, cnt, level);
    sfg_cts *o = nullptr;
    if (cts_new(c, nvec, level + 1, &o)) return -1;
    if (inner_sum_all_dev(c, level, v->d, nvec, cnt, o->d)) {
        sfg_cts_destroy(o);
        return -1;
    }
    *out = o;
    return 0;
}

// ---- local arithmetic of the collective bootstrap (mpc/mhe.go:262-341) ----
int sfg_refresh_gen_shares(sfg_ctx *h, int level, int nct, const uint64_t *c1, const uint64_t *sk_mont, const uint64_t *crp, const uint64_t *mask_mag,
                           const int8_t *mask_sign, int nwords, double in_scale, double out_scale, const int64_t *e0, const int64_t *e1, uint64_t *share_decrypt,
                           uint64_t *share_recrypt) {
    Ctx *c = &h->c;
    if (level < 0 || level >= c->nQ || nct < 1 || nwords < 1) SFG_FAIL(c, "sfg_refresh_gen_shares: bad level / counts");
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t N = c->N, nl = level + 1, nQ = c->nQ, n = nct;
    Buf dc1, dsk, dcrp, dm, dsg, de0, de1, dh0, dh1;
    struct { Buf *b; const void *src; size_t bytes; } in[] = {{&dc1, c1, n * nl * N * 8}, {&dsk, sk_mont, nQ * N * 8}, {&dcrp, crp, n * nQ * N * 8},
                                                              {&dm, mask_mag, n * N * nwords * 8}, {&dsg, mask_sign, n * N}, {&de0, e0, n * N * 8},
                                                              {&de1, e1, n * N * 8}};
    for (auto &x : in) {
        if (x.b->alloc(c, x.bytes)) return -1;
        SFG_CUDA(c, cudaMemcpyAsync(x.b->p, x.src, x.bytes, cudaMemcpyDefault, c->stream));
    }
    if (dh0.alloc(c, n * nl * N * 8) || dh1.alloc(c, n * nQ * N * 8)) return -1;
    if (refresh_gen_shares_dev(c, level, nct, dc1.as<uint64_t>(), dsk.as<uint64_t>(), dcrp.as<uint64_t>(), dm.as<uint64_t>(), dsg.as<int8_t>(), nwords,
                               in_scale, out_scale, de0.as<long long>(), de1.as<long long>(), dh0.as<uint64_t>(), dh1.as<uint64_t>()))
        return -1;
    SFG_CUDA(c, cudaMemcpyAsync(share_decrypt, dh0.p, n * nl * N * 8, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaMemcpyAsync(share_recrypt, dh1.p, n * nQ * N * 8, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
int sfg_refresh_finish(sfg_ctx *h, int level, int nct, const uint64_t *c0, int c0_nl, double in_scale, double out_scale, const uint64_t *agg_decrypt,
                       const uint64_t *agg_recrypt, const uint64_t *crp, uint64_t *out) {
    Ctx *c = &h->c;
    if (level < 0 || level >= c->nQ || nct < 1 || c0_nl < level + 1) SFG_FAIL(c, "sfg_refresh_finish: bad level / counts");
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t N = c->N, nl = level + 1, nQ = c->nQ, n = nct;
    Buf dc0, da0, da1, dcrp, dout;
    struct { Buf *b; const void *src; size_t bytes; } in[] = {{&dc0, c0, n * c0_nl * N * 8}, {&da0, agg_decrypt, n * nl * N * 8},
                                                              {&da1, agg_recrypt, n * nQ * N * 8}, {&dcrp, crp, n * nQ * N * 8}};
    for (auto &x : in) {
        if (x.b->alloc(c, x.bytes)) return -1;
        SFG_CUDA(c, cudaMemcpyAsync(x.b->p, x.src, x.bytes, cudaMemcpyDefault, c->stream));
    }
    if (dout.alloc(c, n * 2 * nQ * N * 8)) return -1;
    if (refresh_finish_dev(c, level, nct, dc0.as<uint64_t>(), c0_nl, in_scale, out_scale, da0.as<uint64_t>(), da1.as<uint64_t>(), dcrp.as<uint64_t>(),
                           dout.as<uint64_t>()))
        return -1;
    SFG_CUDA(c, cudaMemcpyAsync(out, dout.p, n * 2 * nQ * N * 8, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int sfg_encode_slots_i8(sfg_ctx *h, const int8_t *v, int level, int mont, uint64_t *out) { return encode_slots_host(&h->c, v, level, mont != 0, out); }

}  // extern "C"
