// matmult.cu -- host orchestration of MatMult4StreamPreprocess / MatMult4StreamCompute / MatMult4Stream
// (gwas/matmult.go:914-1505) on one GPU, and the block-row-sharded pieces used for multi-GPU runs.
//
// Reference control flow -> device schedule
//   Preprocess : per block row, per active shift, per block column: GetDiag + EncodeNTT + MForm, written to disk.
//                Here: one encode CTA per diagonal polynomial, written into a compact HBM cache (or regenerated per
//                giant-step chunk when the cache would not fit the budget).
//   Compute    : per block row: baby rotations of A (key-switch), then for every cached diagonal a 128-bit lazy MAC into
//                acc[i][giant] under a mutex; afterwards per (i, giant): Montgomery reduce, giant rotation, Add into out.
//                Here: (1) all baby rotations, batched per baby step over every (i, bi) that shares the Galois key;
//                (2) ONE output-stationary MAC launch per giant chunk that streams the diagonals exactly once and
//                emits canonical residues (K1+K2 fused); (3) giant rotations batched per giant step over every (i, bj),
//                accumulated into out in place (K6+K7 fused).
#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "matmult.h"

namespace sfg {

thread_local float g_last_ms[5] = {0, 0, 0, 0, 0};

// ---------------------------------------------------------------------------------------------------------------
// genotype matrix
// ---------------------------------------------------------------------------------------------------------------
int geno_create(Ctx *c, size_t nrows, size_t ncols, Geno **out) {
    if (nrows == 0 || ncols == 0) SFG_FAIL(c, "empty genotype matrix (%zu x %zu)", nrows, ncols);
    SFG_CUDA(c, cudaSetDevice(c->device));
    Geno *g = new Geno();
    g->c = c;
    g->device = c->device;
    g->nrows = nrows;
    g->ncols = ncols;
    if (dev_alloc(c, (void **)&g->d, nrows * ncols, "genotype matrix")) {
        delete g;
        return -1;
    }
    *out = g;
    return 0;
}
int geno_push(Geno *g, const int8_t *rows, size_t n) {
    Ctx *c = g->c;
    if (g->filled + n > g->nrows) SFG_FAIL(c, "geno_push: %zu rows pushed into a %zu-row matrix (already %zu)", n, g->nrows, g->filled);
    SFG_CUDA(c, cudaSetDevice(c->device));
    SFG_CUDA(c, cudaMemcpyAsync(g->d + g->filled * g->ncols, rows, n * g->ncols, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    g->filled += n;
    return 0;
}
void geno_release(Geno *g) {
    if (!g) return;
    if (--g->refs == 0) {
        cudaSetDevice(g->device);
        cudaFree(g->d);
        delete g;
    }
}

// gwas/matmult.go:627-631 GetDiagBool for index = -shift
static inline bool diag_exists(int r, int cdim, int slots, int shift) {
    const int index = (slots - shift) % slots;
    return (slots + 1 - r) <= index || index <= cdim - 1;
}

static int block_rows(const Cache *ca, int bi) {
    return (int)std::min<size_t>((size_t)(bi + 1) * ca->slots, ca->nrows) - bi * ca->slots;
}
static int block_cols(const Cache *ca, int bj) {
    return (int)std::min<size_t>((size_t)(bj + 1) * ca->slots, ca->ncols) - bj * ca->slots;
}

// ---------------------------------------------------------------------------------------------------------------
// Preprocess (gwas/matmult.go:914-1041)
// ---------------------------------------------------------------------------------------------------------------
static int encode_jobs(Ctx *c, const Cache *ca, const std::vector<EncJob> &jobs, void *P) {
    if (jobs.empty()) return 0;
    void *dj;
    if (ws_get(c, WS_META2, jobs.size() * sizeof(EncJob), &dj)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(dj, jobs.data(), jobs.size() * sizeof(EncJob), cudaMemcpyDefault, c->stream));
    // plain residues (mont = false): the byte planes of the image are cut from the canonical value
    if (launch_encode(c, ca->g->d, ca->ncols, (const EncJob *)dj, (int)jobs.size(), ca->lay, false, P, nullptr, c->stream)) return -1;
    return 0;
}

// Encode the diagonals of column tiles [tile_lo, tile_hi) (columns = (active giant, block column) pairs, 128 per tile) for the K
// groups [grp_lo, grp_hi) and lay them out as a tensor-core image holding exactly those tiles and groups.
// img: (grp_hi - grp_lo) * tc_group_bytes(ntiles) bytes.
static long long tc_group_bytes(const Cache *ca, int ntiles) {
    long long b = 0;
    for (int l = 0; l < ca->tc.L; l++) b += (long long)ca->tc.N * ntiles * ca->tc.nb[l] * 128 * ca->tc.Kg;
    return b;
}
static int build_p_tiles(Ctx *c, const Cache *ca, int tile_lo, int tile_hi, int grp_lo, int grp_hi, unsigned char *img) {
    const TcGeomP &tc = ca->tc;
    const int d = ca->d, m_ct = ca->m_ct, slots = ca->slots, Kg = tc.Kg, K = tc.K;
    const int ntl = tile_hi - tile_lo;
    const long long gbytes = tc_group_bytes(ca, ntl);
    // no memset of the image: k_img_build writes EVERY byte of a (tile, K group) block -- zeros for nil diagonals, padding columns and
    // padding K steps (table entry -1) -- so the 55 GB image of config 2 is written exactly once
    void *tmp, *dtab;
    const size_t RB = (size_t)ca->lay.bytes;
    if (ws_get(c, WS_TMPP, (size_t)128 * Kg * RB, &tmp)) return -1;
    if (ws_get(c, WS_POFF, (size_t)128 * Kg * sizeof(long long), &dtab)) return -1;
    std::vector<long long> tab((size_t)128 * Kg);
    std::vector<EncJob> jobs;
    for (int ct = tile_lo; ct < tile_hi; ct++)
        for (int grp = grp_lo; grp < grp_hi; grp++) {
            jobs.clear();
            std::fill(tab.begin(), tab.end(), -1LL);
            for (int cc = 0; cc < 128; cc++) {
                const int col = ct * 128 + cc;
                if (col >= tc.ncols) break;
                const int g = ca->gact[col / m_ct], bj = col % m_ct;
                for (int kk = 0; kk < Kg; kk++) {
                    const int k = grp * Kg + kk;
                    if (k >= K) break;
                    const int bi = ca->kbi[k], b = ca->kb[k];
                    const int shift = g * d + b;
                    if (shift >= slots) continue;
                    if (ca->pidx[((size_t)bi * slots + shift) * m_ct + bj] < 0) continue;
                    const long long off = (long long)((size_t)kk * 128 + cc) * (long long)RB;
                    tab[(size_t)kk * 128 + cc] = off;
                    // EncodeDiagWithEncoder(blockVec, -shift, d*giant, maxLevel, enc)  matmult.go:1024
                    jobs.push_back(EncJob{bi * slots, bj * slots, block_rows(ca, bi), block_cols(ca, bj), shift, d * g, off});
                }
            }
            SFG_CUDA(c, cudaMemcpyAsync(dtab, tab.data(), tab.size() * sizeof(long long), cudaMemcpyDefault, c->stream));
            if (!jobs.empty() && encode_jobs(c, ca, jobs, tmp)) return -1;  // (no diagonal in this block: the all -1 table writes zeros)
            if (launch_img_p(c, tc, ca->lay, tmp, (const long long *)dtab, ntl, ct - tile_lo, img + (size_t)(grp - grp_lo) * gbytes, c->stream))
                return -1;
        }
    return 0;
}

static int cache_meta_finish(Ctx *c, Cache *ca);
// shapes, index tables (gwas/matmult.go:962-974) and image geometry of a cache over an nrows x ncols matrix
static int cache_meta_base(Ctx *c, Cache *ca, size_t nrows, size_t ncols, int maxLevel) {
    ca->c = c;
    ca->device = c->device;
    ca->maxLevel = maxLevel;
    ca->L = maxLevel;  // limb-COUNT quirk: accumulators cover limbs 0..maxLevel-1 (gwas/matmult.go:1125 -> :231, App. A.4)
    if (maxLevel > kMaxLayoutLimbs) SFG_FAIL(c, "maxLevel %d > %d not supported", maxLevel, kMaxLayoutLimbs);
    ca->lay = make_layout(c, ca->L, true);
    ca->slots = c->slots;
    ca->d = c->d;
    ca->nrows = nrows;
    ca->ncols = ncols;
    ca->m_ct = (int)((ncols - 1) / c->slots) + 1;   // matmult.go:920
    ca->nbr = (int)((nrows - 1) / c->slots) + 1;    // matmult.go:921
    const int slots = ca->slots, d = ca->d, m_ct = ca->m_ct, nbr = ca->nbr;
    ca->baby.assign((size_t)nbr * d, 0);
    ca->giant.assign((size_t)nbr * d, 0);
    ca->shiftT.assign((size_t)nbr * slots, 0);
    ca->pidx.assign((size_t)nbr * slots * m_ct, -1);
    return 0;
}
static int cache_init_meta(Ctx *c, Cache *ca, size_t nrows, size_t ncols, int maxLevel) {
    if (cache_meta_base(c, ca, nrows, ncols, maxLevel)) return -1;
    const int slots = ca->slots, d = ca->d, m_ct = ca->m_ct, nbr = ca->nbr;
    size_t npoly = 0;
    for (int bi = 0; bi < nbr; bi++) {
        const int nr = block_rows(ca, bi);
        for (int shift = 0; shift < slots; shift++) {
            bool any = false;
            for (int bj = 0; bj < m_ct; bj++) {
                if (diag_exists(nr, block_cols(ca, bj), slots, shift)) {
                    any = true;
                    ca->pidx[((size_t)bi * slots + shift) * m_ct + bj] = (int)(npoly++);
                }
            }
            if (any) {  // matmult.go:962-974
                ca->baby[(size_t)bi * d + shift % d] = 1;
                ca->giant[(size_t)bi * d + shift / d] = 1;
                ca->shiftT[(size_t)bi * slots + shift] = 1;
            }
        }
    }
    ca->npoly = npoly;
    return cache_meta_finish(c, ca);
}
// K list, active giants and image geometry from the baby / giant tables
static int cache_meta_finish(Ctx *c, Cache *ca) {
    const int d = ca->d, m_ct = ca->m_ct, nbr = ca->nbr;
    const size_t npoly = ca->npoly;
    ca->kidx.assign((size_t)nbr * d, -1);
    if (ca->bi_hi < 0) ca->bi_hi = nbr;
    for (int bi = ca->bi_lo; bi < ca->bi_hi; bi++)  // the K list (and with it the image) covers the cache's own block rows only
        for (int b = 0; b < d; b++)
            if (ca->baby[(size_t)bi * d + b]) {
                ca->kidx[(size_t)bi * d + b] = (int)ca->kbi.size();
                ca->kbi.push_back(bi);
                ca->kb.push_back(b);
            }
    for (int gi = 0; gi < d; gi++) {
        bool any = false;
        for (int bi = 0; bi < nbr; bi++) any |= ca->giant[(size_t)bi * d + gi] != 0;
        if (any) ca->gact.push_back(gi);
    }
    if (ca->g_nparts > 1) {  // contiguous, balanced share of the active giant steps (the first n % parts shares get one more)
        const int n = (int)ca->gact.size(), base = n / ca->g_nparts, extra = n % ca->g_nparts;
        const int lo = ca->g_part * base + std::min(ca->g_part, extra), hi = lo + base + (ca->g_part < extra ? 1 : 0);
        ca->gact = std::vector<int>(ca->gact.begin() + lo, ca->gact.begin() + hi);
    }
    if (npoly > 0x7fffffffULL) SFG_FAIL(c, "too many diagonal polynomials");
    return tc_geom_p(c, ca->L, (int)ca->kbi.size(), (int)ca->gact.size() * m_ct, &ca->tc);
}

int cache_build(Ctx *c, Geno *g, int maxLevel, Cache **out, int bi_lo, int bi_hi, int g_part, int g_nparts) {
    if (g_nparts < 1 || g_part < 0 || g_part >= g_nparts) SFG_FAIL(c, "giant-step share %d of %d", g_part, g_nparts);
    if (g->filled != g->nrows) SFG_FAIL(c, "genotype matrix incomplete: %zu of %zu rows pushed", g->filled, g->nrows);
    if (maxLevel < 1 || maxLevel > c->nQ - 1) SFG_FAIL(c, "maxLevel %d needs %d Q limbs, parameters have %d", maxLevel, maxLevel + 1, c->nQ);
    const int nbr_all = (int)((g->nrows - 1) / c->slots) + 1;
    if (bi_hi < 0) bi_hi = nbr_all;
    if (bi_lo < 0 || bi_hi > nbr_all || bi_lo > bi_hi) SFG_FAIL(c, "block-row range [%d, %d) out of [0, %d)", bi_lo, bi_hi, nbr_all);
    SFG_CUDA(c, cudaSetDevice(c->device));
    Cache *ca = new Cache();
    ca->g = g;
    g->refs++;
    ca->bi_lo = bi_lo;
    ca->bi_hi = bi_hi;
    ca->g_part = g_part;
    ca->g_nparts = g_nparts;
    if (cache_init_meta(c, ca, g->nrows, g->ncols, maxLevel)) {
        cache_destroy(ca);
        return -1;
    }
    const size_t npoly = ca->npoly;
    // materialise if it fits the budget
    size_t budget = c->cache_budget;
    if (budget == 0) {
        size_t fr = 0, tot = 0;
        SFG_CUDA(c, cudaMemGetInfo(&fr, &tot));
        budget = (size_t)(0.70 * (double)fr);
    }
    const size_t bytes = (size_t)ca->tc.group_bytes * ca->tc.ngroups;
    if (bytes == 0) {  // a share without giant steps (more shares than giant steps): nothing to cache, Compute yields zero
        ca->materialised = true;
    } else if (npoly > 0 && bytes <= budget) {
        cudaError_t e = cudaMalloc(&ca->img, bytes);
        if (e == cudaSuccess) {
            ca->img_bytes = bytes;
            poison_fill(c, ca->img, bytes);  // SFG_POISON=1: a byte the image builder failed to write would surface as a parity failure
            if (build_p_tiles(c, ca, 0, ca->tc.ntiles, 0, ca->tc.ngroups, ca->img) || cudaStreamSynchronize(c->stream) != cudaSuccess) {
                if (c->err.empty()) c->err = "diagonal cache build failed";
                cache_destroy(ca);
                return -1;
            }
            ca->materialised = true;
        } else {
            cudaGetLastError();
        }
    }
    *out = ca;
    return 0;
}

void cache_destroy(Cache *ca) {
    if (!ca) return;
    cudaSetDevice(ca->device);
    cudaFree(ca->img);
    geno_release(ca->g);
    delete ca;
}

// ---------------------------------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------------------------------
static int find_key(Ctx *c, int rot_left, const GaloisKey **out) {
    const uint64_t galEl = h_galois_element(c->logN, rot_left);
    std::lock_guard<std::mutex> g(c->mu);
    auto it = c->keys.find(galEl);
    if (it == c->keys.end()) SFG_FAIL(c, "rotation key for left rotation by %d (galEl %llu) not loaded", rot_left, (unsigned long long)galEl);
    *out = &it->second;
    return 0;
}

struct PhaseTimer {
    std::vector<cudaEvent_t> ev;
    std::vector<int> phase;
    cudaStream_t st;
    explicit PhaseTimer(cudaStream_t s) : st(s) {}
    void mark(int ph) {  // ph = phase that STARTS here (-1 = end, 3 = MAC kernel proper)
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back(e);
        phase.push_back(ph);
    }
    void finish(float out[5]) {
        for (int i = 0; i < 5; i++) out[i] = 0;
        if (ev.empty()) return;
        cudaEventSynchronize(ev.back());
        for (size_t i = 0; i + 1 < ev.size(); i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            if (phase[i] == 3) {  // the MAC kernel alone: also part of the MAC phase
                out[4] += ms;
                out[1] += ms;
            } else if (phase[i] >= 0 && phase[i] < 3) {
                out[phase[i]] += ms;
            }
        }
        cudaEventElapsedTime(&out[3], ev.front(), ev.back());
        for (auto e : ev) cudaEventDestroy(e);
        ev.clear();
    }
};
static thread_local PhaseTimer *g_tm = nullptr;

// ---------------------------------------------------------------------------------------------------------------
// rotation batches: every entry = one ciphertext rotated with its own Galois key (kernels_ks.cu)
// ---------------------------------------------------------------------------------------------------------------
struct RotEntry {
    long long in_off;   // element offset of the input ct
    long long out_off;  // BYTE offset of the output ct
    int c2_slot;        // which INTT(c1) slot it reads
    const GaloisKey *key;
};

// device image of the per-entry arrays of a list of batches, uploaded once
struct RotMeta {
    std::vector<long long> in_off, out_off, c2_src;
    std::vector<const uint64_t *> keys;
    std::vector<const uint32_t *> perms;
    std::vector<int> c2_slot;
    std::vector<uint32_t> ginv;  // galEl^-1 mod 2N of every entry (giant-step sums)
    // device pointers after upload
    const long long *d_in_off = nullptr, *d_out_off = nullptr, *d_c2_src = nullptr;
    const uint64_t *const *d_keys = nullptr;
    const uint32_t *const *d_perms = nullptr;
    const int *d_c2_slot = nullptr;
    const uint32_t *d_ginv = nullptr;
    void add(const RotEntry &e) {
        in_off.push_back(e.in_off);
        out_off.push_back(e.out_off);
        c2_slot.push_back(e.c2_slot);
        keys.push_back(e.key ? e.key->key : nullptr);
        perms.push_back(e.key ? e.key->perm : nullptr);
    }
    int upload(Ctx *c, int slot) {
        const size_t n = in_off.size(), m = c2_src.size();
        const size_t bytes = (2 * n + m) * 8 + 2 * n * 8 + n * 4 + ginv.size() * 4 + 64;
        std::vector<unsigned char> h(bytes);
        unsigned char *p = h.data();
        size_t o_in = 0, o_out = n * 8, o_src = 2 * n * 8, o_keys = (2 * n + m) * 8, o_perms = o_keys + n * 8, o_slot = o_perms + n * 8;
        memcpy(p + o_in, in_off.data(), n * 8);
        memcpy(p + o_out, out_off.data(), n * 8);
        memcpy(p + o_src, c2_src.data(), m * 8);
        memcpy(p + o_keys, keys.data(), n * 8);
        memcpy(p + o_perms, perms.data(), n * 8);
        memcpy(p + o_slot, c2_slot.data(), n * 4);
        const size_t o_ginv = o_slot + n * 4;
        if (!ginv.empty()) memcpy(p + o_ginv, ginv.data(), ginv.size() * 4);
        void *d;
        if (ws_get(c, slot, bytes, &d)) return -1;
        SFG_CUDA(c, cudaMemcpyAsync(d, p, bytes, cudaMemcpyDefault, c->stream));
        unsigned char *db = (unsigned char *)d;
        d_in_off = (const long long *)(db + o_in);
        d_out_off = (const long long *)(db + o_out);
        d_c2_src = (const long long *)(db + o_src);
        d_keys = (const uint64_t *const *)(db + o_keys);
        d_perms = (const uint32_t *const *)(db + o_perms);
        d_c2_slot = (const int *)(db + o_slot);
        d_ginv = (const uint32_t *)(db + o_ginv);
        return 0;
    }
};

// scratch for a batch: c2 [n_c2][nl][N], acc [cap][2][nl+nP][N] with cap bounded by an 8 GiB budget
static int fill_scratch(Ctx *c, KsBatch &kb, int max_nct, int max_c2) {
    const size_t N = c->N;
    const int nl = kb.level + 1, nt = nl + c->nP;
    const size_t per_ct = (size_t)2 * nt * N * 8;
    const int cap = (int)std::max<size_t>(1, std::min<size_t>((size_t)max_nct, ((size_t)8 << 30) / per_ct));
    void *pc2, *pacc;
    if (ws_get(c, WS_C2, (size_t)max_c2 * nl * N * 8, &pc2) || ws_get(c, WS_ACC, (size_t)cap * per_ct, &pacc)) return -1;
    kb.c2 = (uint64_t *)pc2;
    kb.acc = (uint64_t *)pacc;
    kb.acc_cap = cap;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// (1) baby-step rotation cache  R[k][row = 2i+c][l < L][N]  for the K entries whose block row is in [bi_lo, bi_hi)
//     gwas/matmult.go:1083-1119 : rotCache[i][baby] = RotateRightWithEvaluator(A[i][bi], -baby)
//     ONE batch over every (baby step, i, bi): each entry carries its own Galois key; INTT(c1) is shared by all baby steps.
// ---------------------------------------------------------------------------------------------------------------
// k_lo / k_hi: only the entries whose position in klist lies in [k_lo, k_hi) are computed (baby-step sharding: every rank rotates its share
// and the shares are all-gathered); R_ext: write into this buffer (laid out like the workspace one) instead of the context's workspace
static int build_rot_cache(Ctx *c, const Cache *ca, const uint64_t *d_A, int s, int levelA, int bi_lo, int bi_hi,
                           std::vector<int> &klist, void **R_out, int k_lo = 0, int k_hi = 0x7fffffff, void *R_ext = nullptr) {
    const int d = ca->d, nbr = ca->nbr, N = c->N, nlA = levelA + 1, nrows = 2 * s;
    klist.clear();
    std::vector<int> klocal((size_t)nbr * d, -1);
    for (size_t k = 0; k < ca->kbi.size(); k++)
        if (ca->kbi[k] >= bi_lo && ca->kbi[k] < bi_hi) {
            klocal[(size_t)ca->kbi[k] * d + ca->kb[k]] = (int)klist.size();
            klist.push_back((int)k);
        }
    const size_t RB = (size_t)ca->lay.bytes;  // bytes of one (k, row) record
    void *R = R_ext;
    if (!R && ws_get(c, WS_R, std::max<size_t>(klist.size(), 1) * nrows * RB, &R)) return -1;
    *R_out = R;
    const size_t ctA = (size_t)2 * nlA * N;
    RotMeta rot, cpy;
    std::vector<int> slot_of((size_t)s * nbr, -1);
    for (int b = 0; b < d; b++) {
        const GaloisKey *key = nullptr;
        if (b > 0) {
            bool any = false;
            for (int bi = bi_lo; bi < bi_hi; bi++) any |= ca->baby[(size_t)bi * d + b] != 0;
            if (!any) continue;
            if (find_key(c, b, &key)) return -1;
        }
        for (int bi = bi_lo; bi < bi_hi; bi++) {
            if (!ca->baby[(size_t)bi * d + b]) continue;
            if (klocal[(size_t)bi * d + b] < k_lo || klocal[(size_t)bi * d + b] >= k_hi) continue;
            for (int i = 0; i < s; i++) {
                const long long in_off = (long long)(((size_t)i * nbr + bi) * ctA);
                const long long out_off = (long long)(((size_t)klocal[(size_t)bi * d + b] * nrows + 2 * i) * RB);
                if (b == 0) {
                    cpy.add(RotEntry{in_off, out_off, 0, nullptr});
                } else {
                    int &sl = slot_of[(size_t)i * nbr + bi];
                    if (sl < 0) {
                        sl = (int)rot.c2_src.size();
                        rot.c2_src.push_back(in_off);
                    }
                    rot.add(RotEntry{in_off, out_off, sl, key});
                }
            }
        }
    }
    KsBatch kb{};
    // A above maxLevel is dropped to maxLevel (only limbs 0..maxLevel are read, crypto/basics.go:806-824); A at maxLevel-1 is used as it
    // is (gwas/matmult.go:1053 drops only when Level() > maxLevel; the accumulators read limbs 0..maxLevel-1, which it still has)
    kb.level = std::min(levelA, ca->maxLevel);
    kb.in = d_A;
    kb.in_nl = nlA;
    kb.out = R;
    kb.out_layout = ca->lay;  // limb index maxLevel of the rotated ct is never read by the MAC (App. A.4)
    kb.accumulate = false;
    if (!cpy.in_off.empty()) {
        if (cpy.upload(c, WS_META2)) return -1;
        kb.nct = (int)cpy.in_off.size();
        kb.in_off = cpy.d_in_off;
        kb.out_off = cpy.d_out_off;
        if (launch_copy_add(c, kb, c->stream)) return -1;
    }
    if (!rot.in_off.empty()) {
        if (rot.upload(c, WS_META)) return -1;
        kb.nct = (int)rot.in_off.size();
        kb.in_off = rot.d_in_off;
        kb.out_off = rot.d_out_off;
        kb.n_c2 = (int)rot.c2_src.size();
        kb.c2_src_off = rot.d_c2_src;
        kb.c2_slot = rot.d_c2_slot;
        kb.keys = rot.d_keys;
        kb.perms = rot.d_perms;
        if (fill_scratch(c, kb, kb.nct, kb.n_c2)) return -1;
        if (launch_rotate(c, kb, c->stream)) return -1;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// (2) MAC for the giant steps gact[gi_lo .. gi_hi) -> d_cv [(gi-gi_lo)*m_ct + bj][row][l][N]
//     gwas/matmult.go:1154-1168 (CPMultAccWithoutMRedV2) + :1203 (ModularReduceV2)
// ---------------------------------------------------------------------------------------------------------------
// One row part of the MAC: rows [row0, row0 + rows) of the 2s ciphertext polynomials with their own R image
struct MacPart {
    TcGeomR gr;
    unsigned char *rimg;
};

static int run_mac(Ctx *c, const Cache *ca, const void *R, const std::vector<int> &klist, int s, int gi_lo, int gi_hi,
                   uint64_t *d_cv) {
    const TcGeomP &tc = ca->tc;
    const int m_ct = ca->m_ct, rows_all = 2 * s, Kg = tc.Kg, K = tc.K;
    const int col_lo = gi_lo * m_ct, col_hi = gi_hi * m_ct;
    if (klist.empty() || col_hi <= col_lo) return 0;
    // When the accumulator region of all 2s rows exceeds half of TMEM the kernel would run single-buffered (MMA stream and epilogue do
    // not overlap).  Two row halves that each fit 256 columns are faster even though the P image is streamed twice (kp = 15 at logN 14:
    // 9 x 32 = 288 columns for 30 rows, 9 x 16 = 144 for 16).
    std::vector<std::pair<int, int>> ranges;  // (row0, rows)
    {
        TcGeomR gr, gh;
        if (tc_geom_r(c, tc, rows_all, &gr)) return -1;
        const int h = (rows_all / 2 + 1) / 2 * 2;  // even: a ciphertext's two polynomials stay together
        if (gr.tbuf_stride == 0 && rows_all >= 4 && getenv("SFG_MAC_NOSPLIT") == nullptr && !tc_geom_r(c, tc, h, &gh) && gh.tbuf_stride != 0) {
            ranges.push_back({0, h});
            ranges.push_back({h, rows_all - h});
        } else {
            ranges.push_back({0, rows_all});
        }
    }
    const int tile_lo = col_lo / 128, tile_hi = (col_hi + 127) / 128;
    // R images of the parts: source-record table [group][Kg][rows]; k outside klist (other block rows / padding) contributes zero
    std::vector<int> kpos((size_t)K, -1);
    for (size_t i = 0; i < klist.size(); i++) kpos[klist[i]] = (int)i;
    const size_t RB = (size_t)ca->lay.bytes;
    std::vector<MacPart> parts(ranges.size());
    size_t rimg_total = 0, tab_total = 0;
    for (size_t pi = 0; pi < ranges.size(); pi++) {
        // a second part takes the first one's row pitch: identical image geometry, so that the two can share the P stream (pair mode)
        if (tc_geom_r(c, tc, ranges[pi].second, &parts[pi].gr, pi ? parts[0].gr.RP : 0)) return -1;
        parts[pi].gr.cv_rows = rows_all;
        parts[pi].gr.cv_row0 = ranges[pi].first;
        rimg_total += (size_t)parts[pi].gr.group_bytes * tc.ngroups;
        tab_total += (size_t)tc.ngroups * Kg * ranges[pi].second;
    }
    std::vector<long long> tab(tab_total, -1);
    {
        size_t o = 0;
        for (size_t pi = 0; pi < ranges.size(); pi++) {
            const int row0 = ranges[pi].first, rows = ranges[pi].second;
            for (int grp = 0; grp < tc.ngroups; grp++)
                for (int kk = 0; kk < Kg; kk++) {
                    const int k = grp * Kg + kk;
                    if (k < K && kpos[k] >= 0)
                        for (int r = 0; r < rows; r++) tab[o + ((size_t)grp * Kg + kk) * rows + r] = (long long)(((size_t)kpos[k] * rows_all + row0 + r) * RB);
                }
            o += (size_t)tc.ngroups * Kg * rows;
        }
    }
    void *dtab, *rimg;
    if (ws_get(c, WS_RTAB, tab.size() * sizeof(long long), &dtab)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(dtab, tab.data(), tab.size() * sizeof(long long), cudaMemcpyDefault, c->stream));
    if (ws_get(c, WS_RIMG, rimg_total, &rimg)) return -1;
    {
        size_t o = 0, ro = 0;
        for (size_t pi = 0; pi < ranges.size(); pi++) {
            const int rows = ranges[pi].second;
            parts[pi].rimg = (unsigned char *)rimg + ro;
            for (int grp = 0; grp < tc.ngroups; grp++)
                if (launch_img_r(c, tc, parts[pi].gr, ca->lay, R, (const long long *)dtab + o + (size_t)grp * Kg * rows,
                                 parts[pi].rimg + (size_t)grp * parts[pi].gr.group_bytes, c->stream))
                    return -1;
            o += (size_t)tc.ngroups * Kg * rows;
            ro += (size_t)parts[pi].gr.group_bytes * tc.ngroups;
        }
    }
    // all K groups accumulate in TMEM inside one launch (many block rows: the transposed orientation, PCA shapes); only when their
    // products could overflow the s32 partial sums is the K range cut into several launches that add into cv mod q
    const int gf = tc_max_fused_groups(tc);
    // pimg holds the K groups [g_lo, g_hi) of some tiles; the first group of the whole K range overwrites cv, the others add into it
    auto mac = [&](const unsigned char *pimg, long long p_gstride, int img_ntiles, int img_tile0, int t0, int t1, int g_lo, int g_hi) -> int {
        if (g_tm) g_tm->mark(3);  // the MAC kernel alone, on the stream it is launched on (bench.py roofline)
        // Two row parts with the same image geometry (kp = 9 .. 16 with 5-6 byte planes: the 2 kp rows do not fit TMEM twice) can run as
        // ONE launch of 2-CTA clusters that share the P stream by multicast (pair mode, SFG_TC_PAIR=1) instead of two launches that each
        // read the whole image.  Bit-identical (test_pair_mode_shares_the_p_stream_bit_exact) and half the HBM traffic, but measured
        // NEUTRAL at logN 14 / kp = 15 (26.9 vs 26.8 ms, profiles/r2/mac14_pair_mode_probe.txt): the kernel is not HBM-bound there -- the
        // u64 epilogue of the two parts takes as long as the two passes over P, and the lock-stepped 4-stage ring of a pair tops out at
        // 20 ms without any epilogue.  Off by default until the epilogue gets more warps.
        const char *pe = getenv("SFG_TC_PAIR");
        bool pair = pe && *pe == '1' && parts.size() == 2 && parts[0].gr.RP == parts[1].gr.RP && parts[0].gr.group_bytes == parts[1].gr.group_bytes &&
                    parts[0].gr.tbuf_stride == parts[1].gr.tbuf_stride;
        for (int l = 0; pair && l < tc.L; l++) pair = parts[0].gr.npad[l] == parts[1].gr.npad[l] && parts[0].gr.rbase[l] == parts[1].gr.rbase[l];
        for (size_t pi = 0; pi < parts.size(); pi += pair ? 2 : 1) {
            const MacPart &pt = parts[pi];
            for (int grp = g_lo; grp < g_hi; grp += gf)
                if (launch_mac_tc(c, tc, pt.gr, pimg + (size_t)(grp - g_lo) * p_gstride, p_gstride, img_ntiles, img_tile0,
                                  pt.rimg + (size_t)grp * pt.gr.group_bytes, pt.gr.group_bytes, std::min(gf, g_hi - grp), t0, t1, col_lo, col_hi,
                                  grp > 0, d_cv, c->stream, pair ? parts[1].rimg + (size_t)grp * parts[1].gr.group_bytes : nullptr,
                                  pair ? parts[1].gr.cv_row0 : 0, pair ? parts[1].gr.rows : 0))
                    return -1;
        }
        if (g_tm) g_tm->mark(1);
        return 0;
    };
    if (ca->materialised) return mac(ca->img, tc.group_bytes, tc.ntiles, 0, tile_lo, tile_hi, 0, tc.ngroups);
    // diagonals regenerated on the fly: a temporary image of at most ~6 GiB (SFG_OTF_IMG_MB overrides: tests) holds a chunk of tiles
    // with all their K groups -- or, when one tile's groups alone exceed it (many block rows: PCA shapes), one tile and as many K
    // groups as fit.  Every diagonal is encoded ONCE and consumed by all row parts.
    long long img_budget = (long long)6 << 30;
    if (const char *e = getenv("SFG_OTF_IMG_MB")) img_budget = std::max(1LL, atoll(e)) << 20;
    const long long per_tg = tc_group_bytes(ca, 1), per_tile = per_tg * tc.ngroups;
    const int tchunk = (int)std::max<long long>(1, std::min<long long>(tile_hi - tile_lo, img_budget / per_tile));
    const int gchunk = per_tile <= img_budget ? tc.ngroups : (int)std::max<long long>(1, img_budget / per_tg);
    void *pimg;
    if (ws_get(c, WS_PIMG, (size_t)per_tg * tchunk * gchunk, &pimg)) return -1;
    for (int t0 = tile_lo; t0 < tile_hi; t0 += tchunk) {
        const int t1 = std::min(tile_hi, t0 + tchunk);
        for (int g0 = 0; g0 < tc.ngroups; g0 += gchunk) {
            const int g1 = std::min(tc.ngroups, g0 + gchunk);
            if (build_p_tiles(c, ca, t0, t1, g0, g1, (unsigned char *)pimg)) return -1;
            if (mac((const unsigned char *)pimg, tc_group_bytes(ca, t1 - t0), t1 - t0, t0, t0, t1, g0, g1)) return -1;
        }
    }
    return 0;
}

int cache_get_diag_dev(Ctx *c, const Cache *ca, int bi, int shift, int bj, uint64_t *d_out, int *present) {
    const int pi = ca->pidx[((size_t)bi * ca->slots + shift) * ca->m_ct + bj];
    *present = pi >= 0;
    if (pi < 0) return 0;
    if (!ca->materialised) SFG_FAIL(c, "cache is not materialised (diagonals are regenerated on the fly)");
    const int g = shift / ca->d, b = shift % ca->d;
    int gi = -1;
    for (size_t i = 0; i < ca->gact.size(); i++)
        if (ca->gact[i] == g) gi = (int)i;
    const int k = ca->kidx[(size_t)bi * ca->d + b];
    if (gi < 0 || k < 0) SFG_FAIL(c, "cache_get_diag: inconsistent activity tables");
    const int grp = k / ca->tc.Kg, kk = k % ca->tc.Kg;
    for (int l = 0; l < ca->L; l++)
        if (launch_img_extract(c, ca->tc, ca->img + (size_t)grp * ca->tc.group_bytes, l, gi * ca->m_ct + bj, kk, d_out + (size_t)l * c->N,
                               c->stream))
            return -1;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// (3) giant-step alignment + accumulation (gwas/matmult.go:1203-1227):
//     out[i][bj] += RotateRightWithEvaluator(cv[i][g][bj], -g*d)   for g in gact[gi_lo .. gi_hi)
//     one batch per giant step (entries of a batch must target distinct outputs), metadata uploaded once for all of them
// ---------------------------------------------------------------------------------------------------------------
static uint32_t inv_mod_pow2(uint64_t a, int bits) {  // a odd
    uint64_t x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - a * x;
    return (uint32_t)(x & ((1ULL << bits) - 1));
}

// Optional host destination of the result: rows of `out` are copied back on a second stream as soon as their giant-step sums are
// final, so the device -> host transfer of the reference-facing entry points overlaps the remaining key-switches.
struct HostSink {
    uint64_t *host = nullptr;    // [s][m_ct][2][L][N] (pinned or pageable host memory of the caller, or the pinned staging buffer)
    uint64_t *const *limbs = nullptr;  // optional: one destination pointer per limb (cgo: one Go slice each); `host` is then the staging buffer
    cudaStream_t copy = nullptr;
    cudaEvent_t ev = nullptr;
    bool used = false;           // set when run_giant issued the copies itself
    struct Piece {               // a range of `host` whose device -> host copy has been enqueued on `copy`
        size_t off, cnt;
        cudaEvent_t done;
    };
    std::vector<Piece> pieces;   // limbs != nullptr only: scattered to the limb pointers by the host while the GPU computes on
};
// device -> host of d_out[off, off + cnt) on the copy stream, after everything enqueued so far on the compute stream
static int sink_copy(Ctx *c, HostSink *sink, const uint64_t *d_out, size_t off, size_t cnt) {
    SFG_CUDA(c, cudaEventRecord(sink->ev, c->stream));
    SFG_CUDA(c, cudaStreamWaitEvent(sink->copy, sink->ev, 0));
    SFG_CUDA(c, cudaMemcpyAsync(sink->host + off, d_out + off, cnt * 8, cudaMemcpyDefault, sink->copy));
    if (sink->limbs) {
        cudaEvent_t done;
        SFG_CUDA(c, cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        SFG_CUDA(c, cudaEventRecord(done, sink->copy));
        sink->pieces.push_back(HostSink::Piece{off, cnt, done});
    }
    sink->used = true;
    return 0;
}
// scatter the finished pieces to the caller's limb pointers (blocks on each piece's copy, not on the compute stream)
static int sink_drain(Ctx *c, HostSink *sink) {
    const size_t N = c->N;
    cudaError_t e = cudaSuccess;
    for (auto &pc : sink->pieces) {
        if (e == cudaSuccess) e = cudaEventSynchronize(pc.done);
        if (e == cudaSuccess) scatter_host_to_limbs(sink->host + pc.off, sink->limbs + pc.off / N, pc.cnt / N, N);
        cudaEventDestroy(pc.done);
    }
    sink->pieces.clear();
    SFG_CUDA(c, e);
    return 0;
}

static int run_giant(Ctx *c, const Cache *ca, int s, const uint64_t *d_cv, int gi_lo, int gi_hi, uint64_t *d_out, HostSink *sink = nullptr) {
    const int d = ca->d, m_ct = ca->m_ct, L = ca->L, N = c->N, nrows = 2 * s;
    const size_t LN = (size_t)L * N;
    const int nout = m_ct * s;
    if (gi_hi <= gi_lo) return 0;
    // giant steps with a rotation, and the entries of giant step 0 (no rotation, crypto/basics.go:203)
    std::vector<int> grot;
    RotMeta cpy;
    for (int gi = gi_lo; gi < gi_hi; gi++) {
        const int g = ca->gact[gi];
        if (g > 0) {
            grot.push_back(gi);
            continue;
        }
        for (int t = 0; t < nout; t++) {  // t = bj*s + i : consecutive ciphertexts of the cv image
            const int bj = t / s, i = t % s;
            cpy.add(RotEntry{(long long)((((size_t)(gi - gi_lo) * m_ct + bj) * nrows + 2 * i) * LN), (long long)(((size_t)i * m_ct + bj) * 2 * LN * 8), 0, nullptr});
        }
    }
    const int nrot = (int)grot.size();
    KsBatch kb{};
    kb.level = L - 1;  // ModularReduceV2 creates the ct at level len(acc0)-1 (gwas/matmult.go:350)
    kb.in = d_cv;
    kb.in_nl = L;
    kb.out = d_out;
    kb.out_layout = make_layout(c, L, false);
    kb.accumulate = true;
    if (!cpy.in_off.empty()) {
        if (cpy.upload(c, WS_META2)) return -1;
        kb.nct = (int)cpy.in_off.size();
        kb.in_off = cpy.d_in_off;
        kb.out_off = cpy.d_out_off;
        if (launch_copy_add(c, kb, c->stream)) return -1;
    }
    if (nrot == 0) return 0;
    // Output-major schedule: the rows i of A are cut into chunks whose nrot * rows * m_ct rotations fit the key-switch scratch, so that a
    // chunk's sums over ALL giant steps finish in one launch sequence (no partial sums carried in HBM) and its rows of `out` are final
    // -- and can start their way to the host -- while the next chunk is computed.  If even one row does not fit, the giant steps of a
    // chunk are processed in groups with the partial sums carried (S1, C0, E).
    kb.nct = m_ct;
    kb.n_c2 = m_ct;
    if (fill_scratch(c, kb, nrot * nout, 1)) return -1;
    if (kb.acc_cap < m_ct) SFG_FAIL(c, "key-switch scratch too small for one giant step of one row (%d ciphertexts)", m_ct);
    const int rows_fit = kb.acc_cap / (nrot * m_ct);
    const int nchunk = (s + std::max(1, rows_fit) - 1) / std::max(1, rows_fit);
    const int rpc = (s + nchunk - 1) / nchunk;                                // rows of A per chunk, balanced (k_md_accum's grid = 10 * rows * m_ct
                                                                              // CTAs on 148 SMs: few large chunks quantise better than many small ones)
    const int G = rows_fit >= 1 ? nrot : std::max(1, kb.acc_cap / m_ct);      // giant steps per launch sequence
    if (fill_scratch(c, kb, std::min(G, nrot) * rpc * m_ct, std::min(G, nrot) * rpc * m_ct)) return -1;
    RotMeta rot;
    std::vector<size_t> chunk_base;
    for (int ch = 0; ch < nchunk; ch++) {
        const int i_lo = ch * rpc, i_hi = std::min(s, i_lo + rpc), nout_c = (i_hi - i_lo) * m_ct;
        chunk_base.push_back(rot.in_off.size());
        for (int a = 0; a < nrot; a++) {
            const int gi = grot[a], g = ca->gact[gi];
            const GaloisKey *key = nullptr;
            if (find_key(c, (g * d) % ca->slots, &key)) return -1;
            // galEl^-1 mod 2N (low 18 bits) and the rotation amount r, galEl = 5^r (k_md_accum: cyclic shift in discrete-log order)
            const uint32_t gin = inv_mod_pow2(key->galEl, c->logN + 1) | ((uint32_t)((g * d) % ca->slots) << 18);
            for (int o = 0; o < nout_c; o++) {  // o = (i - i_lo) * m_ct + bj: the chunk's outputs are contiguous in `out`
                const int i = i_lo + o / m_ct, bj = o % m_ct;
                const long long in_off = (long long)((((size_t)(gi - gi_lo) * m_ct + bj) * nrows + 2 * i) * LN);
                rot.add(RotEntry{in_off, (long long)(((size_t)i * m_ct + bj) * 2 * LN * 8), (int)(((size_t)(a % G) * nout_c + o)), key});
                rot.ginv.push_back(gin);
            }
        }
    }
    rot.c2_src = rot.in_off;
    if (rot.upload(c, WS_META)) return -1;
    void *mdbuf;
    const size_t msz = (size_t)rpc * m_ct * 2 * LN;
    if (ws_get(c, WS_MD, 3 * msz * 8, &mdbuf)) return -1;
    uint64_t *S1 = (uint64_t *)mdbuf, *C0 = S1 + msz, *E = C0 + msz;
    for (int ch = 0; ch < nchunk; ch++) {
        const int i_lo = ch * rpc, i_hi = std::min(s, i_lo + rpc), nout_c = (i_hi - i_lo) * m_ct;
        for (int a0 = 0; a0 < nrot; a0 += G) {
            const int na = std::min(G, nrot - a0);
            const size_t o = chunk_base[ch] + (size_t)a0 * nout_c;
            kb.nct = na * nout_c;
            kb.n_c2 = na * nout_c;
            kb.in_off = rot.d_in_off + o;
            kb.out_off = rot.d_out_off + o;
            kb.c2_src_off = rot.d_c2_src + o;
            kb.c2_slot = rot.d_c2_slot + o;
            kb.keys = rot.d_keys + o;
            kb.perms = rot.d_perms + o;
            if (launch_rotate_sum(c, kb, nout_c, rot.d_ginv + o, S1, C0, E, a0 == 0, c->stream)) return -1;
        }
        if (launch_rotate_sum_final(c, L - 1, nout_c, S1, C0, E, d_out, rot.d_out_off + chunk_base[ch], kb.out_layout, c->stream)) return -1;
        if (sink && sink->host) {  // rows [i_lo, i_hi) of out are final
            const size_t off = (size_t)i_lo * m_ct * 2 * LN, cnt = (size_t)(i_hi - i_lo) * m_ct * 2 * LN;
            if (sink_copy(c, sink, d_out, off, cnt)) return -1;
        }
    }
    return 0;
}

static int check_args(Ctx *c, const Cache *ca, int s, int nbr, int levelA, int maxLevel) {
    if (s < 1) SFG_FAIL(c, "A has no rows");
    if (nbr != ca->nbr) SFG_FAIL(c, "A has %d block rows but the genotype matrix has %d", nbr, ca->nbr);
    if (maxLevel != ca->maxLevel) SFG_FAIL(c, "maxLevel %d differs from the cache's %d", maxLevel, ca->maxLevel);
    // gwas/matmult.go:1053-1056 drops A only when Level() > maxLevel; an input at maxLevel-1 still has the maxLevel limbs the accumulators
    // read (:231-245, :393) and goes through unchanged; below that the reference indexes past Coeffs[] and panics
    if (levelA < maxLevel - 1) SFG_FAIL(c, "input level %d has fewer than the %d limbs the accumulators read (index out of range)", levelA, maxLevel);
    if (levelA > c->nQ - 1) SFG_FAIL(c, "input level %d exceeds the parameter chain", levelA);
    return 0;
}
static int check_rows(Ctx *c, const Cache *ca, int bi_lo, int bi_hi) {
    if (bi_lo < ca->bi_lo || bi_hi > ca->bi_hi)
        SFG_FAIL(c, "the cache holds block rows [%d, %d) only (built for block-row sharding); block rows [%d, %d) were requested", ca->bi_lo, ca->bi_hi,
                 bi_lo, bi_hi);
    return 0;
}

// rows of A per pass: the tensor-core MAC stacks the byte planes of the 2s ciphertext polynomials as one MMA operand (N <= 256 columns,
// (2 nb - 1) RP <= 512 TMEM columns), which holds for 16 rows with every modulus below 2^48
constexpr int kMaxRowsPerPass = 16;

// one pass over rows [0, s) of d_A / d_out (s <= kMaxRowsPerPass); ms accumulates the phase timings
// R_ext (optional): the rotation cache was computed beforehand (baby-step sharding: sfg_matmult4_baby_dev + all-gather); d_A is then unused
static int mm_compute_rows(Ctx *c, const uint64_t *d_A, int s, int nbr, int levelA, Cache *ca, uint64_t *d_out, HostSink *sink, float ms[5],
                           const void *R_ext = nullptr) {
    const int L = ca->L, N = c->N, m_ct = ca->m_ct;
    const size_t LN = (size_t)L * N;
    PhaseTimer tm(c->stream);
    g_tm = &tm;
    struct Reset { ~Reset() { g_tm = nullptr; } } reset;
    void *R;
    std::vector<int> klist;
    tm.mark(0);
    if (R_ext) {
        R = const_cast<void *>(R_ext);
        for (size_t k = 0; k < ca->kbi.size(); k++) klist.push_back((int)k);
    } else if (build_rot_cache(c, ca, d_A, s, levelA, 0, nbr, klist, &R)) {
        return -1;
    }
    SFG_CUDA(c, cudaMemsetAsync(d_out, 0, (size_t)s * m_ct * 2 * LN * 8, c->stream));
    // giant chunks bounded by the cv image size (default 24 GiB)
    const size_t per_g = (size_t)m_ct * 2 * s * LN * 8;
    const int ng = (int)ca->gact.size();
    int gchunk = ng;
    if ((size_t)ng * per_g > c->ws[WS_CV].bytes) {
        size_t fr = 0, tot = 0;
        SFG_CUDA(c, cudaMemGetInfo(&fr, &tot));
        const size_t cv_budget = std::min<size_t>((size_t)24 << 30, (fr + c->ws[WS_CV].bytes) / 3);
        gchunk = std::min(ng, (int)std::max<size_t>(1, cv_budget / per_g));
    }
    void *cv;
    if (ws_get(c, WS_CV, (size_t)gchunk * per_g, &cv)) return -1;
    bool copied = false;
    for (int g0 = 0; g0 < ng; g0 += gchunk) {
        const int g1 = std::min(ng, g0 + gchunk);
        tm.mark(1);
        if (run_mac(c, ca, R, klist, s, g0, g1, (uint64_t *)cv)) return -1;
        tm.mark(2);
        // the rows of out can leave for the host chunk by chunk only when this pass sees every giant step at once
        HostSink *sk = (sink && g0 == 0 && g1 == ng) ? sink : nullptr;
        if (sk) sk->used = false;
        if (run_giant(c, ca, s, (const uint64_t *)cv, g0, g1, d_out, sk)) return -1;
        copied = sk && sk->used;
    }
    if (sink && !copied && sink_copy(c, sink, d_out, 0, (size_t)s * m_ct * 2 * LN)) return -1;  // several giant chunks / nothing to rotate
    tm.mark(-1);
    // per-limb destinations: the host scatters every piece as soon as its copy has landed, while the GPU works on the following rows
    if (sink && sink->limbs && sink_drain(c, sink)) return -1;
    float t[5];
    tm.finish(t);
    for (int i = 0; i < 5; i++) ms[i] += t[i];
    return 0;
}

// host_out: the result also goes to this HOST buffer; out_limbs: ... or to one host pointer per limb (through the pinned staging
// buffer).  Either way rows leave as soon as their giant-step sums are final, overlapped with the remaining key-switches.
int mm_compute_dev(Ctx *c, const uint64_t *d_A, int s, int nbr, int levelA, int maxLevel, Cache *ca, uint64_t *d_out, uint64_t *host_out,
                   uint64_t *const *out_limbs) {
    if (check_args(c, ca, s, nbr, levelA, maxLevel) || check_rows(c, ca, 0, nbr)) return -1;
    SFG_CUDA(c, cudaSetDevice(c->device));
    const int L = ca->L, N = c->N, m_ct = ca->m_ct;
    const size_t LN = (size_t)L * N, ctA = (size_t)2 * (levelA + 1) * N, row_out = (size_t)m_ct * 2 * LN;
    HostSink sink;
    sink.host = host_out;
    sink.limbs = out_limbs;
    if (out_limbs) {
        void *pin;
        if (pinned_get(c, PIN_OUT, (size_t)s * row_out * 8, &pin)) return -1;
        sink.host = (uint64_t *)pin;
    }
    const bool to_host = sink.host != nullptr;
    if (to_host) {
        SFG_CUDA(c, cudaStreamCreateWithFlags(&sink.copy, cudaStreamNonBlocking));
        SFG_CUDA(c, cudaEventCreateWithFlags(&sink.ev, cudaEventDisableTiming));
    }
    // The reference has no limit on len(A) (assoc.go:700-718 passes len(Q) + 2 rows); rows are independent, so more than
    // kMaxRowsPerPass rows run as balanced passes over row ranges.
    const int npass = (s + kMaxRowsPerPass - 1) / kMaxRowsPerPass, rpp = (s + npass - 1) / npass;
    float ms[5] = {0, 0, 0, 0, 0};
    int rc = 0;
    for (int i0 = 0; i0 < s && !rc; i0 += rpp) {
        const int si = std::min(rpp, s - i0);
        HostSink part = sink;  // views of the same streams, offset to this pass's rows
        part.pieces.clear();
        if (part.host) part.host += (size_t)i0 * row_out;
        if (part.limbs) part.limbs += (size_t)i0 * row_out / N;
        rc = mm_compute_rows(c, d_A + (size_t)i0 * nbr * ctA, si, nbr, levelA, ca, d_out + (size_t)i0 * row_out, to_host ? &part : nullptr, ms);
        for (auto &pc : part.pieces) cudaEventDestroy(pc.done);  // only non-empty after an error
    }
    for (int i = 0; i < 5; i++) g_last_ms[i] = ms[i];
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (sink.copy) {
        if (e == cudaSuccess) e = cudaStreamSynchronize(sink.copy);
        cudaStreamDestroy(sink.copy);
        cudaEventDestroy(sink.ev);
    }
    if (rc) return -1;
    SFG_CUDA(c, e);
    return 0;
}

// Baby-step sharding (strong scaling on top of the giant-step sharding: the baby rotations are the part every rank would otherwise
// repeat): share `part` of `nparts` of the K rotation-cache entries, written at their global positions of d_R ([K_pad][2s][record]).
size_t mm_baby_chunk_bytes(const Cache *ca, int s, int nparts) {
    const size_t K = ca->kbi.size(), per = (K + nparts - 1) / nparts;
    return per * 2 * s * (size_t)ca->lay.bytes;
}
int mm_baby_dev(Ctx *c, const uint64_t *d_A, int s, int nbr, int levelA, int maxLevel, Cache *ca, int part, int nparts, void *d_R) {
    if (check_args(c, ca, s, nbr, levelA, maxLevel) || check_rows(c, ca, 0, nbr)) return -1;
    if (s > kMaxRowsPerPass) SFG_FAIL(c, "s = %d ciphertext rows: the sharded pieces take at most %d per call (split A by rows)", s, kMaxRowsPerPass);
    if (nparts < 1 || part < 0 || part >= nparts) SFG_FAIL(c, "baby-step share %d of %d", part, nparts);
    SFG_CUDA(c, cudaSetDevice(c->device));
    const int K = (int)ca->kbi.size(), per = (K + nparts - 1) / nparts;
    PhaseTimer tm(c->stream);
    tm.mark(0);
    std::vector<int> klist;
    void *R;
    if (build_rot_cache(c, ca, d_A, s, levelA, 0, nbr, klist, &R, part * per, std::min(K, (part + 1) * per), d_R)) return -1;
    tm.mark(-1);
    tm.finish(g_last_ms);
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}
// the rest of MatMult4StreamCompute on a complete rotation cache d_R (after the all-gather of the shares)
int mm_compute_r_dev(Ctx *c, const void *d_R, int s, int maxLevel, Cache *ca, uint64_t *d_out) {
    if (check_args(c, ca, s, ca->nbr, maxLevel, maxLevel) || check_rows(c, ca, 0, ca->nbr)) return -1;
    if (s > kMaxRowsPerPass) SFG_FAIL(c, "s = %d ciphertext rows: the sharded pieces take at most %d per call (split A by rows)", s, kMaxRowsPerPass);
    SFG_CUDA(c, cudaSetDevice(c->device));
    float ms[5] = {0, 0, 0, 0, 0};
    if (mm_compute_rows(c, nullptr, s, ca->nbr, maxLevel, ca, d_out, nullptr, ms, d_R)) return -1;
    for (int i = 0; i < 5; i++) g_last_ms[i] = ms[i];
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int mm_partial_dev(Ctx *c, const uint64_t *d_A, int s, int nbr, int levelA, int maxLevel, Cache *ca, int bi_lo, int bi_hi,
                   uint64_t *d_cv) {
    if (check_args(c, ca, s, nbr, levelA, maxLevel)) return -1;
    if (s > kMaxRowsPerPass) SFG_FAIL(c, "s = %d ciphertext rows: the sharded pieces take at most %d per call (split A by rows)", s, kMaxRowsPerPass);
    if (bi_lo < 0 || bi_hi > nbr || bi_lo > bi_hi) SFG_FAIL(c, "block-row range [%d, %d) out of [0, %d)", bi_lo, bi_hi, nbr);
    if (bi_lo < bi_hi && check_rows(c, ca, bi_lo, bi_hi)) return -1;
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t LN = (size_t)ca->L * c->N;
    const size_t total = ca->gact.size() * (size_t)ca->m_ct * 2 * s * LN;
    PhaseTimer tm(c->stream);
    g_tm = &tm;
    struct Reset { ~Reset() { g_tm = nullptr; } } reset;
    void *R;
    std::vector<int> klist;
    tm.mark(0);
    if (build_rot_cache(c, ca, d_A, s, levelA, bi_lo, bi_hi, klist, &R)) return -1;
    tm.mark(1);
    if (klist.empty()) {
        SFG_CUDA(c, cudaMemsetAsync(d_cv, 0, total * 8, c->stream));
    } else if (run_mac(c, ca, R, klist, s, 0, (int)ca->gact.size(), d_cv)) {
        return -1;
    }
    tm.mark(-1);
    tm.finish(g_last_ms);
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int mm_finish_dev(Ctx *c, Cache *ca, int s, int maxLevel, const uint64_t *d_cv, int g_lo, int g_hi, uint64_t *d_out) {
    if (maxLevel != ca->maxLevel) SFG_FAIL(c, "maxLevel mismatch");
    if (s < 1 || s > kMaxRowsPerPass) SFG_FAIL(c, "s = %d ciphertext rows: the sharded pieces take 1..%d per call", s, kMaxRowsPerPass);
    const int ng = (int)ca->gact.size();
    if (g_lo < 0 || g_hi > ng || g_lo > g_hi) SFG_FAIL(c, "giant range [%d, %d) out of [0, %d)", g_lo, g_hi, ng);
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t LN = (size_t)ca->L * c->N;
    PhaseTimer tm(c->stream);
    SFG_CUDA(c, cudaMemsetAsync(d_out, 0, (size_t)s * ca->m_ct * 2 * LN * 8, c->stream));
    tm.mark(2);
    const size_t per_g = (size_t)ca->m_ct * 2 * s * LN;
    if (run_giant(c, ca, s, d_cv + (size_t)g_lo * per_g, g_lo, g_hi, d_out)) return -1;
    tm.mark(-1);
    tm.finish(g_last_ms);
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// crypto/basics.go:201-210
int rotate_right_dev(Ctx *c, int level, const uint64_t *d_in, int nct, int nrot, uint64_t *d_out) {
    SFG_CUDA(c, cudaSetDevice(c->device));
    const int slots = c->slots, N = c->N, nl = level + 1;
    nrot %= slots;
    if (nrot < 0) nrot += slots;
    const size_t ct = (size_t)2 * nl * N;
    const GaloisKey *key = nullptr;
    if (nrot != 0 && find_key(c, slots - nrot, &key)) return -1;  // RotateNew(ct, slots - nrot): left rotation
    RotMeta rot;
    for (int t = 0; t < nct; t++) rot.add(RotEntry{(long long)(t * ct), (long long)(t * ct * 8), t, key});
    rot.c2_src = rot.in_off;
    if (rot.upload(c, WS_META)) return -1;
    KsBatch kb{};
    kb.level = level;
    kb.nct = nct;
    kb.in = d_in;
    kb.in_off = rot.d_in_off;
    kb.in_nl = nl;
    kb.n_c2 = nct;
    kb.c2_src_off = rot.d_c2_src;
    kb.c2_slot = rot.d_c2_slot;
    kb.keys = rot.d_keys;
    kb.perms = rot.d_perms;
    kb.out = d_out;
    kb.out_off = rot.d_out_off;
    kb.out_layout = make_layout(c, nl, false);
    kb.accumulate = false;
    if (nrot == 0) {
        if (launch_copy_add(c, kb, c->stream)) return -1;
    } else {
        if (fill_scratch(c, kb, nct, nct)) return -1;
        if (launch_rotate(c, kb, c->stream)) return -1;
    }
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

int encode_diag_host(Ctx *c, const Geno *g, int bi, int shift, int nrot, int level, bool mont, uint64_t *out, uint8_t *present,
                     int64_t *coeffs) {
    SFG_CUDA(c, cudaSetDevice(c->device));
    const int slots = c->slots, N = c->N, nl = level + 1;
    const int m_ct = (int)((g->ncols - 1) / slots) + 1, nbr = (int)((g->nrows - 1) / slots) + 1;
    if (bi < 0 || bi >= nbr || shift < 0 || shift >= slots) SFG_FAIL(c, "encode_diag: block row %d / shift %d out of range", bi, shift);
    if (level < 0 || level >= c->nQ) SFG_FAIL(c, "encode_diag: level %d out of range", level);
    const int nr = (int)std::min<size_t>((size_t)(bi + 1) * slots, g->nrows) - bi * slots;
    std::vector<EncJob> jobs;
    std::vector<int> which;
    for (int bj = 0; bj < m_ct; bj++) {
        const int nc = (int)std::min<size_t>((size_t)(bj + 1) * slots, g->ncols) - bj * slots;
        present[bj] = diag_exists(nr, nc, slots, shift) ? 1 : 0;
        if (present[bj]) {
            jobs.push_back(EncJob{bi * slots, bj * slots, nr, nc, shift, ((nrot % slots) + slots) % slots, (long long)(jobs.size() * (size_t)nl * N * 8)});
            which.push_back(bj);
        }
    }
    if (jobs.empty()) return 0;
    Buf dj, dout, dco;
    if (dj.alloc(c, jobs.size() * sizeof(EncJob)) || dout.alloc(c, jobs.size() * (size_t)nl * N * 8)) return -1;
    if (coeffs && dco.alloc(c, jobs.size() * (size_t)N * 8)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(dj.p, jobs.data(), jobs.size() * sizeof(EncJob), cudaMemcpyDefault, c->stream));
    if (launch_encode(c, g->d, g->ncols, dj.as<EncJob>(), (int)jobs.size(), make_layout(c, nl, false), mont, dout.p, coeffs ? dco.as<long long>() : nullptr, c->stream)) return -1;
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    for (size_t k = 0; k < jobs.size(); k++) {
        SFG_CUDA(c, cudaMemcpy(out + (size_t)which[k] * nl * N, dout.as<uint64_t>() + k * (size_t)nl * N, (size_t)nl * N * 8, cudaMemcpyDefault));
        if (coeffs) SFG_CUDA(c, cudaMemcpy(coeffs + (size_t)which[k] * N, dco.as<long long>() + k * (size_t)N, (size_t)N * 8, cudaMemcpyDefault));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Ciphertext algebra of the callers around the path: the device side of crypto.CMult / CMultScalar /
// MaskTrunc / InnerSumAll / CSub as QXLazyNormStream and QXtLazyNormStream use them (gwas/matmult.go:27-116).
// All pointers are device pointers; ct k of an operand is [2][*_nl][N] at base + k * 2 * *_nl * N (count 1 = broadcast).
// ---------------------------------------------------------------------------------------------------------------

// evaluator.Rescale, `times` steps of ring.DivRoundByLastModulusNTT: d_in nct cts at `level` -> d_out nct cts at level - times
int rescale_dev(Ctx *c, int level, const uint64_t *d_in, int nct, int times, uint64_t *d_out) {
    SFG_CUDA(c, cudaSetDevice(c->device));
    if (times < 0 || level - times < 0) SFG_FAIL(c, "rescale: cannot drop %d levels from level %d", times, level);
    const size_t N = c->N;
    const int npoly = 2 * nct;
    if (times == 0) {
        SFG_CUDA(c, cudaMemcpyAsync(d_out, d_in, (size_t)npoly * (level + 1) * N * 8, cudaMemcpyDefault, c->stream));
        return 0;
    }
    Buf T, U, ping, pong;
    if (T.alloc(c, (size_t)npoly * N * 8) || U.alloc(c, (size_t)npoly * level * N * 8)) return -1;
    if (times > 1 && (ping.alloc(c, (size_t)npoly * level * N * 8) || pong.alloc(c, (size_t)npoly * level * N * 8))) return -1;
    const uint64_t *src = d_in;
    for (int t = 0; t < times; t++) {
        uint64_t *dst = (t == times - 1) ? d_out : ((t & 1) ? pong.as<uint64_t>() : ping.as<uint64_t>());
        if (launch_rescale(c, level - t, src, npoly, dst, T.as<uint64_t>(), U.as<uint64_t>(), c->stream)) return -1;
        src = dst;
    }
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// evaluator.MulRelinNew on n = max(nx, ny) pairs (a side of 1 is broadcast, crypto/basics.go:386-427) + `times` rescale steps
int mul_relin_dev(Ctx *c, int level, const uint64_t *d_x, int nx, int x_nl, const uint64_t *d_y, int ny, int y_nl, int times, uint64_t *d_out) {
    SFG_CUDA(c, cudaSetDevice(c->device));
    const int N = c->N, nl = level + 1, n = std::max(nx, ny);
    if (level < 0 || level >= c->nQ || x_nl < nl || y_nl < nl) SFG_FAIL(c, "mul_relin: level %d needs %d limbs (operands have %d / %d)", level, nl, x_nl, y_nl);
    if (!((nx == n || nx == 1) && (ny == n || ny == 1)) || n < 1) SFG_FAIL(c, "mul_relin: %d x %d ciphertexts do not broadcast", nx, ny);
    const GaloisKey *rlk = nullptr;
    if (find_key(c, 0, &rlk)) SFG_FAIL(c, "relinearisation key not loaded (sfg_ctx_set_relin_key)");
    const size_t ct = (size_t)2 * nl * N;
    Buf tmp, prod;
    if (tmp.alloc(c, n * ct * 8)) return -1;
    uint64_t *dprod = d_out;
    if (times > 0) {
        if (prod.alloc(c, n * ct * 8)) return -1;
        dprod = prod.as<uint64_t>();
    }
    if (launch_ct_tensor(c, d_x, nx == 1 ? 0 : (long long)2 * x_nl * N, x_nl, d_y, ny == 1 ? 0 : (long long)2 * y_nl * N, y_nl, nl, n,
                         tmp.as<uint64_t>(), dprod, c->stream))
        return -1;
    // relinearise: key-switch d2 with rlk (galEl 1: the identity permutation) and accumulate (d0 + ks0, ks1) into (0, d1)
    RotMeta rot;
    for (int t = 0; t < n; t++) rot.add(RotEntry{(long long)(t * ct), (long long)(t * ct * 8), t, rlk});
    rot.c2_src = rot.in_off;
    if (rot.upload(c, WS_META)) return -1;
    KsBatch kb{};
    kb.level = level;
    kb.nct = n;
    kb.in = tmp.as<uint64_t>();
    kb.in_off = rot.d_in_off;
    kb.in_nl = nl;
    kb.n_c2 = n;
    kb.c2_src_off = rot.d_c2_src;
    kb.c2_slot = rot.d_c2_slot;
    kb.keys = rot.d_keys;
    kb.perms = rot.d_perms;
    kb.out = dprod;
    kb.out_off = rot.d_out_off;
    kb.out_layout = make_layout(c, nl, false);
    kb.accumulate = true;
    if (fill_scratch(c, kb, n, n)) return -1;
    if (launch_rotate(c, kb, c->stream)) return -1;
    if (times > 0) return rescale_dev(c, level, dprod, n, times, d_out);
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// evaluator.MulRelinNew(plaintext, ct) (crypto.MaskTrunc, crypto/basics.go:110-127) + `times` rescale steps
int mul_plain_dev(Ctx *c, int level, const uint64_t *d_pt, int npt, int pt_nl, const uint64_t *d_ct, int nct, int ct_nl, int times, uint64_t *d_out) {
    SFG_CUDA(c, cudaSetDevice(c->device));
    const int N = c->N, nl = level + 1;
    if (level < 0 || level >= c->nQ || pt_nl < nl || ct_nl < nl) SFG_FAIL(c, "mul_plain: level %d needs %d limbs (operands have %d / %d)", level, nl, pt_nl, ct_nl);
    if (!(npt == nct || npt == 1) || nct < 1) SFG_FAIL(c, "mul_plain: %d plaintexts x %d ciphertexts do not broadcast", npt, nct);
    Buf prod;
    uint64_t *dprod = d_out;
    if (times > 0) {
        if (prod.alloc(c, (size_t)nct * 2 * nl * N * 8)) return -1;
        dprod = prod.as<uint64_t>();
    }
    if (launch_pt_mul(c, d_pt, npt == 1 ? 0 : (long long)pt_nl * N, d_ct, (long long)2 * ct_nl * N, ct_nl, nl, nct, dprod, c->stream)) return -1;
    if (times > 0) return rescale_dev(c, level, dprod, nct, times, d_out);
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// evaluator.Add / Sub, operands of matching scale: n = max(na, nb) results at `level`
int addsub_dev(Ctx *c, int level, const uint64_t *d_a, int na, int a_nl, const uint64_t *d_b, int nb, int b_nl, bool sub, uint64_t *d_out) {
    SFG_CUDA(c, cudaSetDevice(c->device));
    const int N = c->N, nl = level + 1, n = std::max(na, nb);
    if (level < 0 || level >= c->nQ || a_nl < nl || b_nl < nl) SFG_FAIL(c, "ct add/sub: level %d needs %d limbs (operands have %d / %d)", level, nl, a_nl, b_nl);
    if (!((na == n || na == 1) && (nb == n || nb == 1)) || n < 1) SFG_FAIL(c, "ct add/sub: %d and %d ciphertexts do not broadcast", na, nb);
    if (launch_ct_addsub(c, d_a, na == 1 ? 0 : (long long)2 * a_nl * N, a_nl, d_b, nb == 1 ? 0 : (long long)2 * b_nl * N, b_nl, nl, n, sub, d_out,
                         c->stream))
        return -1;
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// crypto.InnerSumAll (crypto/basics.go:278-293) for nvec vectors of cnt ciphertexts each (all [2][level+1][N]):
// out[v] = sum of the vector, then out[v] += RotL_r(out[v]) for r = 1, 2, 4, .. < slots (RotateAndAdd :236-246); the nvec rotations of one
// round go through one batched key-switch.
int inner_sum_all_dev(Ctx *c, int level, const uint64_t *d_in, int nvec, int cnt, uint64_t *d_out) {
    SFG_CUDA(c, cudaSetDevice(c->device));
    const int N = c->N, nl = level + 1;
    if (level < 0 || level >= c->nQ || nvec < 1 || cnt < 1) SFG_FAIL(c, "inner_sum_all: bad level / counts");
    Buf rt;
    if (rt.alloc(c, (size_t)nvec * 2 * nl * N * 8)) return -1;
    if (launch_ct_sum(c, d_in, nvec, cnt, nl, d_out, c->stream)) return -1;
    const long long ct = (long long)2 * nl * N;
    for (int r = 1; r < c->slots; r *= 2) {
        if (rotate_right_dev(c, level, d_out, nvec, c->slots - r, rt.as<uint64_t>())) return -1;  // RotateNew(ct, r): left rotation by r
        if (launch_ct_addsub(c, rt.as<uint64_t>(), ct, nl, d_out, ct, nl, nl, nvec, false, d_out, c->stream)) return -1;
    }
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// EncodeNTT of an int8 slot vector (e.g. the 0/1 mask of crypto.MaskTrunc) with the diagonal encoder: correctly rounded
// scale * sigma^-1(v), NTT domain, limbs 0..level, plain or Montgomery form.
int encode_slots_host(Ctx *c, const int8_t *v, int level, bool mont, uint64_t *out) {
    SFG_CUDA(c, cudaSetDevice(c->device));
    const int slots = c->slots, N = c->N, nl = level + 1;
    if (level < 0 || level >= c->nQ) SFG_FAIL(c, "encode_slots: level %d out of range", level);
    Buf dv, dj, dout;
    if (dv.alloc(c, slots) || dj.alloc(c, sizeof(EncJob)) || dout.alloc(c, (size_t)nl * N * 8)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(dv.p, v, slots, cudaMemcpyDefault, c->stream));
    const EncJob job{0, 0, slots, slots, 0, 0, 0};  // diagonal 0 of a block whose every row is v (leading dimension 0)
    SFG_CUDA(c, cudaMemcpyAsync(dj.p, &job, sizeof job, cudaMemcpyDefault, c->stream));
    if (launch_encode(c, dv.as<int8_t>(), 0, dj.as<EncJob>(), 1, make_layout(c, nl, false), mont, dout.p, nullptr, c->stream)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(out, dout.p, (size_t)nl * N * 8, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// The reference's on-disk diagonal cache (gwas/filestream.go:19-282, App. D.2): one file `<prefix>_<bi>.bin` per
// block row = header {vectorLen = m_ct, level = maxLevel, scale, n = N, numModuli = maxLevel+1, rowSize} (6 x u64 LE) + baby / giant
// tables (d bytes each) + records {u64 LE length, u32 LE shift, per block column: u8 isEmpty [, numModuli x N BIG-endian u64]}.
// cache_write_files lets a CPU run of the reference consume a GPU preprocess; cache_load_files builds the HBM image from files the
// reference wrote (records may come in any order: the reference's writer receives them from nproc goroutines).
// ---------------------------------------------------------------------------------------------------------------
static size_t env_size(const char *name, size_t dflt) {
    const char *e = getenv(name);
    return (e && *e) ? (size_t)strtoull(e, nullptr, 10) : dflt;
}

int cache_write_files(Ctx *c, const Cache *ca, const char *prefix) {
    if (!ca->g) SFG_FAIL(c, "cache_write_files: this cache was loaded from files (no genotype matrix to encode limb %d from)", ca->maxLevel);
    SFG_CUDA(c, cudaSetDevice(c->device));
    const int slots = ca->slots, d = ca->d, m_ct = ca->m_ct, N = c->N, nlF = ca->maxLevel + 1;
    const size_t polyB = (size_t)nlF * N * 8;
    const PolyLayout layF = make_layout(c, nlF, false);
    const size_t cap = std::max<size_t>((size_t)m_ct, env_size("SFG_CACHEFILE_CHUNK_POLYS", ((size_t)512 << 20) / polyB));
    Buf dbuf, djobs;
    if (dbuf.alloc(c, cap * polyB) || djobs.alloc(c, cap * sizeof(EncJob))) return -1;
    std::vector<unsigned char> host(cap * polyB);
    for (int bi = 0; bi < ca->nbr; bi++) {
        const std::string fn = std::string(prefix) + "_" + std::to_string(bi) + ".bin";
        FILE *f = fopen(fn.c_str(), "wb");
        if (!f) SFG_FAIL(c, "create %s: %s", fn.c_str(), strerror(errno));
        uint64_t hdr[6] = {(uint64_t)m_ct, (uint64_t)ca->maxLevel, 0, (uint64_t)N, (uint64_t)nlF, 4 + (1 + (uint64_t)polyB) * (uint64_t)m_ct};
        memcpy(&hdr[2], &c->scale, 8);
        bool ok = fwrite(hdr, 8, 6, f) == 6 && fwrite(&ca->baby[(size_t)bi * d], 1, d, f) == (size_t)d && fwrite(&ca->giant[(size_t)bi * d], 1, d, f) == (size_t)d;
        int shift = 0;
        while (ok && shift < slots) {
            // batch of consecutive active shifts whose polynomials fit the staging buffers
            std::vector<EncJob> jobs;
            std::vector<int> shifts;
            while (shift < slots && jobs.size() + m_ct <= cap) {
                if (ca->shiftT[(size_t)bi * slots + shift]) {
                    shifts.push_back(shift);
                    for (int bj = 0; bj < m_ct; bj++)
                        if (ca->pidx[((size_t)bi * slots + shift) * m_ct + bj] >= 0)  // EncodeDiagWithEncoder(blockVec, -shift, d*giant, maxLevel) :1024
                            jobs.push_back(EncJob{bi * slots, bj * slots, block_rows(ca, bi), block_cols(ca, bj), shift, d * (shift / d),
                                                  (long long)(jobs.size() * polyB)});
                }
                shift++;
            }
            if (jobs.empty()) continue;
            SFG_CUDA(c, cudaMemcpyAsync(djobs.p, jobs.data(), jobs.size() * sizeof(EncJob), cudaMemcpyDefault, c->stream));
            if (launch_encode(c, ca->g->d, ca->ncols, djobs.as<EncJob>(), (int)jobs.size(), layF, true /* ToMontgomeryForm :401-409 */, dbuf.p, nullptr,
                              c->stream) ||
                launch_bswap64(c, dbuf.as<uint64_t>(), jobs.size() * (size_t)nlF * N, c->stream)) {
                fclose(f);
                return -1;
            }
            SFG_CUDA(c, cudaMemcpyAsync(host.data(), dbuf.p, jobs.size() * polyB, cudaMemcpyDefault, c->stream));
            SFG_CUDA(c, cudaStreamSynchronize(c->stream));
            size_t k = 0;
            for (int sh : shifts) {
                uint64_t len = 4;
                for (int bj = 0; bj < m_ct; bj++) len += 1 + (ca->pidx[((size_t)bi * slots + sh) * m_ct + bj] >= 0 ? polyB : 0);
                const uint32_t sh32 = (uint32_t)sh;
                ok = ok && fwrite(&len, 8, 1, f) == 1 && fwrite(&sh32, 4, 1, f) == 1;
                for (int bj = 0; bj < m_ct && ok; bj++) {
                    const bool present = ca->pidx[((size_t)bi * slots + sh) * m_ct + bj] >= 0;
                    const unsigned char empty = present ? 0 : 1;
                    ok = fwrite(&empty, 1, 1, f) == 1;
                    if (present && ok) ok = fwrite(host.data() + (k++) * polyB, 1, polyB, f) == polyB;
                }
            }
        }
        if (fclose(f) != 0) ok = false;
        if (!ok) SFG_FAIL(c, "write %s failed: %s", fn.c_str(), strerror(errno));
    }
    return 0;
}

int cache_load_files(Ctx *c, const char *prefix, size_t nrows, size_t ncols, int maxLevel, Cache **out) {
    const bool infer = nrows == 0 || ncols == 0;  // shape unknown (the reference's Compute only has the prefix): take it from the files
    if (maxLevel < 1 || maxLevel > c->nQ - 1) SFG_FAIL(c, "maxLevel %d needs %d Q limbs, parameters have %d", maxLevel, maxLevel + 1, c->nQ);
    SFG_CUDA(c, cudaSetDevice(c->device));
    Cache *ca = new Cache();
    std::vector<FILE *> files;
    auto fail = [&](const std::string &m) {
        for (FILE *f : files)
            if (f) fclose(f);
        cache_destroy(ca);
        if (!m.empty()) c->err = m;
        return -1;
    };
    const int N = c->N, nlF = maxLevel + 1;
    const size_t polyB = (size_t)nlF * N * 8;
    if (infer) {  // number of block rows = number of files, number of block columns = vectorLen of the first header
        int nf = 0;
        uint64_t vlen = 0;
        for (;; nf++) {
            FILE *f = fopen((std::string(prefix) + "_" + std::to_string(nf) + ".bin").c_str(), "rb");
            if (!f) break;
            if (nf == 0 && fread(&vlen, 8, 1, f) != 1) vlen = 0;
            fclose(f);
        }
        if (nf == 0) return fail("open " + std::string(prefix) + "_0.bin: " + strerror(ENOENT));
        if (vlen == 0 || vlen > (1u << 20)) return fail(std::string(prefix) + "_0.bin: bad header");
        if (cache_meta_base(c, ca, (size_t)nf * c->slots, (size_t)vlen * c->slots, maxLevel)) return fail("");
    } else if (cache_init_meta(c, ca, nrows, ncols, maxLevel)) {
        return fail("");
    }
    const int slots = ca->slots, d = ca->d, m_ct = ca->m_ct, nbr = ca->nbr;
    size_t npoly_files = 0;
    struct Rec {
        long long pos = -1;  // file offset of the record payload (after the 8-byte length)
        uint64_t len = 0;
    };
    std::vector<Rec> recs((size_t)nbr * slots);
    for (int bi = 0; bi < nbr; bi++) {
        const std::string fn = std::string(prefix) + "_" + std::to_string(bi) + ".bin";
        FILE *f = fopen(fn.c_str(), "rb");
        if (!f) return fail("open " + fn + ": " + strerror(errno));  // the reference panics in NewDiagCacheStream (gwas/filestream.go:56-61)
        files.push_back(f);
        uint64_t hdr[6];
        std::vector<unsigned char> tab(2 * (size_t)d);
        if (fread(hdr, 8, 6, f) != 6 || fread(tab.data(), 1, tab.size(), f) != tab.size()) return fail(fn + ": truncated header");
        if (hdr[0] != (uint64_t)m_ct || hdr[1] != (uint64_t)maxLevel || hdr[3] != (uint64_t)N || hdr[4] != (uint64_t)nlF)
            return fail(fn + ": header (vectorLen " + std::to_string(hdr[0]) + ", level " + std::to_string(hdr[1]) + ", n " + std::to_string(hdr[3]) +
                        ", numModuli " + std::to_string(hdr[4]) + ") does not match the " + std::to_string(nrows) + " x " + std::to_string(ncols) +
                        " matrix / parameters");
        if (infer) {
            for (int k = 0; k < d; k++) {
                ca->baby[(size_t)bi * d + k] = tab[k] != 0;
                ca->giant[(size_t)bi * d + k] = tab[d + k] != 0;
            }
        } else if (memcmp(tab.data(), &ca->baby[(size_t)bi * d], d) || memcmp(tab.data() + d, &ca->giant[(size_t)bi * d], d)) {
            return fail(fn + ": baby / giant tables do not match the matrix shape");
        }
        long long pos = 48 + 2 * (long long)d;
        for (;;) {
            uint64_t len;
            uint32_t sh;
            if (fread(&len, 8, 1, f) != 1) break;  // EOF
            if (len < 4 || fread(&sh, 4, 1, f) != 1) return fail(fn + ": truncated record");
            if (sh >= (uint32_t)slots || (!infer && !ca->shiftT[(size_t)bi * slots + sh])) return fail(fn + ": unexpected diagonal " + std::to_string(sh));
            if (recs[(size_t)bi * slots + sh].pos >= 0) return fail(fn + ": diagonal " + std::to_string(sh) + " appears twice");
            recs[(size_t)bi * slots + sh] = Rec{pos + 8, len};
            // nil flags of the block columns (isEmpty bytes sit between the polynomials)
            long long p = pos + 12;
            for (int bj = 0; bj < m_ct; bj++) {
                unsigned char empty;
                if (p + 1 > pos + 8 + (long long)len || fseeko(f, p, SEEK_SET) || fread(&empty, 1, 1, f) != 1) return fail(fn + ": malformed record");
                p += 1 + (empty == 1 ? 0 : (long long)polyB);
                int &pi = ca->pidx[((size_t)bi * slots + sh) * m_ct + bj];
                if (infer) {
                    if (empty != 1) pi = 0;  // numbered below
                } else if ((pi >= 0) != (empty != 1)) {
                    return fail(fn + ": diagonal " + std::to_string(sh) + ", block column " + std::to_string(bj) + ": nil flag does not match the matrix shape");
                }
                npoly_files += empty != 1;
            }
            if (p != pos + 8 + (long long)len) return fail(fn + ": record length does not match its contents");
            if (infer) ca->shiftT[(size_t)bi * slots + sh] = 1;
            pos += 8 + (long long)len;
            if (fseeko(f, pos, SEEK_SET)) return fail(fn + ": seek failed");
        }
        for (int sh = 0; sh < slots; sh++)
            if (ca->shiftT[(size_t)bi * slots + sh] && recs[(size_t)bi * slots + sh].pos < 0) return fail(fn + ": diagonal " + std::to_string(sh) + " missing");
    }
    if (infer) {
        int n = 0;
        for (int &pi : ca->pidx)
            if (pi >= 0) pi = n++;
        ca->npoly = (size_t)n;
        if (cache_meta_finish(c, ca)) return fail("");
    }
    if (npoly_files != ca->npoly) return fail("cache files hold " + std::to_string(npoly_files) + " diagonals, expected " + std::to_string(ca->npoly));
    // the image must be resident: there is no genotype matrix to regenerate it from
    const TcGeomP &tc = ca->tc;
    const size_t bytes = (size_t)tc.group_bytes * tc.ngroups;
    if (ca->npoly == 0) return fail("cache_load_files: no diagonals");
    if (cudaMalloc(&ca->img, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail("cache_load_files: the " + std::to_string(bytes >> 20) + " MiB image does not fit in HBM (loaded caches must be resident)");
    }
    ca->img_bytes = bytes;
    if (cudaMemsetAsync(ca->img, 0, bytes, c->stream) != cudaSuccess) return fail("cudaMemset failed");
    const int Kg = tc.Kg, K = tc.K;
    const long long gbytes = tc_group_bytes(ca, tc.ntiles);
    const size_t RB = (size_t)ca->lay.bytes;
    const size_t cap = std::max<size_t>(1, env_size("SFG_CACHEFILE_CHUNK_POLYS", ((size_t)1 << 30) / polyB));
    void *tmp, *dtab;
    if (ws_get(c, WS_TMPP, (size_t)128 * Kg * RB, &tmp) || ws_get(c, WS_POFF, (size_t)128 * Kg * sizeof(long long), &dtab)) return fail("");
    Buf draw, doff;
    if (draw.alloc(c, cap * polyB) || doff.alloc(c, cap * sizeof(long long))) return fail("");
    std::vector<unsigned char> hraw(cap * polyB), rec;
    std::vector<long long> tab((size_t)128 * Kg), hoff;
    auto flush = [&]() -> int {
        if (hoff.empty()) return 0;
        SFG_CUDA(c, cudaMemcpyAsync(draw.p, hraw.data(), hoff.size() * polyB, cudaMemcpyDefault, c->stream));
        SFG_CUDA(c, cudaMemcpyAsync(doff.p, hoff.data(), hoff.size() * sizeof(long long), cudaMemcpyDefault, c->stream));
        if (launch_file_to_rec(c, draw.as<uint64_t>(), doff.as<long long>(), (int)hoff.size(), nlF, ca->lay, tmp, c->stream)) return -1;
        SFG_CUDA(c, cudaStreamSynchronize(c->stream));  // the staging buffers are reused
        hoff.clear();
        return 0;
    };
    for (int ct = 0; ct < tc.ntiles; ct++)
        for (int grp = 0; grp < tc.ngroups; grp++) {
            std::fill(tab.begin(), tab.end(), -1LL);
            bool any = false;
            for (int kk = 0; kk < Kg && grp * Kg + kk < K; kk++) {
                const int k = grp * Kg + kk, bi = ca->kbi[k], b = ca->kb[k];
                int cur_g = -1;
                std::vector<long long> bjpos;  // payload offset of block column bj inside the current record, -1 = nil
                for (int cc = 0; cc < 128; cc++) {
                    const int col = ct * 128 + cc;
                    if (col >= tc.ncols) break;
                    const int g = ca->gact[col / m_ct], bj = col % m_ct, shift = g * d + b;
                    if (shift >= slots || ca->pidx[((size_t)bi * slots + shift) * m_ct + bj] < 0) continue;
                    if (g != cur_g) {  // read the record of (bi, shift) once for all its block columns in this tile
                        const Rec &r = recs[(size_t)bi * slots + shift];
                        rec.resize(r.len);
                        if (fseeko(files[bi], r.pos, SEEK_SET) || fread(rec.data(), 1, r.len, files[bi]) != r.len) return fail("cache file: short read");
                        bjpos.assign(m_ct, -1);
                        uint64_t p = 4;
                        for (int j = 0; j < m_ct; j++) {
                            if (p >= r.len) return fail("cache file: malformed record");
                            const bool empty = rec[p++] == 1;
                            if (!empty) {
                                if (p + polyB > r.len) return fail("cache file: malformed record");
                                bjpos[j] = (long long)p;
                                p += polyB;
                            }
                        }
                        cur_g = g;
                    }
                    if (bjpos[bj] < 0) return fail("cache file: a diagonal the matrix shape requires is nil in the file");
                    const long long off = (long long)((size_t)kk * 128 + cc) * (long long)RB;
                    tab[(size_t)kk * 128 + cc] = off;
                    memcpy(hraw.data() + hoff.size() * polyB, rec.data() + bjpos[bj], polyB);
                    hoff.push_back(off);
                    any = true;
                    if (hoff.size() == cap && flush()) return fail("");
                }
            }
            if (!any) continue;
            if (flush()) return fail("");
            if (cudaMemcpyAsync(dtab, tab.data(), tab.size() * sizeof(long long), cudaMemcpyDefault, c->stream) != cudaSuccess) return fail("cudaMemcpy failed");
            if (launch_img_p(c, tc, ca->lay, tmp, (const long long *)dtab, tc.ntiles, ct, ca->img + (size_t)grp * gbytes, c->stream)) return fail("");
            if (cudaStreamSynchronize(c->stream) != cudaSuccess) return fail("image build failed");  // tab / tmp are reused
        }
    for (FILE *f : files) fclose(f);
    files.clear();
    ca->materialised = true;
    *out = ca;
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Count sketch + column sums of the randomized PCA (gwas/pca.go:124-162) on the HBM-resident genotype matrix.
// rand_index[i] in [0, kp), sgn[i] = +-1 per row (the PRG draws stay with the caller: mpcObj.Network.Rand.CurPRG()).
// d_ms: optional, time of the scan in milliseconds (CUDA events on the context's stream).
// ---------------------------------------------------------------------------------------------------------------
int geno_count_sketch(Ctx *c, const Geno *g, const int32_t *rand_index, const int8_t *sgn, int kp, double *sketch, uint64_t *xsum, uint64_t *x2sum,
                      float *ms) {
    if (g->filled != g->nrows) SFG_FAIL(c, "genotype matrix incomplete: %zu of %zu rows pushed", g->filled, g->nrows);
    if (kp < 1) SFG_FAIL(c, "count sketch: kp = %d", kp);
    SFG_CUDA(c, cudaSetDevice(c->device));
    const size_t nrows = g->nrows, ncols = g->ncols;
    std::vector<int> off(kp + 1, 0), rows(nrows);
    std::vector<int8_t> sg(nrows);
    for (size_t i = 0; i < nrows; i++) {
        if (rand_index[i] < 0 || rand_index[i] >= kp) SFG_FAIL(c, "count sketch: randIndex[%zu] = %d outside [0, %d)", i, rand_index[i], kp);
        if (sgn[i] != 1 && sgn[i] != -1) SFG_FAIL(c, "count sketch: sgn[%zu] = %d is not +-1", i, (int)sgn[i]);
        off[rand_index[i] + 1]++;
    }
    for (int b = 0; b < kp; b++) off[b + 1] += off[b];
    {
        std::vector<int> cur(off.begin(), off.end() - 1);
        for (size_t i = 0; i < nrows; i++) {
            const int p = cur[rand_index[i]]++;
            rows[p] = (int)i;
            sg[p] = sgn[i];
        }
    }
    Buf drows, dsg, doff, dsk, dskf, dx, dbad;
    if (drows.alloc(c, nrows * 4) || dsg.alloc(c, nrows) || doff.alloc(c, (size_t)(kp + 1) * 4) || dsk.alloc(c, (size_t)kp * ncols * 8) ||
        dskf.alloc(c, (size_t)kp * ncols * 8) || dx.alloc(c, 2 * ncols * 8) || dbad.alloc(c, 4))
        return -1;
    SFG_CUDA(c, cudaMemcpyAsync(drows.p, rows.data(), nrows * 4, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaMemcpyAsync(dsg.p, sg.data(), nrows, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaMemcpyAsync(doff.p, off.data(), (size_t)(kp + 1) * 4, cudaMemcpyDefault, c->stream));
    SFG_CUDA(c, cudaMemsetAsync(dsk.p, 0, (size_t)kp * ncols * 8, c->stream));
    SFG_CUDA(c, cudaMemsetAsync(dx.p, 0, 2 * ncols * 8, c->stream));
    SFG_CUDA(c, cudaMemsetAsync(dbad.p, 0, 4, c->stream));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0, c->stream);
    const int rc = launch_count_sketch(c, g->d, nrows, ncols, drows.as<int>(), dsg.as<int8_t>(), doff.as<int>(), kp, dsk.as<long long>(), dskf.as<double>(),
                                       dx.as<unsigned long long>(), dx.as<unsigned long long>() + ncols, dbad.as<int>(), c->stream);
    cudaEventRecord(e1, c->stream);
    int bad = 0;
    cudaError_t e = rc ? cudaSuccess : cudaMemcpyAsync(&bad, dbad.p, 4, cudaMemcpyDefault, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    float t = 0;
    if (!rc && e == cudaSuccess) cudaEventElapsedTime(&t, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (rc) return -1;
    SFG_CUDA(c, e);
    if (bad) SFG_FAIL(c, "count sketch: a genotype outside {0, 1, 2} (missing values must already be replaced, gwas/filestream.go:349-351)");
    if (ms) *ms = t;
    SFG_CUDA(c, cudaMemcpy(sketch, dskf.p, (size_t)kp * ncols * 8, cudaMemcpyDefault));
    SFG_CUDA(c, cudaMemcpy(xsum, dx.p, ncols * 8, cudaMemcpyDefault));
    SFG_CUDA(c, cudaMemcpy(x2sum, dx.as<uint64_t>() + ncols, ncols * 8, cudaMemcpyDefault));
    return 0;
}

}  // namespace sfg
