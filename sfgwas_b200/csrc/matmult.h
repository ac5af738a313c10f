// matmult.h -- host-side orchestration of the stream MatMult entry points (gwas/matmult.go:914-1505).
#pragma once
#include <atomic>
#include <vector>

#include "kernels.h"

namespace sfg {

struct Geno {
    Ctx *c = nullptr;             // may dangle after sfg_ctx_destroy: only `device` is used on release
    int device = 0;
    size_t nrows = 0, ncols = 0, filled = 0;
    int8_t *d = nullptr;  // device, row-major nrows x ncols
    std::atomic<int> refs{1};
};

struct Cache {
    Ctx *c = nullptr;
    int device = 0;
    Geno *g = nullptr;            // retained (needed when diagonals are regenerated on the fly)
    bool own_geno_copy = false;
    int maxLevel = 0, L = 0;      // L = maxLevel limbs are used by the MAC
    int slots = 0, d = 0, m_ct = 0, nbr = 0;
    size_t nrows = 0, ncols = 0;
    std::vector<uint8_t> baby, giant, shiftT;  // [nbr][d], [nbr][d], [nbr][slots]   (matmult.go:962-974)
    bool materialised = false;
    PolyLayout lay{};             // record layout of cached diagonals and of the rotation cache (packed narrow limbs)
    // Diagonals are kept as the byte-plane K-major image the tensor-core MAC streams (kernels_mactc.cu): `tc.ngroups` K groups,
    // each [limb][coefficient superblock][column tile][coefficient][byte plane] blocks of 128 x Kg bytes.
    TcGeomP tc{};
    unsigned char *img = nullptr;
    size_t img_bytes = 0;
    size_t npoly = 0;
    std::vector<int> pidx;        // host [(bi*slots+shift)*m_ct+bj] -> running index of the diagonal or -1 (nil, matmult.go:703-705)
    // K list: (bi, b) pairs with an active baby step, in (bi, b) order; kidx[bi*d+b] -> k or -1
    std::vector<int> kbi, kb, kidx;
    std::vector<int> gact;        // active giant indices (any block row)
    int bi_lo = 0, bi_hi = -1;    // block rows whose diagonals the image holds (-1 = all): block-row sharding builds only a rank's own
    int g_part = 0, g_nparts = 1; // giant-step sharding: gact holds share g_part of g_nparts of the active giant steps
};

struct Buf {  // RAII device buffer
    void *p = nullptr;
    size_t bytes = 0;
    ~Buf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    int alloc(Ctx *c, size_t n) {
        release();
        if (n == 0) n = 16;
        if (dev_alloc(c, &p, n, "temporary buffer")) return -1;  // 0xA5-filled under SFG_POISON=1
        bytes = n;
        return 0;
    }
    template <typename T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

int geno_create(Ctx *c, size_t nrows, size_t ncols, Geno **out);
int geno_push(Geno *g, const int8_t *rows, size_t n);
void geno_release(Geno *g);

// bi_lo / bi_hi: build only the diagonals of block rows [bi_lo, bi_hi) (block-row sharding; such a cache serves mm_partial_dev only)
// g_part / g_nparts: keep only that contiguous share of the active giant steps (giant-step sharding: Compute on such a cache yields the
// partial sum over its giant steps; the per-rank results add up mod q to the full product)
int cache_build(Ctx *c, Geno *g, int maxLevel, Cache **out, int bi_lo = 0, int bi_hi = -1, int g_part = 0, int g_nparts = 1);
// one cached diagonal polynomial (plain canonical residues, [L][N]) read back out of the image; 0 = nil
int cache_get_diag_dev(Ctx *c, const Cache *ca, int bi, int shift, int bj, uint64_t *d_out, int *present);
void cache_destroy(Cache *cache);
// the reference's on-disk cache format (gwas/filestream.go:19-282): <prefix>_<bi>.bin per block row
int cache_write_files(Ctx *c, const Cache *ca, const char *prefix);
int cache_load_files(Ctx *c, const char *prefix, size_t nrows, size_t ncols, int maxLevel, Cache **out);

// full single-GPU compute: d_A device [s][nbr][2][nlA][N] -> d_out device [s][m_ct][2][L][N]
// host_out (optional): the result is also copied to this HOST buffer, rows leaving as soon as their giant-step sums are final
// out_limbs (optional, instead of host_out): one HOST pointer per limb of the result, [((i*m_ct+bj)*2+c)*maxLevel+l] (cgo callers)
int mm_compute_dev(Ctx *c, const uint64_t *d_A, int s, int nbr, int levelA, int maxLevel, Cache *cache, uint64_t *d_out,
                   uint64_t *host_out = nullptr, uint64_t *const *out_limbs = nullptr);
// baby-step sharding (matmult.cu): share `part` of `nparts` of the rotation cache into d_R; the remainder of Compute on a complete d_R
size_t mm_baby_chunk_bytes(const Cache *ca, int s, int nparts);
int mm_baby_dev(Ctx *c, const uint64_t *d_A, int s, int nbr, int levelA, int maxLevel, Cache *cache, int part, int nparts, void *d_R);
int mm_compute_r_dev(Ctx *c, const void *d_R, int s, int maxLevel, Cache *cache, uint64_t *d_out);
// multi-GPU pieces
int mm_partial_dev(Ctx *c, const uint64_t *d_A, int s, int nbr, int levelA, int maxLevel, Cache *cache, int bi_lo, int bi_hi,
                   uint64_t *d_cv);
int mm_finish_dev(Ctx *c, Cache *cache, int s, int maxLevel, const uint64_t *d_cv, int g_lo, int g_hi, uint64_t *d_out);
int rotate_right_dev(Ctx *c, int level, const uint64_t *d_in, int nct, int nrot, uint64_t *d_out);
int encode_diag_host(Ctx *c, const Geno *g, int bi, int shift, int nrot, int level, bool mont, uint64_t *out, uint8_t *present,
                     int64_t *coeffs);

// ciphertext algebra of the callers (gwas/matmult.go:27-116); device pointers, see matmult.cu
int rescale_dev(Ctx *c, int level, const uint64_t *d_in, int nct, int times, uint64_t *d_out);
int mul_relin_dev(Ctx *c, int level, const uint64_t *d_x, int nx, int x_nl, const uint64_t *d_y, int ny, int y_nl, int times, uint64_t *d_out);
int mul_plain_dev(Ctx *c, int level, const uint64_t *d_pt, int npt, int pt_nl, const uint64_t *d_ct, int nct, int ct_nl, int times, uint64_t *d_out);
int addsub_dev(Ctx *c, int level, const uint64_t *d_a, int na, int a_nl, const uint64_t *d_b, int nb, int b_nl, bool sub, uint64_t *d_out);
int inner_sum_all_dev(Ctx *c, int level, const uint64_t *d_in, int nvec, int cnt, uint64_t *d_out);
int geno_count_sketch(Ctx *c, const Geno *g, const int32_t *rand_index, const int8_t *sgn, int kp, double *sketch, uint64_t *xsum, uint64_t *x2sum,
                      float *ms);
int encode_slots_host(Ctx *c, const int8_t *v, int level, bool mont, uint64_t *out);

// local arithmetic of the collective bootstrap (kernels_refresh.cu; mpc/mhe.go:262-341): device pointers
int refresh_gen_shares_dev(Ctx *c, int level, int nct, const uint64_t *d_c1, const uint64_t *d_sk, const uint64_t *d_crp, const uint64_t *d_mask,
                           const int8_t *d_sign, int nwords, double in_scale, double out_scale, const long long *d_e0, const long long *d_e1, uint64_t *d_h0,
                           uint64_t *d_h1);
int refresh_finish_dev(Ctx *c, int level, int nct, const uint64_t *d_c0, int c0_nl, double in_scale, double out_scale, const uint64_t *d_agg0,
                       const uint64_t *d_agg1, const uint64_t *d_crp, uint64_t *d_out);

extern thread_local float g_last_ms[5];  // baby, mac phase, giant, total, mac kernel only

}  // namespace sfg
