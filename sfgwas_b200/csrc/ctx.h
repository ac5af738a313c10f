// ctx.h -- host-side context of libsfgwas_b200 (C++17, CUDA runtime only; no torch types).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "modarith.cuh"
#include "ntt2.cuh"

namespace sfg {

constexpr int kMaxAlpha = 8;   // max #moduli per key-switch digit / #P moduli (Lattigo arrays are size 8 too)
constexpr int kMaxLimbs = 48;

// Fast exact base conversion {s_k} -> t of Lattigo (ring.modUpExact / Decomposer), SURVEY App. B.5.
struct BaseConv {
    int ns;                       // number of source moduli (1 => plain BRedAdd of the representative)
    int src_limb[kMaxAlpha];      // QP index of each source modulus
    uint64_t sinv[kMaxAlpha];     // (S/s_k)^-1 mod s_k
    uint64_t sinv_sh[kMaxAlpha];
    uint64_t fac[kMaxAlpha];      // S/s_k mod t
    uint64_t fac_sh[kMaxAlpha];
    uint64_t smod;                // S mod t
    uint64_t smod_sh;
    double sf[kMaxAlpha];         // float64(s_k)
};

struct GaloisKey {
    uint64_t galEl = 0;
    uint64_t *key = nullptr;      // device [beta][2][nQP][N], NTT + Montgomery form (Lattigo SwitchingKey)
    uint32_t *perm = nullptr;     // device [N] PermuteNTTIndex(galEl)
};

struct Ctx {
    int device = 0;
    int logN = 0, N = 0, slots = 0, d = 0;
    int nQ = 0, nP = 0, nQP = 0, beta = 0;
    double scale = 0;
    std::vector<uint64_t> mod, psi;        // host copies
    std::vector<LimbConst> lc_h;
    LimbConst *lc = nullptr;               // device [nQP]
    uint64_t *tw = nullptr;                // device [nQP][4][N]  (radix-2 tables: encoder, logN > 14)
    TwTab *tw2 = nullptr;                  // device [nQP]: per-class twiddle tables of the register-tiled transforms (ntt2.cuh)
    std::vector<void *> tw2_bufs;          // device allocations behind tw2
    // encoder (special inverse FFT)
    double2 *roots = nullptr;              // device [2N+1] exp(2 pi i k / 2N)
    int *rot5 = nullptr;                   // device [slots] 5^j mod 2N
    double2 *fft_tw = nullptr;             // device [slots]: twiddles of the special inverse FFT per stage, stage with half-length h at [h, 2h)
    // discrete-log order of the NTT evaluation points: coefficient i of an NTT-domain polynomial is the value at psi^(2 brv(i) + 1) =
    // psi^(+-5^t); position q = s * N/2 + t (s = sign = top bit of i).  In that order the automorphism X -> X^(5^r) is a cyclic
    // shift by r inside each half (kernels_ks.cu: the giant-step sums).  dlog_pos[i] = q, dlog_src[q] = i.
    uint32_t *dlog_pos = nullptr, *dlog_src = nullptr;  // device [N]
    double2 *ddcos = nullptr;              // device [2N] cos(2 pi t / 2N) as double-double (hi, lo)
    double enc_delta = 0;                  // half-integer ambiguity window for the FP64 path
    unsigned long long *enc_stats = nullptr;  // device [2]: {rechecked coefficients, unresolved ties}
    // key-switch tables per level: bc_ks[level] -> device array [beta_l][level+1+nP]; bc_md -> [level+1]
    std::map<int, BaseConv *> bc_ks, bc_md;
    std::map<int, uint64_t *> pinv;        // device [level+1][2]: P^-1 mod q_l and Shoup companion
    std::map<uint64_t, GaloisKey> keys;    // galEl -> key
    cudaStream_t stream = nullptr;
    size_t cache_budget = 0;               // bytes of HBM the diagonal cache may take (0 = auto)
    // grow-only device workspace, reused across calls (no cudaMalloc / cudaFree in the steady state)
    struct WsBuf {
        void *p = nullptr;
        size_t bytes = 0;
    };
    WsBuf ws[20];
    // grow-only PINNED host staging (hostio.cu): the "_ptrs" entry points gather / scatter one Go slice per limb through it
    WsBuf pin[2];
    std::mutex mu;
    std::string err;
    // counters
    unsigned long long launches = 0;
};

#define SFG_CUDA(ctx, expr)                                                                             \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            char buf__[512];                                                                            \
            snprintf(buf__, sizeof buf__, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            (ctx)->err = buf__;                                                                         \
            return -1;                                                                                  \
        }                                                                                               \
    } while (0)

#define SFG_FAIL(ctx, ...)                              \
    do {                                                \
        char buf__[512];                                \
        snprintf(buf__, sizeof buf__, __VA_ARGS__);     \
        (ctx)->err = buf__;                             \
        return -1;                                      \
    } while (0)

// after every kernel launch: catch launch errors; with SFG_DEBUG=1 also synchronise and report the failing kernel
int launch_check(Ctx *c, const char *what, cudaStream_t st);
#define SFG_LAUNCHED(ctx, what, st)                  \
    do {                                             \
        (ctx)->launches++;                           \
        if (launch_check((ctx), (what), (st))) return -1; \
    } while (0)

// ---- device memory helpers (ctx.cu) ----
// The context stream is created with cudaStreamNonBlocking: it does NOT synchronise with the legacy default stream, and a synchronous
// cudaMemcpy from pageable host memory may return before its DMA (issued on the legacy stream) has reached the device.  A kernel
// launched on the context stream right after such an upload can read what the allocation held before (root cause of the rare
// first-call mismatch of round 1; profiles/microbench/nullstream_race.cu).  Every host -> device upload of the library
// therefore goes through upload(): ordered on the context stream, then synchronised.
int upload(Ctx *c, void *dst, const void *src, size_t bytes);
// cudaMalloc; under SFG_POISON=1 the allocation is filled with 0xA5 first, so that any read of bytes this library has not written
// (including reads compute-sanitizer's initcheck cannot see: cp.async.bulk / tcgen05 operand fetches) gives a deterministic mismatch
int dev_alloc(Ctx *c, void **p, size_t bytes, const char *what);
bool poison_enabled();
// fills [p, p + bytes) with 0xA5 on the context stream when SFG_POISON=1 (per-call scratch whose contents are declared undefined)
void poison_fill(Ctx *c, void *p, size_t bytes);

// ---- per-limb host pointers <-> device (hostio.cu): what a cgo caller hands over is one pageable Go slice per limb ----
enum PinSlot { PIN_IN = 0, PIN_OUT = 1 };
int pinned_get(Ctx *c, int slot, size_t bytes, void **out);
void pinned_release(Ctx *c);
// np polynomials of N uint64 each, polynomial p at limbs[p] (host or device memory) -> d_dst + p * N; stream-ordered on c->stream,
// the host buffers are free for reuse on return
int gather_limbs_to_device(Ctx *c, const uint64_t *const *limbs, size_t np, uint64_t *d_dst);
// the reverse for results that already sit in the pinned staging buffer `src` (np polynomials): parallel host copies
void scatter_host_to_limbs(const uint64_t *src, uint64_t *const *limbs, size_t np, size_t N);
// plain device -> per-limb pointers (synchronous; used where nothing overlaps)
int scatter_device_to_limbs(Ctx *c, const uint64_t *d_src, uint64_t *const *limbs, size_t np);

// workspace slots
enum WsSlot { WS_C2 = 0, WS_ACC, WS_META, WS_R, WS_CV, WS_POFF, WS_TMPP, WS_A, WS_OUT, WS_META2, WS_RIMG, WS_PIMG, WS_MD, WS_KSB, WS_VQ, WS_EXTD, WS_RTAB, WS_COUNT };
int ws_get(Ctx *c, int slot, size_t bytes, void **out);  // returns a buffer of at least `bytes` (contents undefined)
void ws_release(Ctx *c);

// ---- table builders (ctx.cu) ----
int ctx_build_tables(Ctx *c, const uint64_t *psi_opt);
int ctx_get_ks_tables(Ctx *c, int level, BaseConv **ks, BaseConv **md, uint64_t **pinv);

}  // namespace sfg
