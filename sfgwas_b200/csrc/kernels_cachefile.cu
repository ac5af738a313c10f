// kernels_cachefile.cu -- conversion between the reference's on-disk diagonal cache records (gwas/filestream.go:42-282, SURVEY
// App. D.2: per plaintext numModuli x N coefficients, each a BIG-endian uint64 (Lattigo ring.WriteCoeffsTo), NTT domain,
// Montgomery form, gwas/matmult.go:401-440) and the device record layout the image builder reads (plain residues, narrow limbs
// packed as uint32).  Pure streaming kernels: one thread per coefficient, unit-stride 8-byte accesses.
#include "kernels.h"

namespace sfg {

__device__ __forceinline__ uint64_t bswap64(uint64_t x) {
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | (uint64_t)__byte_perm(hi, 0, 0x0123);
}

__global__ void k_bswap64(uint64_t *__restrict__ x, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = bswap64(x[i]);
}
int launch_bswap64(Ctx *c, uint64_t *x, size_t n, cudaStream_t st) {
    if (n == 0) return 0;
    k_bswap64<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, st>>>(x, n);
    SFG_LAUNCHED(c, "k_bswap64", st);
    return 0;
}

// raw [npoly][file_nl][N] (file bytes) -> record p at out + dst_off[p]: limbs 0..lay.nl-1 as plain residues (InvMForm = MRed(x, 1))
__global__ void k_file_to_rec(const uint64_t *__restrict__ raw, const long long *__restrict__ dst_off, int file_nl, PolyLayout lay, int N,
                              const LimbConst *__restrict__ lcs, unsigned char *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, l = blockIdx.y, p = blockIdx.z;
    if (j >= N) return;
    const LimbConst lc = lcs[l];
    const uint64_t v = mred(bswap64(raw[((size_t)p * file_nl + l) * N + j]), 1, lc);
    unsigned char *o = out + dst_off[p] + lay.off[l];
    if (lay.es[l] == 4)
        reinterpret_cast<uint32_t *>(o)[j] = (uint32_t)v;
    else
        reinterpret_cast<uint64_t *>(o)[j] = v;
}
int launch_file_to_rec(Ctx *c, const uint64_t *raw, const long long *dst_off_dev, int npoly, int file_nl, const PolyLayout &lay, void *out,
                       cudaStream_t st) {
    if (npoly <= 0) return 0;
    k_file_to_rec<<<dim3((c->N + 255) / 256, lay.nl, npoly), 256, 0, st>>>(raw, dst_off_dev, file_nl, lay, c->N, c->lc, (unsigned char *)out);
    SFG_LAUNCHED(c, "k_file_to_rec", st);
    return 0;
}

}  // namespace sfg
