// hostio.cu -- moving ciphertexts between the caller's memory and the device for the "_ptrs" entry points.
//
// The cgo shim (go/gwas/matmult_b200.go) cannot hand over one flat buffer: a Lattigo polynomial is a [][]uint64, i.e. one Go slice per
// limb (gwas/matmult.go:372-375), pageable, 64-128 KB each.  One cudaMemcpyAsync per limb from pageable memory costs a driver staging
// round trip each (config 2: 360 + 2 500 of them per call).  Instead the limbs are gathered by a few host threads into a PINNED,
// grow-only staging buffer that goes over PCIe/C2C as ONE transfer, and results come back the same way (chunk by chunk, overlapped with
// the giant-step key-switches of the following rows: matmult.cu HostSink).
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "ctx.h"

namespace sfg {

int pinned_get(Ctx *c, int slot, size_t bytes, void **out) {
    Ctx::WsBuf &b = c->pin[slot];
    if (bytes == 0) bytes = 16;
    if (b.bytes < bytes) {
        if (b.p) {
            SFG_CUDA(c, cudaStreamSynchronize(c->stream));
            cudaFreeHost(b.p);
            b.p = nullptr;
            b.bytes = 0;
        }
        cudaError_t e = cudaHostAlloc(&b.p, bytes, cudaHostAllocDefault);
        if (e != cudaSuccess) {
            cudaGetLastError();
            b.p = nullptr;
            SFG_FAIL(c, "pinned staging allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
        }
        b.bytes = bytes;
    }
    *out = b.p;
    return 0;
}
void pinned_release(Ctx *c) {
    for (auto &b : c->pin) {
        if (b.p) cudaFreeHost(b.p);
        b.p = nullptr;
        b.bytes = 0;
    }
}

// copies of `np` limbs split over a few host threads (a single thread moves ~10 GB/s; the 164 MB result of config 2 would cost 16 ms)
template <class F>
static void par_limbs(size_t np, size_t bytes_each, F f) {
    const size_t total = np * bytes_each;
    unsigned nt = (unsigned)std::min<size_t>(8, std::max<size_t>(1, total >> 22));  // one thread per 4 MiB, at most 8
    nt = std::min(nt, std::max(1u, std::thread::hardware_concurrency()));
    if (nt <= 1) {
        for (size_t p = 0; p < np; p++) f(p);
        return;
    }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([=] {
            for (size_t p = np * t / nt; p < np * (t + 1) / nt; p++) f(p);
        });
    for (auto &x : th) x.join();
}

static bool is_device_ptr(const void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int gather_limbs_to_device(Ctx *c, const uint64_t *const *limbs, size_t np, uint64_t *d_dst) {
    const size_t N = c->N, B = N * 8;
    if (np == 0) return 0;
    if (is_device_ptr(limbs[0])) {  // device-resident limbs: plain stream-ordered copies
        for (size_t p = 0; p < np; p++) SFG_CUDA(c, cudaMemcpyAsync(d_dst + p * N, limbs[p], B, cudaMemcpyDefault, c->stream));
        SFG_CUDA(c, cudaStreamSynchronize(c->stream));
        return 0;
    }
    void *pin;
    if (pinned_get(c, PIN_IN, np * B, &pin)) return -1;
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));  // a previous transfer out of the staging buffer must have finished
    unsigned char *pb = (unsigned char *)pin;
    par_limbs(np, B, [=](size_t p) { memcpy(pb + p * B, limbs[p], B); });
    SFG_CUDA(c, cudaMemcpyAsync(d_dst, pin, np * B, cudaMemcpyHostToDevice, c->stream));
    return 0;  // stream-ordered; the caller's slices are no longer referenced
}

void scatter_host_to_limbs(const uint64_t *src, uint64_t *const *limbs, size_t np, size_t N) {
    const size_t B = N * 8;
    par_limbs(np, B, [=](size_t p) { memcpy(limbs[p], src + p * N, B); });
}

int scatter_device_to_limbs(Ctx *c, const uint64_t *d_src, uint64_t *const *limbs, size_t np) {
    const size_t N = c->N, B = N * 8;
    if (np == 0) return 0;
    if (is_device_ptr(limbs[0])) {
        for (size_t p = 0; p < np; p++) SFG_CUDA(c, cudaMemcpyAsync(limbs[p], d_src + p * N, B, cudaMemcpyDefault, c->stream));
        SFG_CUDA(c, cudaStreamSynchronize(c->stream));
        return 0;
    }
    void *pin;
    if (pinned_get(c, PIN_OUT, np * B, &pin)) return -1;
    SFG_CUDA(c, cudaMemcpyAsync(pin, d_src, np * B, cudaMemcpyDeviceToHost, c->stream));
    SFG_CUDA(c, cudaStreamSynchronize(c->stream));
    scatter_host_to_limbs((const uint64_t *)pin, limbs, np, N);
    return 0;
}

}  // namespace sfg
