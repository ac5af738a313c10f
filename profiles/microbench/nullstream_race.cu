// Micro-test for the root cause of the rare CUDA-vs-oracle mismatch of round 1.
//
// Hypothesis: the library's context stream is created with cudaStreamNonBlocking, but its lookup tables (key-switch base-conversion
// constants, P^-1, permutation tables, Galois keys) were uploaded with the SYNCHRONOUS cudaMemcpy from PAGEABLE host memory.  The CUDA
// runtime documents that such a copy "may return before the DMA to the final destination has completed", and the copy runs on the
// legacy default stream -- which a non-blocking stream does NOT synchronise with.  A kernel launched on the non-blocking stream right
// after the upload can therefore read the bytes the allocation held BEFORE the copy (a previous process's data, or zeros).
//
// This program poisons a buffer, uploads `bytes` with cudaMemcpy from pageable memory, immediately launches a reader on a non-blocking
// stream and counts how often the reader saw poison.  Variants: (a) as the library did it; (b) the fix: cudaMemcpyAsync on the SAME
// stream as the reader.  Optional contention (a second process hammering the GPU) widens the window.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o nullstream_race nullstream_race.cu
// Run  : ./nullstream_race [iterations] [bytes]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

__global__ void reader(const uint32_t *buf, int n, uint32_t expect, unsigned long long *stale) {
    unsigned long long bad = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) bad += buf[i] != expect;
    if (bad) atomicAdd(stale, bad);
}

int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 20000;
    const size_t bytes = argc > 2 ? (size_t)atol(argv[2]) : 15792;  // the PN13 level-5 BaseConv table of the library
    const int n = (int)(bytes / 4);
    cudaStream_t st;
    cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    uint32_t *d;
    unsigned long long *dstale, hstale;
    cudaMalloc(&d, bytes);
    cudaMalloc(&dstale, 8);
    std::vector<uint32_t> h(n);
    for (int variant = 0; variant < 2; variant++) {
        cudaMemset(dstale, 0, 8);
        cudaDeviceSynchronize();
        int hits = 0;
        for (int it = 0; it < iters; it++) {
            const uint32_t val = 0x10000u + (uint32_t)it;
            for (int i = 0; i < n; i++) h[i] = val;
            cudaMemsetAsync(d, 0xA5, bytes, st);  // poison: what the allocation "held before"
            cudaStreamSynchronize(st);
            unsigned long long before = 0;
            if (variant == 0) {
                cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice);  // legacy default stream, pageable source
            } else {
                cudaMemcpyAsync(d, h.data(), bytes, cudaMemcpyHostToDevice, st);  // the fix: ordered on the reader's stream
            }
            reader<<<1, 256, 0, st>>>(d, n, val, dstale);
            cudaStreamSynchronize(st);
            cudaDeviceSynchronize();
            cudaMemcpy(&hstale, dstale, 8, cudaMemcpyDeviceToHost);
            if (hstale != before) {
                hits++;
                cudaMemset(dstale, 0, 8);
                cudaDeviceSynchronize();
            }
        }
        printf("%s: %d of %d uploads of %zu bytes were read stale by a kernel on a non-blocking stream\n",
               variant == 0 ? "cudaMemcpy (legacy stream) + kernel on cudaStreamNonBlocking" : "cudaMemcpyAsync on the kernel's stream              ",
               hits, iters, bytes);
    }
    return 0;
}
