// Micro-benchmark: peak issue rates of the pipes the key-switch / NTT kernels are bound by on B200 ("integer-pipe peak
// ... must be measured by a micro-benchmark before any fraction of integer roofline is quoted"), and the peak BUTTERFLY rates of the
// four arithmetic classes of ntt2.cuh evaluated from registers (no memory traffic): the roofline the transform kernels are quoted against.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I sfgwas_b200/csrc -o pipe_peaks profiles/microbench/pipe_peaks.cu sfgwas_b200/build/hostmath.cpp.o -lquadmath
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "ntt2.cuh"
using namespace sfg;

constexpr int ILP = 8, ITER = 4096;

template <int MODE>
__global__ void __launch_bounds__(512) k_ops(uint64_t *out, uint32_t seed) {
    uint32_t a[ILP], b = seed | 1u;
    uint64_t w[ILP];
    double d[ILP], e = 1.0000001 + seed * 1e-9;
    for (int i = 0; i < ILP; i++) {
        a[i] = threadIdx.x * 2654435761u + i;
        w[i] = a[i];
        d[i] = (double)a[i];
    }
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE == 0) a[i] = a[i] * b + 12345u;                                                         // IMAD (mad.lo.u32)
            if (MODE == 1) w[i] = (uint64_t)(uint32_t)w[i] * b + w[i];                                       // IMAD.WIDE.U32 (mad.wide.u32)
            if (MODE == 2) d[i] = __fma_rn(d[i], e, 0.5);                                                    // DFMA
            if (MODE == 3) a[i] = __umulhi(a[i], b) + 7u;                                                    // IMAD.HI.U32
            if (MODE == 4) a[i] = min(a[i] + b, a[i] - b);                                                   // IADD3 + IMNMX (ALU pipe)
        }
    }
    uint64_t s = 0;
    for (int i = 0; i < ILP; i++) s ^= a[i] ^ w[i] ^ (uint64_t)__double_as_longlong(d[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ILP independent butterflies per thread per iteration with the class's forward butterfly (the inner operation of every transform)
template <class A>
__global__ void __launch_bounds__(512) k_bfly(uint64_t *out, LimbConst lc, typename A::TW tw) {
    using T = typename A::T;
    const typename A::C c = A::make(lc);
    T x[ILP], y[ILP];
    for (int i = 0; i < ILP; i++) {
        x[i] = A::from_canon((threadIdx.x * 977u + i) % lc.q, c);
        y[i] = A::from_canon((threadIdx.x * 131u + 7 * i + 1) % lc.q, c);
    }
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            A::fwd(x[i], y[i], tw, c);
            if (A::kKind == kArD && (it & 7) == 7) { x[i] = ArD::red((double)x[i], *(const ArD::C *)&c); y[i] = ArD::red((double)y[i], *(const ArD::C *)&c); }
            if (A::kKind == kArW && (it & 15) == 15) { x[i] = A::canon(x[i], c); y[i] = A::canon(y[i], c); }
        }
    }
    uint64_t s = 0;
    for (int i = 0; i < ILP; i++) s ^= (uint64_t)A::canon(x[i], c) ^ (uint64_t)A::canon(y[i], c);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static double time_ms(F launch) {
    launch();
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

static LimbConst mk(uint64_t q) {
    LimbConst lc{};
    lc.q = q;
    uint64_t inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - q * inv;
    lc.qinv = inv;
    lc.bred_hi = (uint64_t)((((unsigned __int128)1) << 127) / q * 2 >> 64);
    lc.bred_hi = (uint64_t)((~(unsigned __int128)0) / q >> 64);
    lc.bred_lo = (uint64_t)((~(unsigned __int128)0) / q);
    lc.ninv = 1;
    lc.ninv_sh = 0;
    return lc;
}

int main() {
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    uint64_t *out;
    cudaMalloc(&out, (size_t)nsm * 4 * 512 * 8);
    const int blocks = nsm * 4, threads = 512;
    const double ops = (double)blocks * threads * ITER * ILP;
    printf("B200: %d SMs, max SM clock %.0f MHz.  lanes/clk/SM = ops / time / SMs / clock\n", nsm, clk / 1e3);
    const char *names[] = {"IMAD (mad.lo.u32)", "IMAD.WIDE.U32", "DFMA", "IMAD.HI.U32", "IADD3+IMNMX (2 ops)"};
    double ms;
    ms = time_ms([&] { k_ops<0><<<blocks, threads>>>(out, 3); });
    printf("%-22s %8.3f ms  %7.2f Tops/s  %6.1f lanes/clk/SM\n", names[0], ms, ops / ms / 1e9, ops / ms / nsm / (double)clk);
    ms = time_ms([&] { k_ops<1><<<blocks, threads>>>(out, 3); });
    printf("%-22s %8.3f ms  %7.2f Tops/s  %6.1f lanes/clk/SM\n", names[1], ms, ops / ms / 1e9, ops / ms / nsm / (double)clk);
    ms = time_ms([&] { k_ops<2><<<blocks, threads>>>(out, 3); });
    printf("%-22s %8.3f ms  %7.2f Tops/s  %6.1f lanes/clk/SM\n", names[2], ms, ops / ms / 1e9, ops / ms / nsm / (double)clk);
    ms = time_ms([&] { k_ops<3><<<blocks, threads>>>(out, 3); });
    printf("%-22s %8.3f ms  %7.2f Tops/s  %6.1f lanes/clk/SM\n", names[3], ms, ops / ms / 1e9, ops / ms / nsm / (double)clk);
    ms = time_ms([&] { k_ops<4><<<blocks, threads>>>(out, 3); });
    printf("%-22s %8.3f ms  %7.2f Tops/s  %6.1f lanes/clk/SM\n", names[4], ms, 2 * ops / ms / 1e9, 2 * ops / ms / nsm / (double)clk);
    printf("\nforward butterflies from registers (the roofline of the transform kernels), Gbutterfly/s over the whole GPU:\n");
    const LimbConst l30 = mk(0x3FFC0001ULL), l31 = mk(0x40020001ULL), ld33 = mk(0x1FFFEC001ULL), ld36 = mk(0x800004001ULL), lw46 = mk(0x200000008001ULL);
    ms = time_ms([&] { k_bfly<ArN30><<<blocks, threads>>>(out, l30, ArN30::make_tw(12345678, l30.q)); });
    printf("%-34s %8.3f ms  %8.1f Gbfly/s\n", "ArN30 (q < 2^30, u32 Shoup lazy)", ms, ops / ms / 1e6);
    ms = time_ms([&] { k_bfly<ArN31><<<blocks, threads>>>(out, l31, ArN31::make_tw(12345678, l31.q)); });
    printf("%-34s %8.3f ms  %8.1f Gbfly/s\n", "ArN31 (q < 2^31, u32 canonical)", ms, ops / ms / 1e6);
    ms = time_ms([&] { k_bfly<ArD><<<blocks, threads>>>(out, ld33, ArD::make_tw(12345678, ld33.q)); });
    printf("%-34s %8.3f ms  %8.1f Gbfly/s\n", "ArD 33-bit (FP64 exact integers)", ms, ops / ms / 1e6);
    ms = time_ms([&] { k_bfly<ArD><<<blocks, threads>>>(out, ld36, ArD::make_tw(12345678, ld36.q)); });
    printf("%-34s %8.3f ms  %8.1f Gbfly/s\n", "ArD 36-bit (FP64 exact integers)", ms, ops / ms / 1e6);
    ms = time_ms([&] { k_bfly<ArW><<<blocks, threads>>>(out, lw46, ArW::make_tw(12345678, lw46.q)); });
    printf("%-34s %8.3f ms  %8.1f Gbfly/s\n", "ArW 46-bit (u64 Shoup lazy)", ms, ops / ms / 1e6);
    cudaFree(out);
    return 0;
}
