// Microbenchmark: issue rate of the integer instructions the alignment kernels are made of, on all SMs.
// SURVEY.md §8(d) asks for the INT roofline to be MEASURED with the same instruction mix rather than derived
// from lane counts.  Each thread runs ILP independent dependency chains of one instruction kind.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peak int_peak.cu && ./int_peak
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ILP = 8, ITERS = 4096;

template <int KIND>
__global__ void __launch_bounds__(256) k(uint32_t *out, uint32_t seed)
{
    uint32_t a[ILP], b = seed | 0x00010001u, c = seed * 3u + 7u;
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 17u + i + seed;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (KIND == 0) a[i] = __viaddmax_s16x2(a[i], b, c);            // VIADDMNMX.S16x2
            if (KIND == 1) a[i] = __vimax3_s16x2(a[i], b, c);              // VIMNMX3.S16x2
            if (KIND == 2) a[i] = __vmaxs2(a[i], b);                       // VIMNMX.S16x2
            if (KIND == 3) a[i] = a[i] + b;                                // 32-bit add (VIADD / IADD3)
            if (KIND == 4) a[i] = __byte_perm(a[i], b, c & 0x7777u);       // PRMT
            if (KIND == 5) a[i] = (a[i] & b) ^ c;                          // LOP3
            if (KIND == 6) a[i] = __funnelshift_r(a[i], b, c);             // SHF
            if (KIND == 7) a[i] = __popc(a[i]) + b;                        // POPC (+ add)
            if (KIND == 8) { a[i] = __viaddmax_s16x2(a[i], b, c); a[i] = __vimax3_s16x2(a[i], b, c); a[i] = a[i] + b;      // the SW cell mix:
                             a[i] = __viaddmax_s16x2(a[i], b, c); a[i] = a[i] + c; a[i] = __viaddmax_s16x2(a[i], c, b);    // 3 VIADDMNMX, 1 VIMNMX3,
                             a[i] = __byte_perm(a[i], b, 0x7531u); }                                                    // 2 adds, 1 PRMT
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= a[i];
    if (s == 0x12345678u) out[0] = s;
}

int main()
{
    uint32_t *out; cudaMalloc(&out, 16);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[] = {"VIADDMNMX.S16x2", "VIMNMX3.S16x2", "VIMNMX.S16x2", "add32", "PRMT", "LOP3", "SHF", "POPC+add", "SW cell mix (7 instr)"};
    const int per[] = {1, 1, 1, 1, 1, 1, 1, 2, 7};
    auto run = [&](int kind, auto kern) {
        float best = 1e30f;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0); kern<<<blocks, 256>>>(out, 12345u + r); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
        }
        const double inst = (double)blocks * 256 * ITERS * ILP * per[kind];
        printf("%-22s %8.3f ms  %7.2f T thread-instr/s  (%.1f lanes/clk/SM at %d MHz, %d SMs)\n", names[kind], best, inst / best / 1e9,
               inst / (best * 1e-3) / p.multiProcessorCount / (p.clockRate * 1e3), p.clockRate / 1000, p.multiProcessorCount);
    };
    run(0, k<0>); run(1, k<1>); run(2, k<2>); run(3, k<3>); run(4, k<4>); run(5, k<5>); run(6, k<6>); run(7, k<7>); run(8, k<8>);
    return 0;
}
