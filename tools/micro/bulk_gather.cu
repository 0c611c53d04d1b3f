// Microbenchmark: how fast can an SM gather many small (64..96 B), 16-byte-aligned, randomly placed
// spans from an L2-resident table into shared memory -- with cp.async.bulk (TMA engine, one copy per
// thread, mbarrier completion) versus plain LDG.128 + STS.128 by four lanes per span.
// Decides whether the ungapped verification kernel should move its window gather off the LSU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_gather bulk_gather.cu && ./bulk_gather
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BYTES>
__global__ void __launch_bounds__(128) k_bulk(const uint8_t *__restrict__ tab, const uint32_t *__restrict__ idx, int iters, uint32_t *sink)
{
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    const int t = threadIdx.x;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    uint32_t acc = 0, phase = 0;
    const uint32_t *my = idx + ((size_t)blockIdx.x * iters) * 128;
    for (int it = 0; it < iters; ++it) {
        if (t == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(128 * BYTES));
        __syncthreads();
        const uint8_t *src = tab + (size_t)my[it * 128 + t] * 16;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sm + t * BYTES)), "l"(src), "r"(BYTES), "r"(smem_u32(&bar)) : "memory");
        // wait
        uint32_t done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
        }
        phase ^= 1;
        acc += reinterpret_cast<const uint32_t *>(sm + t * BYTES)[it & 3];
        __syncthreads();
    }
    if (acc == 0x12345678) sink[0] = acc;
}

template <int BYTES>
__global__ void __launch_bounds__(128) k_ldg(const uint8_t *__restrict__ tab, const uint32_t *__restrict__ idx, int iters, uint32_t *sink)
{
    extern __shared__ __align__(128) uint8_t sm[];
    constexpr int LPS = BYTES / 16;          // lanes per span
    const int t = threadIdx.x;
    uint32_t acc = 0;
    const uint32_t *my = idx + ((size_t)blockIdx.x * iters) * 128;
    for (int it = 0; it < iters; ++it) {
        // 128 spans per iteration, LPS lanes each: 128*LPS lane-tasks over 128 threads
#pragma unroll
        for (int r = 0; r < LPS; ++r) {
            const int task = r * 128 + t;
            const int span = task / LPS, part = task % LPS;
            const uint4 v = *reinterpret_cast<const uint4 *>(tab + (size_t)my[it * 128 + span] * 16 + part * 16);
            *reinterpret_cast<uint4 *>(sm + span * BYTES + part * 16) = v;
        }
        __syncthreads();
        acc += reinterpret_cast<const uint32_t *>(sm + t * BYTES)[it & 3];
        __syncthreads();
    }
    if (acc == 0x12345678) sink[0] = acc;
}

int main()
{
    const size_t TAB = 25u << 20;            // 25 MB: the 50 Mbp mixRef, L2 resident
    const int blocks = 148 * 8, iters = 200;
    uint8_t *tab; uint32_t *idx, *sink;
    cudaMalloc(&tab, TAB + 256); cudaMemset(tab, 1, TAB + 256);
    const size_t n = (size_t)blocks * iters * 128;
    uint32_t *h = (uint32_t *)malloc(n * 4);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (uint32_t)(s % (TAB / 16 - 8)); }
    cudaMalloc(&idx, n * 4); cudaMemcpy(idx, h, n * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&sink, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char *name, auto kern, int bytes) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * bytes);
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            kern<<<blocks, 128, 128 * bytes>>>(tab, idx, iters, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        printf("%-10s %3d B/span: %8.3f ms  %7.2f G spans/s  %7.1f GB/s  %s\n", name, bytes, best, n / best / 1e6, n * (double)bytes / best / 1e6,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    };
    run("bulk", k_bulk<64>, 64); run("bulk", k_bulk<80>, 80); run("bulk", k_bulk<96>, 96);
    run("ldg128", k_ldg<64>, 64); run("ldg128", k_ldg<80>, 80); run("ldg128", k_ldg<96>, 96);
    return 0;
}
