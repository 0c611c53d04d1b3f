// transport.cu -- the compact host->device chunk format (salt_packed_chunk_t, include/salt_b200.h).
//
// A chunk's reads arrive 2 or 4 bits per base and its CSR offsets as per-read counts; these kernels
// rebuild on the device exactly what salt_b200_set_reads + salt_cands_t would have uploaded: the
// byte codes 0..4 (query.c:177-181), the read offsets and the two candidate-offset arrays.  The PCIe
// link, not HBM, bounds the end-to-end verification stage, so bytes saved here are time.
//
//   scan3_*        exclusive prefix sums of up to three count arrays in one launch set (16/32-bit counts, or a
//                  uniform value): block sums -> scan of the block sums -> block-local scan + base
//   unpack_bases   2/4-bit base stream -> one code byte per base, four bases per thread
//   patch_n        2-bit streams carry N as a side list of stream positions
#if !defined(SALT_EMUL)
#include <cuda_runtime.h>
#endif

#include "common.cuh"
#include "kernels.h"

namespace salt {

constexpr int SCAN_T = 256;            // threads per block
constexpr int SCAN_E = 8;              // elements per thread
constexpr int SCAN_TILE = SCAN_T * SCAN_E;

__device__ __forceinline__ uint32_t scan_in(const Scan3 &a, int k, size_t i)
{
    if (i >= a.n) return 0u;
    if (!a.in[k]) return a.uniform[k];
    return a.width[k] == 16 ? (uint32_t)static_cast<const uint16_t *>(a.in[k])[i] : static_cast<const uint32_t *>(a.in[k])[i];
}

// inclusive scan of one value per thread across the block (Hillis-Steele in shared memory); returns the
// thread's inclusive prefix, *total = the block's sum
__device__ __forceinline__ uint32_t block_scan(uint32_t v, uint32_t *sh, uint32_t *total)
{
    const int t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
    for (int o = 1; o < SCAN_T; o <<= 1) {
        const uint32_t add = t >= o ? sh[t - o] : 0u;
        __syncthreads();
        sh[t] += add;
        __syncthreads();
    }
    const uint32_t r = sh[t];
    *total = sh[SCAN_T - 1];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(SCAN_T)
scan3_sums_kernel(Scan3 a)
{
    __shared__ uint32_t sh[SCAN_T];
    const int k = blockIdx.y;
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_E;
    uint32_t v = 0;
#pragma unroll
    for (int e = 0; e < SCAN_E; ++e) v += scan_in(a, k, base + e);
    uint32_t total;
    block_scan(v, sh, &total);
    if (threadIdx.x == 0) a.partial[(size_t)k * a.n_blocks + blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_T)
scan3_top_kernel(Scan3 a)
{
    __shared__ uint32_t sh[SCAN_T];
    const int k = blockIdx.x;
    uint32_t *p = a.partial + (size_t)k * a.n_blocks;
    uint32_t carry = 0;
    for (uint32_t b0 = 0; b0 < a.n_blocks; b0 += SCAN_T) {
        const uint32_t i = b0 + threadIdx.x;
        const uint32_t v = i < a.n_blocks ? p[i] : 0u;
        uint32_t total;
        const uint32_t inc = block_scan(v, sh, &total);
        if (i < a.n_blocks) p[i] = carry + inc - v;              // exclusive
        carry += total;
    }
    if (threadIdx.x == 0) a.out[k][a.n] = carry;
}

__global__ void __launch_bounds__(SCAN_T)
scan3_final_kernel(Scan3 a)
{
    __shared__ uint32_t sh[SCAN_T];
    const int k = blockIdx.y;
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_E;
    uint32_t x[SCAN_E], v = 0;
#pragma unroll
    for (int e = 0; e < SCAN_E; ++e) { x[e] = scan_in(a, k, base + e); v += x[e]; }
    uint32_t total;
    const uint32_t inc = block_scan(v, sh, &total);
    uint32_t run = a.partial[(size_t)k * a.n_blocks + blockIdx.x] + inc - v;
    uint32_t *out = a.out[k];
#pragma unroll
    for (int e = 0; e < SCAN_E; ++e) {
        if (base + e < a.n) out[base + e] = run;
        run += x[e];
    }
}

// Output base q of the chunk is input stream position q + phase: `bits` bits each, lowest bits first within a byte.
__global__ void __launch_bounds__(256)
unpack_bases_kernel(const uint8_t *__restrict__ in, uint32_t phase, int bits, size_t n_bases, uint8_t *__restrict__ codes)
{
    const size_t q0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (q0 >= n_bases) return;
    const size_t p0 = q0 + phase;
    uint32_t w = 0;
    if (bits == 2) {
        const size_t b = p0 >> 2;
        const uint32_t raw = (uint32_t)in[b] | ((uint32_t)in[b + 1] << 8);          // buffers carry slack past the end
        const uint32_t v = raw >> (2u * (uint32_t)(p0 & 3));
#pragma unroll
        for (int e = 0; e < 4; ++e) w |= ((v >> (2 * e)) & 3u) << (8 * e);
    } else {
        const size_t b = p0 >> 1;
        const uint32_t raw = (uint32_t)in[b] | ((uint32_t)in[b + 1] << 8) | ((uint32_t)in[b + 2] << 16);
        const uint32_t v = raw >> (4u * (uint32_t)(p0 & 1));
#pragma unroll
        for (int e = 0; e < 4; ++e) { uint32_t c = (v >> (4 * e)) & 15u; c = c > 4u ? 4u : c; w |= c << (8 * e); }
    }
    *reinterpret_cast<uint32_t *>(codes + q0) = w;       // codes is allocated in whole words
}

__global__ void __launch_bounds__(256)
patch_n_kernel(const uint32_t *__restrict__ n_pos, size_t n_n, uint32_t origin, size_t n_bases, uint8_t *__restrict__ codes)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_n) return;
    const uint32_t q = n_pos[i] - origin;
    if (q < n_bases) codes[q] = 4;
}

// The CIGARs of a chunk's gapped primaries sit in stride-byte rows (the caller's stride, 128 in the reference); nearly all
// are a dozen characters.  The rows that go back to the host without waiting for their count are slimmed to 32 bytes
// (0xFF in byte 0: does not fit, fetched on demand), a quarter of the download.
__global__ void __launch_bounds__(256)
cig_slim_kernel(const char *__restrict__ rows, int stride, const uint32_t *__restrict__ count, uint32_t eager, char *__restrict__ slim)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= eager || i >= *count) return;
    const char *__restrict__ src = rows + (size_t)i * (size_t)stride;
    char *__restrict__ dst = slim + (size_t)i * 32;
    int len = 0;
    while (len < stride - 1 && len < 32 && src[len] != '\0') ++len;
    if (len >= 32) { dst[0] = (char)0xFF; return; }
    for (int k = 0; k < len; ++k) dst[k] = src[k];
    dst[len] = '\0';
}

cudaError_t launch_cig_slim(const char *rows, int stride, const uint32_t *count, uint32_t eager, char *slim, cudaStream_t st)
{
    if (!eager) return cudaSuccess;
    SALT_LAUNCH(cig_slim_kernel, (eager + 255) / 256, 256, 0, st, rows, stride, count, eager, slim);
    return cudaGetLastError();
}

uint32_t scan3_blocks(size_t n) { return (uint32_t)((n + SCAN_TILE - 1) / SCAN_TILE); }

cudaError_t launch_scan3(Scan3 a, int n_arrays, cudaStream_t st)
{
    if (n_arrays < 1 || n_arrays > 3) return cudaErrorInvalidValue;
    a.n_blocks = scan3_blocks(a.n);
    if (a.n_blocks == 0) {                              // empty chunk: just the terminating zero
        for (int k = 0; k < n_arrays; ++k) {
            cudaError_t e = cudaMemsetAsync(a.out[k], 0, 4, st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    SALT_LAUNCH(scan3_sums_kernel, dim3(a.n_blocks, (unsigned)n_arrays), SCAN_T, 0, st, a);
    SALT_LAUNCH(scan3_top_kernel, (unsigned)n_arrays, SCAN_T, 0, st, a);
    SALT_LAUNCH(scan3_final_kernel, dim3(a.n_blocks, (unsigned)n_arrays), SCAN_T, 0, st, a);
    return cudaGetLastError();
}

cudaError_t launch_unpack_bases(const uint8_t *in, uint32_t phase, int bits, size_t n_bases, uint8_t *codes,
                                const uint32_t *n_pos, size_t n_n, uint32_t origin, cudaStream_t st)
{
    if (!n_bases) return cudaSuccess;
    const size_t threads = (n_bases + 3) / 4;
    SALT_LAUNCH(unpack_bases_kernel, (unsigned)((threads + 255) / 256), 256, 0, st, in, phase, bits, n_bases, codes);
    if (n_n) SALT_LAUNCH(patch_n_kernel, (unsigned)((n_n + 255) / 256), 256, 0, st, n_pos, n_n, origin, n_bases, codes);
    return cudaGetLastError();
}

}  // namespace salt
