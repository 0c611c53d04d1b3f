// mixref.cu -- SNP-aware reference construction on the device.
// Restates build_mixRef (Index_src/mixRef.c:93-197): every base becomes a 4-bit allele mask
// (A=1 C=2 G=4 T=8, anything else 0; table at mixRef.c:36-53), 8 bases per uint32, base p in
// bits 4*(p%8)..; then each SNP row ORs its allele mask into its position (mixRef.c:147-151).
#if !defined(SALT_EMUL)
#include <cuda_runtime.h>
#endif

#include "common.cuh"
#include "kernels.h"

namespace salt {

__device__ __forceinline__ uint32_t base_mask(unsigned char ch)
{
    switch (ch) {
    case 'A': case 'a': return 1u;
    case 'C': case 'c': return 2u;
    case 'G': case 'g': return 4u;
    case 'T': case 't': return 8u;
    default: return 0u;
    }
}

__global__ void __launch_bounds__(256)
mixref_pack_kernel(const char *__restrict__ bases, uint32_t l, uint32_t *__restrict__ words, size_t n_words)
{
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const size_t p = w * 8 + b;
        if (p < l) v |= base_mask((unsigned char)bases[p]) << (4 * b);
    }
    words[w] = v;
}

__global__ void __launch_bounds__(256)
mixref_snp_kernel(const uint32_t *__restrict__ pos, const uint8_t *__restrict__ mask, size_t n, uint32_t l,
                  uint32_t *__restrict__ words)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = pos[i];
    if (p < l) atomicOr(&words[p >> 3], (uint32_t)(mask[i] & 15u) << (4 * (p & 7u)));
}

cudaError_t launch_build_mixref(const char *bases, uint32_t l, const uint32_t *snp_pos, const uint8_t *snp_mask,
                                size_t n_snp, uint32_t *words, cudaStream_t st)
{
    const size_t n_words = ((size_t)l + 7) / 8;
    if (n_words) SALT_LAUNCH(mixref_pack_kernel, (unsigned)((n_words + 255) / 256), 256, 0, st, bases, l, words, n_words);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (n_snp) SALT_LAUNCH(mixref_snp_kernel, (unsigned)((n_snp + 255) / 256), 256, 0, st, snp_pos, snp_mask, n_snp, l, words);
    return cudaGetLastError();
}

}  // namespace salt
