// common.cuh -- shared definitions for the sm_100a kernels and their host-side logic checks.
//
// Functions marked SALT_HD compile both as device code (nvcc) and as plain C++ (g++), so the
// per-diagonal / per-cell logic can be exercised on a CPU-only box by tests/emul (test
// infrastructure).  The product path is the CUDA build only.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SALT_HD __host__ __device__ __forceinline__
#define SALT_D __device__ __forceinline__
#else
#define SALT_HD inline
#endif

#include "../../include/salt_b200.h"

// Kernel launch / dynamic shared memory spelled through macros so that tests/emul (a CPU SIMT
// emulator used only by the CPU test-suite) can compile these same sources as plain C++.
#if !defined(SALT_EMUL)
#define SALT_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define SALT_DYN_SMEM(type, name) extern __shared__ type name[]
#endif

namespace salt {

// Device-visible view of everything resident in HBM.  Passed to kernels by value.
struct DevCtx {
    const uint32_t *mixref;   // 4-bit allele masks, padded with >= 64 zero bytes
    uint32_t l;               // bases in mixref
    const uint8_t *pac;       // 2-bit bases, padded; may be null
    int64_t l_pac;
    const uint64_t *rd4;      // packed reads: one-hot nibbles (N = 15), [(rid*2+strand)*W64 + w]
    const uint16_t *rd_len;   // per read
    uint32_t n_reads;
    uint32_t W64;             // 64-bit words per packed read (>= ceil(Lmax/16) + 1, last word zero)
    uint32_t l_max;           // longest read in the chunk
};

// 8 nibbles starting at nibble offset `off` of a little-endian nibble array held in 32-bit words
SALT_HD uint32_t nib8(const uint32_t *a, int off)
{
    int w = off >> 3, sh = (off & 7) * 4;
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(a[w], a[w + 1], sh);
#else
    return sh ? (a[w] >> sh) | (a[w + 1] << (32 - sh)) : a[w];
#endif
}

// bit 4i set <=> nibble i of m is zero
SALT_HD uint32_t zero_nibbles(uint32_t m)
{
    m |= m >> 1;
    m |= m >> 2;
    return ~m & 0x11111111u;
}

SALT_HD int first_set_nibble(uint32_t z)   // z != 0, bits only at 4i
{
#if defined(__CUDA_ARCH__)
    return (__ffs((int)z) - 1) >> 2;
#else
    return __builtin_ctz(z) >> 2;
#endif
}

SALT_HD int imin(int a, int b) { return a < b ? a : b; }
SALT_HD int imax(int a, int b) { return a > b ? a : b; }
// three-input unsigned minimum: one VIMNMX3 on sm_100a
SALT_HD uint32_t salt_min3u(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __vimin3_u32(a, b, c);
#else
    const uint32_t t = a < b ? a : b;
    return t < c ? t : c;
#endif
}

}  // namespace salt
