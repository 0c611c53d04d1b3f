// cuda_shim.h -- TEST INFRASTRUCTURE ONLY.
//
// A tiny CPU SIMT emulator: it lets tests/emul compile salt_b200/csrc/*.cu as plain C++ and
// run the kernels with one OS thread per CUDA thread, so that the kernels' control logic
// (group shuffles, ballots, worklists, systolic wavefront, tracebacks) can be checked against
// the oracle on a machine without a GPU.  Warp collectives are rendezvous points on a
// per-warp monitor; CTAs run one after another.  Nothing in salt_b200/ loads or links this;
// the product library is the nvcc build and has no CPU path.
#pragma once
#define SALT_EMUL 1
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { uint4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
static inline uint2 make_uint2(unsigned a, unsigned b) { uint2 r; r.x = a; r.y = b; return r; }

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2 };
typedef void *cudaStream_t;
enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount; };

namespace emu {

struct WarpMon {
    std::mutex m;
    std::condition_variable cv;
    uint32_t vals[32];
    uint32_t arrived = 0, ready = 0, done = 0;
};

struct CtaState {
    std::vector<WarpMon> warps;
    std::mutex bm; std::condition_variable bcv; unsigned bcount = 0, bgen = 0, nthreads = 0;
    unsigned char *dyn = nullptr;
};

extern thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
extern thread_local CtaState *t_cta;

// all lanes named in `mask` deposit v; every lane gets the whole table back
static inline void gather(unsigned mask, uint32_t v, uint32_t out[32])
{
    const unsigned tid = t_threadIdx.x;
    WarpMon &w = t_cta->warps[tid / 32];
    const unsigned bit = 1u << (tid % 32);
    std::unique_lock<std::mutex> lk(w.m);
    w.cv.wait(lk, [&] { return !(w.arrived & bit); });
    w.vals[tid % 32] = v;
    w.arrived |= bit;
    if ((w.arrived & mask) == mask) { w.ready |= mask; w.cv.notify_all(); }
    w.cv.wait(lk, [&] { return (w.ready & bit) != 0; });
    for (int i = 0; i < 32; ++i) out[i] = w.vals[i];
    w.done |= bit;
    if ((w.done & mask) == mask) { w.arrived &= ~mask; w.ready &= ~mask; w.done &= ~mask; w.cv.notify_all(); }
}

static inline void syncthreads()
{
    CtaState &c = *t_cta;
    std::unique_lock<std::mutex> lk(c.bm);
    const unsigned gen = c.bgen;
    if (++c.bcount == c.nthreads) { c.bcount = 0; ++c.bgen; c.bcv.notify_all(); }
    else c.bcv.wait(lk, [&] { return c.bgen != gen; });
}

// kernels of different host threads (several handles in one process) run one after the other: __shared__ variables are
// plain statics here
static std::mutex g_launch_mu;

template <class F>
static inline void launch(dim3 grid, dim3 block, size_t smem, F body)
{
    std::lock_guard<std::mutex> one_at_a_time(g_launch_mu);
    std::vector<unsigned char> dyn(smem + 64);
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned b = 0; b < grid.x; ++b) {
        CtaState cta;
        cta.warps = std::vector<WarpMon>((block.x + 31) / 32);
        cta.nthreads = block.x;
        cta.dyn = dyn.data();
        std::vector<std::thread> th;
        th.reserve(block.x);
        for (unsigned t = 0; t < block.x; ++t)
            th.emplace_back([&, t, b, by] {
                t_threadIdx = dim3(t); t_blockIdx = dim3(b, by); t_blockDim = block; t_gridDim = grid; t_cta = &cta;
                body();
            });
        for (auto &x : th) x.join();
    }
}

}  // namespace emu

#define threadIdx emu::t_threadIdx
#define blockIdx emu::t_blockIdx
#define blockDim emu::t_blockDim
#define gridDim emu::t_gridDim

#define SALT_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&] { kern(__VA_ARGS__); })
#define SALT_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(emu::t_cta->dyn)

// ---- warp collectives -------------------------------------------------------------
namespace emu {
// value of lane `s` (absolute lane index in the warp) for every participating lane; 4- or 8-byte T
template <class T> static inline T pick(unsigned mask, T v, int s)
{
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "4- or 8-byte shuffles only");
    uint32_t part[2] = {0, 0}, res[2] = {0, 0};
    memcpy(part, &v, sizeof(T));
    for (unsigned k = 0; k < sizeof(T) / 4; ++k) { uint32_t tab[32]; gather(mask, part[k], tab); res[k] = tab[s]; }
    T r; memcpy(&r, res, sizeof(T)); return r;
}
}  // namespace emu
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    const int lane = threadIdx.x % 32, base = lane / width * width;
    return emu::pick(mask, v, base + ((src % width) + width) % width);
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32)
{
    const int lane = threadIdx.x % 32, base = lane / width * width;
    const int s = lane - (int)d;
    return emu::pick(mask, v, s < base ? lane : s);
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32)
{
    const int lane = threadIdx.x % 32, base = lane / width * width;
    const int s = lane + (int)d;
    return emu::pick(mask, v, s >= base + width ? lane : s);
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32)
{
    const int lane = threadIdx.x % 32, base = lane / width * width;
    const int s = lane ^ x;
    return emu::pick(mask, v, (s >= base && s < base + width) ? s : lane);
}
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
    uint32_t tab[32]; emu::gather(mask, pred ? 1u : 0u, tab);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) if ((mask >> i & 1u) && tab[i]) r |= 1u << i;
    return r;
}
static inline int __reduce_max_sync(unsigned mask, int v)
{
    uint32_t tab[32]; emu::gather(mask, (uint32_t)v, tab);
    int r = (int)0x80000000;
    for (int i = 0; i < 32; ++i) if (mask >> i & 1u) r = std::max(r, (int)tab[i]);
    return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { uint32_t tab[32]; emu::gather(mask, 0, tab); }
static inline void __syncthreads() { emu::syncthreads(); }

// ---- scalar intrinsics ------------------------------------------------------------
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh)
{
    sh &= 31u; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh)
{
    sh &= 31u; return sh ? (hi << sh) | (lo >> (32 - sh)) : hi;
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
    unsigned long long src = (unsigned long long)a | ((unsigned long long)b << 32);
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((src >> (8 * ((sel >> (4 * i)) & 7u))) & 255u) << (8 * i);
    return r;
}
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
using std::min;
using std::max;

// ---- runtime API (host memory stands in for device memory) -------------------------
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
static inline const char *cudaGetErrorName(cudaError_t) { return "emulated"; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) { p->multiProcessorCount = 2; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, int) { *s = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
#ifndef SALT_EMUL_SLACK
#define SALT_EMUL_SLACK 64
#endif
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { *p = (T *)calloc(n + SALT_EMUL_SLACK, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { *p = (T *)calloc(n + SALT_EMUL_SLACK, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
typedef void *cudaEvent_t;
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = (void *)1; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
