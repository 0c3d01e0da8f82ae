// cuda_emu.h -- a small CUDA-on-CPU emulator for the CPU test suite: the product's KERNELS (csrc/*.cu, compiled by g++ with FZ_HOST_EMULATION) and
// their host launch code run unmodified, one CTA after the other; the threads of a CTA are fibers of one host thread, __syncthreads / __syncwarp / warp
// shuffles and votes are cooperative barriers (threads which have returned from the kernel drop out of them, as on the device).  Shared memory is `static` storage
// (one CTA runs at a time); constant and device globals are plain globals; cudaMemcpyToSymbol is memcpy.  Test infrastructure only: nothing in the
// product includes this file.
#pragma once
#define FZ_HOST_EMULATION 1

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

#include <cuda_runtime.h>   // vector types (int4, uint4, dim3 ...), cudaError_t, cudaStream_t: declarations only, nothing is linked

// ---- execution-space and memory-space specifiers
#undef __global__
#define __global__
#undef __device__
#define __device__
#undef __host__
#define __host__
#undef __constant__
#define __constant__
#undef __shared__
#define __shared__ static
#define FZ_DYN_SHARED extern          // dynamic shared memory: the harness defines the arrays
#undef __forceinline__
#define __forceinline__ inline
#undef __noinline__
#define __noinline__ __attribute__((noinline))
#undef __launch_bounds__
#define __launch_bounds__(...)

namespace fz_emu {

// The threads of a CTA are FIBERS of one host thread, scheduled round-robin; a barrier is "count me in, let the others run until everybody has arrived".
// The scheduler and the context switch (a dozen instructions of x86-64 assembly) are in emu/emu_runtime.cpp.
struct Idx { unsigned x = 0, y = 0, z = 0; };
struct Bar { int expected = 0, arrived = 0; unsigned gen = 0; };
struct Fiber { void* sp = nullptr; Idx tid; bool done = false; };
struct Cta {
    int n_threads = 0;
    Bar block_bar;
    Bar warp_bar[32];
    uint32_t slot[1024];     // one exchange word per thread
};
extern Cta g_cta;            // the CTA that is running (one at a time)
extern Fiber* g_cur;         // the thread that is running
extern Idx g_blockIdx, g_blockDim, g_gridDim;
void bar_wait(Bar& b);       // arrive and wait; threads that have left the kernel no longer count (as exited threads on the device)
void run_cta(int n_threads, const std::function<void()>& body);

inline int lane() { return (int)(g_cur->tid.x & 31u); }
inline int warp() { return (int)(g_cur->tid.x >> 5); }
inline void sync_warp() { bar_wait(g_cta.warp_bar[warp()]); }
inline void sync_block() { bar_wait(g_cta.block_bar); }

inline uint32_t shfl(uint32_t v, int src) {
    uint32_t* s = g_cta.slot + 32 * warp();
    s[lane()] = v;
    sync_warp();
    const uint32_t r = s[src & 31];
    sync_warp();
    return r;
}

// grid / block: unsigned, int or dim3
inline Idx to_idx(unsigned v) { Idx i; i.x = v; i.y = 1; i.z = 1; return i; }
inline Idx to_idx(int v) { return to_idx((unsigned)v); }
inline Idx to_idx(const dim3& d) { Idx i; i.x = d.x; i.y = d.y; i.z = d.z; return i; }

template <typename G, typename B>
void launch(G grid, B block, size_t /*dynamic shared memory: the arrays are defined by the harness*/, const std::function<void()>& body) {
    g_gridDim = to_idx(grid);
    g_blockDim = to_idx(block);
    const int n = (int)(g_blockDim.x * g_blockDim.y * g_blockDim.z);
    for (unsigned bz = 0; bz < g_gridDim.z; ++bz)
        for (unsigned by = 0; by < g_gridDim.y; ++by)
            for (unsigned bx = 0; bx < g_gridDim.x; ++bx) {
                g_blockIdx.x = bx; g_blockIdx.y = by; g_blockIdx.z = bz;
                run_cta(n, body);
            }
}

}  // namespace fz_emu

#define threadIdx (::fz_emu::g_cur->tid)
#define blockIdx (::fz_emu::g_blockIdx)
#define blockDim (::fz_emu::g_blockDim)
#define gridDim (::fz_emu::g_gridDim)

// ---- synchronisation, shuffles, votes (full masks only, as the product uses them)
static inline void __syncthreads() { ::fz_emu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { ::fz_emu::sync_warp(); }
template <typename T> static inline T fz_emu_shfl_t(T v, int src) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "shuffle of a 4- or 8-byte value");
    if constexpr (sizeof(T) == 4) {
        uint32_t w; std::memcpy(&w, &v, 4);
        w = ::fz_emu::shfl(w, src);
        T r; std::memcpy(&r, &w, 4); return r;
    } else {
        uint32_t w[2]; std::memcpy(w, &v, 8);
        w[0] = ::fz_emu::shfl(w[0], src); w[1] = ::fz_emu::shfl(w[1], src);
        T r; std::memcpy(&r, w, 8); return r;
    }
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) { return fz_emu_shfl_t(v, src); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return fz_emu_shfl_t(v, ::fz_emu::lane() ^ m); }
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) { return fz_emu_shfl_t(v, ::fz_emu::lane() >= d ? ::fz_emu::lane() - d : ::fz_emu::lane()); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) { return fz_emu_shfl_t(v, ::fz_emu::lane() + d < 32 ? ::fz_emu::lane() + d : ::fz_emu::lane()); }
static inline int __any_sync(unsigned, int p) {
    int r = 0;
    for (int l = 0; l < 32; ++l) r |= fz_emu_shfl_t(p ? 1 : 0, l);   // every lane reads every lane: simple, and the same on all lanes
    return r;
}
static inline int __all_sync(unsigned, int p) {
    int r = 1;
    for (int l = 0; l < 32; ++l) r &= fz_emu_shfl_t(p ? 1 : 0, l);
    return r;
}

// ---- arithmetic intrinsics
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((unsigned long long)a * b) >> 32); }
static inline int32_t __mulhi(int32_t a, int32_t b) { return (int32_t)(((long long)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t n) {   // high word of (hi:lo) << (n & 31)
    n &= 31;
    return n ? (hi << n) | (lo >> (32 - n)) : hi;
}
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline long long clock64() { return 0; }
using std::max;
using std::min;

// ---- atomics (on shared or global memory alike)
template <typename T, typename V> static inline T atomicOr(T* p, V v) { return __atomic_fetch_or(p, (T)v, __ATOMIC_SEQ_CST); }
template <typename T, typename V> static inline T atomicAnd(T* p, V v) { return __atomic_fetch_and(p, (T)v, __ATOMIC_SEQ_CST); }
template <typename T, typename V> static inline T atomicAdd(T* p, V v) { return __atomic_fetch_add(p, (T)v, __ATOMIC_SEQ_CST); }
template <typename T, typename V> static inline T atomicMax(T* p, V v) {
    T old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < (T)v && !__atomic_compare_exchange_n(p, &old, (T)v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}

// ---- memory intrinsics: cache hints mean nothing here
template <typename T> static inline T __ldg(const T* p) { return *p; }
template <typename T> static inline T __ldcs(const T* p) { return *p; }
template <typename T> static inline void __stcs(T* p, const T& v) { *p = v; }

// ---- the few runtime calls the library makes: there is one "device", its memory is the host's
#define cudaFuncSetAttribute(...) cudaSuccess
#define cudaGetLastError() cudaSuccess
namespace fz_emu {
constexpr int SM_COUNT = 2;     // small grids: every CTA is a set of host threads
inline cudaError_t get_device_count(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t get_device_properties(cudaDeviceProp* p, int) { std::memset(p, 0, sizeof(*p)); p->multiProcessorCount = SM_COUNT; p->major = 10; return cudaSuccess; }
inline cudaError_t device_get_attribute(int* v, cudaDeviceAttr, int) { *v = 227 * 1024; return cudaSuccess; }
inline cudaError_t dev_malloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t dev_free(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t memcpy_async(void* d, const void* s, size_t n) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t stream_create(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
template <typename K> inline cudaError_t occupancy(int* occ, K, int, size_t) { *occ = 1; return cudaSuccess; }
}  // namespace fz_emu
#define cudaGetDeviceCount(n) ::fz_emu::get_device_count(n)
#define cudaSetDevice(d) cudaSuccess
#define cudaGetDeviceProperties(p, d) ::fz_emu::get_device_properties((p), (d))
#define cudaDeviceGetAttribute(v, a, d) ::fz_emu::device_get_attribute((v), (a), (d))
#define cudaMalloc(p, n) ::fz_emu::dev_malloc((void**)(p), (n))
#define cudaFree(p) ::fz_emu::dev_free((void*)(p))
#define cudaMemcpyAsync(d, s, n, kind, st) ::fz_emu::memcpy_async((d), (s), (n))
#define cudaMemcpy(d, s, n, kind) ::fz_emu::memcpy_async((d), (s), (n))
#define cudaStreamCreateWithFlags(s, f) ::fz_emu::stream_create(s)
#define cudaStreamCreate(s) ::fz_emu::stream_create(s)
#define cudaStreamSynchronize(s) cudaSuccess
#define cudaDeviceSynchronize() cudaSuccess
#define cudaStreamDestroy(s) cudaSuccess
#define cudaGetErrorString(e) "emulated CUDA runtime"
#define cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k, t, s) ::fz_emu::occupancy((occ), (k), (t), (s))
template <typename S> static inline cudaError_t fz_emu_to_symbol(S& symbol, const void* src, size_t count, size_t offset = 0) {
    std::memcpy(reinterpret_cast<char*>(&symbol) + offset, src, count);
    return cudaSuccess;
}
#define cudaMemcpyToSymbol(...) fz_emu_to_symbol(__VA_ARGS__)

// kernel launches of the product's host code: FZ_LAUNCH(kernel, grid, block, dynamic_smem, stream, args...)
#define FZ_LAUNCH(kernel, grid, block, smem, stream, ...) ::fz_emu::launch((grid), (block), (size_t)(smem), [&] { kernel(__VA_ARGS__); })
