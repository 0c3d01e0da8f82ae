// Host build of csrc/fz_mt.cuh: MT19937 seeding (init_by_array in its three forms), output regeneration and the per-thread draw loops of both sampler
// kernels, run one emulated thread after the other on state columns laid out as the kernels lay them out.  Test infrastructure only
// (tests/test_sampler_host.py); the library never sees this file.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((unsigned long long)a * b) >> 32); }
static inline int32_t __mulhi(int32_t a, int32_t b) { return (int32_t)(((long long)a * b) >> 32); }
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
using std::min;
#undef __constant__
#define __constant__
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

#include "../../fusion_cryptography_b200/csrc/fz_mt.cuh"

namespace {
bool g_ready = false;
void init() {
    if (!g_ready) { fz::mt_build_seed_table(fz::c_mt0); g_ready = true; }
}

template <int STAGE>
void dense(const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, int d, uint32_t bound, int32_t* out) {
    using Cfg = fz::SdCfg<STAGE>;
    constexpr int G = Cfg::G;
    std::vector<uint32_t> sm((size_t)G * fz::MT_N + 256 * Cfg::SB / 4 + 64, 0xA5A5A5A5u);   // the kernel's dynamic shared memory: G state columns + the byte tile
    uint8_t* stage = reinterpret_cast<uint8_t*>(sm.data() + (size_t)G * fz::MT_N);
    for (size_t i = 0; i < n; ++i) {
        const int g = (int)(i % G);                  // the thread's column, as in the kernel (warp * PW + lane)
        uint32_t key[8];
        int kl = klen[i] > 8 ? 8 : klen[i];
        for (int j = 0; j < 8; ++j) key[j] = j < kl ? keys[i * key_stride + j] : 0u;
        uint32_t* m = sm.data() + g;
        if (kl <= 2) fz::mt_seed2<G>(m, key[0], kl == 2 ? key[1] + 1u : key[0]);
        else fz::mt_seed_any<G>(m, key, kl);
        int32_t* o = out + i * d;
        fz::dense_draws<STAGE>(m, g, d, bound, stage, o);
        if (STAGE) {                                 // the kernel's write-out of the staged bytes
            for (int k = 0; k < d; ++k) {
                const uint32_t b = stage[k * Cfg::SB + g];
                const int32_t mg = (int32_t)(b & 0x7fu);
                o[k] = (b & 0x80u) ? -mg : mg;
            }
        }
    }
}
}  // namespace

extern "C" {

// sample_kernel: mode 0 secret-key coefficients (dense or Fisher-Yates branch), mode 1 public-matrix values
void h_sample(int mode, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, int d, uint32_t q, uint32_t bound, int weight, int32_t* out) {
    init();
    std::vector<uint32_t> sm((size_t)fz::SMP_STREAMS * fz::SMP_WORDS, 0x5A5A5A5Au);   // the kernel's dynamic shared memory
    for (size_t i = 0; i < n; ++i) {
        fz::Mt s;
        s.m = sm.data() + (i % fz::SMP_STREAMS);
        uint32_t key[8];
        int kl = klen[i] > 8 ? 8 : klen[i];
        for (int j = 0; j < 8; ++j) key[j] = j < kl ? keys[i * key_stride + j] : 0u;
        fz::mt_seed(s, key, kl);
        fz::sample_draws(s, mode, d, q, bound, weight, out + i * d);
    }
}

// sample_dense_kernel<1> (byte tile) / <0> (16-byte stores); out must be 16-byte aligned for stage 0
void h_sample_dense(int stage, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, int d, uint32_t bound, int32_t* out) {
    init();
    if (stage) dense<1>(keys, key_stride, klen, n, d, bound, out);
    else dense<0>(keys, key_stride, klen, n, d, bound, out);
}

}  // extern "C"
