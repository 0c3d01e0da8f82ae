// Host build of csrc/fz_hashch.cuh: the challenge hash exactly as one sponge thread of hash_ch_stream_kernel runs it (token generator over the key and
// the prehash, absorb, squeeze, sparse ternary decode), one emulated thread after the other on the CTA's staging tile.  Test infrastructure only
// (tests/test_hashch_host.py); the library never sees this file.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((unsigned long long)a * b) >> 32); }
static inline int32_t __mulhi(int32_t a, int32_t b) { return (int32_t)(((long long)a * b) >> 32); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t n) {   // high word of (hi:lo) << (n & 31)
    n &= 31;
    return n ? (hi << n) | (lo >> (32 - n)) : hi;
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
using std::min;
#undef __constant__
#define __constant__

#include "../../fusion_cryptography_b200/csrc/fz_hashch.cuh"

extern "C" {

// level 0: secpar 128, 1: secpar 256.  vk_l, vk_r: [n][d]; digests: [n][32]; coef_out: [n][d] (the challenge's coefficients, not yet transformed)
int h_hash_ch(int level, int d, int weight, int bpc, int bpi, uint32_t n_used, const int32_t* vk_l, const int32_t* vk_r, const uint8_t* digests, size_t n,
              int32_t* coef_out) {
    static bool ready = false;
    if (!ready) {
        if (!fz::hashstream_build_tables(fz::c_magic2, fz::c_seg, &fz::g_text[0][0])) return -1;
        ready = true;
    }
    std::vector<uint32_t> tile(2 * fz::KECCAK_RATE_LANES * fz::HS_THREADS, 0xC3C3C3C3u);   // the CTA's __shared__ staging tile
    for (size_t i = 0; i < n; ++i)
        fz::hash_ch_thread(vk_l + i * d, vk_r + i * d, digests + i * 32, level, d, weight, bpc, bpi, n_used, coef_out + i * d,
                           tile.data() + (i % fz::HS_THREADS));
    return 0;
}

}  // extern "C"
