// Host build of csrc/fz_keccak_lanes.cuh on an EMULATED WARP: 32 host threads are the 32 lanes, a warp shuffle is "publish my value, barrier, read
// the source lane's value, barrier".  The device functions run unmodified.  Test infrastructure only (tests/test_keccak_lanes_host.py); the library
// never sees this file.
#include <pthread.h>

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __device__
#define __forceinline__ inline
#define __constant__
#define __restrict__

namespace warp_emu {
pthread_barrier_t bar;
uint32_t slot[32];
thread_local int lane = 0;
inline uint32_t shfl(uint32_t v, int src) {
    slot[lane] = v;
    pthread_barrier_wait(&bar);
    const uint32_t r = slot[src & 31];
    pthread_barrier_wait(&bar);
    return r;
}
void run(const std::function<void(int)>& body) {
    pthread_barrier_init(&bar, nullptr, 32);
    std::vector<std::thread> th;
    for (int l = 0; l < 32; ++l) th.emplace_back([l, &body] { lane = l; body(l); });
    for (auto& t : th) t.join();
    pthread_barrier_destroy(&bar);
}
}  // namespace warp_emu

static inline uint32_t __shfl_sync(unsigned, uint32_t v, int src) { return warp_emu::shfl(v, src); }
static inline uint32_t __shfl_xor_sync(unsigned, uint32_t v, int m) { return warp_emu::shfl(v, warp_emu::lane ^ m); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t n) {   // high word of (hi:lo) << (n & 31)
    n &= 31;
    return n ? (hi << n) | (lo >> (32 - n)) : hi;
}

#include "../../fusion_cryptography_b200/csrc/fz_keccak_lanes.cuh"

extern "C" {

// reference: the one-thread permutation of fz_keccak.cuh on n_states states of 25 lanes
void h_keccak_f(unsigned long long* states, size_t n_states) {
    for (size_t s = 0; s < n_states; ++s) {
        unsigned long long a[25];
        std::memcpy(a, states + 25 * s, sizeof(a));
        fz::keccak_f(a);
        std::memcpy(states + 25 * s, a, sizeof(a));
    }
}

// warpag: one state on the warp, lane L < 25 holds Keccak lane L.  variant 0: keccak_f_warp (loop), 1: keccak_f_warp_u<24> (the shipped, unrolled one),
// 2: keccak_f_warp_u<1>.  `times` permutations in a row.
void h_keccak_f_warp(unsigned long long* state, int variant, int times) {
    warp_emu::run([&](int L) {
        const fz::warpag::Lane c = fz::warpag::lane_init(L);
        uint32_t lo = L < 25 ? (uint32_t)state[L] : 0xDEAD0000u + L, hi = L < 25 ? (uint32_t)(state[L] >> 32) : 0xBEEF0000u + L;   // idle lanes hold garbage
        for (int t = 0; t < times; ++t) {
            if (variant == 0) fz::warpag::keccak_f_warp(lo, hi, c, L);
            else if (variant == 1) fz::warpag::keccak_f_warp_u<24>(lo, hi, c, L);
            else fz::warpag::keccak_f_warp_u<1>(lo, hi, c, L);
        }
        if (L < 25) state[L] = ((unsigned long long)hi << 32) | lo;
    });
}

// pairag: 16 states on the warp, lanes 2k / 2k + 1 hold the low / high words of state k.  `times` permutations in a row.
void h_keccak_f_pair(unsigned long long* states16, int times) {
    warp_emu::run([&](int L) {
        const uint32_t hi = (uint32_t)(L & 1);
        unsigned long long* st = states16 + 25 * (L >> 1);
        uint32_t a[25];
        for (int k = 0; k < 25; ++k) a[k] = hi ? (uint32_t)(st[k] >> 32) : (uint32_t)st[k];
        for (int t = 0; t < times; ++t) fz::pairag::keccak_f_pair(a, hi);
        pthread_barrier_wait(&warp_emu::bar);      // everybody has read its half before anybody writes
        for (int k = 0; k < 25; ++k) {
            uint32_t* w = reinterpret_cast<uint32_t*>(st + k);
            w[hi] = a[k];                           // little-endian host: word 0 = low half
        }
    });
}

}  // extern "C"
