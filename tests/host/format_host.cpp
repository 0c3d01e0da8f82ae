// Host build of csrc/fz_format.cuh on an emulated warp (32 host threads; shuffles and __syncwarp through a barrier): the text one warp of the format
// kernels writes for a verification key, a challenge and the decimal prehash.  The assembly order of the pieces is format_kernel's (fz_hash.cu);
// the value lists, heads and scans are the header's own code.  Test infrastructure only (tests/test_format_host.py).
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
int slot[32];
thread_local int lane = 0;
inline int shfl(int v, int src) {
    slot[lane] = v;
    pthread_barrier_wait(&bar);
    const int r = slot[src & 31];
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

static inline int __shfl_sync(unsigned, int v, int src) { return warp_emu::shfl(v, src); }
static inline int __shfl_up_sync(unsigned, int v, int delta) { return warp_emu::shfl(v, warp_emu::lane >= delta ? warp_emu::lane - delta : warp_emu::lane); }
static inline int __shfl_xor_sync(unsigned, int v, int m) { return warp_emu::shfl(v, warp_emu::lane ^ m); }
static inline void __syncwarp() { pthread_barrier_wait(&warp_emu::bar); }

#include "../../fusion_cryptography_b200/csrc/fz_format.cuh"

namespace {
void init() {
    static bool ready = false;
    if (!ready) { fz::format_build_tables(fz::c_fmt); ready = true; }
}
}  // namespace

extern "C" {

// str(OneTimeVerificationKey) of one key: returns the length; buf must hold 11,264 bytes
int h_format_vk(int level, int d, const int32_t* vk_l, const int32_t* vk_r, uint8_t* buf) {
    init();
    int len = 0;
    warp_emu::run([&](int lane) {
        const fz::FmtConst& fc = fz::c_fmt[level];
        int o = 0;
        const char h1[] = "OneTimeVerificationKey(left_vk_hat=";
        const char h2[] = ", right_vk_hat=";
        o += fz::put_str(buf + o, h1, 35, lane);
        o += fz::warp_format_gm1(buf + o, fc, vk_l, d, lane);
        o += fz::put_str(buf + o, h2, 15, lane);
        o += fz::warp_format_gm1(buf + o, fc, vk_r, d, lane);
        if (lane == 0) buf[o] = ')';
        o += 1;
        __syncwarp();
        if (lane == 0) len = o;
    });
    return len;
}

// str(SignatureChallenge) of one challenge
int h_format_ch(int level, int d, const int32_t* c_hat, uint8_t* buf) {
    init();
    int len = 0;
    warp_emu::run([&](int lane) {
        const fz::FmtConst& fc = fz::c_fmt[level];
        const char h3[] = "SignatureChallenge(c_hat=";
        int o = fz::put_str(buf, h3, 25, lane);
        o += fz::warp_format_poly(buf + o, fc, c_hat, d, lane);
        if (lane == 0) buf[o] = ')';
        o += 1;
        __syncwarp();
        if (lane == 0) len = o;
    });
    return len;
}

// str(int.from_bytes(digest, "little")) as one lane writes it
int h_format_le256(const uint8_t* digest, uint8_t* buf) { return fz::format_le256(buf, digest); }

// the length agw_len_kernel computes for a value list without formatting it
int h_values_len(const int32_t* vals, int d) {
    int total = 0;
    for (int k = 0; k < d; ++k) total += fz::dec_len(vals[k]);
    return total + 2 * (d - 1);
}

}  // extern "C"
