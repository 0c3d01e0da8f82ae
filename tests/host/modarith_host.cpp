// Host build of the product's arithmetic headers (fz_common.cuh, fz_keccak.cuh) for the CPU test suite: the device intrinsics they use are
// defined here in plain C++, everything else is the header's own code.  Test infrastructure only (tests/test_modarith_host.py); the library
// never sees this file.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((unsigned long long)a * b) >> 32); }
static inline int32_t __mulhi(int32_t a, int32_t b) { return (int32_t)(((long long)a * b) >> 32); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t n) {   // high word of (hi:lo) << (n & 31)
    n &= 31;
    return n ? (hi << n) | (lo >> (32 - n)) : hi;
}
using std::min;
#undef __constant__
#define __constant__

#include "../../fusion_cryptography_b200/csrc/fz_common.cuh"
#include "../../fusion_cryptography_b200/csrc/fz_keccak.cuh"

extern "C" {

uint32_t h_q() { return fz::Q; }

// op: 0 sh_mul(a, w, wq)  1 pl_mul(a, lo, hi)  2 pl_mul_cent(a, lo, hi)  3 red_q(a)  4 mont_mul(a, (int32) b)  5 to_lazy((int32) a)  6 to_red((int32) a)
//     7 cent_u(a)  8 mul_cent((int32) a, (int32) b)
void h_op32(int op, const uint32_t* a, const uint32_t* b, const uint32_t* c, int64_t* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        switch (op) {
            case 0: out[i] = fz::sh_mul(a[i], b[i], c[i]); break;
            case 1: out[i] = fz::pl_mul(a[i], b[i], c[i]); break;
            case 2: out[i] = fz::pl_mul_cent(a[i], b[i], c[i]); break;
            case 3: out[i] = fz::red_q(a[i]); break;
            case 4: out[i] = fz::mont_mul(a[i], (int32_t)b[i]); break;
            case 5: out[i] = fz::to_lazy((int32_t)a[i]); break;
            case 6: out[i] = fz::to_red((int32_t)a[i]); break;
            case 7: out[i] = fz::cent_u(a[i]); break;
            case 8: out[i] = fz::mul_cent((int32_t)a[i], (int32_t)b[i]); break;
            default: out[i] = 0;
        }
    }
}

// op: 0 cent_i64  1 fold_i64
void h_op64(int op, const int64_t* a, int64_t* out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = op == 0 ? (int64_t)fz::cent_i64((long long)a[i]) : (int64_t)fz::fold_i64((long long)a[i]);
}

// SHA3-256 (suffix 0x06) / SHAKE256 (suffix 0x1f) over the header's keccak_f, rate 136: the sponge every hashing kernel of the library builds around it
void h_sponge(const uint8_t* msg, size_t len, uint8_t suffix, uint8_t* out, size_t out_len) {
    unsigned long long a[25];
    for (int k = 0; k < 25; ++k) a[k] = 0;
    uint8_t blk[fz::KECCAK_RATE];
    size_t off = 0;
    for (;;) {
        const size_t take = len - off < (size_t)fz::KECCAK_RATE ? len - off : (size_t)fz::KECCAK_RATE;
        std::memset(blk, 0, sizeof(blk));
        std::memcpy(blk, msg + off, take);
        const bool last = take < (size_t)fz::KECCAK_RATE;
        if (last) { blk[take] ^= suffix; blk[fz::KECCAK_RATE - 1] ^= 0x80; }
        for (int k = 0; k < fz::KECCAK_RATE_LANES; ++k) {
            unsigned long long w;
            std::memcpy(&w, blk + 8 * k, 8);
            a[k] ^= w;
        }
        fz::keccak_f(a);
        off += take;
        if (last) break;
    }
    size_t produced = 0;
    while (produced < out_len) {
        const size_t take = out_len - produced < (size_t)fz::KECCAK_RATE ? out_len - produced : (size_t)fz::KECCAK_RATE;
        std::memcpy(out + produced, a, take);
        produced += take;
        if (produced < out_len) fz::keccak_f(a);
    }
}

}  // extern "C"
