// Host build of csrc/fz_text.cuh (the byte queue and the decimal formatter of the streaming hash kernels) for the CPU test suite.
// Test infrastructure only (tests/test_text_host.py); the library never sees this file.
#include <cstddef>
#include <cstdint>
#include <cstring>

#define __device__
#define __forceinline__ inline
#define __restrict__
#define __constant__
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t n) {   // high word of (hi:lo) << (n & 31)
    n &= 31;
    return n ? (hi << n) | (lo >> (32 - n)) : hi;
}

#include "../../fusion_cryptography_b200/csrc/fz_text.cuh"

extern "C" {

// decimal text of v, followed by ", " when sep: returns the length, text in out[0..16)
int h_fmt_value(int32_t v, int sep, uint8_t* out) {
    unsigned long long lo, hi;
    const int n = fz::fmt_value(v, sep != 0, lo, hi);
    std::memcpy(out, &lo, 8);
    std::memcpy(out + 8, &hi, 8);
    return n;
}

// decimal text of int.from_bytes(digest, "little"): returns the length, text in buf[0..96)
int h_fmt_le256(const uint8_t* digest, uint8_t* buf) { return fz::fmt_le256(digest, buf); }

// SHAKE256 over a token stream, driven exactly like a sponge thread of hash_ch_stream_kernel: tokens are pushed while fewer than 136 bytes are
// queued, full blocks are absorbed from the queue, the last block is padded by q_absorb_last.
//   kind[i] = 0: raw chunk, raw_len[i] (<= 16) bytes at raw + 16 * i;  1: value[i] with ", ";  2: value[i] without separator
void h_shake_tokens(const int32_t* kind, const int32_t* value, const uint8_t* raw, const int32_t* raw_len, size_t n_tokens, uint8_t* out, size_t out_len) {
    unsigned long long a[25];
    for (int k = 0; k < 25; ++k) a[k] = 0;
    fz::Queue q;
    fz::q_reset(q);
    size_t t = 0;
    for (;;) {
        while (fz::q_bytes(q) < fz::KECCAK_RATE && t < n_tokens) {
            unsigned long long lo = 0, hi = 0;
            int n;
            if (kind[t] == 0) {
                n = raw_len[t];
                uint8_t tmp[16] = {0};
                std::memcpy(tmp, raw + 16 * t, (size_t)n);
                std::memcpy(&lo, tmp, 8);
                std::memcpy(&hi, tmp + 8, 8);
            } else {
                n = fz::fmt_value(value[t], kind[t] == 1, lo, hi);
            }
            fz::q_push(q, lo, hi, n);
            ++t;
        }
        if (fz::q_bytes(q) >= fz::KECCAK_RATE) {
            fz::q_absorb_full(q, a);
            fz::keccak_f(a);
        } else {
            fz::q_absorb_last(q, a);
            fz::keccak_f(a);
            break;
        }
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
