// fz_decoder.cuh -- decode_bytes_to_polynomial_coefficients (fusion.py:422-481) as a byte-driven state machine: signum bits, `weight` magnitudes
// (int(bytes) % bound + 1), then the partial Fisher-Yates shuffle; bytes that run out leave truncated / empty slices exactly as Python's slicing
// does.  One thread owns one polynomial.  fz_hash.cu wraps it in decode_kernel / shake_decode_kernel; the CPU suite compiles the
// same functions for the host and replays the reference's decoder cases (tests/test_decoder_host.py).  The __constant__ table is DEFINED here:
// include this header in exactly one translation unit.
#pragma once
#include <cstdint>

namespace fz {

// sparse ternary decoder (fusion.py:422-481) as a byte-driven state machine.
// Phases: signum bytes -> weight magnitudes (bpc bytes each) -> partial Fisher-Yates (bpi bytes each).
// c_magic[m] = ceil(2^32 / m): for m <= 256 and v < 2^24, umulhi(v, c_magic[m]) == floor(v / m) exactly.
__constant__ uint32_t c_magic[257];

struct Decoder {
    int d, weight, bpc, bpi, bfs;
    uint32_t bound;
    int phase;      // 0 signums, 1 magnitudes, 2 shuffle, 3 done
    int item;       // index within the phase
    int left;       // bytes left in the current item
    unsigned long long acc;   // magnitude accumulator (general bound)
    uint32_t acc32, m, magic; // shuffle accumulator, modulus i+1 and its reciprocal
    unsigned long long sg[4]; // big-endian integer of the signum bytes (weight <= 256 bits)
};

__device__ __forceinline__ int sign_of(const Decoder& s, int i) {  // bit i (LSB first) of the big-endian signum integer
    return 2 * (int)((s.sg[(i >> 6) & 3] >> (i & 63)) & 1ULL) - 1;
}

__device__ __forceinline__ void dec_set_shuffle_item(Decoder& s, int i) {
    s.item = i; s.left = s.bpi; s.acc32 = 0;
    s.m = (uint32_t)(i + 1);
    s.magic = c_magic[s.m <= 256 ? s.m : 0];
}

// finish the current item (called when its bytes are exhausted OR the byte string ended: Python slices
// past the end yield shorter/empty byte strings and int.from_bytes(b"") == 0 -- SURVEY quirk Q4)
__device__ __forceinline__ void dec_finish_item(Decoder& s, int32_t* c) {
    if (s.phase == 0) {
        s.phase = 1; s.item = 0; s.left = s.bpc; s.acc = 0;
        if (s.weight == 0) { s.phase = 2; dec_set_shuffle_item(s, s.d - 1); }
    } else if (s.phase == 1) {
        const int i = s.item;
        // magnitude = (int(bytes) % bound) + 1, times signum (fusion.py:463-465)
        const long long mag = (long long)(s.bound == 1 ? 0ULL : s.acc % s.bound) + 1;
        if (i < s.d) c[i] = (int32_t)(mag * sign_of(s, i));
        s.item++; s.left = s.bpc; s.acc = 0;
        if (s.item >= s.weight) { s.phase = 2; dec_set_shuffle_item(s, s.d - 1); }
    } else if (s.phase == 2) {
        const int i = s.item;
        const int j = (int)s.acc32;  // already reduced mod i + 1
        const int32_t t = c[i]; c[i] = c[j]; c[j] = t;
        dec_set_shuffle_item(s, i - 1);
    }
    // fusion.py:472-473: shuffle only if num_coefs < degree, for i = d-1 ... weight+1
    if (s.phase == 2 && (s.item <= s.weight || s.weight >= s.d)) s.phase = 3;
}

__device__ __forceinline__ void dec_init(Decoder& s, int32_t* c, int d, int weight, uint32_t bound, int bpc, int bpi) {
    s.d = d; s.weight = weight; s.bound = bound; s.bpc = bpc; s.bpi = bpi;
    s.bfs = (weight + 7) / 8;
    s.phase = 0; s.item = 0; s.left = s.bfs; s.acc = 0; s.acc32 = 0; s.m = 1; s.magic = 0;
    s.sg[0] = s.sg[1] = s.sg[2] = s.sg[3] = 0;
    while (s.left == 0 && s.phase != 3) dec_finish_item(s, c);
}

__device__ __forceinline__ void dec_byte(Decoder& s, int32_t* c, uint32_t b) {
    if (s.phase == 2) {
        // Horner step with exact reciprocal: acc32 < m <= 256  =>  v < 2^16 + 2^8
        const uint32_t v = (s.acc32 << 8) | b;
        s.acc32 = v - __umulhi(v, s.magic) * s.m;
    } else if (s.phase == 1) {
        if (s.bound != 1) s.acc = ((s.acc << 8) | b) % s.bound;
    } else {
        s.sg[3] = (s.sg[3] << 8) | (s.sg[2] >> 56);
        s.sg[2] = (s.sg[2] << 8) | (s.sg[1] >> 56);
        s.sg[1] = (s.sg[1] << 8) | (s.sg[0] >> 56);
        s.sg[0] = (s.sg[0] << 8) | b;
    }
    if (--s.left == 0) dec_finish_item(s, c);
}

// drain: the byte string is exhausted; remaining items see empty / truncated slices
__device__ __forceinline__ void dec_drain(Decoder& s, int32_t* c) {
    while (s.phase != 3) dec_finish_item(s, c);
}

// host: the reciprocals the shuffle's Horner steps multiply by
static inline void decoder_build_magic(uint32_t* magic) {
    magic[0] = 0; magic[1] = 0;  // m = 1 never occurs in the shuffle (i >= 1)
    for (uint32_t m = 2; m <= 256; ++m) magic[m] = 0xFFFFFFFFu / m + 1u;
}

}  // namespace fz
