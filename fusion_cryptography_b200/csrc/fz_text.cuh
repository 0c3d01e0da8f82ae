// fz_text.cuh -- the byte queue and the decimal formatter of the streaming hash kernels (fz_hashstream.cu): every sponge thread generates the text
// of Python reprs (str(OneTimeVerificationKey), the decimal prehash; fusion/fusion.py:412-419, 328-329, algebra/matrices.py:40-41,
// algebra/polynomials.py:257-258) 16 bytes at a time and absorbs it 136 bytes at a time.  Pure device functions, kept in a header so that the CPU
// suite can compile them for the host and check them against Python's str() and hashlib (tests/test_text_host.py).
#pragma once
#include <cstdint>

#include "fz_keccak.cuh"

namespace fz {

// ---------------------------------------------------------------------------------- byte queue
// Whole 64-bit words go to a small local array with plain stores (no read-modify-write on the serial path); the
// partial word lives in registers.
constexpr int Q_WORDS = 26;  // 17 rate lanes + room for one 4-value step (<= 52 bytes) plus a 16-byte token of overshoot
struct Queue {
    unsigned long long blk[Q_WORDS];
    unsigned long long cur;  // partial word, nb valid low bytes
    int nb, wi;              // bytes in cur, complete words in blk
};

__device__ __forceinline__ void q_reset(Queue& q) { q.cur = 0; q.nb = 0; q.wi = 0; }
__device__ __forceinline__ int q_bytes(const Queue& q) { return q.wi * 8 + q.nb; }

// append the low n (<= 16) bytes of (lo, hi); bytes above n must be zero
__device__ __forceinline__ void q_push(Queue& q, unsigned long long lo, unsigned long long hi, int n) {
    unsigned long long w0, w1, w2;
    if (q.nb == 0) {
        w0 = lo; w1 = hi; w2 = 0;
    } else {
        const int sh = q.nb * 8;
        w0 = q.cur | (lo << sh);
        w1 = (lo >> (64 - sh)) | (hi << sh);
        w2 = hi >> (64 - sh);
    }
    const int tot = q.nb + n;
    if (tot >= 16) {
        q.blk[q.wi] = w0; q.blk[q.wi + 1] = w1; q.wi += 2; q.cur = w2;
    } else if (tot >= 8) {
        q.blk[q.wi] = w0; q.wi += 1; q.cur = w1;
    } else {
        q.cur = w0;
    }
    q.nb = tot & 7;
}

// absorb the 17 full lanes and keep what is beyond them
__device__ __forceinline__ void q_absorb_full(Queue& q, unsigned long long (&a)[25]) {
#pragma unroll
    for (int k = 0; k < KECCAK_RATE_LANES; ++k) a[k] ^= q.blk[k];
    const int extra = q.wi - KECCAK_RATE_LANES;
    for (int k = 0; k < extra; ++k) q.blk[k] = q.blk[KECCAK_RATE_LANES + k];
    q.wi = extra;
}

// final (partial) block with SHAKE padding; fewer than 136 bytes are buffered
__device__ __forceinline__ void q_absorb_last(Queue& q, unsigned long long (&a)[25]) {
    q.blk[q.wi] = q.cur;
    for (int k = q.wi + 1; k < KECCAK_RATE_LANES; ++k) q.blk[k] = 0;
    q.blk[q.wi] ^= 0x1FULL << (q.nb * 8);
    q.blk[16] ^= 0x8000000000000000ULL;
#pragma unroll
    for (int k = 0; k < KECCAK_RATE_LANES; ++k) a[k] ^= q.blk[k];
}

// ---------------------------------------------------------------------------------- value formatting
// decimal text of v followed by ", " (sep) -> up to 13 bytes, first character in the lowest byte
__device__ __forceinline__ int fmt_value(int32_t v, bool sep, unsigned long long& lo, unsigned long long& hi) {
    uint32_t u = v < 0 ? (uint32_t)(-(long long)v) : (uint32_t)v;
    // zero padded digits m0..m9 (m0 most significant) from nine INDEPENDENT constant divisions (short dependency chain)
    const uint32_t q1 = u / 10u, q2 = u / 100u, q3 = u / 1000u, q4 = u / 10000u, q5 = u / 100000u;
    const uint32_t q6 = u / 1000000u, q7 = u / 10000000u, q8 = u / 100000000u, q9 = u / 1000000000u;
    const uint32_t m9 = u - 10u * q1, m8 = q1 - 10u * q2, m7 = q2 - 10u * q3, m6 = q3 - 10u * q4, m5 = q4 - 10u * q5;
    const uint32_t m4 = q5 - 10u * q6, m3 = q6 - 10u * q7, m2 = q7 - 10u * q8, m1 = q8 - 10u * q9, m0 = q9;
    const uint32_t w0 = (m0 | (m1 << 8) | (m2 << 16) | (m3 << 24)) + 0x30303030u;
    const uint32_t w1 = (m4 | (m5 << 8) | (m6 << 16) | (m7 << 24)) + 0x30303030u;
    const uint32_t w2 = (m8 | (m9 << 8)) + 0x3030u;
    int nd = 1;
    if (u >= 10u) nd = 2;
    if (u >= 100u) nd = 3;
    if (u >= 1000u) nd = 4;
    if (u >= 10000u) nd = 5;
    if (u >= 100000u) nd = 6;
    if (u >= 1000000u) nd = 7;
    if (u >= 10000000u) nd = 8;
    if (u >= 100000000u) nd = 9;
    if (u >= 1000000000u) nd = 10;
    // drop the 10 - nd leading zeros: 80-bit right shift by whole bytes
    unsigned long long s_lo = ((unsigned long long)w1 << 32) | w0, s_hi = w2;
    const int skip = 10 - nd;
    if (skip >= 8) {
        s_lo = s_hi >> ((skip - 8) * 8);
        s_hi = 0;
    } else if (skip > 0) {
        s_lo = (s_lo >> (skip * 8)) | (s_hi << (64 - skip * 8));
        s_hi = s_hi >> (skip * 8);
    }
    int len = nd;
    if (v < 0) {  // prepend '-'
        s_hi = (s_hi << 8) | (s_lo >> 56);
        s_lo = (s_lo << 8) | 0x2dULL;
        ++len;
    }
    if (sep) {  // append ", "
        if (len < 7) s_lo |= 0x202cULL << (len * 8);
        else if (len == 7) { s_lo |= 0x2cULL << 56; s_hi |= 0x20ULL; }
        else s_hi |= 0x202cULL << ((len - 8) * 8);
        len += 2;
    }
    lo = s_lo;
    hi = s_hi;
    return len;
}

// decimal text of int.from_bytes(digest, "little") into buf (<= 78 bytes), returns length
__device__ int fmt_le256(const uint8_t* __restrict__ dg, uint8_t* buf) {
    uint32_t limb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        limb[i] = (uint32_t)dg[4 * i] | ((uint32_t)dg[4 * i + 1] << 8) | ((uint32_t)dg[4 * i + 2] << 16) | ((uint32_t)dg[4 * i + 3] << 24);
    uint8_t tmp[88];
    int n = 0;
    for (;;) {
        uint32_t nz = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) nz |= limb[i];
        if (!nz) break;
        unsigned long long rem = 0;
#pragma unroll
        for (int i = 7; i >= 0; --i) {
            const unsigned long long cur = (rem << 32) | limb[i];
            limb[i] = (uint32_t)(cur / 1000000000ULL);
            rem = cur - (unsigned long long)limb[i] * 1000000000ULL;
        }
        uint32_t r = (uint32_t)rem;
        for (int k = 0; k < 9; ++k) { tmp[n++] = (uint8_t)('0' + r % 10u); r /= 10u; }
    }
    while (n > 1 && tmp[n - 1] == '0') --n;
    if (n == 0) tmp[n++] = '0';
    for (int i = 0; i < n; ++i) buf[i] = tmp[n - 1 - i];
    for (int i = n; i < 96; ++i) buf[i] = 0;
    return n;
}

}  // namespace fz
