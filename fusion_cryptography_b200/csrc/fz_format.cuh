// fz_format.cuh -- the text of Python reprs written by ONE WARP into a byte buffer: the decimal value lists of PolynomialNTTRepresentation /
// GeneralMatrix / OneTimeVerificationKey / SignatureChallenge as str() prints them (algebra/polynomials.py:257-258, algebra/matrices.py:40-41,
// fusion/fusion.py:328-329, 388-389) and the decimal prehash.  fz_hash.cu's format kernels (the materialised hash_ch message, the pieces of the
// aggregation hash's message) are built from these; the CPU suite runs them unmodified on an emulated warp against the reference's own strings
// (tests/test_format_host.py).  c_fmt is DEFINED here: include this header in exactly one translation unit.
#pragma once
#include <cstdint>
#include <cstdio>

namespace fz {

struct FmtConst {
    // fixed text fragments of one level, assembled on the host at init
    char poly_head[160];   // "PolynomialNTTRepresentation(modulus=..., values=["
    int poly_head_len;
    char gm_head[112];     // "GeneralMatrix(elem_class=<class 'algebra.polynomials.PolynomialNTTRepresentation'>, matrix=[["
    int gm_head_len;
};
__constant__ FmtConst c_fmt[2];  // [0] secpar 128, [1] secpar 256

__device__ __forceinline__ int dec_len(int32_t v) {
    uint32_t u = v < 0 ? (uint32_t)(-(long long)v) : (uint32_t)v;
    int n = 1;
    if (u >= 10u) n = 2;
    if (u >= 100u) n = 3;
    if (u >= 1000u) n = 4;
    if (u >= 10000u) n = 5;
    if (u >= 100000u) n = 6;
    if (u >= 1000000u) n = 7;
    if (u >= 10000000u) n = 8;
    if (u >= 100000000u) n = 9;
    if (u >= 1000000000u) n = 10;
    return n + (v < 0 ? 1 : 0);
}

// writes the decimal text of v (Python str(int)) at dst, returns its length
__device__ __forceinline__ int dec_write(uint8_t* dst, int32_t v) {
    const int len = dec_len(v);
    uint32_t u = v < 0 ? (uint32_t)(-(long long)v) : (uint32_t)v;
    int p = len;
    do {
        const uint32_t q10 = u / 10u;
        dst[--p] = (uint8_t)('0' + (u - q10 * 10u));
        u = q10;
    } while (u);
    if (v < 0) dst[0] = '-';
    return len;
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total) {
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    total = __shfl_sync(0xffffffffu, x, 31);
    return x - v;
}

__device__ __forceinline__ int put_str(uint8_t* dst, const char* s, int len, int lane) {
    for (int i = lane; i < len; i += 32) dst[i] = (uint8_t)s[i];
    return len;
}

// One warp writes "v0, v1, ..., v_{d-1}" (the inside of a Python list of ints) into shared memory.
__device__ int warp_format_values(uint8_t* dst, const int32_t* __restrict__ vals, int d, int lane) {
    const int per = d / 32;  // 8 (d = 256) or 2 (d = 64)
    int32_t v[8];
    int my = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j < per) {
            v[j] = vals[lane * per + j];
            my += dec_len(v[j]) + 2;  // ", " after every value; the trailing one is dropped below
        }
    int total;
    int off = warp_excl_scan(my, lane, total);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j < per) {
            off += dec_write(dst + off, v[j]);
            if (!(lane == 31 && j == per - 1)) { dst[off] = ','; dst[off + 1] = ' '; }
            off += 2;
        }
    return total - 2;
}

// "PolynomialNTTRepresentation(...values=[...])"
__device__ int warp_format_poly(uint8_t* dst, const FmtConst& fc, const int32_t* vals, int d, int lane) {
    int o = put_str(dst, fc.poly_head, fc.poly_head_len, lane);
    o += warp_format_values(dst + o, vals, d, lane);
    if (lane == 0) { dst[o] = ']'; dst[o + 1] = ')'; }
    return o + 2;
}

// "GeneralMatrix(elem_class=..., matrix=[[POLY]])"
__device__ int warp_format_gm1(uint8_t* dst, const FmtConst& fc, const int32_t* vals, int d, int lane) {
    int o = put_str(dst, fc.gm_head, fc.gm_head_len, lane);
    o += warp_format_poly(dst + o, fc, vals, d, lane);
    if (lane == 0) { dst[o] = ']'; dst[o + 1] = ']'; dst[o + 2] = ')'; }
    return o + 3;
}

// decimal of int.from_bytes(digest, "little") (fusion.py:409 then str()), one lane, returns length (<= 78)
__device__ int format_le256(uint8_t* dst, const uint8_t* __restrict__ dg) {
    uint32_t limb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        limb[i] = (uint32_t)dg[4 * i] | ((uint32_t)dg[4 * i + 1] << 8) | ((uint32_t)dg[4 * i + 2] << 16) | ((uint32_t)dg[4 * i + 3] << 24);
    uint8_t tmp[88];  // up to 9 groups of 9 digits before the leading zeros are stripped
    int n = 0;
    for (;;) {
        uint32_t nz = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) nz |= limb[i];
        if (!nz) break;
        // divide by 10^9, emit 9 digits
        unsigned long long rem = 0;
#pragma unroll
        for (int i = 7; i >= 0; --i) {
            unsigned long long cur = (rem << 32) | limb[i];
            limb[i] = (uint32_t)(cur / 1000000000ULL);
            rem = cur - (unsigned long long)limb[i] * 1000000000ULL;
        }
        uint32_t r = (uint32_t)rem;
        for (int k = 0; k < 9; ++k) { tmp[n++] = (uint8_t)('0' + r % 10u); r /= 10u; }
    }
    while (n > 1 && tmp[n - 1] == '0') --n;  // strip leading zeros of the most significant group
    if (n == 0) tmp[n++] = '0';
    for (int i = 0; i < n; ++i) dst[i] = tmp[n - 1 - i];
    return n;
}

// host: the fixed text of both levels ([0] secpar 128, [1] secpar 256; fusion.py:36-37, 76-77, 102-103)
static inline void format_build_tables(FmtConst* h) {
    const int degs[2] = {64, 256};
    const unsigned roots[2] = {23584283u, 3337519u}, iroots[2] = {540632852u, 1978410468u};
    for (int L = 0; L < 2; ++L) {
        h[L].poly_head_len = snprintf(h[L].poly_head, sizeof(h[L].poly_head),
                                      "PolynomialNTTRepresentation(modulus=2147465729, degree=%d, root=%u, inv_root=%u, root_order=%d, values=[",
                                      degs[L], roots[L], iroots[L], 2 * degs[L]);
        h[L].gm_head_len = snprintf(h[L].gm_head, sizeof(h[L].gm_head),
                                    "GeneralMatrix(elem_class=<class 'algebra.polynomials.PolynomialNTTRepresentation'>, matrix=[[");
    }
}

}  // namespace fz
