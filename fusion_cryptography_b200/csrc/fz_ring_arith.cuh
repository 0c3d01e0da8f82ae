// fz_ring_arith.cuh -- arithmetic of the general-ring kernels (fz_ring.cu): any odd prime modulus q < 2^31 as run-time data.  Residues live in [0, q);
// inputs may be any int32 (Python's %), outputs are the canonical centred residues of cent() (algebra/ntt.py:93-123).  Pure device functions; the CPU
// suite compiles them for the host and checks them against Python integers on many moduli (tests/test_ring_arith_host.py).
#pragma once
#include "fz_common.cuh"

namespace fz {

struct RingK {
    uint32_t q, half;
    int logd;
    Sh ninv;
    unsigned long long mu;
    uint32_t mu32;   // floor(2^32 / q)
    uint32_t c31;    // 2^31 mod q
};

__device__ __forceinline__ uint32_t rq_mul(uint32_t a, Sh w, uint32_t q) {   // a*w mod q in [0, q); a any 32-bit value
    uint32_t r = a * w.w - __umulhi(a, w.wq) * q;                            // [0, 2q)
    return min(r, r - q);
}
__device__ __forceinline__ uint32_t rq_add(uint32_t a, uint32_t b, uint32_t q) {   // a, b in [0, q)
    uint32_t s = a + b;
    return min(s, s - q);
}
__device__ __forceinline__ uint32_t rq_sub(uint32_t a, uint32_t b, uint32_t q) {
    uint32_t s = a - b + q;
    return min(s, s - q);
}
// any int32 -> [0, q) (Python's %).  x + 2^31 is non-negative; floor((x + 2^31) * mu32 / 2^32) is floor((x + 2^31) / q) or one less, so one
// fused add-min finishes the reduction, and 2^31 mod q is taken off again: six instructions instead of a 32-bit division.
__device__ __forceinline__ uint32_t rq_from(int32_t x, const RingK& k) {
    const uint32_t u = (uint32_t)x ^ 0x80000000u;
    uint32_t r = u - __umulhi(u, k.mu32) * k.q;      // [0, 2q)
    r = min(r, r - k.q);
    return rq_sub(r, k.c31, k.q);
}
// [0, q) -> canonical centred residue, algebra/ntt.py:120-123
__device__ __forceinline__ int32_t rq_cent(uint32_t y, uint32_t q, uint32_t half) { return y > half ? (int32_t)(y - q) : (int32_t)y; }

// cent(x * y) / cent(x + y), algebra/polynomials.py:341-385, :282-318 (and :114-148 for the coefficient representation's +)
template <bool MUL>
__device__ __forceinline__ int32_t ring_pw1(const RingK& k, int32_t a, int32_t b) {
    const uint32_t q = k.q, x = rq_from(a, k), y = rq_from(b, k);
    uint32_t r;
    if (MUL) {
        const unsigned long long P = (unsigned long long)x * y;             // < 2^62
        const unsigned long long est = __umul64hi(P, k.mu);                 // floor(P / q) or one less
        r = (uint32_t)(P - est * q);                                        // [0, 2q), exact in 32 bits since 2q < 2^32
        r = min(r, r - q);
    } else {
        r = rq_add(x, y, q);
    }
    return rq_cent(r, q, k.half);
}

}  // namespace fz
