// fz_common.cuh -- modular arithmetic over Z_q, q = 2^31 - 17919 (fusion/fusion.py:17 PRIME),
// shared by every kernel of libfusion_b200.  sm_100a only.
//
// Representation on the wire (HBM): int32 canonical centred residues in
// [-(q-1)/2, (q-1)/2] -- exactly what the reference's cent() returns
// (algebra/ntt.py:93-123).  Inside kernels values are lazy unsigned residues in
// [0, 2q] (2q < 2^32) and constants are multiplied with Plantard's word-size
// modular multiplication: for a 64-bit precomputed W = (w * (-2^64) mod q) * q^-1 mod 2^64,
//     u1 = hi32((a * W) mod 2^64);   c = hi32(u1 * q + q)
// gives c == a*w (mod q) with c in [0, q] for ANY a in [0, 2^32) because 2q < 2^32
// (derivation and an exhaustive-style randomized check are in DESIGN.md / tests).
// With u1 taken as a SIGNED 32-bit value the same three instructions return the
// canonical centred residue directly (used in the last stage of each transform).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

// the dynamic shared memory of a kernel: `FZ_DYN_SHARED __align__(16) uint8_t name[];` (the CPU test suite's emulator, tests/host/cuda_emu.h, gives the
// arrays static storage and defines this macro itself)
#ifndef FZ_DYN_SHARED
#define FZ_DYN_SHARED extern __shared__
#endif

namespace fz {

constexpr uint32_t Q = 2147465729u;
constexpr uint32_t Q2 = 4294931458u;            // 2q, fits 32 bits
constexpr int32_t HALF = 1073732864;            // (q-1)/2
constexpr uint32_t MERS_C = 17919u;             // 2^31 - q
constexpr unsigned long long QINV64 = 0x686a7c4b6b244601ull;  // q^-1 mod 2^64
constexpr uint32_t NEG_N2 = 863103485u;         // (-2^64) mod q

struct Tw {  // 64-bit Plantard constant of one twiddle, split in two words
    uint32_t lo, hi;
};

struct Sh {  // Shoup pair of one twiddle: w in [0, q) and wq = floor(w * 2^32 / q)
    uint32_t w, wq;
};

// Shoup/Harvey multiplication by a precomputed constant: a*w mod q in [0, 2q) for ANY 32-bit a.
// hi = floor(a*wq / 2^32) under-estimates floor(a*w/q) by at most 1, and 2q < 2^32, so the
// wrapped 32-bit difference is exact.  Measured on B200: IMAD.HI issues every 4 cycles per
// SM sub-partition, IMAD every 2 -- this form (one HI + two IMAD = 8 cycles of the FMA-heavy
// pipe) is cheaper than Plantard's (two HI + one IMAD = 10 cycles).
__device__ __forceinline__ uint32_t sh_mul(uint32_t a, uint32_t w, uint32_t wq) {
    uint32_t hi = __umulhi(a, wq);
    return a * w - hi * Q;
}

// a*w mod q in [0, q]; a is any 32-bit unsigned value
__device__ __forceinline__ uint32_t pl_mul(uint32_t a, uint32_t wlo, uint32_t whi) {
    uint32_t u1 = __umulhi(a, wlo) + a * whi;
    return (uint32_t)(((unsigned long long)u1 * Q + Q) >> 32);
}

// cent(a*w): canonical centred residue; a is any 32-bit unsigned value
__device__ __forceinline__ int32_t pl_mul_cent(uint32_t a, uint32_t wlo, uint32_t whi) {
    int32_t u1 = (int32_t)(__umulhi(a, wlo) + a * whi);
    return (int32_t)(((long long)u1 * (long long)Q + (long long)Q) >> 32);
}

// [0, 2q] -> [0, q]
__device__ __forceinline__ uint32_t red_q(uint32_t x) { return min(x, x - Q); }

// Montgomery product of a reduced residue and ANY int32: yr * b * 2^-32 mod q in [0, q], yr in [0, q].
// m = lo32(T) * q^-1 makes T - m*q divisible by 2^32, so the quotient is hi32(T) - hi32(m*q) with no borrow; |T| <= q * 2^31
// puts both high words inside (-q/2, q/2], hence t in (-q, q).  Five instructions (IMAD.WIDE, IMAD, IMAD.HI, IADD3, VIADDMNMX) against the
// twelve of the pseudo-Mersenne folds for a product of two variables; the caller owes a factor 2^32 (folded into a later constant).
constexpr uint32_t QINV32 = 1797539329u;   // q^-1 mod 2^32
__device__ __forceinline__ uint32_t mont_mul(uint32_t yr, int32_t b) {
    const long long T = (long long)(int32_t)yr * (long long)b;
    const int32_t m = (int32_t)((uint32_t)T * QINV32);
    const int32_t t = (int32_t)(T >> 32) - __mulhi(m, (int32_t)Q);
    return red_q((uint32_t)t + Q);
}

// any int32 -> lazy unsigned residue in [0, 2q)
__device__ __forceinline__ uint32_t to_lazy(int32_t x) {
    uint32_t v = (uint32_t)x;
#ifdef __CUDA_ARCH__
    asm("{ .reg .pred p; setp.lt.s32 p, %1, 0; @p add.u32 %0, %0, %2; }" : "+r"(v) : "r"(x), "r"(Q2));
#else   // host build of the arithmetic unit tests (tests/test_modarith_host.py); never part of the library
    if (x < 0) v += Q2;
#endif
    return v;
}

// any int32 -> [0, q]
__device__ __forceinline__ uint32_t to_red(int32_t x) { return red_q(to_lazy(x)); }

// [0, 2q] -> canonical centred
__device__ __forceinline__ int32_t cent_u(uint32_t s) {
    uint32_t c = red_q(s);                 // [0, q]
    if (c > (uint32_t)HALF) c -= Q;
    return (int32_t)c;
}

// any signed 64-bit accumulator -> canonical centred residue.
// Two pseudo-Mersenne folds (2^31 == 17919 mod q): |acc| < 2^63 gives |f1| < 2^46.2 and
// f2 in (-6.5e8, 2^31 + 6.5e8), so a single conditional subtraction lands in [-(q-1)/2, (q-1)/2].
__device__ __forceinline__ int32_t cent_i64(long long acc) {
    long long f1 = (acc >> 31) * (long long)MERS_C + (acc & 0x7fffffffLL);
    long long f2 = (f1 >> 31) * (long long)MERS_C + (f1 & 0x7fffffffLL);
    if (f2 > (long long)HALF) f2 -= (long long)Q;
    return (int32_t)f2;
}

// one fold of a running signed accumulator back under 2^47 (keeps congruence mod q)
__device__ __forceinline__ long long fold_i64(long long acc) {
    return (acc >> 31) * (long long)MERS_C + (acc & 0x7fffffffLL);
}

// cent(a*b) for canonical (or any int32) a, b: |a*b| < 2^62
__device__ __forceinline__ int32_t mul_cent(int32_t a, int32_t b) { return cent_i64((long long)a * (long long)b); }

__device__ __forceinline__ int4 ld_stream(const int4* p) {
    int4 r;
#ifdef __CUDA_ARCH__
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#else
    r = *p;
#endif
    return r;
}
__device__ __forceinline__ void st_stream(int4* p, const int4& v) {
#ifdef __CUDA_ARCH__
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
#else
    *p = v;
#endif
}

}  // namespace fz
