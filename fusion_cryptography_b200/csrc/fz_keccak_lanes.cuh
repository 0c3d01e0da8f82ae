// fz_keccak_lanes.cuh -- Keccak-f[1600] spread over the lanes of a warp, the two cooperative layouts of the serial SHAKE256 chains of hash_ag
// (fusion.py:573-591: one sponge per aggregate).  pairag: TWO lanes per sponge -- the even lane holds the low 32 bits of all 25 Keccak lanes, the odd
// lane the high 32 bits, 16 sponges per warp; every 64-bit rotation is one exchange with the partner lane plus a funnel shift.  warpag: ONE sponge
// per warp -- lane L < 25 holds Keccak lane L as (lo, hi); theta, pi and chi are warp shuffles.  fz_hashstream.cu builds the ring-fed absorb /
// squeeze kernels around them; the CPU suite runs the same functions on an emulated warp (32 host threads, shuffles through a barrier) against the
// one-thread permutation of fz_keccak.cuh (tests/test_keccak_lanes_host.py).  warpag::c_rot is DEFINED here: include in exactly one translation unit.
#pragma once
#include <cstdint>

#include "fz_keccak.cuh"

namespace fz {

namespace pairag {

template <int N>
__device__ __forceinline__ uint32_t rot(uint32_t v) {
    if constexpr (N == 0) return v;
    const uint32_t p = __shfl_xor_sync(0xffffffffu, v, 1);
    if constexpr (N == 32) return p;
    // n < 32: lo' = (lo << n) | (hi >> (32 - n)), hi' = (hi << n) | (lo >> (32 - n)): funnel(partner, mine, n) on both lanes; 32 + n swaps the halves first
    constexpr int n = N & 31;
    if constexpr (N < 32) return __funnelshift_l(p, v, n);
    else return __funnelshift_l(v, p, n);
}
__device__ __forceinline__ uint32_t x5(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e) { return a ^ b ^ c ^ d ^ e; }
__device__ __forceinline__ uint32_t ch(uint32_t a, uint32_t b, uint32_t c) { return a ^ (~b & c); }

__device__ __forceinline__ void round(uint32_t (&a)[25], uint32_t rc) {
    const uint32_t c0 = x5(a[0], a[5], a[10], a[15], a[20]);
    const uint32_t c1 = x5(a[1], a[6], a[11], a[16], a[21]);
    const uint32_t c2 = x5(a[2], a[7], a[12], a[17], a[22]);
    const uint32_t c3 = x5(a[3], a[8], a[13], a[18], a[23]);
    const uint32_t c4 = x5(a[4], a[9], a[14], a[19], a[24]);
    const uint32_t r0 = rot<1>(c0), r1 = rot<1>(c1), r2 = rot<1>(c2), r3 = rot<1>(c3), r4 = rot<1>(c4);
#define TH0(v) ((v) ^ c4 ^ r1)
#define TH1(v) ((v) ^ c0 ^ r2)
#define TH2(v) ((v) ^ c1 ^ r3)
#define TH3(v) ((v) ^ c2 ^ r4)
#define TH4(v) ((v) ^ c3 ^ r0)
    const uint32_t b0 = TH0(a[0]);
    const uint32_t b10 = rot<1>(TH1(a[1]));
    const uint32_t b20 = rot<62>(TH2(a[2]));
    const uint32_t b5 = rot<28>(TH3(a[3]));
    const uint32_t b15 = rot<27>(TH4(a[4]));
    const uint32_t b16 = rot<36>(TH0(a[5]));
    const uint32_t b1 = rot<44>(TH1(a[6]));
    const uint32_t b11 = rot<6>(TH2(a[7]));
    const uint32_t b21 = rot<55>(TH3(a[8]));
    const uint32_t b6 = rot<20>(TH4(a[9]));
    const uint32_t b7 = rot<3>(TH0(a[10]));
    const uint32_t b17 = rot<10>(TH1(a[11]));
    const uint32_t b2 = rot<43>(TH2(a[12]));
    const uint32_t b12 = rot<25>(TH3(a[13]));
    const uint32_t b22 = rot<39>(TH4(a[14]));
    const uint32_t b23 = rot<41>(TH0(a[15]));
    const uint32_t b8 = rot<45>(TH1(a[16]));
    const uint32_t b18 = rot<15>(TH2(a[17]));
    const uint32_t b3 = rot<21>(TH3(a[18]));
    const uint32_t b13 = rot<8>(TH4(a[19]));
    const uint32_t b14 = rot<18>(TH0(a[20]));
    const uint32_t b24 = rot<2>(TH1(a[21]));
    const uint32_t b9 = rot<61>(TH2(a[22]));
    const uint32_t b19 = rot<56>(TH3(a[23]));
    const uint32_t b4 = rot<14>(TH4(a[24]));
#undef TH0
#undef TH1
#undef TH2
#undef TH3
#undef TH4
    a[0] = ch(b0, b1, b2) ^ rc; a[1] = ch(b1, b2, b3); a[2] = ch(b2, b3, b4); a[3] = ch(b3, b4, b0); a[4] = ch(b4, b0, b1);
    a[5] = ch(b5, b6, b7); a[6] = ch(b6, b7, b8); a[7] = ch(b7, b8, b9); a[8] = ch(b8, b9, b5); a[9] = ch(b9, b5, b6);
    a[10] = ch(b10, b11, b12); a[11] = ch(b11, b12, b13); a[12] = ch(b12, b13, b14); a[13] = ch(b13, b14, b10); a[14] = ch(b14, b10, b11);
    a[15] = ch(b15, b16, b17); a[16] = ch(b16, b17, b18); a[17] = ch(b17, b18, b19); a[18] = ch(b18, b19, b15); a[19] = ch(b19, b15, b16);
    a[20] = ch(b20, b21, b22); a[21] = ch(b21, b22, b23); a[22] = ch(b22, b23, b24); a[23] = ch(b23, b24, b20); a[24] = ch(b24, b20, b21);
}

// Keccak-f[1600] on the half of the state this lane holds (hi = 0 / 1); must be executed by all 32 lanes of the warp
__device__ __forceinline__ void keccak_f_pair(uint32_t (&a)[25], uint32_t hi) {
    const uint32_t* rc = reinterpret_cast<const uint32_t*>(c_keccak_rc) + hi;
#pragma unroll 4
    for (int r = 0; r < 24; ++r) round(a, rc[2 * r]);
}

}  // namespace pairag

namespace warpag {

__constant__ int c_rot[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

struct Lane {  // per-lane constants of the cooperative permutation
    int src5, src10, src15, src20;  // same column, other rows
    int xm1, xp1, xp2;              // same row, columns x-1, x+1, x+2
    int pi0, pi1, pi2;              // lanes whose rotated values become B[x][y], B[x+1][y], B[x+2][y] after pi
    int rot;
    bool active;
};

__device__ __forceinline__ Lane lane_init(int L) {
    Lane c;
    const int l = L < 25 ? L : 0;
    const int x = l % 5, y = l / 5;
    c.src5 = (l + 5) % 25; c.src10 = (l + 10) % 25; c.src15 = (l + 15) % 25; c.src20 = (l + 20) % 25;
    c.xm1 = (x + 4) % 5 + 5 * y; c.xp1 = (x + 1) % 5 + 5 * y; c.xp2 = (x + 2) % 5 + 5 * y;
    // pi: B[x'][y'] = rot(A[(x' + 3y') % 5][x']); chi reads B at columns x, x+1, x+2 of the same row, so the three
    // sources are gathered straight from the rotated values (pi and chi in one shuffle level)
    const int x1 = (x + 1) % 5, x2 = (x + 2) % 5;
    c.pi0 = ((x + 3 * y) % 5) + 5 * x;
    c.pi1 = ((x1 + 3 * y) % 5) + 5 * x1;
    c.pi2 = ((x2 + 3 * y) % 5) + 5 * x2;
    c.rot = c_rot[l];
    c.active = L < 25;
    return c;
}

#define FZ_SHFL(v, s) __shfl_sync(0xffffffffu, (v), (s))

__device__ __forceinline__ void keccak_f_warp(uint32_t& lo, uint32_t& hi, const Lane& c, int L) {
#pragma unroll 1
    for (int r = 0; r < 24; ++r) {
        // theta
        uint32_t cl = lo ^ FZ_SHFL(lo, c.src5) ^ FZ_SHFL(lo, c.src10) ^ FZ_SHFL(lo, c.src15) ^ FZ_SHFL(lo, c.src20);
        uint32_t ch = hi ^ FZ_SHFL(hi, c.src5) ^ FZ_SHFL(hi, c.src10) ^ FZ_SHFL(hi, c.src15) ^ FZ_SHFL(hi, c.src20);
        const uint32_t ml = FZ_SHFL(cl, c.xm1), mh = FZ_SHFL(ch, c.xm1);
        const uint32_t pl = FZ_SHFL(cl, c.xp1), ph = FZ_SHFL(ch, c.xp1);
        lo ^= ml ^ __funnelshift_l(ph, pl, 1);
        hi ^= mh ^ __funnelshift_l(pl, ph, 1);
        // rho: rotate left by c.rot
        {
            const bool sw = c.rot >= 32;
            const uint32_t a = sw ? hi : lo, b = sw ? lo : hi;   // value = (b:a) rotated by rot & 31
            const int s = c.rot & 31;
            lo = __funnelshift_l(b, a, s);
            hi = __funnelshift_l(a, b, s);
        }
        // pi + chi
        const uint32_t b0l = FZ_SHFL(lo, c.pi0), b0h = FZ_SHFL(hi, c.pi0);
        const uint32_t b1l = FZ_SHFL(lo, c.pi1), b1h = FZ_SHFL(hi, c.pi1);
        const uint32_t b2l = FZ_SHFL(lo, c.pi2), b2h = FZ_SHFL(hi, c.pi2);
        lo = b0l ^ (~b1l & b2l);
        hi = b0h ^ (~b1h & b2h);
        // iota
        if (L == 0) {
            const unsigned long long rc = c_keccak_rc[r];
            lo ^= (uint32_t)rc;
            hi ^= (uint32_t)(rc >> 32);
        }
    }
}

// The same permutation, fully unrolled (round constants become immediates under a lane-0 mask, no predicated constant load), with
// the shuffles written in the order that lets the scheduler run the low half ahead of the high half: the low half's stage-2 shuffles
// are issued while the high half's stage-1 shuffles are still in flight, and the low half of the next round starts while the high
// half finishes chi.  The schedule ptxas emits follows the source order of the shuffles closely; measured in tools/keccak_lab.cu
// (cycles per round at 1 / 8 / 12 chains per SM): this order 168 / 197 / 240, the plain loop 200 / 257 / 287, the plain loop unrolled
// 198 / 219 / 266, unrolled by eight 177 / 254 / 288.
template <int UNROLL>
__device__ __forceinline__ void keccak_f_warp_u(uint32_t& lo, uint32_t& hi, const Lane& c, int L) {
    constexpr unsigned long long RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
        0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
        0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    const uint32_t is0 = L == 0 ? 0xffffffffu : 0u;
#pragma unroll UNROLL
    for (int r = 0; r < 24; ++r) {
        // theta: column parity of the low half, then of the high half
        const uint32_t l1 = FZ_SHFL(lo, c.src5), l2 = FZ_SHFL(lo, c.src10), l3 = FZ_SHFL(lo, c.src15), l4 = FZ_SHFL(lo, c.src20);
        const uint32_t h1 = FZ_SHFL(hi, c.src5), h2 = FZ_SHFL(hi, c.src10), h3 = FZ_SHFL(hi, c.src15), h4 = FZ_SHFL(hi, c.src20);
        const uint32_t cl = lo ^ l1 ^ l2 ^ l3 ^ l4;
        const uint32_t ml = FZ_SHFL(cl, c.xm1), pl = FZ_SHFL(cl, c.xp1);
        const uint32_t ch = hi ^ h1 ^ h2 ^ h3 ^ h4;
        const uint32_t mh = FZ_SHFL(ch, c.xm1), ph = FZ_SHFL(ch, c.xp1);
        lo ^= ml ^ __funnelshift_l(ph, pl, 1);
        hi ^= mh ^ __funnelshift_l(pl, ph, 1);
        {   // rho
            const bool sw = c.rot >= 32;
            const uint32_t a = sw ? hi : lo, b = sw ? lo : hi;
            const int s = c.rot & 31;
            lo = __funnelshift_l(b, a, s);
            hi = __funnelshift_l(a, b, s);
        }
        // pi + chi + iota, low half first
        const uint32_t b0l = FZ_SHFL(lo, c.pi0), b1l = FZ_SHFL(lo, c.pi1), b2l = FZ_SHFL(lo, c.pi2);
        const uint32_t b0h = FZ_SHFL(hi, c.pi0), b1h = FZ_SHFL(hi, c.pi1), b2h = FZ_SHFL(hi, c.pi2);
        lo = b0l ^ (~b1l & b2l) ^ ((uint32_t)RC[r] & is0);
        hi = b0h ^ (~b1h & b2h) ^ ((uint32_t)(RC[r] >> 32) & is0);
    }
}

}  // namespace warpag

}  // namespace fz
