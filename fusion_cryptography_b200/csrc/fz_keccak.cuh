// fz_keccak.cuh -- Keccak-f[1600] with the 25 lanes of one sponge in the registers of one
// thread (FIPS 202).  Replaces hashlib.sha3_256 / hashlib.shake_256 as called from
// fusion/fusion.py:408, 419, 591.  Rate 136 bytes (17 lanes) for both SHA3-256 and SHAKE256.
//
// B200 notes: XOR/AND-NOT are LOP3 and rotations are funnel shifts, all on the ALU pipe
// (0.5 warp-instruction/cycle/sub-partition measured), so this kernel family is bound by that
// pipe, not by HBM and not by the FMA/tensor pipes.
#pragma once
#include <cstdint>

namespace fz {

__device__ __constant__ unsigned long long c_keccak_rc[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

template <int N>
__device__ __forceinline__ unsigned long long rol64(unsigned long long x) {
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    if constexpr (N == 0) {
        return x;
    } else if constexpr (N == 32) {
        return ((unsigned long long)lo << 32) | hi;
    } else if constexpr (N < 32) {
        uint32_t nh = __funnelshift_l(lo, hi, N), nl = __funnelshift_l(hi, lo, N);
        return ((unsigned long long)nh << 32) | nl;
    } else {
        uint32_t nh = __funnelshift_l(hi, lo, N - 32), nl = __funnelshift_l(lo, hi, N - 32);
        return ((unsigned long long)nh << 32) | nl;
    }
}

__device__ __forceinline__ unsigned long long xor5(unsigned long long a, unsigned long long b, unsigned long long c, unsigned long long d,
                                                   unsigned long long e) {
    return a ^ b ^ c ^ d ^ e;
}
__device__ __forceinline__ unsigned long long chi(unsigned long long a, unsigned long long b, unsigned long long c) { return a ^ (~b & c); }

// one round; lane index = x + 5y
__device__ __forceinline__ void keccak_round(unsigned long long (&a)[25], unsigned long long rc) {
    unsigned long long c0 = xor5(a[0], a[5], a[10], a[15], a[20]);
    unsigned long long c1 = xor5(a[1], a[6], a[11], a[16], a[21]);
    unsigned long long c2 = xor5(a[2], a[7], a[12], a[17], a[22]);
    unsigned long long c3 = xor5(a[3], a[8], a[13], a[18], a[23]);
    unsigned long long c4 = xor5(a[4], a[9], a[14], a[19], a[24]);
    // theta folded into the rho inputs: a ^ D[x] = a ^ C[x-1] ^ rol(C[x+1], 1) is ONE three-input LOP3 per half, so D is never formed
    const unsigned long long r0 = rol64<1>(c0), r1 = rol64<1>(c1), r2 = rol64<1>(c2), r3 = rol64<1>(c3), r4 = rol64<1>(c4);
#define TH0(v) ((v) ^ c4 ^ r1)
#define TH1(v) ((v) ^ c0 ^ r2)
#define TH2(v) ((v) ^ c1 ^ r3)
#define TH3(v) ((v) ^ c2 ^ r4)
#define TH4(v) ((v) ^ c3 ^ r0)
    // theta + rho + pi:  b[y + 5*((2x + 3y) % 5)] = rol(a[x + 5y] ^ d[x], r[x + 5y])
    unsigned long long b0 = TH0(a[0]);
    unsigned long long b10 = rol64<1>(TH1(a[1]));
    unsigned long long b20 = rol64<62>(TH2(a[2]));
    unsigned long long b5 = rol64<28>(TH3(a[3]));
    unsigned long long b15 = rol64<27>(TH4(a[4]));
    unsigned long long b16 = rol64<36>(TH0(a[5]));
    unsigned long long b1 = rol64<44>(TH1(a[6]));
    unsigned long long b11 = rol64<6>(TH2(a[7]));
    unsigned long long b21 = rol64<55>(TH3(a[8]));
    unsigned long long b6 = rol64<20>(TH4(a[9]));
    unsigned long long b7 = rol64<3>(TH0(a[10]));
    unsigned long long b17 = rol64<10>(TH1(a[11]));
    unsigned long long b2 = rol64<43>(TH2(a[12]));
    unsigned long long b12 = rol64<25>(TH3(a[13]));
    unsigned long long b22 = rol64<39>(TH4(a[14]));
    unsigned long long b23 = rol64<41>(TH0(a[15]));
    unsigned long long b8 = rol64<45>(TH1(a[16]));
    unsigned long long b18 = rol64<15>(TH2(a[17]));
    unsigned long long b3 = rol64<21>(TH3(a[18]));
    unsigned long long b13 = rol64<8>(TH4(a[19]));
    unsigned long long b14 = rol64<18>(TH0(a[20]));
    unsigned long long b24 = rol64<2>(TH1(a[21]));
    unsigned long long b9 = rol64<61>(TH2(a[22]));
    unsigned long long b19 = rol64<56>(TH3(a[23]));
    unsigned long long b4 = rol64<14>(TH4(a[24]));
#undef TH0
#undef TH1
#undef TH2
#undef TH3
#undef TH4
    // chi + iota
    a[0] = chi(b0, b1, b2) ^ rc; a[1] = chi(b1, b2, b3); a[2] = chi(b2, b3, b4); a[3] = chi(b3, b4, b0); a[4] = chi(b4, b0, b1);
    a[5] = chi(b5, b6, b7); a[6] = chi(b6, b7, b8); a[7] = chi(b7, b8, b9); a[8] = chi(b8, b9, b5); a[9] = chi(b9, b5, b6);
    a[10] = chi(b10, b11, b12); a[11] = chi(b11, b12, b13); a[12] = chi(b12, b13, b14); a[13] = chi(b13, b14, b10); a[14] = chi(b14, b10, b11);
    a[15] = chi(b15, b16, b17); a[16] = chi(b16, b17, b18); a[17] = chi(b17, b18, b19); a[18] = chi(b18, b19, b15); a[19] = chi(b19, b15, b16);
    a[20] = chi(b20, b21, b22); a[21] = chi(b21, b22, b23); a[22] = chi(b22, b23, b24); a[23] = chi(b23, b24, b20); a[24] = chi(b24, b20, b21);
}

__device__ __forceinline__ void keccak_f(unsigned long long (&a)[25]) {
#pragma unroll 1
    for (int r = 0; r < 24; ++r) keccak_round(a, c_keccak_rc[r]);
}

constexpr int KECCAK_RATE = 136;        // bytes
constexpr int KECCAK_RATE_LANES = 17;   // 64-bit lanes

}  // namespace fz
