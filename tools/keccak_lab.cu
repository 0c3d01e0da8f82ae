// keccak_lab.cu -- development microbenchmark: cycles per Keccak-f round of one SERIAL sponge chain for several
// ways of spreading the 25 lanes over a warp (the aggregation hash of fusion.py:573-591 is one serial SHAKE256 per
// aggregate, so chain latency -- not ALU throughput -- bounds hash_ag at capacity), and the ALU-bound
// thread-per-sponge rate used as the roofline of hash_ch.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/keccak_lab tools/keccak_lab.cu
//   tools/keccak_lab [perms]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "../fusion_cryptography_b200/csrc/fz_keccak.cuh"

using namespace fz;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__constant__ int c_rot[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

struct Lane {
    int src5, src10, src15, src20, xm1, xp1, pi0, pi1, pi2, rot, x;
    unsigned colmask;
};
__device__ __forceinline__ Lane lane_init(int L) {
    Lane c;
    const int l = L < 25 ? L : 0;
    const int x = l % 5, y = l / 5;
    c.x = x;
    c.src5 = (l + 5) % 25; c.src10 = (l + 10) % 25; c.src15 = (l + 15) % 25; c.src20 = (l + 20) % 25;
    c.xm1 = (x + 4) % 5 + 5 * y; c.xp1 = (x + 1) % 5 + 5 * y;
    const int x1 = (x + 1) % 5, x2 = (x + 2) % 5;
    c.pi0 = ((x + 3 * y) % 5) + 5 * x;
    c.pi1 = ((x1 + 3 * y) % 5) + 5 * x1;
    c.pi2 = ((x2 + 3 * y) % 5) + 5 * x2;
    c.rot = c_rot[l];
    c.colmask = L < 25 ? (0x108421u << x) : 0xFE000000u;
    return c;
}
#define SH(v, s) __shfl_sync(0xffffffffu, (v), (s))

__device__ __forceinline__ void rho(uint32_t& lo, uint32_t& hi, int rot) {
    const bool sw = rot >= 32;
    const uint32_t a = sw ? hi : lo, b = sw ? lo : hi;
    const int s = rot & 31;
    lo = __funnelshift_l(b, a, s);
    hi = __funnelshift_l(a, b, s);
}
__device__ __forceinline__ void iota(uint32_t& lo, uint32_t& hi, int r, int L) {
    if (L == 0) {
        const unsigned long long rc = c_keccak_rc[r];
        lo ^= (uint32_t)rc;
        hi ^= (uint32_t)(rc >> 32);
    }
}

// V0: 18 shuffles per round (the shipped hash_ag_warp permutation)
__device__ __forceinline__ void f_v0(uint32_t& lo, uint32_t& hi, const Lane& c, int L, uint2*, uint2*) {
#pragma unroll 1
    for (int r = 0; r < 24; ++r) {
        uint32_t cl = lo ^ SH(lo, c.src5) ^ SH(lo, c.src10) ^ SH(lo, c.src15) ^ SH(lo, c.src20);
        uint32_t ch = hi ^ SH(hi, c.src5) ^ SH(hi, c.src10) ^ SH(hi, c.src15) ^ SH(hi, c.src20);
        const uint32_t ml = SH(cl, c.xm1), mh = SH(ch, c.xm1), pl = SH(cl, c.xp1), ph = SH(ch, c.xp1);
        lo ^= ml ^ __funnelshift_l(ph, pl, 1);
        hi ^= mh ^ __funnelshift_l(pl, ph, 1);
        rho(lo, hi, c.rot);
        const uint32_t b0l = SH(lo, c.pi0), b0h = SH(hi, c.pi0), b1l = SH(lo, c.pi1), b1h = SH(hi, c.pi1), b2l = SH(lo, c.pi2), b2h = SH(hi, c.pi2);
        lo = b0l ^ (~b1l & b2l);
        hi = b0h ^ (~b1h & b2h);
        iota(lo, hi, r, L);
    }
}

// V1: column parity with redux.sync over per-column member masks: 2 REDUX + 10 SHFL
__device__ __forceinline__ void f_v1(uint32_t& lo, uint32_t& hi, const Lane& c, int L, uint2*, uint2*) {
#pragma unroll 1
    for (int r = 0; r < 24; ++r) {
        const uint32_t cl = __reduce_xor_sync(c.colmask, lo), ch = __reduce_xor_sync(c.colmask, hi);
        const uint32_t ml = SH(cl, c.xm1), mh = SH(ch, c.xm1), pl = SH(cl, c.xp1), ph = SH(ch, c.xp1);
        lo ^= ml ^ __funnelshift_l(ph, pl, 1);
        hi ^= mh ^ __funnelshift_l(pl, ph, 1);
        rho(lo, hi, c.rot);
        const uint32_t b0l = SH(lo, c.pi0), b0h = SH(hi, c.pi0), b1l = SH(lo, c.pi1), b1h = SH(hi, c.pi1), b2l = SH(lo, c.pi2), b2h = SH(hi, c.pi2);
        lo = b0l ^ (~b1l & b2l);
        hi = b0h ^ (~b1h & b2h);
        iota(lo, hi, r, L);
    }
}

// V2: shared memory, three stages (own column's 4 other rows -> C; C of the two neighbour columns; pi/chi gather)
__device__ __forceinline__ void f_v2(uint32_t& lo, uint32_t& hi, const Lane& c, int L, uint2* bufA, uint2* bufB) {
    uint2* bufC = bufB + 32;
#pragma unroll 1
    for (int r = 0; r < 24; ++r) {
        bufA[L] = make_uint2(lo, hi);
        __syncwarp();
        const uint2 a1 = bufA[c.src5], a2 = bufA[c.src10], a3 = bufA[c.src15], a4 = bufA[c.src20];
        const uint32_t cl = lo ^ a1.x ^ a2.x ^ a3.x ^ a4.x, ch = hi ^ a1.y ^ a2.y ^ a3.y ^ a4.y;
        bufC[L] = make_uint2(cl, ch);
        __syncwarp();
        const uint2 m = bufC[c.xm1], p = bufC[c.xp1];
        lo ^= m.x ^ __funnelshift_l(p.y, p.x, 1);
        hi ^= m.y ^ __funnelshift_l(p.x, p.y, 1);
        rho(lo, hi, c.rot);
        bufB[L] = make_uint2(lo, hi);
        __syncwarp();
        const uint2 b0 = bufB[c.pi0], b1 = bufB[c.pi1], b2 = bufB[c.pi2];
        lo = b0.x ^ (~b1.x & b2.x);
        hi = b0.y ^ (~b1.y & b2.y);
        iota(lo, hi, r, L);
    }
}

// V3: shared memory, two stages: theta reads the ten lanes of columns x-1 and x+1 directly
__device__ __forceinline__ void f_v3(uint32_t& lo, uint32_t& hi, const Lane& c, int L, uint2* bufA, uint2* bufB) {
    const int cm = (c.x + 4) % 5, cp = (c.x + 1) % 5;
#pragma unroll 1
    for (int r = 0; r < 24; ++r) {
        bufA[L] = make_uint2(lo, hi);
        __syncwarp();
        uint2 m = bufA[cm], p = bufA[cp];
#pragma unroll
        for (int y = 1; y < 5; ++y) {
            const uint2 mm = bufA[cm + 5 * y], pp = bufA[cp + 5 * y];
            m.x ^= mm.x; m.y ^= mm.y; p.x ^= pp.x; p.y ^= pp.y;
        }
        lo ^= m.x ^ __funnelshift_l(p.y, p.x, 1);
        hi ^= m.y ^ __funnelshift_l(p.x, p.y, 1);
        rho(lo, hi, c.rot);
        bufB[L] = make_uint2(lo, hi);
        __syncwarp();
        const uint2 b0 = bufB[c.pi0], b1 = bufB[c.pi1], b2 = bufB[c.pi2];
        lo = b0.x ^ (~b1.x & b2.x);
        hi = b0.y ^ (~b1.y & b2.y);
        iota(lo, hi, r, L);
    }
}

// V5: REDUX parity, then shared memory for the two remaining exchanges
__device__ __forceinline__ void f_v5(uint32_t& lo, uint32_t& hi, const Lane& c, int L, uint2* bufA, uint2* bufB) {
#pragma unroll 1
    for (int r = 0; r < 24; ++r) {
        const uint32_t cl = __reduce_xor_sync(c.colmask, lo), ch = __reduce_xor_sync(c.colmask, hi);
        bufA[L] = make_uint2(cl, ch);
        __syncwarp();
        const uint2 m = bufA[c.xm1], p = bufA[c.xp1];
        lo ^= m.x ^ __funnelshift_l(p.y, p.x, 1);
        hi ^= m.y ^ __funnelshift_l(p.x, p.y, 1);
        rho(lo, hi, c.rot);
        bufB[L] = make_uint2(lo, hi);
        __syncwarp();
        const uint2 b0 = bufB[c.pi0], b1 = bufB[c.pi1], b2 = bufB[c.pi2];
        lo = b0.x ^ (~b1.x & b2.x);
        hi = b0.y ^ (~b1.y & b2.y);
        iota(lo, hi, r, L);
    }
}



// V8 / V9: the V0 round with the 24 rounds fully unrolled (round constants become immediates, the scheduler may overlap the tail of
// one round with the head of the next) / unrolled by 4
template <int UNROLL>
__device__ __forceinline__ void f_v8(uint32_t& lo, uint32_t& hi, const Lane& c, int L) {
    constexpr unsigned long long RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
        0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
        0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    const uint32_t is0 = L == 0 ? 0xffffffffu : 0u;
#pragma unroll UNROLL
    for (int r = 0; r < 24; ++r) {
        uint32_t cl = lo ^ SH(lo, c.src5) ^ SH(lo, c.src10) ^ SH(lo, c.src15) ^ SH(lo, c.src20);
        uint32_t ch = hi ^ SH(hi, c.src5) ^ SH(hi, c.src10) ^ SH(hi, c.src15) ^ SH(hi, c.src20);
        const uint32_t ml = SH(cl, c.xm1), mh = SH(ch, c.xm1), pl = SH(cl, c.xp1), ph = SH(ch, c.xp1);
        lo ^= ml ^ __funnelshift_l(ph, pl, 1);
        hi ^= mh ^ __funnelshift_l(pl, ph, 1);
        rho(lo, hi, c.rot);
        const uint32_t b0l = SH(lo, c.pi0), b0h = SH(hi, c.pi0), b1l = SH(lo, c.pi1), b1h = SH(hi, c.pi1), b2l = SH(lo, c.pi2), b2h = SH(hi, c.pi2);
        lo = b0l ^ (~b1l & b2l) ^ ((uint32_t)RC[r] & is0);
        hi = b0h ^ (~b1h & b2h) ^ ((uint32_t)(RC[r] >> 32) & is0);
    }
}


// V16 / V17: source-level reorderings of the V8 round (the compiler's schedule follows the source order of the shuffles closely)
template <int UNROLL, int ORDER>
__device__ __forceinline__ void f_v16(uint32_t& lo, uint32_t& hi, const Lane& c, int L) {
    constexpr unsigned long long RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
        0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
        0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    const uint32_t is0 = L == 0 ? 0xffffffffu : 0u;
#pragma unroll UNROLL
    for (int r = 0; r < 24; ++r) {
        uint32_t cl, ch, ml, mh, pl, ph;
        if (ORDER == 0) {   // interleave the halves
            const uint32_t l1 = SH(lo, c.src5), h1 = SH(hi, c.src5), l2 = SH(lo, c.src10), h2 = SH(hi, c.src10);
            const uint32_t l3 = SH(lo, c.src15), h3 = SH(hi, c.src15), l4 = SH(lo, c.src20), h4 = SH(hi, c.src20);
            cl = lo ^ l1 ^ l2 ^ l3 ^ l4; ch = hi ^ h1 ^ h2 ^ h3 ^ h4;
            ml = SH(cl, c.xm1); pl = SH(cl, c.xp1); mh = SH(ch, c.xm1); ph = SH(ch, c.xp1);
        } else {            // low half first, its stage-2 shuffles issued before the high half's stage 1 completes
            const uint32_t l1 = SH(lo, c.src5), l2 = SH(lo, c.src10), l3 = SH(lo, c.src15), l4 = SH(lo, c.src20);
            const uint32_t h1 = SH(hi, c.src5), h2 = SH(hi, c.src10), h3 = SH(hi, c.src15), h4 = SH(hi, c.src20);
            cl = lo ^ l1 ^ l2 ^ l3 ^ l4;
            ml = SH(cl, c.xm1); pl = SH(cl, c.xp1);
            ch = hi ^ h1 ^ h2 ^ h3 ^ h4;
            mh = SH(ch, c.xm1); ph = SH(ch, c.xp1);
        }
        lo ^= ml ^ __funnelshift_l(ph, pl, 1);
        hi ^= mh ^ __funnelshift_l(pl, ph, 1);
        rho(lo, hi, c.rot);
        const uint32_t b0l = SH(lo, c.pi0), b1l = SH(lo, c.pi1), b2l = SH(lo, c.pi2);
        const uint32_t b0h = SH(hi, c.pi0), b1h = SH(hi, c.pi1), b2h = SH(hi, c.pi2);
        lo = b0l ^ (~b1l & b2l) ^ ((uint32_t)RC[r] & is0);
        hi = b0h ^ (~b1h & b2h) ^ ((uint32_t)(RC[r] >> 32) & is0);
    }
}

// V7: two shuffle stages: after the column parity, every lane gathers -- for each of its three pi sources -- the raw
// lane and the parities of the source's neighbour columns, and applies theta + rho for those three lanes itself.
struct Lane7 { int s[3], cm[3], cp[3], rot[3]; };
__device__ __forceinline__ Lane7 lane7_init(int L) {
    Lane7 q;
    const int l = L < 25 ? L : 0, x = l % 5, y = l / 5;
    for (int k = 0; k < 3; ++k) {
        const int xk = (x + k) % 5;                 // column of B read by chi
        const int src = ((xk + 3 * y) % 5) + 5 * xk; // A lane that pi moves to B[xk][y]
        const int sx = src % 5;
        q.s[k] = src; q.cm[k] = (sx + 4) % 5; q.cp[k] = (sx + 1) % 5; q.rot[k] = c_rot[src];
    }
    return q;
}
__device__ __forceinline__ void f_v7(uint32_t& lo, uint32_t& hi, const Lane& c, const Lane7& q, int L) {
#pragma unroll 1
    for (int r = 0; r < 24; ++r) {
        const uint32_t cl = lo ^ SH(lo, c.src5) ^ SH(lo, c.src10) ^ SH(lo, c.src15) ^ SH(lo, c.src20);
        const uint32_t ch = hi ^ SH(hi, c.src5) ^ SH(hi, c.src10) ^ SH(hi, c.src15) ^ SH(hi, c.src20);
        uint32_t bl[3], bh[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            uint32_t al = SH(lo, q.s[k]), ah = SH(hi, q.s[k]);
            const uint32_t ml = SH(cl, q.cm[k]), mh = SH(ch, q.cm[k]), pl = SH(cl, q.cp[k]), ph = SH(ch, q.cp[k]);
            al ^= ml ^ __funnelshift_l(ph, pl, 1);
            ah ^= mh ^ __funnelshift_l(pl, ph, 1);
            rho(al, ah, q.rot[k]);
            bl[k] = al; bh[k] = ah;
        }
        lo = bl[0] ^ (~bl[1] & bl[2]);
        hi = bh[0] ^ (~bh[1] & bh[2]);
        iota(lo, hi, r, L);
    }
}

// V6: V0 on two independent sponges held by the same warp (instruction-level parallelism across chains)
__device__ __forceinline__ void f_v6(uint32_t& lo, uint32_t& hi, uint32_t& lo2, uint32_t& hi2, const Lane& c, int L) {
#pragma unroll 1
    for (int r = 0; r < 24; ++r) {
        uint32_t cl = lo ^ SH(lo, c.src5) ^ SH(lo, c.src10) ^ SH(lo, c.src15) ^ SH(lo, c.src20);
        uint32_t ch = hi ^ SH(hi, c.src5) ^ SH(hi, c.src10) ^ SH(hi, c.src15) ^ SH(hi, c.src20);
        uint32_t cl2 = lo2 ^ SH(lo2, c.src5) ^ SH(lo2, c.src10) ^ SH(lo2, c.src15) ^ SH(lo2, c.src20);
        uint32_t ch2 = hi2 ^ SH(hi2, c.src5) ^ SH(hi2, c.src10) ^ SH(hi2, c.src15) ^ SH(hi2, c.src20);
        const uint32_t ml = SH(cl, c.xm1), mh = SH(ch, c.xm1), pl = SH(cl, c.xp1), ph = SH(ch, c.xp1);
        const uint32_t ml2 = SH(cl2, c.xm1), mh2 = SH(ch2, c.xm1), pl2 = SH(cl2, c.xp1), ph2 = SH(ch2, c.xp1);
        lo ^= ml ^ __funnelshift_l(ph, pl, 1);
        hi ^= mh ^ __funnelshift_l(pl, ph, 1);
        lo2 ^= ml2 ^ __funnelshift_l(ph2, pl2, 1);
        hi2 ^= mh2 ^ __funnelshift_l(pl2, ph2, 1);
        rho(lo, hi, c.rot);
        rho(lo2, hi2, c.rot);
        const uint32_t b0l = SH(lo, c.pi0), b0h = SH(hi, c.pi0), b1l = SH(lo, c.pi1), b1h = SH(hi, c.pi1), b2l = SH(lo, c.pi2), b2h = SH(hi, c.pi2);
        const uint32_t d0l = SH(lo2, c.pi0), d0h = SH(hi2, c.pi0), d1l = SH(lo2, c.pi1), d1h = SH(hi2, c.pi1), d2l = SH(lo2, c.pi2), d2h = SH(hi2, c.pi2);
        lo = b0l ^ (~b1l & b2l);
        hi = b0h ^ (~b1h & b2h);
        lo2 = d0l ^ (~d1l & d2l);
        hi2 = d0h ^ (~d1h & d2h);
        iota(lo, hi, r, L);
        iota(lo2, hi2, r, L);
    }
}

template <int V>
__global__ void coop2_kernel(int perms, unsigned long long* state_out, long long* cycles_out) {
    const int L = threadIdx.x & 31, w = threadIdx.x >> 5;
    const Lane c = lane_init(L);
    const Lane7 q = lane7_init(L);
    const int chain = blockIdx.x * (blockDim.x >> 5) + w;
    unsigned long long v = L < 25 ? (unsigned long long)(L + 1) * 0x9E3779B97F4A7C15ULL + (unsigned long long)chain : 0ULL;
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32), lo2 = lo ^ 0x55u, hi2 = hi;
    const long long t0 = clock64();
    for (int p = 0; p < perms; ++p) {
        if (V == 6) f_v6(lo, hi, lo2, hi2, c, L);
        if (V == 7) f_v7(lo, hi, c, q, L);
        if (V == 8) f_v8<24>(lo, hi, c, L);
        if (V == 9) f_v8<4>(lo, hi, c, L);
        if (V == 10) f_v8<2>(lo, hi, c, L);
        if (V == 11) f_v8<3>(lo, hi, c, L);
        if (V == 12) f_v8<6>(lo, hi, c, L);
        if (V == 13) f_v8<8>(lo, hi, c, L);
        if (V == 14) f_v8<12>(lo, hi, c, L);
        if (V == 15) f_v8<1>(lo, hi, c, L);
        if (V == 16) f_v16<24, 0>(lo, hi, c, L);
        if (V == 17) f_v16<24, 1>(lo, hi, c, L);
        if (V == 18) f_v16<8, 0>(lo, hi, c, L);
        if (V == 19) f_v16<8, 1>(lo, hi, c, L);
    }
    const long long t1 = clock64();
    if (chain == 0 && L < 25) state_out[L] = ((unsigned long long)hi << 32) | lo;
    if (V == 6 && lo2 == 0x12345u && hi2 == 7u) state_out[0] = 0;   // keep the second chain alive
    if (L == 0) cycles_out[chain] = t1 - t0;
}


// V20: V17 with the column parity (stage 1) through shared memory: 1 STS.64 + 4 LDS.64 instead of 8 shuffles
template <int UNROLL>
__device__ __forceinline__ void f_v20(uint32_t& lo, uint32_t& hi, const Lane& c, int L, uint2* buf) {
    constexpr unsigned long long RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
        0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
        0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    const uint32_t is0 = L == 0 ? 0xffffffffu : 0u;
#pragma unroll UNROLL
    for (int r = 0; r < 24; ++r) {
        uint2* b = buf + (r & 1) * 32;
        b[L] = make_uint2(lo, hi);
        __syncwarp();
        const uint2 a1 = b[c.src5], a2 = b[c.src10], a3 = b[c.src15], a4 = b[c.src20];
        const uint32_t cl = lo ^ a1.x ^ a2.x ^ a3.x ^ a4.x;
        const uint32_t ml = SH(cl, c.xm1), pl = SH(cl, c.xp1);
        const uint32_t ch = hi ^ a1.y ^ a2.y ^ a3.y ^ a4.y;
        const uint32_t mh = SH(ch, c.xm1), ph = SH(ch, c.xp1);
        lo ^= ml ^ __funnelshift_l(ph, pl, 1);
        hi ^= mh ^ __funnelshift_l(pl, ph, 1);
        rho(lo, hi, c.rot);
        const uint32_t b0l = SH(lo, c.pi0), b1l = SH(lo, c.pi1), b2l = SH(lo, c.pi2);
        const uint32_t b0h = SH(hi, c.pi0), b1h = SH(hi, c.pi1), b2h = SH(hi, c.pi2);
        lo = b0l ^ (~b1l & b2l) ^ ((uint32_t)RC[r] & is0);
        hi = b0h ^ (~b1h & b2h) ^ ((uint32_t)(RC[r] >> 32) & is0);
    }
}
template <int UNROLL>
__global__ void coop20_kernel(int perms, unsigned long long* state_out, long long* cycles_out) {
    __shared__ uint2 sm[4 * 64];
    const int L = threadIdx.x & 31, w = threadIdx.x >> 5;
    const Lane c = lane_init(L);
    const int chain = blockIdx.x * (blockDim.x >> 5) + w;
    unsigned long long v = L < 25 ? (unsigned long long)(L + 1) * 0x9E3779B97F4A7C15ULL + (unsigned long long)chain : 0ULL;
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    const long long t0 = clock64();
    for (int p = 0; p < perms; ++p) f_v20<UNROLL>(lo, hi, c, L, sm + w * 64);
    const long long t1 = clock64();
    if (chain == 0 && L < 25) state_out[L] = ((unsigned long long)hi << 32) | lo;
    if (L == 0) cycles_out[chain] = t1 - t0;
}

// SHFL characteristics: dependent chain (latency) and batches of 8 independent shuffles (issue interval)
__global__ void shfl_kernel(int iters, int mode, long long* cycles_out, unsigned* sink) {
    const int L = threadIdx.x & 31;
    unsigned x = threadIdx.x * 2654435761u, y[8];
    const int src = (L * 7 + 3) & 31;
    for (int k = 0; k < 8; ++k) y[k] = x + k;
    const long long t0 = clock64();
    if (mode == 0) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x = SH(x, src);
        }
    } else {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < 8; ++k) y[k] = SH(y[k], src);
        }
    }
    const long long t1 = clock64();
    for (int k = 0; k < 8; ++k) x ^= y[k];
    if (x == 0x1234567u) sink[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles_out[0] = t1 - t0;
}

template <int V>
__global__ void coop_kernel(int perms, unsigned long long* state_out, long long* cycles_out) {
    extern __shared__ uint2 sm[];
    const int L = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint2* bufA = sm + w * 96;
    uint2* bufB = bufA + 32;
    const Lane c = lane_init(L);
    const int chain = blockIdx.x * (blockDim.x >> 5) + w;
    unsigned long long v = L < 25 ? (unsigned long long)(L + 1) * 0x9E3779B97F4A7C15ULL + (unsigned long long)chain : 0ULL;
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    const long long t0 = clock64();
    for (int p = 0; p < perms; ++p) {
        if (V == 0) f_v0(lo, hi, c, L, bufA, bufB);
        if (V == 1) f_v1(lo, hi, c, L, bufA, bufB);
        if (V == 2) f_v2(lo, hi, c, L, bufA, bufB);
        if (V == 3) f_v3(lo, hi, c, L, bufA, bufB);
        if (V == 5) f_v5(lo, hi, c, L, bufA, bufB);
    }
    const long long t1 = clock64();
    if (chain == 0 && L < 25) state_out[L] = ((unsigned long long)hi << 32) | lo;
    if (L == 0) cycles_out[chain] = t1 - t0;
}

// V4: thread per sponge
__global__ void thread_kernel(int perms, unsigned long long* state_out, long long* cycles_out) {
    const int chain = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long a[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) a[k] = (unsigned long long)(k + 1) * 0x9E3779B97F4A7C15ULL + (unsigned long long)chain;
    const long long t0 = clock64();
    for (int p = 0; p < perms; ++p) keccak_f(a);
    const long long t1 = clock64();
    if (chain == 0)
        for (int k = 0; k < 25; ++k) state_out[k] = a[k];
    if ((threadIdx.x & 31) == 0) cycles_out[chain >> 5] = t1 - t0;
}

int main(int argc, char** argv) {
    const int perms = argc > 1 ? atoi(argv[1]) : 2000;
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    printf("device %s, %d SMs, %d perms per chain\n", pr.name, sms, perms);
    unsigned long long* d_state; long long* d_cyc;
    CK(cudaMalloc(&d_state, 25 * 8));
    CK(cudaMalloc(&d_cyc, 1 << 20));
    std::vector<unsigned long long> ref(25), got(25);
    std::vector<long long> cyc(1 << 17);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // reference state from the thread kernel
    thread_kernel<<<1, 32>>>(perms, d_state, d_cyc);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(ref.data(), d_state, 200, cudaMemcpyDeviceToHost));
    auto report = [&](const char* name, int warps_per_sm, int n_chains_warps, float ms, int sponges_per_warp) {
        CK(cudaMemcpy(cyc.data(), d_cyc, n_chains_warps * 8, cudaMemcpyDeviceToHost));
        long long mx = 0; double avg = 0;
        for (int i = 0; i < n_chains_warps; ++i) { mx = cyc[i] > mx ? cyc[i] : mx; avg += (double)cyc[i]; }
        avg /= n_chains_warps;
        CK(cudaMemcpy(got.data(), d_state, 200, cudaMemcpyDeviceToHost));
        bool ok = true;
        for (int k = 0; k < 25; ++k) ok = ok && got[k] == ref[k];
        const double total_perms = (double)n_chains_warps * sponges_per_warp * perms;
        printf("%-28s warps/SM %2d  cycles/round avg %7.1f max %7.1f  time %8.3f ms  %7.3f Gperm/s  %s\n", name, warps_per_sm, avg / perms / 24.0,
               (double)mx / perms / 24.0, ms, total_perms / (ms * 1e-3) / 1e9, ok ? "ok" : "MISMATCH");
    };
    const int wps[] = {1, 2, 4, 8, 12, 16};
#define RUN_COOP(V, NAME)                                                                     \
    for (int wi = 0; wi < 6; ++wi) {                                                          \
        const int w = wps[wi];                                                                \
        const int bw = w < 4 ? w : 4, blocks = sms * (w / bw);                                \
        coop_kernel<V><<<blocks, bw * 32, bw * 96 * 8>>>(10, d_state, d_cyc);                  \
        CK(cudaEventRecord(e0));                                                              \
        coop_kernel<V><<<blocks, bw * 32, bw * 96 * 8>>>(perms, d_state, d_cyc);               \
        CK(cudaEventRecord(e1));                                                              \
        CK(cudaDeviceSynchronize());                                                          \
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));                                      \
        report(NAME, w, blocks * bw, ms, 1);                                                  \
    }
    RUN_COOP(0, "V0 shfl x18")
    RUN_COOP(1, "V1 redux x2 + shfl x10")
    RUN_COOP(2, "V2 smem 3-stage")
    RUN_COOP(3, "V3 smem 2-stage")
    RUN_COOP(5, "V5 redux + smem")

#define RUN_COOP2(V, NAME)                                                                    \
    for (int wi = 0; wi < 6; ++wi) {                                                          \
        const int w = wps[wi];                                                                \
        const int bw = w < 4 ? w : 4, blocks = sms * (w / bw);                                \
        coop2_kernel<V><<<blocks, bw * 32>>>(10, d_state, d_cyc);                              \
        CK(cudaEventRecord(e0));                                                              \
        coop2_kernel<V><<<blocks, bw * 32>>>(perms, d_state, d_cyc);                           \
        CK(cudaEventRecord(e1));                                                              \
        CK(cudaDeviceSynchronize());                                                          \
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));                                      \
        report(NAME, w, blocks * bw, ms, V == 6 ? 2 : 1);                                     \
    }
    RUN_COOP2(6, "V6 shfl x18, 2 chains/warp")
    RUN_COOP2(7, "V7 shfl x26, 2 stages")
    RUN_COOP2(8, "V8 = V0 unrolled x24")
    RUN_COOP2(9, "V9 = V0 unrolled x4")
    RUN_COOP2(16, "V16 x24 interleaved halves")
    RUN_COOP2(17, "V17 x24 low half first")
    RUN_COOP2(18, "V18 x8 interleaved halves")
    RUN_COOP2(19, "V19 x8 low half first")
    RUN_COOP2(15, "V15 = V0 mask-iota, x1")
    RUN_COOP2(10, "V10 = V0 unrolled x2")
    RUN_COOP2(11, "V11 = V0 unrolled x3")
    RUN_COOP2(12, "V12 = V0 unrolled x6")
    RUN_COOP2(13, "V13 = V0 unrolled x8")
    RUN_COOP2(14, "V14 = V0 unrolled x12")

    {   // CTA shape: the same 1250 chains as 1-, 2- and 4-warp CTAs (how evenly do they spread over SMs and sub-partitions?)
        const int shapes[][2] = {{1250, 32}, {625, 64}, {313, 128}, {1184, 32}, {296, 128}, {592, 32}, {148, 128}, {1776, 32}};
        for (auto& sh : shapes) {
            coop2_kernel<17><<<sh[0], sh[1]>>>(10, d_state, d_cyc);
            CK(cudaEventRecord(e0));
            coop2_kernel<17><<<sh[0], sh[1]>>>(perms, d_state, d_cyc);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const int nw = sh[0] * sh[1] / 32;
            CK(cudaMemcpy(cyc.data(), d_cyc, nw * 8, cudaMemcpyDeviceToHost));
            long long mx = 0; double avg = 0;
            for (int i = 0; i < nw; ++i) { mx = cyc[i] > mx ? cyc[i] : mx; avg += (double)cyc[i]; }
            printf("V17 as %4d CTAs x %3d threads: cycles/round avg %.1f max %.1f, %.3f ms\n", sh[0], sh[1], avg / nw / perms / 24.0, (double)mx / perms / 24.0, ms);
        }
    }

    {
        const int shapes[][2] = {{1250, 32}, {1184, 32}, {592, 32}, {148, 32}, {1776, 32}};
        for (auto& sh : shapes) {
            coop20_kernel<24><<<sh[0], sh[1]>>>(10, d_state, d_cyc);
            CK(cudaEventRecord(e0));
            coop20_kernel<24><<<sh[0], sh[1]>>>(perms, d_state, d_cyc);
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const int nw = sh[0] * sh[1] / 32;
            CK(cudaMemcpy(cyc.data(), d_cyc, nw * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(got.data(), d_state, 200, cudaMemcpyDeviceToHost));
            bool ok = true; for (int k = 0; k < 25; ++k) ok = ok && got[k] == ref[k];
            long long mx = 0; double avg = 0;
            for (int i = 0; i < nw; ++i) { mx = cyc[i] > mx ? cyc[i] : mx; avg += (double)cyc[i]; }
            printf("V20 (smem parity) as %4d CTAs x %3d threads: cycles/round avg %.1f max %.1f, %.3f ms %s\n", sh[0], sh[1], avg / nw / perms / 24.0, (double)mx / perms / 24.0, ms, ok ? "ok" : "MISMATCH");
        }
    }
    {
        unsigned* d_sink; CK(cudaMalloc(&d_sink, 4));
        const int it = 10000;
        for (int mode = 0; mode < 2; ++mode)
            for (int w = 1; w <= 16; w *= 2) {
                const int bw = w < 4 ? w : 4, blocks = sms * (w / bw);
                shfl_kernel<<<blocks, bw * 32>>>(it, mode, d_cyc, d_sink);
                CK(cudaDeviceSynchronize());
                long long cy; CK(cudaMemcpy(&cy, d_cyc, 8, cudaMemcpyDeviceToHost));
                printf("SHFL %s warps/SM %2d: %.2f cycles per shuffle (one warp's view)\n", mode ? "8 independent" : "dependent chain", w, (double)cy / it / 8.0);
            }
    }
    const int tw[] = {4, 8, 16, 24, 32};
    for (int wi = 0; wi < 5; ++wi) {
        const int w = tw[wi];
        const int bw = 4, blocks = sms * (w / bw);
        thread_kernel<<<blocks, bw * 32>>>(10, d_state, d_cyc);
        CK(cudaEventRecord(e0));
        thread_kernel<<<blocks, bw * 32>>>(perms / 4, d_state, d_cyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        // the reference state was taken at `perms`; recompute it for perms / 4 is not needed: only report the rate
        CK(cudaMemcpy(cyc.data(), d_cyc, blocks * bw * 8, cudaMemcpyDeviceToHost));
        double avg = 0;
        for (int i = 0; i < blocks * bw; ++i) avg += (double)cyc[i];
        avg /= blocks * bw;
        printf("%-28s warps/SM %2d  cycles/round avg %7.1f  time %8.3f ms  %7.3f Gperm/s\n", "V4 thread per sponge", w, avg / (perms / 4) / 24.0, ms,
               (double)blocks * bw * 32 * (perms / 4) / (ms * 1e-3) / 1e9);
    }
    return 0;
}
