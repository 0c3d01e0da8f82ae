// fz_ntt_core.cuh -- the arithmetic of the scheme's negacyclic NTT / INTT (algebra/ntt.py:216-377): the twiddle tables, the two butterflies and the
// stage functions a lane runs on its 16 coefficients, plus the host routine that builds the tables.  fz_ntt.cu wraps them in the kernels (loads,
// the warp-private transpose, stores); the CPU suite compiles the same functions for the host and runs whole transforms through them against the
// reference's vectors (tests/test_ntt_host.py).  The __constant__ tables are DEFINED here: include this header in exactly one translation unit.
#pragma once
#include "fz_common.cuh"

namespace fz {

__constant__ Sh c_fwd256[256];
__constant__ Sh c_inv256[256];
__constant__ Sh c_fwd64[64];
__constant__ Sh c_inv64[64];
// [0] = n^-1, [1] = n^-1 * itw[1]  (last inverse stage with the scaling folded in); [2], [3] = the same times 2^32 (fused pass: pays the
// 2^-32 of the Montgomery pointwise product)
__constant__ Tw c_last256[4];
__constant__ Tw c_last64[4];

template <int LOGD> struct Cfg;
template <> struct Cfg<8> {
    static constexpr int D = 256, TPP = 16, PPW = 2, WARP_WORDS = 2 * 336;
    static __device__ __forceinline__ const Sh* fwd() { return c_fwd256; }
    static __device__ __forceinline__ const Sh* inv() { return c_inv256; }
    static __device__ __forceinline__ const Tw* last() { return c_last256; }
    // row of 16 coefficients padded to 20 words; second polynomial of the warp shifted by 16 banks
    static __device__ __forceinline__ int base(int p) { return p * 336; }
};
template <> struct Cfg<6> {
    static constexpr int D = 64, TPP = 4, PPW = 8, WARP_WORDS = 4 * 164;
    static __device__ __forceinline__ const Sh* fwd() { return c_fwd64; }
    static __device__ __forceinline__ const Sh* inv() { return c_inv64; }
    static __device__ __forceinline__ const Tw* last() { return c_last64; }
    static __device__ __forceinline__ int base(int p) { return (p >> 1) * 164 + (p & 1) * 80; }
};

__device__ __forceinline__ int row_off(int e) { return (e >> 4) * 20 + (e & 15); }

// Cooley-Tukey butterfly, algebra/ntt.py:283-289: (u, v*s) -> (u + v*s, u - v*s)
__device__ __forceinline__ void bf_ct(uint32_t& X, uint32_t& Y, uint32_t w, uint32_t wq) {
    uint32_t t = red_q(sh_mul(Y, w, wq));
    uint32_t xr = red_q(X);
    X = xr + t;
    Y = xr - t + Q;
}

// Gentleman-Sande butterfly, algebra/ntt.py:363-369: (u, v) -> (u + v, (u - v)*s); u, v in [0, q]
__device__ __forceinline__ void bf_gs(uint32_t& U, uint32_t& V, uint32_t w, uint32_t wq) {
    uint32_t a = U + V;
    uint32_t b = U - V + Q;
    U = red_q(a);
    V = red_q(sh_mul(b, w, wq));
}

// twiddles of the contiguous pass for lane-in-polynomial tau: slot 8/t - 1 + j  <->  table[d/(2t) + (8/t)*tau + j]
template <int LOGD>
__device__ __forceinline__ void load_tw2(const Sh* table, int tau, Sh (&tw2)[15]) {
    constexpr int D = Cfg<LOGD>::D;
#pragma unroll
    for (int t = 8; t >= 1; t >>= 1) {
        if (t * 32 > D) continue;  // stage does not exist for this degree (d = 64: only t = 2, 1)
#pragma unroll
        for (int j = 0; j < 8 / t; ++j) tw2[8 / t - 1 + j] = table[D / (2 * t) + (8 / t) * tau + j];
    }
}

// forward, long-stride stages (m = 1, 2, 4, 8); lane holds e = tau + TPP*k
template <int LOGD>
__device__ __forceinline__ void fwd_strided(uint32_t (&x)[16]) {
    const Sh* tw = Cfg<LOGD>::fwd();
#pragma unroll
    for (int s = 1; s <= 4; ++s) {
        const int dk = 16 >> s, m = 1 << (s - 1);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k & dk) continue;
            const int i = k >> (5 - s);
            bf_ct(x[k], x[k + dk], tw[m + i].w, tw[m + i].wq);
        }
    }
}

// forward, short-stride stages on 16 contiguous coefficients
template <int LOGD>
__device__ __forceinline__ void fwd_contig(uint32_t (&y)[16], const Sh (&tw2)[15]) {
#pragma unroll
    for (int t = 8; t >= 1; t >>= 1) {
        if (t * 32 > Cfg<LOGD>::D) continue;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k & t) continue;
            const int j = k / (2 * t);
            bf_ct(y[k], y[k + t], tw2[8 / t - 1 + j].w, tw2[8 / t - 1 + j].wq);
        }
    }
}

// inverse, short-stride stages first (t = 1, 2, 4, 8) on 16 contiguous coefficients in [0, q]
template <int LOGD>
__device__ __forceinline__ void inv_contig(uint32_t (&y)[16], const Sh (&tw2)[15]) {
#pragma unroll
    for (int t = 1; t <= 8; t <<= 1) {
        if (t * 32 > Cfg<LOGD>::D) continue;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k & t) continue;
            const int j = k / (2 * t);
            bf_gs(y[k], y[k + t], tw2[8 / t - 1 + j].w, tw2[8 / t - 1 + j].wq);
        }
    }
}

// inverse, long-stride stages; the last one carries n^-1 and emits canonical centred values
template <int LOGD, int LAST = 0>
__device__ __forceinline__ void inv_strided(uint32_t (&x)[16], int32_t (&o)[16]) {
    const Sh* tw = Cfg<LOGD>::inv();
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int dk = 1 << j, h = 8 >> j;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k & dk) continue;
            const int i = k >> (j + 1);
            bf_gs(x[k], x[k + dk], tw[h + i].w, tw[h + i].wq);
        }
    }
    const Tw* lw = Cfg<LOGD>::last();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        uint32_t a = x[k] + x[k + 8];
        uint32_t b = x[k] - x[k + 8] + Q;
        o[k] = pl_mul_cent(a, lw[LAST].lo, lw[LAST].hi);
        o[k + 8] = pl_mul_cent(b, lw[LAST + 1].lo, lw[LAST + 1].hi);
    }
}

// ---------------------------------------------------------------- table construction (host)
static unsigned long long h_powmod(unsigned long long b, unsigned long long e) {
    unsigned __int128 r = 1, x = b % Q;
    while (e) {
        if (e & 1) r = (r * x) % Q;
        x = (x * x) % Q;
        e >>= 1;
    }
    return (unsigned long long)r;
}

static Tw h_plantard(unsigned long long w) {
    unsigned long long b = (unsigned long long)(((unsigned __int128)(w % Q) * NEG_N2) % Q);
    unsigned long long W = b * QINV64;  // mod 2^64
    Tw t;
    t.lo = (uint32_t)W;
    t.hi = (uint32_t)(W >> 32);
    return t;
}

static int h_brv(int i, int bits) {
    int r = 0;
    for (int k = 0; k < bits; ++k)
        if (i & (1 << k)) r |= 1 << (bits - 1 - k);
    return r;
}

// tw[k] = psi^brv(k) mod q (algebra/polynomials.py:396-397, algebra/ntt.py:74-90) as Shoup pairs, and the four constants of the last inverse stage;
// f, iv: d entries each, last: 4 entries.  levels: (logd, psi) = (6, 23584283), (8, 3337519)   (fusion.py:36-37)
static inline void ntt_build_tables(int logd, unsigned long long psi, Sh* f, Sh* iv, Tw* last) {
    const int d = 1 << logd;
    const unsigned long long ipsi = h_powmod(psi, Q - 2), ninv = h_powmod(d, Q - 2);
    unsigned long long itw1 = 0;
    for (int k = 0; k < d; ++k) {
        unsigned long long w = h_powmod(psi, h_brv(k, logd)), iw = h_powmod(ipsi, h_brv(k, logd));
        f[k].w = (uint32_t)w;
        f[k].wq = (uint32_t)((((unsigned __int128)w) << 32) / Q);
        iv[k].w = (uint32_t)iw;
        iv[k].wq = (uint32_t)((((unsigned __int128)iw) << 32) / Q);
        if (k == 1) itw1 = iw;
    }
    last[0] = h_plantard(ninv);
    last[1] = h_plantard((unsigned long long)(((unsigned __int128)ninv * itw1) % Q));
    const unsigned long long r32 = (1ULL << 32) % Q;
    last[2] = h_plantard((unsigned long long)(((unsigned __int128)ninv * r32) % Q));
    last[3] = h_plantard((unsigned long long)(((unsigned __int128)ninv * itw1 % Q) * r32 % Q));
}

}  // namespace fz
