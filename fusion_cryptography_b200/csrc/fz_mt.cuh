// fz_mt.cuh -- MT19937 as CPython's `random` module drives it (Modules/_randommodule.c init_by_array, Lib/random.py _randbelow_with_getrandbits) and the
// per-thread draw loops of the samplers (algebra/polynomials.py:436-488).  fz_sample.cu wraps them in the kernels (which generator a thread owns, the
// shared-memory state columns, the coalesced write-out); the CPU suite compiles the same functions for the host and checks them against CPython's
// `random` (tests/test_sampler_host.py).  The __constant__ table is DEFINED here: include this header in exactly one translation unit.
#pragma once
#include "fz_common.cuh"

namespace fz {

constexpr int MT_N = 624, MT_M = 397;
constexpr int SMP_STREAMS = 90;             // generators per CTA
constexpr int SMP_THREADS = 96;
#ifndef FZ_SMP_TRIP
#define FZ_SMP_TRIP 16
#endif
constexpr int SMP_TRIP = FZ_SMP_TRIP;   // outputs regenerated per trip of the dense draw loop
constexpr int SMP_WORDS = MT_N + 8;         // state + the (key[j] + j) table of init_by_array
constexpr int SMP_SMEM = SMP_STREAMS * SMP_WORDS * 4;

__constant__ uint32_t c_mt0[MT_N];          // init_genrand(19650218)
static inline void mt_build_seed_table(uint32_t* mt0) {   // host: the seed-independent first pass of init_by_array
    mt0[0] = 19650218u;
    for (int i = 1; i < MT_N; ++i) mt0[i] = 1812433253u * (mt0[i - 1] ^ (mt0[i - 1] >> 30)) + (uint32_t)i;   // init_genrand
}

struct Mt {
    uint32_t* m;   // word k of this generator is m[k * SMP_STREAMS]
    int idx;       // next state word to regenerate / output
};
#define MTW(s, k) (s).m[(k) * SMP_STREAMS]

__device__ __forceinline__ uint32_t mt_mix(uint32_t prev) { return prev ^ (prev >> 30); }

// Modules/_randommodule.c init_by_array for klen <= 8 limbs (the key loop runs max(624, klen) = 624 times)
__device__ __forceinline__ void mt_seed(Mt& s, const uint32_t* key, int klen) {
    for (int j = 0; j < 8; ++j) MTW(s, MT_N + j) = key[j] + (uint32_t)j;
    uint32_t prev = 19650218u;
    int j = 0;
#pragma unroll 8
    for (int i = 1; i < MT_N; ++i) {
        const uint32_t v = (c_mt0[i] ^ (mt_mix(prev) * 1664525u)) + MTW(s, MT_N + j);
        MTW(s, i) = v;
        prev = v;
        if (++j >= klen) j = 0;
    }
    {   // 624th step: i wrapped to 1 with mt[0] = mt[623]
        const uint32_t v = (MTW(s, 1) ^ (mt_mix(prev) * 1664525u)) + MTW(s, MT_N + j);
        MTW(s, 1) = v;
        prev = v;
    }
    for (int i0 = 2; i0 < MT_N; i0 += 8) {   // 622 steps, i = 2..623, eight state words prefetched per trip
        uint32_t cur[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) cur[e] = i0 + e < MT_N ? MTW(s, i0 + e) : 0u;
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (i0 + e < MT_N) {
                const uint32_t v = (cur[e] ^ (mt_mix(prev) * 1566083941u)) - (uint32_t)(i0 + e);
                MTW(s, i0 + e) = v;
                prev = v;
            }
    }
    MTW(s, 1) = (MTW(s, 1) ^ (mt_mix(prev) * 1566083941u)) - 1u;   // 623rd step: i wrapped to 1
    MTW(s, 0) = 0x80000000u;
    s.idx = 0;   // CPython leaves index = 624: the first draw regenerates; here regeneration is on demand from word 0
}

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

__device__ __forceinline__ uint32_t mt_regen(uint32_t a, uint32_t b, uint32_t far) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}

// one output (scalar path: public matrix, sparse branch)
__device__ __forceinline__ uint32_t mt_next(Mt& s) {
    const int k = s.idx, k1 = k + 1 == MT_N ? 0 : k + 1, km = k + MT_M < MT_N ? k + MT_M : k + MT_M - MT_N;
    const uint32_t v = mt_regen(MTW(s, k), MTW(s, k1), MTW(s, km));
    MTW(s, k) = v;
    s.idx = k1;
    return mt_temper(v);
}

// B outputs per trip (B = 8 or 16; 624 = 78 * 8 = 39 * 16, so s.idx stays a multiple of B): all 2B + 1 state words are loaded before the first is
// regenerated -- with three warps per SM the loads of a trip are the only memory-level parallelism there is
template <int B>
__device__ __forceinline__ void mt_next_b(Mt& s, uint32_t (&y)[B]) {
    const int k = s.idx;
    uint32_t a[B + 1], f[B];
#pragma unroll
    for (int e = 0; e < B; ++e) a[e] = MTW(s, k + e);
    a[B] = MTW(s, k + B == MT_N ? 0 : k + B);   // word 0 was regenerated at the start of this round, as in the sequential twist
#pragma unroll
    for (int e = 0; e < B; ++e) {
        const int km = k + e + MT_M < MT_N ? k + e + MT_M : k + e + MT_M - MT_N;   // old word above k, new word below
        f[e] = MTW(s, km);
    }
#pragma unroll
    for (int e = 0; e < B; ++e) {
        const uint32_t v = mt_regen(a[e], a[e + 1], f[e]);
        MTW(s, k + e) = v;
        y[e] = mt_temper(v);
    }
    s.idx = k + B == MT_N ? 0 : k + B;
}

__device__ __forceinline__ uint32_t mt_randbelow(Mt& s, uint32_t n) {
    const int k = 32 - __clz(n);  // n.bit_length()
    uint32_t r = mt_next(s) >> (32 - k);
    while (r >= n) r = mt_next(s) >> (32 - k);
    return r;
}

// what a thread of sample_kernel does once its generator is seeded:
// mode 0: secret-key coefficients (1 + randrange(bound)) * (1 - 2*randrange(2)), `weight` of them, zero padded,
//         then the Fisher-Yates branch when weight < d (polynomials.py:449-459)
// mode 1: public-matrix values randrange(q) - q//2 (polynomials.py:480)
__device__ __forceinline__ void sample_draws(Mt& s, int mode, int d, uint32_t q, uint32_t bound, int weight, int32_t* __restrict__ o) {
    if (mode == 1) {
        for (int k = 0; k < d; ++k) o[k] = (int32_t)((long long)mt_randbelow(s, q) - (long long)(q / 2));   // q = 2147465729 for the scheme
        return;
    }
    if (weight >= d) {
        // dense branch (omega_sk = d, the scheme's only use): a convergent draw loop.  randrange(bound) keeps the top
        // bit_length(bound) bits of an output and retries while >= bound; randrange(2) keeps the top 2 bits.
        const int shb = __clz(bound);   // 32 - bit_length(bound)
        // branch-free two-state machine: per output, either a magnitude draw is accepted, or a
        // signum draw is accepted and one coefficient is emitted (a predicated store), or nothing happens; measured against
        // two packed-store variants with tools/sample_times.py: 1.55 ms vs 1.66 / 2.05 ms for 262,144 generators
        int k = 0;
        uint32_t have_mag = 0, mag = 0;
        while (k < d) {
            uint32_t y[SMP_TRIP];
            mt_next_b<SMP_TRIP>(s, y);
#pragma unroll
            for (int e = 0; e < SMP_TRIP; ++e) {
                const uint32_t rm = y[e] >> shb, rs = y[e] >> 30;
                const uint32_t live = k < d ? 1u : 0u;
                const uint32_t take_mag = live & (have_mag ^ 1u) & (rm < bound ? 1u : 0u);
                const uint32_t take_sg = live & have_mag & (rs < 2u ? 1u : 0u);
                mag = take_mag ? rm + 1u : mag;
                const int32_t c = rs ? -(int32_t)mag : (int32_t)mag;
                if (take_sg) o[k] = c;
                k += (int)take_sg;
                have_mag = take_mag | (have_mag & (take_sg ^ 1u));
            }
        }
        return;
    }
    const int ncoef = weight < 0 ? 0 : weight;
    for (int k = 0; k < ncoef; ++k) {
        const int32_t mag = 1 + (int32_t)mt_randbelow(s, bound);
        const int32_t sg = 1 - 2 * (int32_t)mt_randbelow(s, 2u);
        o[k] = mag * sg;
    }
    for (int k = ncoef; k < d; ++k) o[k] = 0;
    for (int k = d - 1; k > 0; --k) {
        const int j = (int)mt_randbelow(s, (uint32_t)(k + 1));
        const int32_t t = o[k]; o[k] = o[j]; o[j] = t;
    }
}

constexpr int SD_THREADS = 128;
template <int STAGE> struct SdCfg {
    static constexpr int PW = STAGE ? 21 : 23;          // generators per warp
    static constexpr int G = 4 * PW;                    // generators per CTA: 84 / 92
    static constexpr int SB = STAGE ? G : 0;            // bytes per staged coefficient row (84 = 21 words: odd, so a column read is conflict-free)
    static constexpr int SMEM = G * MT_N * 4 + 256 * SB;
};

// init_by_array with key length <= 2: the term of step s is key[s % klen] + s % klen, i.e. addA on even and addB on odd steps (addB = addA for one limb)
template <int G>
__device__ __forceinline__ void mt_seed2(uint32_t* m, uint32_t addA, uint32_t addB) {
    uint32_t prev = 19650218u, first = 0;
#pragma unroll 8
    for (int i = 1; i < MT_N; ++i) {   // step s = i - 1
        const uint32_t v = (c_mt0[i] ^ (mt_mix(prev) * 1664525u)) + ((i & 1) ? addA : addB);
        if (i == 1) first = v; else m[i * G] = v;
        prev = v;
    }
    // step 623: i wrapped to 1 with mt[0] = mt[623]
    prev = (first ^ (mt_mix(prev) * 1664525u)) + addB;
    const uint32_t w1 = prev;
    // second recurrence, i = 2..623, two blocks of eight state words in flight
    uint32_t a[8], b[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = m[(2 + e) * G];
    for (int i0 = 2; i0 < MT_N; i0 += 16) {
#pragma unroll
        for (int e = 0; e < 8; ++e) b[e] = i0 + 8 + e < MT_N ? m[(i0 + 8 + e) * G] : 0u;
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (i0 + e < MT_N) {
                const uint32_t v = (a[e] ^ (mt_mix(prev) * 1566083941u)) - (uint32_t)(i0 + e);
                m[(i0 + e) * G] = v;
                prev = v;
            }
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = i0 + 16 + e < MT_N ? m[(i0 + 16 + e) * G] : 0u;
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (i0 + 8 + e < MT_N) {
                const uint32_t v = (b[e] ^ (mt_mix(prev) * 1566083941u)) - (uint32_t)(i0 + 8 + e);
                m[(i0 + 8 + e) * G] = v;
                prev = v;
            }
    }
    m[G] = (w1 ^ (mt_mix(prev) * 1566083941u)) - 1u;   // 623rd step: i wrapped to 1
    m[0] = 0x80000000u;
}

// any key length (seeds of more than 64 bits): the terms come from a per-thread table (local memory; rare)
template <int G>
__device__ __noinline__ void mt_seed_any(uint32_t* m, const uint32_t* key, int klen) {
    uint32_t kj[8];
    for (int j = 0; j < 8; ++j) kj[j] = key[j] + (uint32_t)j;
    uint32_t prev = 19650218u;
    int j = 0;
    for (int i = 1; i < MT_N; ++i) {
        const uint32_t v = (c_mt0[i] ^ (mt_mix(prev) * 1664525u)) + kj[j];
        m[i * G] = v;
        prev = v;
        if (++j >= klen) j = 0;
    }
    prev = (m[G] ^ (mt_mix(prev) * 1664525u)) + kj[j];
    m[G] = prev;
    for (int i = 2; i < MT_N; ++i) {
        prev = (m[i * G] ^ (mt_mix(prev) * 1566083941u)) - (uint32_t)i;
        m[i * G] = prev;
    }
    m[G] = (m[G] ^ (mt_mix(prev) * 1566083941u)) - 1u;
    m[0] = 0x80000000u;
}

// B outputs of generator column m (word k at m[k * G]), tempered without the last step: bits 31..14 are exact
template <int G, int B>
__device__ __forceinline__ void mt_next_top(uint32_t* m, int& idx, uint32_t (&y)[B]) {
    const int k = idx;
    uint32_t a[B + 1], f[B];
#pragma unroll
    for (int e = 0; e < B; ++e) a[e] = m[(k + e) * G];
    a[B] = m[(k + B == MT_N ? 0 : k + B) * G];
#pragma unroll
    for (int e = 0; e < B; ++e) {
        const int km = k + e + MT_M < MT_N ? k + e + MT_M : k + e + MT_M - MT_N;
        f[e] = m[km * G];
    }
#pragma unroll
    for (int e = 0; e < B; ++e) {
        uint32_t v = mt_regen(a[e], a[e + 1], f[e]);
        m[(k + e) * G] = v;
        v ^= v >> 11;
        v ^= (v << 7) & 0x9d2c5680u;
        v ^= (v << 15) & 0xefc60000u;
        y[e] = v;
    }
    idx = k + B == MT_N ? 0 : k + B;
}

// what a thread of sample_dense_kernel does once its generator (column m) is seeded: STAGE = 1 stages one byte per coefficient (magnitude | sign << 7)
// at stage[k * SB + g]; STAGE = 0 writes four coefficients per 16-byte store to o
template <int STAGE>
__device__ __forceinline__ void dense_draws(uint32_t* m, int g, int d, uint32_t bound, uint8_t* stage, int32_t* o) {
    using Cfg = SdCfg<STAGE>;
    constexpr int G = Cfg::G;
    int idx = 0;
    const int shb = __clz(bound);   // a draw keeps the top bit_length(bound) bits: 32 - shb <= 7 of them here
    uint32_t have_mag = 0, mag = 0;
    if (STAGE) {
        uint32_t pos = (uint32_t)g;                               // byte offset of the next coefficient: k * SB + g
        const uint32_t end = (uint32_t)g + (uint32_t)d * Cfg::SB;
        while (pos < end) {
            uint32_t y[SMP_TRIP];
            mt_next_top<G, SMP_TRIP>(m, idx, y);
#pragma unroll
            for (int e = 0; e < SMP_TRIP; ++e) {
                const uint32_t rm = y[e] >> shb;
                const bool take_mag = !have_mag && rm < bound;
                const bool take_sg = have_mag && (int32_t)y[e] >= 0 && pos < end;   // randrange(2): top two bits < 2; bit 30 is the draw
                mag = take_mag ? rm + 1u : mag;
                if (take_sg) stage[pos] = (uint8_t)(mag | ((y[e] >> 23) & 0x80u));
                pos += take_sg ? (uint32_t)Cfg::SB : 0u;
                have_mag = take_mag ? 1u : (take_sg ? 0u : have_mag);
            }
        }
    } else {
        int k = 0;
        int32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0;
        while (k < d) {
            uint32_t y[SMP_TRIP];
            mt_next_top<G, SMP_TRIP>(m, idx, y);
#pragma unroll
            for (int e = 0; e < SMP_TRIP; ++e) {
                const uint32_t rm = y[e] >> shb;
                const bool take_mag = !have_mag && rm < bound;
                const bool take_sg = have_mag && (int32_t)y[e] >= 0 && k < d;
                mag = take_mag ? rm + 1u : mag;
                const int32_t c = (y[e] & 0x40000000u) ? -(int32_t)mag : (int32_t)mag;
                if (take_sg) { r0 = r1; r1 = r2; r2 = r3; r3 = c; ++k; }
                if (take_sg && (k & 3) == 0) *reinterpret_cast<int4*>(o + k - 4) = make_int4(r0, r1, r2, r3);
                have_mag = take_mag ? 1u : (take_sg ? 0u : have_mag);
            }
        }
    }
}

}  // namespace fz
