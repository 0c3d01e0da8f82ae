// fz_sample.cu -- seeded samplers, bit-exact with CPython's `random` module as driven by
// algebra/polynomials.py:436-488 (sample_polynomial_coefficient_representation,
// sample_polynomial_ntt_representation): MT19937 seeded with init_by_array over the 32-bit limbs
// of abs(seed) (Modules/_randommodule.c), randrange(n) = top bit_length(n) bits of a 32-bit
// output, retried while >= n (Lib/random.py _randbelow_with_getrandbits).
//
// One thread per polynomial.  The 624-word state of each generator lives in SHARED memory, word-interleaved
// across the threads of the CTA (word k of thread t at k * S + t: every lock-step access is conflict-free), 90
// generators per CTA (90 * 632 words = 222 KiB of the 227 KiB an sm_100a CTA may own).  Seeding is two serial
// recurrences of 624 + 623 steps per generator (init_by_array); the first, seed-independent pass of CPython's
// init_genrand(19650218) is a constant table.  Outputs are produced on demand, in place, eight at a time (the
// regeneration is sequential in k anyway), so no word that is never drawn is ever twisted, and the rejection loops of
// randrange run as one convergent draw loop over a two-state machine.
#include "fz_common.cuh"
#include "fz_internal.h"

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

// mode 0: secret-key coefficients (1 + randrange(bound)) * (1 - 2*randrange(2)), `weight` of them, zero padded,
//         then the Fisher-Yates branch when weight < d (polynomials.py:449-459)
// mode 1: public-matrix values randrange(q) - q//2 (polynomials.py:480)
// keys: [n][key_stride] 32-bit limbs, klen[n] limbs used (>= 1)
__global__ void __launch_bounds__(SMP_THREADS) sample_kernel(int mode, const uint32_t* __restrict__ keys, int key_stride, const int32_t* __restrict__ klen,
                                                              size_t n, int d, uint32_t q, uint32_t bound, int weight, int32_t* __restrict__ out) {
    extern __shared__ __align__(16) uint32_t smp_sm[];
    const size_t i = (size_t)blockIdx.x * SMP_STREAMS + threadIdx.x;
    if (threadIdx.x >= SMP_STREAMS || i >= n) return;   // no block-wide barrier below
    Mt s;
    s.m = smp_sm + threadIdx.x;
    uint32_t key[8];
    int kl = klen[i];
    if (kl > 8) kl = 8;
    for (int j = 0; j < 8; ++j) key[j] = j < kl ? keys[i * key_stride + j] : 0u;
    mt_seed(s, key, kl);
    int32_t* o = out + i * d;
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

// ---- dense secret-key sampler, second version (mode 0 with weight >= d, d <= 256, bound <= 127: the scheme's only use) ----
// tools/sample_lab.cu split the first kernel's time per warp (profiles/r02_sample_lab_v1.txt): seeding 40 K cycles, regeneration + tempering 47 K,
// the accept / reject machine 21 K, the per-lane scattered 4-byte global stores 33 K -- and one of the four sub-partitions of the SM idle (three warps).
// This kernel (i) spreads the CTA's generators over FOUR warps, one per sub-partition; (ii) keeps the (key[j] + j) terms of init_by_array in two registers
// when every seed of the warp has at most two limbs (any u64 seed), which takes the shared-memory load out of the dependent chain of the first
// recurrence, and double-buffers the state words of the second; (iii) skips the last tempering step (it only changes bits 13..0, a draw keeps
// bits 31..shb); (iv) STAGE = 1: emits each accepted coefficient as ONE byte (magnitude | sign << 7) into a shared-memory tile and writes the
// polynomial out afterwards with coalesced 128-byte stores, a warp per polynomial; STAGE = 0: collects four coefficients in registers per 16-byte store.
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

template <int STAGE>
__global__ void __launch_bounds__(SD_THREADS) sample_dense_kernel(const uint32_t* __restrict__ keys, int key_stride, const int32_t* __restrict__ klen, size_t n, int d,
                                                                  uint32_t bound, int32_t* __restrict__ out, long long* __restrict__ cyc) {
    using Cfg = SdCfg<STAGE>;
    constexpr int G = Cfg::G, PW = Cfg::PW;
    extern __shared__ __align__(16) uint32_t smp_sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = warp * PW + lane;
    const size_t base = (size_t)blockIdx.x * G + (size_t)warp * PW;   // first generator of this warp
    const size_t i = base + lane;
    const bool active = lane < PW && i < n;
    uint32_t key[8];
    int kl = 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) key[j] = 0u;
    if (active) {
        kl = klen[i];
        if (kl > 8) kl = 8;
        for (int j = 0; j < 8; ++j) key[j] = j < kl ? keys[i * key_stride + j] : 0u;
    }
    const bool two = __all_sync(0xffffffffu, kl <= 2);
    uint32_t* m = smp_sm + g;
    uint8_t* stage = reinterpret_cast<uint8_t*>(smp_sm + G * MT_N);
    long long t0 = 0, t1 = 0;   // cyc != nullptr (tools/sample_lab.cu only): cycles of the seeding and of the rest, per warp
    if (cyc) t0 = clock64();
    if (active) {
        if (two) mt_seed2<G>(m, key[0], kl == 2 ? key[1] + 1u : key[0]);
        else mt_seed_any<G>(m, key, kl);
        if (cyc) t1 = clock64();
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
            int32_t* o = out + i * d;
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
    if (STAGE) {
        __syncwarp();
        // a warp per polynomial: lane l reads the staged bytes of coefficients l, l + 32, ... and stores them as int32, 128 bytes per instruction
        for (int p = 0; p < PW && base + p < n; ++p) {
            int32_t* o = out + (base + p) * d;
            const uint8_t* col = stage + warp * PW + p;
            for (int k = lane; k < d; k += 32) {
                const uint32_t b = col[k * Cfg::SB];
                const int32_t mg = (int32_t)(b & 0x7fu);
                o[k] = (b & 0x80u) ? -mg : mg;
            }
        }
    }
    if (cyc && lane == 0) {
        const size_t w = (size_t)blockIdx.x * 4 + warp;
        cyc[w * 2] = active ? t1 - t0 : 0;
        cyc[w * 2 + 1] = active ? clock64() - t1 : 0;
    }
}

// u64 seeds -> limb keys for (seed, seed + 1): keys[2i] = limbs(seed_i), keys[2i+1] = limbs(seed_i + 1)   (fusion.py:340, 352)
__global__ void seeds_to_keys_kernel(const unsigned long long* __restrict__ seeds, size_t n, uint32_t* __restrict__ keys, int32_t* __restrict__ klen) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long s0 = seeds[i];
    uint32_t* k0 = keys + (2 * i) * 4;
    uint32_t* k1 = keys + (2 * i + 1) * 4;
    k0[0] = (uint32_t)s0; k0[1] = (uint32_t)(s0 >> 32); k0[2] = 0; k0[3] = 0;
    klen[2 * i] = k0[1] ? 2 : 1;
    const unsigned long long s1 = s0 + 1ULL;
    if (s1 == 0ULL) { k1[0] = 0; k1[1] = 0; k1[2] = 1; k1[3] = 0; klen[2 * i + 1] = 3; }
    else { k1[0] = (uint32_t)s1; k1[1] = (uint32_t)(s1 >> 32); k1[2] = 0; k1[3] = 0; klen[2 * i + 1] = k1[1] ? 2 : 1; }
}

// u64 row seeds [n][2][rows] -> limb keys, side-major [2][n][rows] (general ell-row keygen: every row of a key matrix has its own seed;
// side-major so that each side is one contiguous sampler launch writing straight into sk_l / sk_r)
__global__ void row_seeds_to_keys_kernel(const unsigned long long* __restrict__ seeds, size_t n, int rows, uint32_t* __restrict__ keys,
                                         int32_t* __restrict__ klen) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * n * rows) return;
    const size_t i = t / (2 * (size_t)rows), r = t % (2 * (size_t)rows);
    const size_t side = r / rows, k = r % rows;
    const size_t o = side * n * rows + i * rows + k;
    const unsigned long long s0 = seeds[t];
    uint32_t* k0 = keys + o * 4;
    k0[0] = (uint32_t)s0; k0[1] = (uint32_t)(s0 >> 32); k0[2] = 0; k0[3] = 0;
    klen[o] = k0[1] ? 2 : 1;
}

// dst[b][r][:] = src[b][:] for r < rows  (quirk Q1: every row of a seeded key matrix is the same polynomial)
__global__ void __launch_bounds__(256) replicate_rows_kernel(const int4* __restrict__ src, int4* __restrict__ dst, size_t batch0, int rows, int d4,
                                                               size_t src_batch_stride4) {
    const int per = rows * d4;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per) return;
    const size_t b = batch0 + blockIdx.y;
    st_stream(dst + b * (size_t)per + i, __ldg(src + b * src_batch_stride4 + (i & (d4 - 1))));
}

int sample_init_tables(fz_ctx* ctx) {
    uint32_t mt0[MT_N];
    mt0[0] = 19650218u;
    for (int i = 1; i < MT_N; ++i) mt0[i] = 1812433253u * (mt0[i - 1] ^ (mt0[i - 1] >> 30)) + (uint32_t)i;   // init_genrand
    if (check(cudaMemcpyToSymbol(c_mt0, mt0, sizeof(mt0)), "MT19937 seed table upload")) return -2;
    (void)ctx;
    if (check(cudaFuncSetAttribute(sample_dense_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SdCfg<0>::SMEM), "sample_dense_kernel<0> smem attr")) return -2;
    if (check(cudaFuncSetAttribute(sample_dense_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SdCfg<1>::SMEM), "sample_dense_kernel<1> smem attr")) return -2;
    return check(cudaFuncSetAttribute(sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMP_SMEM), "sample_kernel smem attr");
}

// which kernel serves the dense secret-key case (fz_set_option "sample_impl"): 0 the general kernel, 1 sample_dense_kernel<1> (byte tile, the default),
// 2 sample_dense_kernel<0> (16-byte stores).  262,144 generators through fz_sample_coef into a buffer that is not in L2: 1.098 ms (1), 1.428 ms (2),
// 1.520 ms (0); the 16-byte stores only look as good as the tile (1.079 ms, profiles/r02_sample_lab_v2.txt) when the output still sits in L2 from a
// memset -- half-sector writes to lines that are not resident cost a fill each, the tile's 128-byte coalesced stores do not.
static int g_sample_impl = 1;
void set_sample_impl(int v) { g_sample_impl = v; }

// true when the dense kernel took the launch: dense weights, d <= 256 (the byte tile holds 256 rows), magnitudes that fit seven bits
static bool launch_sample_dense(int d, int mode, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, uint32_t bound, int weight, int32_t* out,
                                cudaStream_t st) {
    if (g_sample_impl == 0 || mode != 0 || weight < d || d > 256 || d < 4 || (d & 3) || bound < 1 || bound > 127) return false;
    if (g_sample_impl == 2 && (reinterpret_cast<uintptr_t>(out) & 15) == 0)
        sample_dense_kernel<0><<<(unsigned)((n + SdCfg<0>::G - 1) / SdCfg<0>::G), SD_THREADS, SdCfg<0>::SMEM, st>>>(keys, key_stride, klen, n, d, bound, out, nullptr);
    else
        sample_dense_kernel<1><<<(unsigned)((n + SdCfg<1>::G - 1) / SdCfg<1>::G), SD_THREADS, SdCfg<1>::SMEM, st>>>(keys, key_stride, klen, n, d, bound, out, nullptr);
    return true;
}

int launch_sample(const fz_ctx* ctx, int mode, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, uint32_t bound, int weight,
                  int32_t* out, cudaStream_t st) {
    if (n == 0) return 0;
    if (launch_sample_dense(ctx->d, mode, keys, key_stride, klen, n, bound, weight, out, st)) return check(cudaGetLastError(), "dense sample launch");
    sample_kernel<<<(unsigned)((n + SMP_STREAMS - 1) / SMP_STREAMS), SMP_THREADS, SMP_SMEM, st>>>(mode, keys, key_stride, klen, n, ctx->d, Q, bound, weight, out);
    return check(cudaGetLastError(), "sample launch");
}

// the same samplers for any ring (fz_ring_sample_*): degree and modulus are runtime arguments of the kernel
int launch_sample_ring(int d, uint32_t q, int mode, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, uint32_t bound, int weight,
                       int32_t* out, cudaStream_t st) {
    if (n == 0) return 0;
    if (launch_sample_dense(d, mode, keys, key_stride, klen, n, bound, weight, out, st)) return check(cudaGetLastError(), "dense ring sample launch");
    sample_kernel<<<(unsigned)((n + SMP_STREAMS - 1) / SMP_STREAMS), SMP_THREADS, SMP_SMEM, st>>>(mode, keys, key_stride, klen, n, d, q, bound, weight, out);
    return check(cudaGetLastError(), "ring sample launch");
}

int launch_seeds_to_keys(const fz_ctx* ctx, const unsigned long long* seeds, size_t n, uint32_t* keys, int32_t* klen, cudaStream_t st) {
    if (n == 0) return 0;
    seeds_to_keys_kernel<<<(int)((n + 127) / 128), 128, 0, st>>>(seeds, n, keys, klen);
    (void)ctx;
    return check(cudaGetLastError(), "seeds_to_keys launch");
}

int launch_row_seeds_to_keys(const fz_ctx* ctx, const unsigned long long* seeds, size_t n, int rows, uint32_t* keys, int32_t* klen, cudaStream_t st) {
    if (n == 0) return 0;
    row_seeds_to_keys_kernel<<<(unsigned)((2 * n * rows + 127) / 128), 128, 0, st>>>(seeds, n, rows, keys, klen);
    (void)ctx;
    return check(cudaGetLastError(), "row_seeds_to_keys launch");
}

int launch_replicate_rows(const fz_ctx* ctx, const int32_t* src, int32_t* dst, size_t batch, int rows, size_t src_batch_stride, cudaStream_t st) {
    if (batch == 0) return 0;
    const int d4 = ctx->d / 4;
    const unsigned gx = (unsigned)((rows * d4 + 255) / 256);
    for (size_t b0 = 0; b0 < batch; b0 += 65535) {
        const size_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
        replicate_rows_kernel<<<dim3(gx, (unsigned)nb), 256, 0, st>>>((const int4*)src, (int4*)dst, b0, rows, d4, src_batch_stride / 4);
        if (check(cudaGetLastError(), "replicate_rows launch")) return -2;
    }
    return 0;
}

}  // namespace fz
