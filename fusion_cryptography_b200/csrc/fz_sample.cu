// fz_sample.cu -- seeded samplers, bit-exact with CPython's `random` module as driven by
// algebra/polynomials.py:436-488 (sample_polynomial_coefficient_representation,
// sample_polynomial_ntt_representation): MT19937 seeded with init_by_array over the 32-bit limbs
// of abs(seed) (Modules/_randommodule.c), randrange(n) = top bit_length(n) bits of a 32-bit
// output, retried while >= n (Lib/random.py _randbelow_with_getrandbits).
//
// One thread per polynomial; the 624-word state lives in local memory, which the hardware
// interleaves across the threads of a warp, so every state access of the (lock-step) seeding
// and twist loops is a coalesced 128-byte transaction.
#include "fz_common.cuh"
#include "fz_internal.h"

namespace fz {

struct Mt {
    uint32_t mt[624];
    int idx;
};

__device__ __forceinline__ void mt_seed(Mt& s, const uint32_t* key, int klen) {
    s.mt[0] = 19650218u;
    for (int i = 1; i < 624; ++i) s.mt[i] = 1812433253u * (s.mt[i - 1] ^ (s.mt[i - 1] >> 30)) + (uint32_t)i;
    int i = 1, j = 0;
    for (int k = klen > 624 ? klen : 624; k; --k) {
        s.mt[i] = (s.mt[i] ^ ((s.mt[i - 1] ^ (s.mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        ++i; ++j;
        if (i >= 624) { s.mt[0] = s.mt[623]; i = 1; }
        if (j >= klen) j = 0;
    }
    for (int k = 623; k; --k) {
        s.mt[i] = (s.mt[i] ^ ((s.mt[i - 1] ^ (s.mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        ++i;
        if (i >= 624) { s.mt[0] = s.mt[623]; i = 1; }
    }
    s.mt[0] = 0x80000000u;
    s.idx = 624;
}

__device__ __forceinline__ void mt_twist(Mt& s) {
    for (int kk = 0; kk < 624; ++kk) {
        const uint32_t y = (s.mt[kk] & 0x80000000u) | (s.mt[kk == 623 ? 0 : kk + 1] & 0x7fffffffu);
        const int k2 = kk + 397 < 624 ? kk + 397 : kk + 397 - 624;
        s.mt[kk] = s.mt[k2] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    s.idx = 0;
}

__device__ __forceinline__ uint32_t mt_next(Mt& s) {
    if (s.idx >= 624) mt_twist(s);
    uint32_t y = s.mt[s.idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
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
__global__ void __launch_bounds__(64) sample_kernel(int mode, const uint32_t* __restrict__ keys, int key_stride, const int32_t* __restrict__ klen, size_t n,
                                                     int d, uint32_t bound, int weight, int32_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Mt s;
    uint32_t key[8];
    int kl = klen[i];
    if (kl > 8) kl = 8;
    for (int j = 0; j < 8; ++j) key[j] = j < kl ? keys[i * key_stride + j] : 0u;
    mt_seed(s, key, kl);
    int32_t* o = out + i * d;
    if (mode == 1) {
        for (int k = 0; k < d; ++k) o[k] = (int32_t)((long long)mt_randbelow(s, Q) - (long long)HALF);
        return;
    }
    const int ncoef = weight < d ? (weight < 0 ? 0 : weight) : d;
    for (int k = 0; k < ncoef; ++k) {
        const int32_t mag = 1 + (int32_t)mt_randbelow(s, bound);
        const int32_t sg = 1 - 2 * (int32_t)mt_randbelow(s, 2u);
        o[k] = mag * sg;
    }
    for (int k = ncoef; k < d; ++k) o[k] = 0;
    if (ncoef < d) {
        for (int k = d - 1; k > 0; --k) {
            const int j = (int)mt_randbelow(s, (uint32_t)(k + 1));
            const int32_t t = o[k]; o[k] = o[j]; o[j] = t;
        }
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

// dst[b][r][:] = src[b][:] for r < rows  (quirk Q1: every row of a seeded key matrix is the same polynomial)
__global__ void __launch_bounds__(256) replicate_rows_kernel(const int4* __restrict__ src, int4* __restrict__ dst, size_t batch0, int rows, int d4,
                                                               size_t src_batch_stride4) {
    const int per = rows * d4;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per) return;
    const size_t b = batch0 + blockIdx.y;
    st_stream(dst + b * (size_t)per + i, __ldg(src + b * src_batch_stride4 + (i & (d4 - 1))));
}

int launch_sample(const fz_ctx* ctx, int mode, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, uint32_t bound, int weight,
                  int32_t* out, cudaStream_t st) {
    if (n == 0) return 0;
    sample_kernel<<<(int)((n + 63) / 64), 64, 0, st>>>(mode, keys, key_stride, klen, n, ctx->d, bound, weight, out);
    return check(cudaGetLastError(), "sample launch");
}

int launch_seeds_to_keys(const fz_ctx* ctx, const unsigned long long* seeds, size_t n, uint32_t* keys, int32_t* klen, cudaStream_t st) {
    if (n == 0) return 0;
    seeds_to_keys_kernel<<<(int)((n + 127) / 128), 128, 0, st>>>(seeds, n, keys, klen);
    (void)ctx;
    return check(cudaGetLastError(), "seeds_to_keys launch");
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
