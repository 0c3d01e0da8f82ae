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
#include "fz_mt.cuh"   // the generator, its seeding and the per-thread draw loops

namespace fz {

// mode 0: secret-key coefficients (1 + randrange(bound)) * (1 - 2*randrange(2)), `weight` of them, zero padded,
//         then the Fisher-Yates branch when weight < d (polynomials.py:449-459)
// mode 1: public-matrix values randrange(q) - q//2 (polynomials.py:480)
// keys: [n][key_stride] 32-bit limbs, klen[n] limbs used (>= 1)
__global__ void __launch_bounds__(SMP_THREADS) sample_kernel(int mode, const uint32_t* __restrict__ keys, int key_stride, const int32_t* __restrict__ klen,
                                                              size_t n, int d, uint32_t q, uint32_t bound, int weight, int32_t* __restrict__ out) {
    FZ_DYN_SHARED __align__(16) uint32_t smp_sm[];
    const size_t i = (size_t)blockIdx.x * SMP_STREAMS + threadIdx.x;
    if (threadIdx.x >= SMP_STREAMS || i >= n) return;   // no block-wide barrier below
    Mt s;
    s.m = smp_sm + threadIdx.x;
    uint32_t key[8];
    int kl = klen[i];
    if (kl > 8) kl = 8;
    for (int j = 0; j < 8; ++j) key[j] = j < kl ? keys[i * key_stride + j] : 0u;
    mt_seed(s, key, kl);
    sample_draws(s, mode, d, q, bound, weight, out + i * d);
}

// ---- dense secret-key sampler, second version (mode 0 with weight >= d, d <= 256, bound <= 127: the scheme's only use) ----
// tools/sample_lab.cu split the first kernel's time per warp (profiles/r02_sample_lab_v1.txt): seeding 40 K cycles, regeneration + tempering 47 K,
// the accept / reject machine 21 K, the per-lane scattered 4-byte global stores 33 K -- and one of the four sub-partitions of the SM idle (three warps).
// This kernel (i) spreads the CTA's generators over FOUR warps, one per sub-partition; (ii) keeps the (key[j] + j) terms of init_by_array in two registers
// when every seed of the warp has at most two limbs (any u64 seed), which takes the shared-memory load out of the dependent chain of the first
// recurrence, and double-buffers the state words of the second; (iii) skips the last tempering step (it only changes bits 13..0, a draw keeps
// bits 31..shb); (iv) STAGE = 1: emits each accepted coefficient as ONE byte (magnitude | sign << 7) into a shared-memory tile and writes the
// polynomial out afterwards with coalesced 128-byte stores, a warp per polynomial; STAGE = 0: collects four coefficients in registers per 16-byte store.
template <int STAGE>
__global__ void __launch_bounds__(SD_THREADS) sample_dense_kernel(const uint32_t* __restrict__ keys, int key_stride, const int32_t* __restrict__ klen, size_t n, int d,
                                                                  uint32_t bound, int32_t* __restrict__ out, long long* __restrict__ cyc) {
    using Cfg = SdCfg<STAGE>;
    constexpr int G = Cfg::G, PW = Cfg::PW;
    FZ_DYN_SHARED __align__(16) uint32_t smp_sm[];
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
        dense_draws<STAGE>(m, g, d, bound, stage, out + i * d);
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
    mt_build_seed_table(mt0);
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
        FZ_LAUNCH(sample_dense_kernel<0>, (unsigned)((n + SdCfg<0>::G - 1) / SdCfg<0>::G), SD_THREADS, SdCfg<0>::SMEM, st, keys, key_stride, klen, n, d, bound, out, nullptr);
    else
        FZ_LAUNCH(sample_dense_kernel<1>, (unsigned)((n + SdCfg<1>::G - 1) / SdCfg<1>::G), SD_THREADS, SdCfg<1>::SMEM, st, keys, key_stride, klen, n, d, bound, out, nullptr);
    return true;
}

int launch_sample(const fz_ctx* ctx, int mode, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, uint32_t bound, int weight,
                  int32_t* out, cudaStream_t st) {
    if (n == 0) return 0;
    if (launch_sample_dense(ctx->d, mode, keys, key_stride, klen, n, bound, weight, out, st)) return check(cudaGetLastError(), "dense sample launch");
    FZ_LAUNCH(sample_kernel, (unsigned)((n + SMP_STREAMS - 1) / SMP_STREAMS), SMP_THREADS, SMP_SMEM, st, mode, keys, key_stride, klen, n, ctx->d, Q, bound, weight, out);
    return check(cudaGetLastError(), "sample launch");
}

// the same samplers for any ring (fz_ring_sample_*): degree and modulus are runtime arguments of the kernel
int launch_sample_ring(int d, uint32_t q, int mode, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, uint32_t bound, int weight,
                       int32_t* out, cudaStream_t st) {
    if (n == 0) return 0;
    if (launch_sample_dense(d, mode, keys, key_stride, klen, n, bound, weight, out, st)) return check(cudaGetLastError(), "dense ring sample launch");
    FZ_LAUNCH(sample_kernel, (unsigned)((n + SMP_STREAMS - 1) / SMP_STREAMS), SMP_THREADS, SMP_SMEM, st, mode, keys, key_stride, klen, n, d, q, bound, weight, out);
    return check(cudaGetLastError(), "ring sample launch");
}

int launch_seeds_to_keys(const fz_ctx* ctx, const unsigned long long* seeds, size_t n, uint32_t* keys, int32_t* klen, cudaStream_t st) {
    if (n == 0) return 0;
    FZ_LAUNCH(seeds_to_keys_kernel, (int)((n + 127) / 128), 128, 0, st, seeds, n, keys, klen);
    (void)ctx;
    return check(cudaGetLastError(), "seeds_to_keys launch");
}

int launch_row_seeds_to_keys(const fz_ctx* ctx, const unsigned long long* seeds, size_t n, int rows, uint32_t* keys, int32_t* klen, cudaStream_t st) {
    if (n == 0) return 0;
    FZ_LAUNCH(row_seeds_to_keys_kernel, (unsigned)((2 * n * rows + 127) / 128), 128, 0, st, seeds, n, rows, keys, klen);
    (void)ctx;
    return check(cudaGetLastError(), "row_seeds_to_keys launch");
}

int launch_replicate_rows(const fz_ctx* ctx, const int32_t* src, int32_t* dst, size_t batch, int rows, size_t src_batch_stride, cudaStream_t st) {
    if (batch == 0) return 0;
    const int d4 = ctx->d / 4;
    const unsigned gx = (unsigned)((rows * d4 + 255) / 256);
    for (size_t b0 = 0; b0 < batch; b0 += 65535) {
        const size_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
        FZ_LAUNCH(replicate_rows_kernel, dim3(gx, (unsigned)nb), 256, 0, st, (const int4*)src, (int4*)dst, b0, rows, d4, src_batch_stride / 4);
        if (check(cudaGetLastError(), "replicate_rows launch")) return -2;
    }
    return 0;
}

}  // namespace fz
