// fz_ring.cu -- the transforms and pointwise operations of algebra/ntt.py and algebra/polynomials.py for ANY ring
// Z_q[X]/(X^d + 1) the reference accepts on 32-bit words: q an odd prime below 2^31, d = 2^logd with 2 <= d <= 2^14,
// and caller-supplied bit-reversed power tables (the reference's cooley_tukey_ntt / gentleman_sande_intt take the
// table as an argument and never look at the root itself, algebra/ntt.py:216-291, :294-377).
//
// The scheme's two parameter sets have their own compile-time kernels (fz_ntt.cu: q, d and the long-stride twiddles are
// constants of the instruction stream).  This module is the general form of the same butterflies, so that every call the
// reference's algebra layer answers is answered here on the GPU as well -- the reference's own unit tests run on moduli
// below 2^17 and degrees 4..64 (tests/test_ntt.py:18-50), not on the scheme's ring.
//
// Layout: a CTA owns a TILE of whole polynomials in shared memory (as many as fit in 4,096 coefficients, one polynomial
// when d is larger), loads them with coalesced accesses and runs the log2(d) butterfly stages FOUR AT A TIME: in each pass a
// thread gathers the 16 coefficients of one radix-16 sub-transform into registers (indices that differ in four consecutive
// address bits), does its 32 butterflies there and scatters them back, so a degree-256 transform is two shared-memory round
// trips and two barriers instead of eight.  One pad word per 32 keeps both access shapes (stride >= 32 across the lanes of a
// warp, and 16 contiguous words per lane in the last pass) free of bank conflicts.  Twiddles are Shoup pairs
// (w, floor(w * 2^32 / q)) built once per ring on the device and read through L1; a*w mod q is one IMAD.HI and two IMADs for
// any 32-bit a because 2q < 2^32 (see fz_common.cuh:sh_mul).
// Results are the canonical centred residues cent() returns (algebra/ntt.py:93-123), whatever logmod the caller of
// the reference passes: for 0 <= y < q the shifted term is -1 up to q//2 and 0 above.
#include <cstdio>
#include <new>

#include "../../include/fusion_b200.h"
#include "fz_common.cuh"
#include "fz_internal.h"
#include "fz_ring_arith.cuh"

struct fz_ring {
    int device;
    uint32_t q, half;        // half = q // 2
    int logd, d;
    fz::Sh* fwd;             // device, d Shoup pairs (nullptr: no forward table was given)
    fz::Sh* inv;
    fz::Sh ninv;             // n^-1 = pow(d, q - 2, q), algebra/ntt.py:353
    unsigned long long mu;   // floor(2^64 / q): Barrett constant of the 62-bit pointwise products
    int sm_count;
};

namespace fz {

constexpr int RING_THREADS = 256;
constexpr int RING_TILE = 4096;   // coefficients per CTA tile when d <= 4096
constexpr int RING_MAX_LOGD = 14;

// RingK, rq_mul / rq_add / rq_sub / rq_from / rq_cent, ring_pw1: fz_ring_arith.cuh

__device__ __forceinline__ int phys(int e) { return e + (e >> 5); }   // one pad word per 32

__global__ void ring_table_kernel(const uint32_t* __restrict__ table, Sh* out, int d, uint32_t q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    const uint32_t w = table[i] % q;
    out[i].w = w;
    out[i].wq = (uint32_t)(((unsigned long long)w << 32) / q);
}

// One pass = R consecutive butterfly stages on the 2^R coefficients whose (polynomial-local) indices are
//     (hi << (lo_bits + R)) | (k << lo_bits) | lo,   k = 0 .. 2^R - 1,
// held in registers by one thread.
// Forward (algebra/ntt.py:271-291): stage with t = 2^lt pairs (j, j + t) and multiplies by tw[d / 2^(lt+1) + (j >> (lt + 1))]; the pass runs
// lt = lo_bits + R - 1 down to lo_bits, butterfly (u, v) -> (u + v*s, u - v*s).
// Inverse (algebra/ntt.py:354-372): lt = lo_bits up to lo_bits + R - 1, (u, v) -> (u + v, (u - v)*s), s = itw[d / 2^(lt+1) + (j >> (lt + 1))].
// The FIRST pass of a transform gathers its coefficients straight from global memory (any int32 -> [0, q)) and the LAST one scatters canonical
// centred residues straight back (times n^-1 for the inverse), so only the exchanges BETWEEN passes go through the shared-memory tile: a
// degree <= 16 transform never touches it, degree 256 crosses it once.  Across the lanes of a warp the global accesses are either consecutive
// 4-byte words (lo_bits > 0) or one 2^R-word run per lane (lo_bits == 0), moved as 16-byte pieces when the operands are aligned (VEC).
// The pass with lo_bits == 0 is the one whose twiddles differ from lane to lane (hi = the item index); fetched from global memory they cost a
// scattered 8-byte load per twiddle and item (the LSU data pipe was the busiest unit, 75 %).  When `twc` is not NULL it holds exactly those
// twiddles, copied once per CTA into shared memory as [slot][item of the tile]: consecutive lanes read consecutive 8-byte words.
template <int R, bool INVERSE, int GS, bool FROM_GLOBAL, bool TO_GLOBAL, bool VEC>
__device__ __forceinline__ void ring_pass(uint32_t* tile, const Sh* __restrict__ tw, const RingK& kk, int lo_bits, int n_items, int tid, const int32_t* gin,
                                          int32_t* gout, const Sh* twc = nullptr, int twc_items = 0) {
    constexpr int N = 1 << R;
    const uint32_t q = kk.q;
    const int logd = kk.logd;
    for (int w = tid; w < n_items; w += GS) {
        const int p = w >> (logd - R), wi = w & ((1 << (logd - R)) - 1);
        const int lo = wi & ((1 << lo_bits) - 1), hi = wi >> lo_bits;
        const int base = (p << logd) + (hi << (lo_bits + R)) + lo;
        uint32_t x[N];
        if (FROM_GLOBAL) {
            if (VEC && R >= 2 && lo_bits == 0) {
                const int4* g4 = reinterpret_cast<const int4*>(gin + base);
#pragma unroll
                for (int c = 0; c < N / 4; ++c) {
                    const int4 v = ld_stream(g4 + c);
                    x[4 * c] = rq_from(v.x, kk); x[4 * c + 1] = rq_from(v.y, kk); x[4 * c + 2] = rq_from(v.z, kk); x[4 * c + 3] = rq_from(v.w, kk);
                }
            } else {
                int32_t raw[N];
#pragma unroll
                for (int k = 0; k < N; ++k) raw[k] = gin[base + (k << lo_bits)];
#pragma unroll
                for (int k = 0; k < N; ++k) x[k] = rq_from(raw[k], kk);
            }
        } else {
#pragma unroll
            for (int k = 0; k < N; ++k) x[k] = tile[phys(base + (k << lo_bits))];
        }
#pragma unroll
        for (int s = 0; s < R; ++s) {
            if (!INVERSE) {
                const int lt = lo_bits + R - 1 - s, dk = N >> (s + 1);
                const Sh* row = tw + ((1 << logd) >> (lt + 1)) + (hi << s);
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    if (k & dk) continue;
                    const Sh tws = twc ? twc[(((1 << s) - 1) + (k >> (R - s))) * twc_items + w] : row[k >> (R - s)];
                    const uint32_t u = min(x[k], x[k] - q), v = rq_mul(x[k + dk], tws, q);   // lazy: x in [0, 2q) between the stages (2q < 2^32)
                    x[k] = u + v;
                    x[k + dk] = u - v + q;
                }
            } else {
                const int lt = lo_bits + s, dk = 1 << s;
                const Sh* row = tw + ((1 << logd) >> (lt + 1)) + (hi << (R - 1 - s));
#pragma unroll
                for (int k = 0; k < N; ++k) {
                    if (k & dk) continue;
                    const Sh tws = twc ? twc[((N - (N >> s)) + (k >> (s + 1))) * twc_items + w] : row[k >> (s + 1)];
                    const uint32_t u = x[k], v = x[k + dk];
                    x[k] = rq_add(u, v, q);
                    x[k + dk] = rq_mul(u - v + q, tws, q);
                }
            }
        }
        if (TO_GLOBAL) {
            int32_t o[N];
#pragma unroll
            for (int k = 0; k < N; ++k) o[k] = rq_cent(INVERSE ? rq_mul(x[k], kk.ninv, q) : min(x[k], x[k] - q), q, kk.half);   // forward values are lazy
            if (VEC && R >= 2 && lo_bits == 0) {
                int4* g4 = reinterpret_cast<int4*>(gout + base);
#pragma unroll
                for (int c = 0; c < N / 4; ++c) st_stream(g4 + c, make_int4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]));
            } else {
#pragma unroll
                for (int k = 0; k < N; ++k) gout[base + (k << lo_bits)] = o[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < N; ++k) tile[phys(base + (k << lo_bits))] = x[k];
        }
    }
}

// the pass shapes a transform can need: r stages with r = 1..4; only a 4-stage pass can be a first-but-not-last or a middle pass
template <bool INVERSE, int GS, bool VEC, bool FROM_GLOBAL, bool TO_GLOBAL>
__device__ __forceinline__ void ring_pass_r(int r, uint32_t* tile, const Sh* __restrict__ tw, const RingK& k, int lo_bits, int n_items, int tid,
                                            const int32_t* gin, int32_t* gout, const Sh* twc = nullptr, int twc_items = 0) {
    if (!TO_GLOBAL) {   // r == 4 by construction
        ring_pass<4, INVERSE, GS, FROM_GLOBAL, false, VEC>(tile, tw, k, lo_bits, n_items, tid, gin, gout, twc, twc_items);
        return;
    }
    switch (r) {
        case 4: ring_pass<4, INVERSE, GS, FROM_GLOBAL, true, VEC>(tile, tw, k, lo_bits, n_items, tid, gin, gout, twc, twc_items); break;
        case 3: ring_pass<3, INVERSE, GS, FROM_GLOBAL, true, VEC>(tile, tw, k, lo_bits, n_items, tid, gin, gout, twc, twc_items); break;
        case 2: ring_pass<2, INVERSE, GS, FROM_GLOBAL, true, VEC>(tile, tw, k, lo_bits, n_items, tid, gin, gout, twc, twc_items); break;
        default: ring_pass<1, INVERSE, GS, FROM_GLOBAL, true, VEC>(tile, tw, k, lo_bits, n_items, tid, gin, gout, twc, twc_items); break;
    }
}

// WARP = false: one tile per CTA (any degree), barriers between passes.  WARP = true (d <= 512): every warp owns a private tile of 512
// coefficients (512 / d whole polynomials: one radix-16 item per lane and pass) and only ever waits for itself (__syncwarp).
constexpr int RING_WARP_TILE = 512;

template <bool INVERSE, bool VEC, bool WARP>
__global__ void __launch_bounds__(RING_THREADS) ring_ntt_kernel(RingK k, const Sh* __restrict__ tw, const int32_t* in, int32_t* out, size_t n_polys,
                                                                 int polys_per_tile, int tile_words, int twc_items) {
    FZ_DYN_SHARED __align__(16) uint32_t ring_smem[];
    constexpr int GS = WARP ? 32 : RING_THREADS;
    const int tid = WARP ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
    uint32_t* tile = WARP ? ring_smem + (threadIdx.x >> 5) * tile_words : ring_smem;   // tile_words includes the pad word per 32
    auto sync = [] { if (WARP) __syncwarp(); else __syncthreads(); };
    const int logd = k.logd;
    const int n_pass = (logd + 3) >> 2, r_odd = logd - 4 * (n_pass - 1);   // n_pass - 1 four-stage passes and one of r_odd = 1..4 stages
    // twiddles of the lo_bits == 0 pass (forward: the last pass, r_odd stages; inverse: the first pass), [slot][item] behind the tiles
    const Sh* twc = nullptr;
    if (twc_items > 0) {
        Sh* fill = reinterpret_cast<Sh*>(ring_smem + (size_t)tile_words * (WARP ? RING_THREADS / 32 : 1));
        const int rc = INVERSE ? (logd < 4 ? logd : 4) : r_odd, nc = 1 << rc, per_poly = 1 << (logd - rc);
        for (int e = threadIdx.x; e < (nc - 1) * twc_items; e += RING_THREADS) {
            const int slot = e / twc_items, w = e - slot * twc_items, hi = w & (per_poly - 1);
            int s2 = 0, i2 = 0, idx;
            if (!INVERSE) {                      // slot = (2^s - 1) + i: stage s of the pass has lt = rc - 1 - s and 2^s twiddles per item
                while (((2 << s2) - 1) <= slot) ++s2;
                i2 = slot - ((1 << s2) - 1);
                idx = ((1 << logd) >> (rc - s2)) + (hi << s2) + i2;
            } else {                             // slot = (nc - (nc >> s)) + i: stage s has lt = s and nc >> (s + 1) twiddles per item
                while (nc - (nc >> (s2 + 1)) <= slot) ++s2;
                i2 = slot - (nc - (nc >> s2));
                idx = ((1 << logd) >> (s2 + 1)) + (hi << (rc - 1 - s2)) + i2;
            }
            fill[e] = tw[idx];
        }
        __syncthreads();
        twc = fill;
    }
    const size_t n_tiles = (n_polys + polys_per_tile - 1) / polys_per_tile;
    const size_t first = WARP ? (size_t)blockIdx.x * (RING_THREADS / 32) + (threadIdx.x >> 5) : (size_t)blockIdx.x;
    const size_t step = WARP ? (size_t)gridDim.x * (RING_THREADS / 32) : (size_t)gridDim.x;
    for (size_t tl = first; tl < n_tiles; tl += step) {
        const size_t p0 = tl * polys_per_tile;
        const int live = (int)min((size_t)polys_per_tile, n_polys - p0);
        const int32_t* gin = in + (p0 << logd);
        int32_t* gout = out + (p0 << logd);
        if (n_pass == 1) {
            ring_pass_r<INVERSE, GS, VEC, true, true>(logd, tile, tw, k, 0, live, tid, gin, gout, twc, twc_items);
            continue;
        }
        // forward: stages lt = logd - 1 downwards, the odd-sized pass last (lo_bits = 0); inverse: lt = 0 upwards, the odd-sized pass last (on top)
        int a = INVERSE ? 0 : logd - 4;
        ring_pass_r<INVERSE, GS, VEC, true, false>(4, tile, tw, k, a, live << (logd - 4), tid, gin, gout, INVERSE ? twc : nullptr, twc_items);
        sync();
        for (int ps = 1; ps < n_pass - 1; ++ps) {
            a += INVERSE ? 4 : -4;
            ring_pass_r<INVERSE, GS, VEC, false, false>(4, tile, tw, k, a, live << (logd - 4), tid, gin, gout);
            sync();
        }
        ring_pass_r<INVERSE, GS, VEC, false, true>(r_odd, tile, tw, k, INVERSE ? a + 4 : 0, live << (logd - r_odd), tid, gin, gout, INVERSE ? nullptr : twc,
                                                   twc_items);
        sync();   // the tile is reused by the next iteration
    }
}

// n4 int4 groups (16-byte aligned operands), then the scalar kernel below finishes the tail / unaligned calls
template <bool MUL>
__global__ void __launch_bounds__(256) ring_pointwise4_kernel(RingK k, const int4* __restrict__ a, const int4* __restrict__ b, int4* __restrict__ out, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const int4 va = ld_stream(a + i), vb = ld_stream(b + i);
        st_stream(out + i, make_int4(ring_pw1<MUL>(k, va.x, vb.x), ring_pw1<MUL>(k, va.y, vb.y), ring_pw1<MUL>(k, va.z, vb.z), ring_pw1<MUL>(k, va.w, vb.w)));
    }
}

template <bool MUL>
__global__ void __launch_bounds__(256) ring_pointwise_kernel(RingK k, const int32_t* a, const int32_t* b, int32_t* out, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = ring_pw1<MUL>(k, a[i], b[i]);
}

// fusion/fusion.py:422-481 (decode_bytes_to_polynomial_coefficients) for any degree: one thread per byte string, coefficients in place in the
// output.  Signum i is bit i (LSB first) of the big-endian integer of the first ceil(weight / 8) bytes (:446-453); magnitudes are
// (int(bpc bytes, big-endian) % bound) + 1 (:458-465); then the partial Fisher-Yates for i = d - 1 down to weight + 1 with
// j = int(bpi bytes) % (i + 1) (:472-480).  Slices past the end of the byte string are shorter or empty, as Python's are.
__global__ void __launch_bounds__(64) ring_decode_kernel(const uint8_t* __restrict__ bytes, size_t stride, uint32_t blen, size_t n, int d, int weight,
                                                          uint32_t bound, int bpc, int bpi, int32_t* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* b = bytes + i * stride;
    int32_t* o = out + i * d;
    const uint32_t bfs = (uint32_t)(weight + 7) / 8;
    uint32_t pos = bfs;
    for (int k = 0; k < weight; ++k) {
        unsigned long long acc = 0;
        for (int t = 0; t < bpc; ++t, ++pos)
            if (pos < blen) acc = ((acc << 8) | b[pos]) % bound;
        const int bit = (b[bfs - 1 - (uint32_t)(k >> 3)] >> (k & 7)) & 1;
        o[k] = (int32_t)(acc + 1) * (2 * bit - 1);
    }
    for (int k = weight; k < d; ++k) o[k] = 0;
    const int num = weight < 1 ? 1 : weight;   // max(1, min(degree, weight_bound)), weight <= d
    if (num < d)
        for (int k = d - 1; k > weight; --k) {
            unsigned long long acc = 0;
            const uint32_t m = (uint32_t)k + 1;
            for (int t = 0; t < bpi; ++t, ++pos)
                if (pos < blen) acc = ((acc << 8) | b[pos]) % m;
            const int32_t tmp = o[k]; o[k] = o[acc]; o[acc] = tmp;
        }
}

static uint32_t host_powmod(uint32_t b, uint32_t e, uint32_t m) {
    unsigned long long r = 1, x = b % m;
    while (e) {
        if (e & 1) r = r * x % m;
        x = x * x % m;
        e >>= 1;
    }
    return (uint32_t)r;
}

static RingK ring_k(const fz_ring* r) {
    RingK k;
    k.q = r->q; k.half = r->half; k.logd = r->logd; k.ninv = r->ninv; k.mu = r->mu;
    k.mu32 = (uint32_t)((1ULL << 32) / r->q);
    k.c31 = (uint32_t)((1ULL << 31) % r->q);
    return k;
}

static int upload_table(fz_ring* r, const uint32_t* host_table, Sh** out) {
    uint32_t* raw = nullptr;
    if (check(cudaMalloc(&raw, sizeof(uint32_t) * r->d), "ring table staging") || check(cudaMalloc(out, sizeof(Sh) * r->d), "ring table")) {
        if (raw) cudaFree(raw);
        return -2;
    }
    int rc = check(cudaMemcpy(raw, host_table, sizeof(uint32_t) * r->d, cudaMemcpyHostToDevice), "ring table upload");
    if (!rc) {
        FZ_LAUNCH(ring_table_kernel, (r->d + 255) / 256, 256, 0, 0, raw, *out, r->d, r->q);
        rc = check(cudaGetLastError(), "ring table kernel");
        if (!rc) rc = check(cudaDeviceSynchronize(), "ring table kernel");
    }
    cudaFree(raw);
    return rc;
}

template <bool INVERSE, bool VEC, bool WARP>
static int launch_ring_ntt_v(fz_ring* r, const int32_t* in, int32_t* out, size_t n_polys, cudaStream_t st) {
    const size_t tile_coefs = WARP ? (size_t)RING_WARP_TILE : (size_t)(r->d >= RING_TILE ? r->d : RING_TILE);
    const int polys_per_tile = (size_t)r->d >= tile_coefs ? 1 : (int)(tile_coefs / r->d);
    const size_t tile_words = tile_coefs + tile_coefs / 32;   // one pad word per 32
    // twiddle cache of the lo_bits == 0 pass: (2^rc - 1) Shoup pairs per item of a tile, kept only while it stays small (d <= 4,096)
    const int n_pass = (r->logd + 3) / 4, r_odd = r->logd - 4 * (n_pass - 1), rc = INVERSE ? (r->logd < 4 ? r->logd : 4) : r_odd;
    int twc_items = (int)(tile_coefs >> rc);
    size_t twc_bytes = sizeof(Sh) * (size_t)((1 << rc) - 1) * twc_items;
    if (twc_bytes > 32 * 1024) twc_items = 0, twc_bytes = 0;
    const size_t smem = sizeof(uint32_t) * tile_words * (WARP ? RING_THREADS / 32 : 1) + twc_bytes;
    auto kern = ring_ntt_kernel<INVERSE, VEC, WARP>;
    if (smem > 48 * 1024 && check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ring NTT shared memory")) return -2;
    int per_sm = 0;
    if (check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RING_THREADS, smem), "ring NTT occupancy")) return -2;
    const size_t n_tiles = (n_polys + polys_per_tile - 1) / polys_per_tile;
    const size_t want = WARP ? (n_tiles + RING_THREADS / 32 - 1) / (RING_THREADS / 32) : n_tiles;
    const size_t cap = (size_t)r->sm_count * (per_sm > 0 ? per_sm : 1);
    FZ_LAUNCH(kern, (unsigned)(want < cap ? want : cap), RING_THREADS, smem, st, ring_k(r), INVERSE ? r->inv : r->fwd, in, out, n_polys, polys_per_tile,
                                                                          (int)tile_words, twc_items);
    return check(cudaGetLastError(), "ring NTT launch");
}

template <bool INVERSE>
static int launch_ring_ntt(fz_ring* r, const int32_t* in, int32_t* out, size_t n_polys, cudaStream_t st) {
    if (n_polys == 0) return 0;
    // every tile starts on a multiple of 16 bytes when the operands do (a tile is 512 or >= 4,096 coefficients)
    const bool vec = (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
    if (r->d <= RING_WARP_TILE)
        return vec ? launch_ring_ntt_v<INVERSE, true, true>(r, in, out, n_polys, st) : launch_ring_ntt_v<INVERSE, false, true>(r, in, out, n_polys, st);
    return vec ? launch_ring_ntt_v<INVERSE, true, false>(r, in, out, n_polys, st) : launch_ring_ntt_v<INVERSE, false, false>(r, in, out, n_polys, st);
}

template <bool MUL>
static int launch_ring_pointwise(fz_ring* r, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, cudaStream_t st) {
    const size_t n = n_polys << r->logd;
    if (n == 0) return 0;
    const size_t cap = (size_t)r->sm_count * 8;
    size_t done = 0;
    if ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0 && n >= 4) {
        const size_t n4 = n / 4, want = (n4 + 255) / 256;
        FZ_LAUNCH(ring_pointwise4_kernel<MUL>, (unsigned)(want < cap ? want : cap), 256, 0, st, ring_k(r), (const int4*)a, (const int4*)b, (int4*)out, n4);
        if (check(cudaGetLastError(), "ring pointwise launch")) return -2;
        done = n4 * 4;
    }
    if (done < n) {
        const size_t rest = n - done, want = (rest + 255) / 256;
        FZ_LAUNCH(ring_pointwise_kernel<MUL>, (unsigned)(want < cap ? want : cap), 256, 0, st, ring_k(r), a + done, b + done, out + done, rest);
        if (check(cudaGetLastError(), "ring pointwise launch")) return -2;
    }
    return 0;
}

}  // namespace fz

using namespace fz;

#define FZR_REQUIRE(cond, msg)           \
    do {                                 \
        if (!(cond)) {                   \
            fz::set_error("%s", msg);    \
            return -1;                   \
        }                                \
    } while (0)

extern "C" {

int fz_ring_create(int device, uint32_t q, int logd, const uint32_t* fwd_table, const uint32_t* inv_table, fz_ring** out) {
    FZR_REQUIRE(out != nullptr, "fz_ring_create: out is NULL");
    FZR_REQUIRE(q >= 3 && (q & 1) && q < (1u << 31), "fz_ring_create: the modulus must be odd, at least 3 and below 2^31");
    FZR_REQUIRE(logd >= 1 && logd <= RING_MAX_LOGD, "fz_ring_create: the degree must be a power of two between 2 and 16384");
    int ndev = 0;
    if (check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount")) return -2;
    FZR_REQUIRE(device >= 0 && device < ndev, "fz_ring_create: no such CUDA device (libfusion_b200 has no CPU fallback)");
    if (check(cudaSetDevice(device), "cudaSetDevice")) return -2;
    fz_ring* r = new (std::nothrow) fz_ring();
    FZR_REQUIRE(r != nullptr, "fz_ring_create: out of host memory");
    r->device = device; r->q = q; r->half = q / 2; r->logd = logd; r->d = 1 << logd;
    r->fwd = r->inv = nullptr;
    const uint32_t ninv = host_powmod((uint32_t)r->d % q, q - 2, q);
    r->ninv.w = ninv;
    r->ninv.wq = (uint32_t)(((unsigned long long)ninv << 32) / q);
    r->mu = (unsigned long long)((((unsigned __int128)1) << 64) / q);
    cudaDeviceProp prop;
    int rc = check(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties");
    r->sm_count = rc ? 1 : prop.multiProcessorCount;
    if (!rc) rc = sample_init_tables(nullptr);   // MT19937 seed table + the sampler's shared-memory opt-in (idempotent)
    if (!rc && fwd_table) rc = upload_table(r, fwd_table, &r->fwd);
    if (!rc && inv_table) rc = upload_table(r, inv_table, &r->inv);
    if (rc) {
        fz_ring_destroy(r);
        return rc;
    }
    *out = r;
    return 0;
}

int fz_ring_destroy(fz_ring* r) {
    if (!r) return 0;
    cudaSetDevice(r->device);
    if (r->fwd) cudaFree(r->fwd);
    if (r->inv) cudaFree(r->inv);
    delete r;
    return 0;
}

int fz_ring_ntt_fwd(fz_ring* r, const int32_t* in, int32_t* out, size_t n_polys, void* stream) {
    FZR_REQUIRE(r && (n_polys == 0 || (in && out)), "fz_ring_ntt_fwd: NULL argument");
    FZR_REQUIRE(r->fwd, "fz_ring_ntt_fwd: the ring was created without a forward table");
    return launch_ring_ntt<false>(r, in, out, n_polys, (cudaStream_t)stream);
}

int fz_ring_ntt_inv(fz_ring* r, const int32_t* in, int32_t* out, size_t n_polys, void* stream) {
    FZR_REQUIRE(r && (n_polys == 0 || (in && out)), "fz_ring_ntt_inv: NULL argument");
    FZR_REQUIRE(r->inv, "fz_ring_ntt_inv: the ring was created without an inverse table");
    return launch_ring_ntt<true>(r, in, out, n_polys, (cudaStream_t)stream);
}

int fz_ring_pointwise_mul(fz_ring* r, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, void* stream) {
    FZR_REQUIRE(r && (n_polys == 0 || (a && b && out)), "fz_ring_pointwise_mul: NULL argument");
    return launch_ring_pointwise<true>(r, a, b, out, n_polys, (cudaStream_t)stream);
}

int fz_ring_pointwise_add(fz_ring* r, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, void* stream) {
    FZR_REQUIRE(r && (n_polys == 0 || (a && b && out)), "fz_ring_pointwise_add: NULL argument");
    return launch_ring_pointwise<false>(r, a, b, out, n_polys, (cudaStream_t)stream);
}

// algebra/polynomials.py:436-467 (mode of fz_sample_coef) and :470-488 (fz_sample_ntt) for this ring: keys are the 32-bit limbs of
// abs(seed) as CPython's random.seed takes them ([n][key_stride], klen[n] limbs used, at most 8)
int fz_ring_sample_coef(fz_ring* r, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, int norm_bound, int weight_bound, int32_t* out,
                        void* stream) {
    FZR_REQUIRE(r && (n == 0 || (keys && klen && out)) && key_stride >= 1, "fz_ring_sample_coef: bad argument");
    const long long bound = norm_bound < (long long)r->half ? norm_bound : (long long)r->half;   // polynomials.py:450
    FZR_REQUIRE(bound >= 1, "fz_ring_sample_coef: norm_bound must be >= 1 (randrange(0) raises in the reference)");
    const int weight = weight_bound < 0 ? 0 : weight_bound > r->d ? r->d : weight_bound;          // polynomials.py:449
    return launch_sample_ring(r->d, r->q, 0, keys, key_stride, klen, n, (uint32_t)bound, weight, out, (cudaStream_t)stream);
}

int fz_ring_sample_ntt(fz_ring* r, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, int32_t* out, void* stream) {
    FZR_REQUIRE(r && (n == 0 || (keys && klen && out)) && key_stride >= 1, "fz_ring_sample_ntt: bad argument");
    return launch_sample_ring(r->d, r->q, 1, keys, key_stride, klen, n, 0, 0, out, (cudaStream_t)stream);
}

// fusion/fusion.py:422-481 for this ring's degree and modulus.  bytes: [n][stride] device bytes of which the first blen of each row are decoded;
// bytes_per_coefficient / bytes_per_index are the caller's evaluation of :433-434 (they involve a floating-point log2 in the reference).
// Returns -3 with the reference's message when blen is below the length check of :435-442.
int fz_ring_decode_poly(fz_ring* r, const uint8_t* bytes, size_t stride, size_t blen, size_t n, long long norm_bound, int weight_bound,
                        int bytes_per_coefficient, int bytes_per_index, int32_t* out, void* stream) {
    FZR_REQUIRE(r && (n == 0 || (bytes && out)), "fz_ring_decode_poly: NULL argument");
    FZR_REQUIRE(weight_bound >= 0 && weight_bound <= r->d, "fz_ring_decode_poly: weight_bound must be in [0, degree]");
    FZR_REQUIRE(bytes_per_coefficient >= 1 && bytes_per_index >= 1 && blen < (1ull << 32), "fz_ring_decode_poly: bad byte counts");
    long long bound = norm_bound < (long long)r->half ? norm_bound : (long long)r->half;   // :432
    if (bound < 1) bound = 1;
    const size_t total = (size_t)((weight_bound + 7) / 8) + (size_t)(bytes_per_coefficient + bytes_per_index) * weight_bound;   // :435-438
    if (blen < total) {
        set_error("Too few bytes to decode polynomial. Expected %zu but got %zu", total, blen);
        return -3;
    }
    if (n == 0) return 0;
    FZ_LAUNCH(ring_decode_kernel, (unsigned)((n + 63) / 64), 64, 0, (cudaStream_t)stream, bytes, stride, (uint32_t)blen, n, r->d, weight_bound, (uint32_t)bound,
                                                                                   bytes_per_coefficient, bytes_per_index, out);
    return check(cudaGetLastError(), "ring decode launch");
}

}  // extern "C"
