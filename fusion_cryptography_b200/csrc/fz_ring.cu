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
// when d is larger), loads them with coalesced 4-byte accesses, runs the log2(d) butterfly stages with one barrier per
// stage and stores canonical centred residues.  Twiddles are Shoup pairs (w, floor(w * 2^32 / q)) built once per ring
// on the device; a*w mod q is one IMAD.HI and two IMADs for any 32-bit a because 2q < 2^32 (see fz_common.cuh:sh_mul).
// Results are the canonical centred residues cent() returns (algebra/ntt.py:93-123), whatever logmod the caller of
// the reference passes: for 0 <= y < q the shifted term is -1 up to q//2 and 0 above.
#include <cstdio>
#include <new>

#include "../../include/fusion_b200.h"
#include "fz_common.cuh"
#include "fz_internal.h"

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

struct RingK {
    uint32_t q, half;
    int logd;
    Sh ninv;
    unsigned long long mu;
};

__device__ __forceinline__ uint32_t rq_mul(uint32_t a, Sh w, uint32_t q) {   // a*w mod q in [0, q); a any 32-bit value
    uint32_t r = a * w.w - __umulhi(a, w.wq) * q;                            // [0, 2q)
    return min(r, r - q);
}
__device__ __forceinline__ uint32_t rq_add(uint32_t a, uint32_t b, uint32_t q) {   // a, b in [0, q)
    uint32_t s = a + b;
    return min(s, s - q);
}
__device__ __forceinline__ uint32_t rq_sub(uint32_t a, uint32_t b, uint32_t q) {
    uint32_t s = a - b + q;
    return min(s, s - q);
}
// any int32 -> [0, q)   (Python's %: the result takes the sign of the modulus)
__device__ __forceinline__ uint32_t rq_from(int32_t x, uint32_t q) {
    int32_t r = x % (int32_t)q;
    return (uint32_t)(r < 0 ? r + (int32_t)q : r);
}
// [0, q) -> canonical centred residue, algebra/ntt.py:120-123
__device__ __forceinline__ int32_t rq_cent(uint32_t y, uint32_t q, uint32_t half) { return y > half ? (int32_t)(y - q) : (int32_t)y; }

__global__ void ring_table_kernel(const uint32_t* __restrict__ table, Sh* out, int d, uint32_t q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    const uint32_t w = table[i] % q;
    out[i].w = w;
    out[i].wq = (uint32_t)(((unsigned long long)w << 32) / q);
}

// One tile = `polys` whole polynomials of d coefficients in shared memory.
// Forward (algebra/ntt.py:271-291): stages m = 1, 2, ..., d/2 with t = d/(2m); butterfly (u, v) -> (u + v*s, u - v*s), s = tw[m + i].
// Inverse (algebra/ntt.py:354-377): stages t = 1, 2, ..., d/2 with h = d/(2t); (u, v) -> (u + v, (u - v)*s), s = itw[h + i]; then * n^-1.
template <bool INVERSE>
__global__ void __launch_bounds__(RING_THREADS) ring_ntt_kernel(RingK k, const Sh* __restrict__ tw, const int32_t* in, int32_t* out, size_t n_polys,
                                                                 int polys_per_tile) {
    extern __shared__ __align__(16) uint32_t tile[];
    const uint32_t q = k.q;
    const int logd = k.logd, d = 1 << logd;
    const size_t n_tiles = (n_polys + polys_per_tile - 1) / polys_per_tile;
    for (size_t tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
        const size_t p0 = tl * polys_per_tile;
        const int live = (int)min((size_t)polys_per_tile, n_polys - p0);
        const int n_coef = live << logd;
        const int32_t* src = in + (p0 << logd);
        for (int e = threadIdx.x; e < n_coef; e += RING_THREADS) tile[e] = rq_from(src[e], q);
        __syncthreads();
        const int n_bf = n_coef >> 1;
        if (!INVERSE) {
            for (int lt = logd - 1; lt >= 0; --lt) {           // t = 2^lt, m = d / (2t)
                const int t = 1 << lt, m = d >> (lt + 1);
                for (int b = threadIdx.x; b < n_bf; b += RING_THREADS) {
                    const int p = b >> (logd - 1), kk = b & ((d >> 1) - 1);
                    const int i = kk >> lt, j = kk & (t - 1);
                    const int idx = (p << logd) + ((i << (lt + 1)) | j);
                    const Sh w = tw[m + i];
                    const uint32_t u = tile[idx], v = rq_mul(tile[idx + t], w, q);
                    tile[idx] = rq_add(u, v, q);
                    tile[idx + t] = rq_sub(u, v, q);
                }
                __syncthreads();
            }
            int32_t* dst = out + (p0 << logd);
            for (int e = threadIdx.x; e < n_coef; e += RING_THREADS) dst[e] = rq_cent(tile[e], q, k.half);
        } else {
            for (int lt = 0; lt < logd; ++lt) {                // t = 2^lt, h = d / (2t)
                const int t = 1 << lt, h = d >> (lt + 1);
                for (int b = threadIdx.x; b < n_bf; b += RING_THREADS) {
                    const int p = b >> (logd - 1), kk = b & ((d >> 1) - 1);
                    const int i = kk >> lt, j = kk & (t - 1);
                    const int idx = (p << logd) + ((i << (lt + 1)) | j);
                    const Sh w = tw[h + i];
                    const uint32_t u = tile[idx], v = tile[idx + t];
                    tile[idx] = rq_add(u, v, q);
                    tile[idx + t] = rq_mul(u - v + q, w, q);
                }
                __syncthreads();
            }
            int32_t* dst = out + (p0 << logd);
            for (int e = threadIdx.x; e < n_coef; e += RING_THREADS) dst[e] = rq_cent(rq_mul(tile[e], k.ninv, q), q, k.half);
        }
        __syncthreads();   // the tile is reused by the next iteration
    }
}

// cent(x * y) / cent(x + y), algebra/polynomials.py:341-385, :282-318 (and :114-148 for the coefficient representation's +)
template <bool MUL>
__global__ void __launch_bounds__(256) ring_pointwise_kernel(RingK k, const int32_t* __restrict__ a, const int32_t* __restrict__ b, int32_t* __restrict__ out,
                                                              size_t n) {
    const uint32_t q = k.q;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t x = rq_from(a[i], q), y = rq_from(b[i], q);
        uint32_t r;
        if (MUL) {
            const unsigned long long P = (unsigned long long)x * y;             // < 2^62
            const unsigned long long est = __umul64hi(P, k.mu);                 // floor(P / q) or one less
            r = (uint32_t)(P - est * q);                                        // [0, 2q), exact in 32 bits since 2q < 2^32
            r = min(r, r - q);
        } else {
            r = rq_add(x, y, q);
        }
        out[i] = rq_cent(r, q, k.half);
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
        ring_table_kernel<<<(r->d + 255) / 256, 256>>>(raw, *out, r->d, r->q);
        rc = check(cudaGetLastError(), "ring table kernel");
        if (!rc) rc = check(cudaDeviceSynchronize(), "ring table kernel");
    }
    cudaFree(raw);
    return rc;
}

template <bool INVERSE>
static int launch_ring_ntt(fz_ring* r, const int32_t* in, int32_t* out, size_t n_polys, cudaStream_t st) {
    if (n_polys == 0) return 0;
    const int polys_per_tile = r->d >= RING_TILE ? 1 : RING_TILE / r->d;
    const size_t smem = sizeof(uint32_t) * (size_t)(r->d >= RING_TILE ? r->d : RING_TILE);
    auto kern = ring_ntt_kernel<INVERSE>;
    if (smem > 48 * 1024 && check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ring NTT shared memory")) return -2;
    int per_sm = 0;
    if (check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, RING_THREADS, smem), "ring NTT occupancy")) return -2;
    const size_t n_tiles = (n_polys + polys_per_tile - 1) / polys_per_tile;
    const size_t cap = (size_t)r->sm_count * (per_sm > 0 ? per_sm : 1);
    kern<<<(unsigned)(n_tiles < cap ? n_tiles : cap), RING_THREADS, smem, st>>>(ring_k(r), INVERSE ? r->inv : r->fwd, in, out, n_polys, polys_per_tile);
    return check(cudaGetLastError(), "ring NTT launch");
}

template <bool MUL>
static int launch_ring_pointwise(fz_ring* r, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, cudaStream_t st) {
    const size_t n = n_polys << r->logd;
    if (n == 0) return 0;
    const size_t want = (n + 255) / 256, cap = (size_t)r->sm_count * 8;
    ring_pointwise_kernel<MUL><<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(ring_k(r), a, b, out, n);
    return check(cudaGetLastError(), "ring pointwise launch");
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

}  // extern "C"
