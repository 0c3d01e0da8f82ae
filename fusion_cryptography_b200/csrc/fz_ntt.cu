// fz_ntt.cu -- batched negacyclic NTT / INTT / pointwise product over Z_q for d = 64 and 256.
//
// Replaces (bit-exactly) algebra/ntt.py:216-291 (cooley_tukey_ntt: natural in, bit-reversed
// out), algebra/ntt.py:294-377 (gentleman_sande_intt: bit-reversed in, natural out, x n^-1),
// algebra/polynomials.py:282-318/341-385 (pointwise add / mul with cent) and
// algebra/polynomials.py:391-433 (transform).
//
// Decomposition (B200: no tensor cores -- modular integer butterflies; the aim is to keep the
// INT32 pipes busy while every coefficient crosses HBM exactly once in and once out):
//   * a polynomial is owned by TPP = d/16 lanes of ONE warp; each lane keeps 16 coefficients
//     in registers, so 4 butterfly stages run with no data movement at all;
//   * the "strided" pass (lane holds e = lane + TPP*k) covers the 4 long-stride stages; its
//     15 twiddles are the same for every lane, so they are compile-time constant-bank
//     operands of the IMADs (no loads, no registers);
//   * one padded shared-memory transpose (warp-private, __syncwarp only) turns the strided
//     ownership into 16 contiguous coefficients per lane for the remaining log2(d)-4 stages,
//     whose lane-dependent twiddles are fetched once per thread and stay in registers while
//     the persistent CTA loops over polynomials;
//   * butterflies multiply by precomputed Shoup pairs (IMAD.HI + 2 IMAD, result in [0, 2q), one
//     fused add-min to [0, q)); measured on B200 the FMA-heavy pipe issues IMAD every 2 and
//     IMAD.HI every 4 cycles per sub-partition, which makes this cheaper than Plantard or
//     Montgomery forms.  The last inverse stage, which also carries n^-1, uses the signed
//     Plantard product: it yields the canonical centred residue with no fix-up at all;
//   * the next tile's coefficients are fetched into registers before the current tile's
//     butterflies, so global-load latency is covered by arithmetic.
#include "fz_common.cuh"
#include "fz_internal.h"
#include "fz_ntt_core.cuh"   // tables, butterflies, stage functions, table construction

namespace fz {

// strided ownership -> contiguous ownership through the warp-private padded tile
template <int LOGD>
__device__ __forceinline__ void xchg_s2c(uint32_t* tile, int tau, const uint32_t (&x)[16], uint32_t (&y)[16]) {
    constexpr int TPP = Cfg<LOGD>::TPP;
#pragma unroll
    for (int k = 0; k < 16; ++k) tile[row_off(tau + TPP * k)] = x[k];
    __syncwarp();
    const uint4* r = reinterpret_cast<const uint4*>(tile + tau * 20);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint4 v = r[c];
        y[4 * c] = v.x; y[4 * c + 1] = v.y; y[4 * c + 2] = v.z; y[4 * c + 3] = v.w;
    }
    __syncwarp();
}

template <int LOGD>
__device__ __forceinline__ void xchg_c2s(uint32_t* tile, int tau, const uint32_t (&y)[16], uint32_t (&x)[16]) {
    constexpr int TPP = Cfg<LOGD>::TPP;
    uint4* w = reinterpret_cast<uint4*>(tile + tau * 20);
#pragma unroll
    for (int c = 0; c < 4; ++c) w[c] = make_uint4(y[4 * c], y[4 * c + 1], y[4 * c + 2], y[4 * c + 3]);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = tile[row_off(tau + TPP * k)];
    __syncwarp();
}

constexpr int NTT_THREADS = 128;
#ifndef NTT_MIN_BLOCKS
#define NTT_MIN_BLOCKS 1
#endif

template <int LOGD>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) ntt_fwd_kernel(const int32_t* in, int32_t* out, size_t n_polys) {
    using C = Cfg<LOGD>;
    __shared__ __align__(16) uint32_t smem[(NTT_THREADS / 32) * C::WARP_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tau = lane % C::TPP, pw = lane / C::TPP;
    uint32_t* tile = smem + warp * C::WARP_WORDS + C::base(pw);
    Sh tw2[15];
    load_tw2<LOGD>(C::fwd(), tau, tw2);
    const size_t per_cta = (NTT_THREADS / 32) * C::PPW, stride = (size_t)gridDim.x * per_cta;
    size_t p = (size_t)blockIdx.x * per_cta + warp * C::PPW + pw;
    int32_t raw[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) raw[k] = p < n_polys ? in[p * C::D + tau + C::TPP * k] : 0;
    for (size_t p0 = (size_t)blockIdx.x * per_cta; p0 < n_polys; p0 += stride, p += stride) {
        const bool live = p < n_polys;
        uint32_t x[16], y[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = to_lazy(raw[k]);
        {   // prefetch the next tile while this one is transformed
            const size_t pn = p + stride;
            const int32_t* src = in + pn * C::D + tau;
#pragma unroll
            for (int k = 0; k < 16; ++k) raw[k] = pn < n_polys ? src[C::TPP * k] : 0;
        }
        fwd_strided<LOGD>(x);
        xchg_s2c<LOGD>(tile, tau, x, y);
        fwd_contig<LOGD>(y, tw2);
        if (live) {
            int4* dst = reinterpret_cast<int4*>(out + p * C::D + tau * 16);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                dst[c] = make_int4(cent_u(y[4 * c]), cent_u(y[4 * c + 1]), cent_u(y[4 * c + 2]), cent_u(y[4 * c + 3]));
        }
    }
}

template <int LOGD>
__global__ void __launch_bounds__(NTT_THREADS, NTT_MIN_BLOCKS) ntt_inv_kernel(const int32_t* in, int32_t* out, size_t n_polys) {
    using C = Cfg<LOGD>;
    __shared__ __align__(16) uint32_t smem[(NTT_THREADS / 32) * C::WARP_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tau = lane % C::TPP, pw = lane / C::TPP;
    uint32_t* tile = smem + warp * C::WARP_WORDS + C::base(pw);
    Sh tw2[15];
    load_tw2<LOGD>(C::inv(), tau, tw2);
    const size_t per_cta = (NTT_THREADS / 32) * C::PPW, stride = (size_t)gridDim.x * per_cta;
    size_t p = (size_t)blockIdx.x * per_cta + warp * C::PPW + pw;
    int4 raw[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) raw[c] = p < n_polys ? reinterpret_cast<const int4*>(in + p * C::D + tau * 16)[c] : make_int4(0, 0, 0, 0);
    for (size_t p0 = (size_t)blockIdx.x * per_cta; p0 < n_polys; p0 += stride, p += stride) {
        const bool live = p < n_polys;
        uint32_t x[16], y[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            y[4 * c] = to_red(raw[c].x); y[4 * c + 1] = to_red(raw[c].y); y[4 * c + 2] = to_red(raw[c].z); y[4 * c + 3] = to_red(raw[c].w);
        }
        {
            const size_t pn = p + stride;
            const int4* src = reinterpret_cast<const int4*>(in + pn * C::D + tau * 16);
#pragma unroll
            for (int c = 0; c < 4; ++c) raw[c] = pn < n_polys ? src[c] : make_int4(0, 0, 0, 0);
        }
        inv_contig<LOGD>(y, tw2);
        xchg_c2s<LOGD>(tile, tau, y, x);
        int32_t o[16];
        inv_strided<LOGD>(x, o);
        if (live) {
            int32_t* dst = out + p * C::D + tau;
#pragma unroll
            for (int k = 0; k < 16; ++k) dst[C::TPP * k] = o[k];
        }
    }
}

// out = INTT(NTT(a) o NTT(b)): algebra/ntt.py:450-471 (ntt_poly_mult) in one pass over HBM
template <int LOGD>
__global__ void __launch_bounds__(NTT_THREADS) polymul_kernel(const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys) {
    using C = Cfg<LOGD>;
    __shared__ __align__(16) uint32_t smem[(NTT_THREADS / 32) * C::WARP_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tau = lane % C::TPP, pw = lane / C::TPP;
    uint32_t* tile = smem + warp * C::WARP_WORDS + C::base(pw);
    // the two sets of lane-dependent twiddles live in shared memory, one row per lane-in-polynomial (as in ntt_pass_kernel): 115 -> fewer registers
    __shared__ __align__(8) Sh tw_sm[2][C::TPP][15];
    if (threadIdx.x < C::TPP) {
        Sh t[15];
        load_tw2<LOGD>(C::fwd(), (int)threadIdx.x, t);
#pragma unroll
        for (int i = 0; i < 15; ++i) tw_sm[0][threadIdx.x][i] = t[i];
        load_tw2<LOGD>(C::inv(), (int)threadIdx.x, t);
#pragma unroll
        for (int i = 0; i < 15; ++i) tw_sm[1][threadIdx.x][i] = t[i];
    }
    __syncthreads();
    const size_t per_cta = (NTT_THREADS / 32) * C::PPW;
    for (size_t p0 = (size_t)blockIdx.x * per_cta; p0 < n_polys; p0 += (size_t)gridDim.x * per_cta) {
        const size_t p = p0 + warp * C::PPW + pw;
        const bool live = p < n_polys;
        uint32_t x[16], ya[16], yb[16];
        const int32_t* sa = a + p * C::D + tau;
        const int32_t* sb = b + p * C::D + tau;
        int32_t rb[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = live ? to_lazy(sa[C::TPP * k]) : 0u;
#pragma unroll
        for (int k = 0; k < 16; ++k) rb[k] = live ? sb[C::TPP * k] : 0;   // the second operand is in flight while the first is transformed
        fwd_strided<LOGD>(x);
        xchg_s2c<LOGD>(tile, tau, x, ya);
        {
            Sh tw[15];
#pragma unroll
            for (int i = 0; i < 15; ++i) tw[i] = tw_sm[0][tau][i];
            fwd_contig<LOGD>(ya, tw);
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = to_lazy(rb[k]);
            fwd_strided<LOGD>(x);
            xchg_s2c<LOGD>(tile, tau, x, yb);
            fwd_contig<LOGD>(yb, tw);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) ya[k] = mont_mul(red_q(ya[k]), (int32_t)red_q(yb[k]));   // [0, q] fits int32; the 2^-32 is repaid by c_last[2..3]
        {
            Sh tw[15];
#pragma unroll
            for (int i = 0; i < 15; ++i) tw[i] = tw_sm[1][tau][i];
            inv_contig<LOGD>(ya, tw);
        }
        xchg_c2s<LOGD>(tile, tau, ya, x);
        int32_t o[16];
        inv_strided<LOGD, 2>(x, o);
        if (live) {
            int32_t* dst = out + p * C::D + tau;
#pragma unroll
            for (int k = 0; k < 16; ++k) dst[C::TPP * k] = o[k];
        }
    }
}

// out = INTT(NTT(a) o b_hat): transform(transform(a) * b_hat) of algebra/polynomials.py:391-433 / :341-385 in ONE pass over HBM --
// a in the coefficient domain, b_hat already in the NTT domain (a public matrix row, a challenge): 3 KiB of traffic per polynomial
// instead of the 7 KiB of forward transform + pointwise product + inverse transform as three kernels.
template <int LOGD>
__global__ void __launch_bounds__(NTT_THREADS) ntt_pass_kernel(const int32_t* a, const int32_t* __restrict__ bhat, int32_t* out,   // a == out is allowed
                                                               
                                                                size_t n_polys) {
    using C = Cfg<LOGD>;
    __shared__ __align__(16) uint32_t smem[(NTT_THREADS / 32) * C::WARP_WORDS];
    // The lane-dependent twiddles of the two contiguous passes (15 Shoup pairs each) would occupy 60 registers if both sets stayed
    // resident; they live in shared memory instead, one row per lane-in-polynomial (15 x 8 bytes: the TPP rows fall on distinct
    // banks), and each set is fetched right before its pass.
    __shared__ __align__(8) Sh tw_sm[2][C::TPP][15];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tau = lane % C::TPP, pw = lane / C::TPP;
    uint32_t* tile = smem + warp * C::WARP_WORDS + C::base(pw);
    if (threadIdx.x < C::TPP) {
        Sh t[15];
        load_tw2<LOGD>(C::fwd(), (int)threadIdx.x, t);
#pragma unroll
        for (int i = 0; i < 15; ++i) tw_sm[0][threadIdx.x][i] = t[i];
        load_tw2<LOGD>(C::inv(), (int)threadIdx.x, t);
#pragma unroll
        for (int i = 0; i < 15; ++i) tw_sm[1][threadIdx.x][i] = t[i];
    }
    __syncthreads();
    const size_t per_cta = (NTT_THREADS / 32) * C::PPW, stride = (size_t)gridDim.x * per_cta;
    size_t p = (size_t)blockIdx.x * per_cta + warp * C::PPW + pw;
    int32_t raw[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) raw[k] = p < n_polys ? a[p * C::D + tau + C::TPP * k] : 0;
    for (size_t p0 = (size_t)blockIdx.x * per_cta; p0 < n_polys; p0 += stride, p += stride) {
        const bool live = p < n_polys;
        uint32_t x[16], y[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = to_lazy(raw[k]);
        {   // prefetch the next tile's coefficients while this one is transformed
            const size_t pn = p + stride;
            const int32_t* src = a + pn * C::D + tau;
#pragma unroll
            for (int k = 0; k < 16; ++k) raw[k] = pn < n_polys ? src[C::TPP * k] : 0;
        }
        int4 bv[4];   // the 16 NTT-domain coefficients this lane owns after the transpose: issued early, used after the forward transform
        {
            const int4* sb = reinterpret_cast<const int4*>(bhat + p * C::D + tau * 16);
#pragma unroll
            for (int c = 0; c < 4; ++c) bv[c] = live ? __ldg(sb + c) : make_int4(0, 0, 0, 0);   // L1-allocating: the four 16-byte pieces of a lane share two sectors (a no-allocate load fetched them twice from DRAM: 2.89 GB read per launch against 2.15 GB algorithmic)
        }
        fwd_strided<LOGD>(x);
        xchg_s2c<LOGD>(tile, tau, x, y);
        {
            Sh tw[15];
#pragma unroll
            for (int i = 0; i < 15; ++i) tw[i] = tw_sm[0][tau][i];
            fwd_contig<LOGD>(y, tw);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {   // Montgomery products: the factor 2^-32 is paid back by the constants of the last inverse stage
            y[4 * c] = mont_mul(red_q(y[4 * c]), bv[c].x); y[4 * c + 1] = mont_mul(red_q(y[4 * c + 1]), bv[c].y);
            y[4 * c + 2] = mont_mul(red_q(y[4 * c + 2]), bv[c].z); y[4 * c + 3] = mont_mul(red_q(y[4 * c + 3]), bv[c].w);
        }
        {
            Sh tw[15];
#pragma unroll
            for (int i = 0; i < 15; ++i) tw[i] = tw_sm[1][tau][i];
            inv_contig<LOGD>(y, tw);
        }
        xchg_c2s<LOGD>(tile, tau, y, x);
        int32_t o[16];
        inv_strided<LOGD, 2>(x, o);
        if (live) {
            int32_t* dst = out + p * C::D + tau;
#pragma unroll
            for (int k = 0; k < 16; ++k) dst[C::TPP * k] = o[k];
        }
    }
}

// out = cent(a + b); polynomials.py:282-318 (__add__), also the coefficient representation's :114-148
__global__ void __launch_bounds__(256) pointwise_add_kernel(const int4* __restrict__ a, const int4* __restrict__ b, int4* __restrict__ out, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        int4 va = ld_stream(a + i), vb = ld_stream(b + i), r;
        r.x = cent_i64((long long)va.x + vb.x); r.y = cent_i64((long long)va.y + vb.y);
        r.z = cent_i64((long long)va.z + vb.z); r.w = cent_i64((long long)va.w + vb.w);
        st_stream(out + i, r);
    }
}

// out = cent(a*b) or cent(cent(a*b) + c); n4 = number of int4 groups.  polynomials.py:341-385, 282-318
template <bool ADD>
__global__ void __launch_bounds__(256) pointwise_kernel(const int4* __restrict__ a, const int4* __restrict__ b, const int4* __restrict__ c,
                                                          int4* __restrict__ out, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        int4 va = ld_stream(a + i), vb = ld_stream(b + i), r;
        if (ADD) {
            int4 vc = ld_stream(c + i);
            r.x = cent_i64((long long)va.x * vb.x + vc.x);
            r.y = cent_i64((long long)va.y * vb.y + vc.y);
            r.z = cent_i64((long long)va.z * vb.z + vc.z);
            r.w = cent_i64((long long)va.w * vb.w + vc.w);
        } else {
            r.x = mul_cent(va.x, vb.x); r.y = mul_cent(va.y, vb.y); r.z = mul_cent(va.z, vb.z); r.w = mul_cent(va.w, vb.w);
        }
        st_stream(out + i, r);
    }
}

// ---------------------------------------------------------------- host side
int ntt_init_tables(fz_ctx* ctx) {
    const int levels[2][2] = {{6, 23584283}, {8, 3337519}};  // fusion.py:36-37
    for (int L = 0; L < 2; ++L) {
        const int logd = levels[L][0], d = 1 << logd;
        Sh f[256], iv[256];
        Tw last[4];
        ntt_build_tables(logd, (unsigned long long)levels[L][1], f, iv, last);
        cudaError_t e;
        if (logd == 8) {
            e = cudaMemcpyToSymbol(c_fwd256, f, sizeof(Sh) * d);
            if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_inv256, iv, sizeof(Sh) * d);
            if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_last256, last, sizeof(last));
        } else {
            e = cudaMemcpyToSymbol(c_fwd64, f, sizeof(Sh) * d);
            if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_inv64, iv, sizeof(Sh) * d);
            if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_last64, last, sizeof(last));
        }
        if (check(e, "twiddle upload")) return -1;
    }
    (void)ctx;
    return 0;
}

template <typename K>
static int persistent_grid(const fz_ctx* ctx, K kernel, int threads, size_t work_items) {
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0);
    if (occ < 1) occ = 1;
    size_t g = (size_t)ctx->sm_count * occ;
    if (work_items < g) g = work_items;
    return g < 1 ? 1 : (int)g;
}

int launch_ntt(const fz_ctx* ctx, int inverse, const int32_t* in, int32_t* out, size_t n_polys, cudaStream_t st) {
    if (n_polys == 0) return 0;
    if (ctx->logd == 8) {
        const size_t items = (n_polys + 7) / 8;
        if (inverse) FZ_LAUNCH(ntt_inv_kernel<8>, persistent_grid(ctx, ntt_inv_kernel<8>, NTT_THREADS, items), NTT_THREADS, 0, st, in, out, n_polys);
        else FZ_LAUNCH(ntt_fwd_kernel<8>, persistent_grid(ctx, ntt_fwd_kernel<8>, NTT_THREADS, items), NTT_THREADS, 0, st, in, out, n_polys);
    } else {
        const size_t items = (n_polys + 31) / 32;
        if (inverse) FZ_LAUNCH(ntt_inv_kernel<6>, persistent_grid(ctx, ntt_inv_kernel<6>, NTT_THREADS, items), NTT_THREADS, 0, st, in, out, n_polys);
        else FZ_LAUNCH(ntt_fwd_kernel<6>, persistent_grid(ctx, ntt_fwd_kernel<6>, NTT_THREADS, items), NTT_THREADS, 0, st, in, out, n_polys);
    }
    return check(cudaGetLastError(), inverse ? "ntt_inv launch" : "ntt_fwd launch");
}

int launch_polymul(const fz_ctx* ctx, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, cudaStream_t st) {
    if (n_polys == 0) return 0;
    if (ctx->logd == 8)
        FZ_LAUNCH(polymul_kernel<8>, persistent_grid(ctx, polymul_kernel<8>, NTT_THREADS, (n_polys + 7) / 8), NTT_THREADS, 0, st, a, b, out, n_polys);
    else
        FZ_LAUNCH(polymul_kernel<6>, persistent_grid(ctx, polymul_kernel<6>, NTT_THREADS, (n_polys + 31) / 32), NTT_THREADS, 0, st, a, b, out, n_polys);
    return check(cudaGetLastError(), "polymul launch");
}

int launch_ntt_pass(const fz_ctx* ctx, const int32_t* a, const int32_t* bhat, int32_t* out, size_t n_polys, cudaStream_t st) {
    if (n_polys == 0) return 0;
    if (ctx->logd == 8)
        FZ_LAUNCH(ntt_pass_kernel<8>, persistent_grid(ctx, ntt_pass_kernel<8>, NTT_THREADS, (n_polys + 7) / 8), NTT_THREADS, 0, st, a, bhat, out, n_polys);
    else
        FZ_LAUNCH(ntt_pass_kernel<6>, persistent_grid(ctx, ntt_pass_kernel<6>, NTT_THREADS, (n_polys + 31) / 32), NTT_THREADS, 0, st, a, bhat, out, n_polys);
    return check(cudaGetLastError(), "ntt_pass launch");
}

int launch_pointwise_add(const fz_ctx* ctx, const int32_t* a, const int32_t* b, int32_t* out, size_t n_coef, cudaStream_t st) {
    if (n_coef == 0) return 0;
    const size_t n4 = n_coef / 4;
    size_t blocks = (n4 + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    FZ_LAUNCH(pointwise_add_kernel, (int)blocks, 256, 0, st, (const int4*)a, (const int4*)b, (int4*)out, n4);
    return check(cudaGetLastError(), "pointwise_add launch");
}

int launch_pointwise(const fz_ctx* ctx, const int32_t* a, const int32_t* b, const int32_t* c, int32_t* out, size_t n_coef, cudaStream_t st) {
    if (n_coef == 0) return 0;
    const size_t n4 = n_coef / 4;  // d is a multiple of 4
    size_t blocks = (n4 + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (c) FZ_LAUNCH(pointwise_kernel<true>, (int)blocks, 256, 0, st, (const int4*)a, (const int4*)b, (const int4*)c, (int4*)out, n4);
    else FZ_LAUNCH(pointwise_kernel<false>, (int)blocks, 256, 0, st, (const int4*)a, (const int4*)b, nullptr, (int4*)out, n4);
    return check(cudaGetLastError(), "pointwise launch");
}

}  // namespace fz
