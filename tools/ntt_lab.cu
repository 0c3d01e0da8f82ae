// ntt_lab.cu -- development harness (not part of the product): times variants of the forward
// NTT kernel to separate memory-side and compute-side limits.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ntt_lab ntt_lab.cu
#include <cstdarg>
#include <cstdio>
#include <vector>
#include "../fusion_cryptography_b200/csrc/fz_ntt.cu"

namespace fz {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fprintf(stderr, "\n"); }
int check(cudaError_t e, const char* what) { if (e == cudaSuccess) return 0; fprintf(stderr, "%s: %s\n", what, cudaGetErrorString(e)); return -2; }

// V4 (experiment, not shipped): measured 0.91 ms for 2^20 polynomials against 0.43 ms for the register-prefetch kernel -- one
// 1 KiB bulk copy per polynomial makes the TMA unit's per-operation rate (about one bulk copy per 120 cycles per SM) the limit.
// ---------------------------------------------------------------------------------------------------------------------
// TMA-fed forward transform: the warp's next tile (PPW polynomials) is brought into shared memory by bulk asynchronous
// copies (cp.async.bulk -> SASS UBLKCP) that complete on an mbarrier, two stages per warp, instead of being prefetched
// into 16 registers per lane.  The stage buffer doubles as the warp's transpose tile once its coefficients are in
// registers, and is refilled (tile + 2) right after the transpose.  Polynomial p of the warp sits at p * (D + PAD) words so
// that the strided reads (lane tau reads e = tau + TPP * k) are bank-conflict free.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

template <int LOGD>
__global__ void __launch_bounds__(NTT_THREADS, 8) ntt_fwd_bulk_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, size_t n_polys) {
    using C = Cfg<LOGD>;
    constexpr int WARPS = NTT_THREADS / 32;
    constexpr int PAD = LOGD == 8 ? 16 : 4;
    static_assert(C::PPW * (C::D + PAD) <= C::WARP_WORDS, "stage fits in the transpose tile");
    __shared__ __align__(128) uint32_t smem[WARPS * 2 * C::WARP_WORDS];
    __shared__ __align__(8) unsigned long long bars[WARPS * 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tau = lane % C::TPP, pw = lane / C::TPP;
    uint32_t* wbuf = smem + warp * 2 * C::WARP_WORDS;
    unsigned long long* wbar = bars + warp * 2;
    Sh tw2[15];
    load_tw2<LOGD>(C::fwd(), tau, tw2);
    const size_t per_cta = WARPS * C::PPW, stride = (size_t)gridDim.x * per_cta;
    const size_t first = (size_t)blockIdx.x * per_cta + warp * C::PPW;   // first polynomial of this warp's first tile
    if (lane == 0) { mbar_init(wbar, 1); mbar_init(wbar + 1, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    auto issue = [&](size_t p0, int stage) {   // lane 0: fetch polynomials p0 .. p0 + PPW - 1 (those that exist) into `stage`
        if (p0 >= n_polys) return;
        const int live = n_polys - p0 < (size_t)C::PPW ? (int)(n_polys - p0) : C::PPW;
        mbar_expect_tx(wbar + stage, (uint32_t)live * C::D * 4);
        for (int q = 0; q < live; ++q) bulk_g2s(wbuf + stage * C::WARP_WORDS + q * (C::D + PAD), in + (p0 + q) * C::D, C::D * 4, wbar + stage);
    };
    if (lane == 0) { issue(first, 0); issue(first + stride, 1); }
    int it = 0;
    for (size_t p0 = first; p0 < n_polys; p0 += stride, ++it) {
        const int stage = it & 1;
        const size_t p = p0 + pw;
        const bool live = p < n_polys;
        uint32_t* buf = wbuf + stage * C::WARP_WORDS;
        mbar_wait(wbar + stage, (uint32_t)(it >> 1) & 1u);
        uint32_t x[16], y[16];
        {
            const uint32_t* src = buf + pw * (C::D + PAD) + tau;
#pragma unroll
            for (int k = 0; k < 16; ++k) x[k] = live ? to_lazy((int32_t)src[C::TPP * k]) : 0u;
        }
        __syncwarp();
        fwd_strided<LOGD>(x);
        xchg_s2c<LOGD>(buf + C::base(pw), tau, x, y);
        if (lane == 0) issue(p0 + 2 * stride, stage);   // the buffer is free again (xchg_s2c ends with __syncwarp)
        fwd_contig<LOGD>(y, tw2);
        if (live) {
            int4* dst = reinterpret_cast<int4*>(out + p * C::D + tau * 16);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                dst[c] = make_int4(cent_u(y[4 * c]), cent_u(y[4 * c + 1]), cent_u(y[4 * c + 2]), cent_u(y[4 * c + 3]));
        }
    }
}

// V1: memory pattern only (strided 4-byte loads, smem transpose, 16-byte stores), no butterflies
template <int LOGD>
__global__ void __launch_bounds__(NTT_THREADS) v1_mem_only(const int32_t* in, int32_t* out, size_t n_polys) {
    using C = Cfg<LOGD>;
    __shared__ __align__(16) uint32_t smem[(NTT_THREADS / 32) * C::WARP_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tau = lane % C::TPP, pw = lane / C::TPP;
    uint32_t* tile = smem + warp * C::WARP_WORDS + C::base(pw);
    const size_t per_cta = (NTT_THREADS / 32) * C::PPW, stride = (size_t)gridDim.x * per_cta;
    for (size_t p = (size_t)blockIdx.x * per_cta + warp * C::PPW + pw; p < n_polys; p += stride) {
        uint32_t x[16], y[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = in[p * C::D + tau + C::TPP * k];
        xchg_s2c<LOGD>(tile, tau, x, y);
        int4* dst = reinterpret_cast<int4*>(out + p * C::D + tau * 16);
#pragma unroll
        for (int c = 0; c < 4; ++c) dst[c] = make_int4(y[4 * c], y[4 * c + 1], y[4 * c + 2], y[4 * c + 3]);
    }
}

// V2: compute only: same loop trip count, global traffic only on the first and last iteration
template <int LOGD>
__global__ void __launch_bounds__(NTT_THREADS) v2_compute_only(const int32_t* in, int32_t* out, size_t n_polys) {
    using C = Cfg<LOGD>;
    __shared__ __align__(16) uint32_t smem[(NTT_THREADS / 32) * C::WARP_WORDS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tau = lane % C::TPP, pw = lane / C::TPP;
    uint32_t* tile = smem + warp * C::WARP_WORDS + C::base(pw);
    Sh tw2[15];
    load_tw2<LOGD>(C::fwd(), tau, tw2);
    const size_t per_cta = (NTT_THREADS / 32) * C::PPW, stride = (size_t)gridDim.x * per_cta;
    size_t p = (size_t)blockIdx.x * per_cta + warp * C::PPW + pw;
    uint32_t x[16], y[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) y[k] = p < n_polys ? (uint32_t)in[p * C::D + tau + C::TPP * k] : 0;
    for (size_t p0 = (size_t)blockIdx.x * per_cta; p0 < n_polys; p0 += stride) {
#pragma unroll
        for (int k = 0; k < 16; ++k) x[k] = to_lazy((int32_t)y[k]);
        fwd_strided<LOGD>(x);
        xchg_s2c<LOGD>(tile, tau, x, y);
        fwd_contig<LOGD>(y, tw2);
#pragma unroll
        for (int k = 0; k < 16; ++k) y[k] = (uint32_t)cent_u(y[k]);
    }
    if (p < n_polys) {
        int4* dst = reinterpret_cast<int4*>(out + p * C::D + tau * 16);
#pragma unroll
        for (int c = 0; c < 4; ++c) dst[c] = make_int4(y[4 * c], y[4 * c + 1], y[4 * c + 2], y[4 * c + 3]);
    }
}

// V3: plain vectorised copy with the same persistent grid
__global__ void __launch_bounds__(NTT_THREADS) v3_copy(const int4* in, int4* out, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}
}  // namespace fz

using namespace fz;

template <typename F>
float time_it(F f, int reps = 20) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; i++) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; i++) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char** argv) {
    const size_t n = 1 << 20;
    int32_t *in, *out;
    cudaMalloc(&in, n * 1024); cudaMalloc(&out, n * 1024);
    std::vector<int32_t> h(n * 256);
    uint64_t s = 88172645463325252ull;
    for (auto& v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (int32_t)(s % 2147465729ull) - 1073732864; }
    cudaMemcpy(in, h.data(), n * 1024, cudaMemcpyHostToDevice);
    fz_ctx ctx{}; ctx.logd = 8; ctx.d = 256; ctx.sm_count = 148;
    ntt_init_tables(&ctx);
    const double gb = 2.0 * n * 1024 / 1e9;
    for (int occ : {4, 6, 8, 10}) {
        int grid = 148 * occ;
        float t0 = time_it([&] { ntt_fwd_kernel<8><<<grid, NTT_THREADS>>>(in, out, n); });
        float t1 = time_it([&] { v1_mem_only<8><<<grid, NTT_THREADS>>>(in, out, n); });
        float t2 = time_it([&] { v2_compute_only<8><<<grid, NTT_THREADS>>>(in, out, n); });
        float t3 = time_it([&] { v3_copy<<<grid, NTT_THREADS>>>((const int4*)in, (int4*)out, n * 64); });
        float t4 = time_it([&] { ntt_inv_kernel<8><<<grid, NTT_THREADS>>>(in, out, n); });
        float t5 = time_it([&] { ntt_fwd_bulk_kernel<8><<<grid, NTT_THREADS>>>(in, out, n); });
        printf("grid=148x%d  fwd (TMA bulk) %.3f ms (%.0f GB/s)\n", occ, t5, gb / t5 * 1e3);
        printf("grid=148x%d  fwd %.3f ms (%.0f GB/s) | mem-only %.3f (%.0f) | compute-only %.3f | copy %.3f (%.0f) | inv %.3f (%.0f)\n", occ, t0, gb / t0 * 1e3, t1,
               gb / t1 * 1e3, t2, t3, gb / t3 * 1e3, t4, gb / t4 * 1e3);
    }
    {   // the TMA-fed kernel against the shipped one, on a ragged count
        const size_t m = n - 3;
        int32_t* out2; cudaMalloc(&out2, n * 1024);
        cudaMemset(out, 0, n * 1024); cudaMemset(out2, 0, n * 1024);
        ntt_fwd_kernel<8><<<148 * 6, NTT_THREADS>>>(in, out, m);
        ntt_fwd_bulk_kernel<8><<<148 * 6, NTT_THREADS>>>(in, out2, m);
        std::vector<int32_t> a(n * 256), b(n * 256);
        cudaMemcpy(a.data(), out, n * 1024, cudaMemcpyDeviceToHost);
        cudaMemcpy(b.data(), out2, n * 1024, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (size_t i = 0; i < a.size(); ++i) bad += a[i] != b[i];
        printf("TMA-fed forward vs shipped forward: %zu mismatches\n", bad);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
