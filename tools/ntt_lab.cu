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
    for (int occ : {4, 6, 8}) {
        int grid = 148 * occ;
        float t0 = time_it([&] { ntt_fwd_kernel<8><<<grid, NTT_THREADS>>>(in, out, n); });
        float t1 = time_it([&] { v1_mem_only<8><<<grid, NTT_THREADS>>>(in, out, n); });
        float t2 = time_it([&] { v2_compute_only<8><<<grid, NTT_THREADS>>>(in, out, n); });
        float t3 = time_it([&] { v3_copy<<<grid, NTT_THREADS>>>((const int4*)in, (int4*)out, n * 64); });
        float t4 = time_it([&] { ntt_inv_kernel<8><<<grid, NTT_THREADS>>>(in, out, n); });
        printf("grid=148x%d  fwd %.3f ms (%.0f GB/s) | mem-only %.3f (%.0f) | compute-only %.3f | copy %.3f (%.0f) | inv %.3f (%.0f)\n", occ, t0, gb / t0 * 1e3, t1,
               gb / t1 * 1e3, t2, t3, gb / t3 * 1e3, t4, gb / t4 * 1e3);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
