// sample_lab.cu -- development microbenchmark for the MT19937 sampler (fz_sample.cu): cycles per generator spent in seeding
// (init_by_array: two serial recurrences) and in the draw loop, at the shipped CTA shape (90 generators per SM).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/sample_lab tools/sample_lab.cu
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <vector>
#include "../fusion_cryptography_b200/csrc/fz_sample.cu"

namespace fz {
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
int check(cudaError_t e, const char* what) { if (e == cudaSuccess) return 0; set_error("%s: %s", what, cudaGetErrorString(e)); return -2; }

template <int MODE>   // 0: as shipped; 1: no global stores (checksum instead); 2: regeneration only (no state machine, no stores)
__global__ void __launch_bounds__(SMP_THREADS) probe_kernel(const uint32_t* __restrict__ keys, size_t n, int d, uint32_t bound, int32_t* __restrict__ out,
                                                           long long* __restrict__ cyc) {
    extern __shared__ __align__(16) uint32_t smp_sm[];
    const size_t i = (size_t)blockIdx.x * SMP_STREAMS + threadIdx.x;
    if (threadIdx.x >= SMP_STREAMS || i >= n) return;
    Mt s;
    s.m = smp_sm + threadIdx.x;
    uint32_t key[8];
    for (int j = 0; j < 8; ++j) key[j] = j < 1 ? keys[i * 8 + j] : 0u;
    const long long t0 = clock64();
    mt_seed(s, key, 1);
    const long long t1 = clock64();
    int32_t* o = out + i * d;
    const int shb = __clz(bound);
    int k = 0, trips = 0;
    uint32_t have_mag = 0, mag = 0;
    while (k < d) {
        uint32_t y[SMP_TRIP];
        mt_next_b<SMP_TRIP>(s, y);
        ++trips;
#pragma unroll
        for (int e = 0; e < SMP_TRIP; ++e) {
            const uint32_t rm = y[e] >> shb, rs = y[e] >> 30;
            const uint32_t live = k < d ? 1u : 0u;
            const uint32_t take_mag = live & (have_mag ^ 1u) & (rm < bound ? 1u : 0u);
            const uint32_t take_sg = live & have_mag & (rs < 2u ? 1u : 0u);
            mag = take_mag ? rm + 1u : mag;
            const int32_t c = rs ? -(int32_t)mag : (int32_t)mag;
            if (MODE == 0) { if (take_sg) o[k] = c; }
            else if (MODE == 1) { mag ^= take_sg ? (uint32_t)c * 31u : 0u; }
            else { mag += y[e]; }
            k += MODE == 2 ? (e == 0 ? 5 : 0) : (int)take_sg;
            have_mag = take_mag | (have_mag & (take_sg ^ 1u));
        }
    }
    if (MODE != 0 && mag == 0x12345678u) o[0] = (int32_t)mag;
    const long long t2 = clock64();
    if ((threadIdx.x & 31) == 0) { cyc[(i / 32) * 3] = t1 - t0; cyc[(i / 32) * 3 + 1] = t2 - t1; cyc[(i / 32) * 3 + 2] = trips; }
}
}  // namespace fz

int main(int argc, char** argv) {
    using namespace fz;
    const size_t n = argc > 1 ? atol(argv[1]) : 262144;
    fz_ctx ctx{};
    ctx.d = 256;
    if (sample_init_tables(&ctx)) return 1;
    cudaFuncSetAttribute(probe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMP_SMEM);
    cudaFuncSetAttribute(probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMP_SMEM);
    cudaFuncSetAttribute(probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMP_SMEM);
    std::vector<uint32_t> keys(n * 8, 0);
    for (size_t i = 0; i < n; ++i) keys[i * 8] = 1000000u + (uint32_t)i;
    uint32_t* d_keys; int32_t* d_out; long long* d_cyc;
    cudaMalloc(&d_keys, n * 32); cudaMalloc(&d_out, n * 1024); cudaMalloc(&d_cyc, (n / 32 + 1) * 24);
    cudaMemcpy(d_keys, keys.data(), n * 32, cudaMemcpyHostToDevice);
    const unsigned grid = (unsigned)((n + SMP_STREAMS - 1) / SMP_STREAMS);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < 6; ++it) {
        cudaEventRecord(e0);
        if (it % 3 == 0) probe_kernel<0><<<grid, SMP_THREADS, SMP_SMEM>>>(d_keys, n, 256, 52u, d_out, d_cyc);
        if (it % 3 == 1) probe_kernel<1><<<grid, SMP_THREADS, SMP_SMEM>>>(d_keys, n, 256, 52u, d_out, d_cyc);
        if (it % 3 == 2) probe_kernel<2><<<grid, SMP_THREADS, SMP_SMEM>>>(d_keys, n, 256, 52u, d_out, d_cyc);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> cyc((n / 32 + 1) * 3);
        cudaMemcpy(cyc.data(), d_cyc, cyc.size() * 8, cudaMemcpyDeviceToHost);
        double a = 0, b = 0, t = 0; size_t m = 0;
        for (size_t w = 0; w + 1 < cyc.size() / 3; ++w) if (cyc[w * 3] > 0) { a += cyc[w * 3]; b += cyc[w * 3 + 1]; t += cyc[w * 3 + 2]; ++m; }
        printf("mode %d: probe %zu generators: %.3f ms (%.1f M gen/s); per warp: seed %.0f cycles, draws %.0f cycles (%.1f trips of %d)\n", it % 3, n, ms, n / ms / 1e3, a / m, b / m,
               t / m, SMP_TRIP);
    }
    // the dense kernels (sample_dense_kernel<STAGE>): time, cycle split and equality with the general kernel's output
    {
        int32_t* d_ref; cudaMalloc(&d_ref, n * 1024);
        std::vector<int32_t> kl(n, 1);
        int32_t* d_kl; cudaMalloc(&d_kl, n * 4); cudaMemcpy(d_kl, kl.data(), n * 4, cudaMemcpyHostToDevice);
        sample_kernel<<<grid, SMP_THREADS, SMP_SMEM>>>(0, d_keys, 8, d_kl, n, 256, Q, 52u, 256, d_ref);
        cudaDeviceSynchronize();
        std::vector<int32_t> ref(n * 256), got(n * 256);
        cudaMemcpy(ref.data(), d_ref, n * 1024, cudaMemcpyDeviceToHost);
        for (int stage = 1; stage >= 0; --stage) {
            const int G = stage ? SdCfg<1>::G : SdCfg<0>::G;
            const unsigned dgrid = (unsigned)((n + G - 1) / G);
            long long* d_c2; cudaMalloc(&d_c2, (size_t)dgrid * 4 * 16);
            for (int it = 0; it < 4; ++it) {
                cudaMemset(d_out, 0xff, n * 1024);
                long long* cp = it >= 2 ? d_c2 : nullptr;
                cudaEventRecord(e0);
                if (stage) sample_dense_kernel<1><<<dgrid, SD_THREADS, SdCfg<1>::SMEM>>>(d_keys, 8, d_kl, n, 256, 52u, d_out, cp);
                else sample_dense_kernel<0><<<dgrid, SD_THREADS, SdCfg<0>::SMEM>>>(d_keys, 8, d_kl, n, 256, 52u, d_out, cp);
                cudaEventRecord(e1); cudaDeviceSynchronize();
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                cudaMemcpy(got.data(), d_out, n * 1024, cudaMemcpyDeviceToHost);
                size_t bad = 0;
                for (size_t i = 0; i < n * 256; ++i) bad += got[i] != ref[i];
                double a = 0, b = 0;
                if (cp) {
                    std::vector<long long> c((size_t)dgrid * 8);
                    cudaMemcpy(c.data(), d_c2, c.size() * 8, cudaMemcpyDeviceToHost);
                    for (size_t w = 0; w < (size_t)dgrid * 4; ++w) { a += c[w * 2]; b += c[w * 2 + 1]; }
                    a /= dgrid * 4.0; b /= dgrid * 4.0;
                }
                printf("dense STAGE=%d (%d generators per CTA): %zu generators: %.3f ms (%.1f M gen/s), mismatching coefficients vs the general kernel: %zu", stage, G, n, ms,
                       n / ms / 1e3, bad);
                if (cp) printf("; per warp: seed %.0f cycles, draws + write-out %.0f cycles", a, b);
                printf("\n");
            }
        }
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
