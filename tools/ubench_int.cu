// ubench_int.cu -- integer pipe throughput micro-benchmark for sm_100a (development tool,
// not part of the product).  Measures warp-instructions per cycle per SM for the opcodes the
// NTT butterfly is built from.  Build: nvcc -arch=sm_100a -O3 -o ubench_int ubench_int.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
#define NCHAIN 8

template <int OP>
__global__ void k(uint32_t* out, uint32_t seed, long long* cyc) {
    uint32_t a[NCHAIN], b = seed | 1, c = seed * 7 + 3;
    unsigned long long w[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; i++) { a[i] = threadIdx.x * 31 + i * 977 + seed; w[i] = a[i]; }
    long long t0 = clock64();
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; i++) {
            if (OP == 0) a[i] = a[i] * b + c;                                   // IMAD
            if (OP == 1) a[i] = __umulhi(a[i], b) + c;                          // IMAD.HI (+add folded?)
            if (OP == 2) { w[i] = (unsigned long long)(uint32_t)w[i] * b + w[i]; }  // IMAD.WIDE accumulate
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));   // IADD
            if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));  // LOP3
            if (OP == 5) a[i] = min(a[i], a[i] - 2147465729u) + c;              // VIADDMNMX + add
            if (OP == 6) a[i] = __funnelshift_l(a[i], c, 7);                    // SHF

            if (OP == 8) { uint32_t h = __umulhi(a[i], b); a[i] = a[i] * c + h; }  // HI + IMAD (plantard first half)
            if (OP == 9) { unsigned long long p = (unsigned long long)a[i] * b; a[i] = (uint32_t)(p >> 32) + (uint32_t)p; } // WIDE use both
            if (OP == 10) a[i] = (uint32_t)(((unsigned long long)a[i] * 2147465729u + 2147465729u) >> 32) + c;  // HI with addend
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NCHAIN; i++) s += a[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int OP>
void run(const char* name, int per_iter_instr) {
    uint32_t* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    k<OP><<<148, 1024>>>(out, 12345, cyc);
    cudaDeviceSynchronize();
    k<OP><<<148, 1024>>>(out, 12345, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double winstr = 32.0 * ITER * NCHAIN * per_iter_instr;  // warp-instructions per SM (32 warps)
    printf("%-28s cycles=%lld  warp-instr/cycle/SM=%.3f  (per SMSP %.3f)\n", name, h, winstr / h, winstr / h / 4);
    cudaFree(out); cudaFree(cyc);
}

// type-pun helper for OP 7
int main() {
    run<0>("IMAD", 1);
    run<1>("IMAD.HI(+c)", 1);
    run<2>("IMAD.WIDE acc", 1);
    run<3>("IADD", 1);
    run<4>("LOP3", 1);
    run<5>("VIADDMNMX+IADD", 2);
    run<6>("SHF", 1);
    run<8>("IMAD.HI+IMAD", 2);
    run<9>("IMAD.WIDE+IADD", 2);
    run<10>("IMAD.HI addend64 +IADD", 2);
    return 0;
}
