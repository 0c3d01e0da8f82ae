// keccak_lab_fma.cu -- development microbenchmark: thread-per-sponge Keccak-f[1600] with part of the 64-bit rotations moved from the ALU pipe
// (SHF funnel shifts) to the FMA pipe (IMAD + IMAD.HI by a power of two held in constant memory, so that ptxas cannot turn it back into a shift).
// The thread-per-sponge kernels (hash_ch, the thread-per-aggregate chain kernels) are bound by the ALU pipe: 180 LOP3 / SHF per round, one warp
// instruction per 2 cycles and sub-partition = 360 cycles per round, while the FMA pipe idles.  A half rotation (one 32-bit output word) is
// 1 SHF on the ALU pipe (2 cycles) or IMAD (2) + IMAD.HI (4) on the FMA pipe; with x of the 58 half rotations moved, ALU = 360 - 2x and FMA = 6x cycles:
// balanced at x = 45 (270 cycles per round, 1.33x) if the two pipes overlap perfectly.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/keccak_lab_fma tools/keccak_lab_fma.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "../fusion_cryptography_b200/csrc/fz_keccak.cuh"

using namespace fz;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__constant__ uint32_t c_pow2[32];
__constant__ uint32_t c_pow2b[32];   // the same table again: low and high products take their multiplier from different tables, so ptxas cannot fuse them into IMAD.WIDE (which costs register-pair moves)

// rotation on the FMA pipe: (lo << n) | (hi >> (32 - n)) = lo * 2^n + mulhi(hi, 2^n)
template <int N>
__device__ __forceinline__ unsigned long long rol64_fma(unsigned long long x) {
    uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    if constexpr (N == 0) return x;
    if constexpr (N >= 32) { const uint32_t t = lo; lo = hi; hi = t; }
    constexpr int n = N & 31;
    if constexpr (n == 0) return ((unsigned long long)hi << 32) | lo;
    const uint32_t m = c_pow2[n], mb = c_pow2b[n];
    uint32_t nl, nh;
    const uint32_t tl = lo * m, th = hi * m;
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(nl) : "r"(hi), "r"(mb), "r"(tl));
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(nh) : "r"(lo), "r"(mb), "r"(th));
    return ((unsigned long long)nh << 32) | nl;
}

// lane k (in the order the round lists them) rotates on the FMA pipe when k < NF
template <int N, int K, int NF>
__device__ __forceinline__ unsigned long long rot(unsigned long long x) {
    if constexpr (K < NF) return rol64_fma<N>(x);
    else return rol64<N>(x);
}

template <int NF, int NT>   // NF of the 24 rho rotations and NT of the 5 theta rotations on the FMA pipe
__device__ __forceinline__ void round_fma(unsigned long long (&a)[25], unsigned long long rc) {
    unsigned long long c0 = xor5(a[0], a[5], a[10], a[15], a[20]);
    unsigned long long c1 = xor5(a[1], a[6], a[11], a[16], a[21]);
    unsigned long long c2 = xor5(a[2], a[7], a[12], a[17], a[22]);
    unsigned long long c3 = xor5(a[3], a[8], a[13], a[18], a[23]);
    unsigned long long c4 = xor5(a[4], a[9], a[14], a[19], a[24]);
    const unsigned long long r0 = rot<1, 0, NT>(c0), r1 = rot<1, 1, NT>(c1), r2 = rot<1, 2, NT>(c2), r3 = rot<1, 3, NT>(c3), r4 = rot<1, 4, NT>(c4);
#define TH0(v) ((v) ^ c4 ^ r1)
#define TH1(v) ((v) ^ c0 ^ r2)
#define TH2(v) ((v) ^ c1 ^ r3)
#define TH3(v) ((v) ^ c2 ^ r4)
#define TH4(v) ((v) ^ c3 ^ r0)
    unsigned long long b0 = TH0(a[0]);
    unsigned long long b10 = rot<1, 0, NF>(TH1(a[1]));
    unsigned long long b20 = rot<62, 1, NF>(TH2(a[2]));
    unsigned long long b5 = rot<28, 2, NF>(TH3(a[3]));
    unsigned long long b15 = rot<27, 3, NF>(TH4(a[4]));
    unsigned long long b16 = rot<36, 4, NF>(TH0(a[5]));
    unsigned long long b1 = rot<44, 5, NF>(TH1(a[6]));
    unsigned long long b11 = rot<6, 6, NF>(TH2(a[7]));
    unsigned long long b21 = rot<55, 7, NF>(TH3(a[8]));
    unsigned long long b6 = rot<20, 8, NF>(TH4(a[9]));
    unsigned long long b7 = rot<3, 9, NF>(TH0(a[10]));
    unsigned long long b17 = rot<10, 10, NF>(TH1(a[11]));
    unsigned long long b2 = rot<43, 11, NF>(TH2(a[12]));
    unsigned long long b12 = rot<25, 12, NF>(TH3(a[13]));
    unsigned long long b22 = rot<39, 13, NF>(TH4(a[14]));
    unsigned long long b23 = rot<41, 14, NF>(TH0(a[15]));
    unsigned long long b8 = rot<45, 15, NF>(TH1(a[16]));
    unsigned long long b18 = rot<15, 16, NF>(TH2(a[17]));
    unsigned long long b3 = rot<21, 17, NF>(TH3(a[18]));
    unsigned long long b13 = rot<8, 18, NF>(TH4(a[19]));
    unsigned long long b14 = rot<18, 19, NF>(TH0(a[20]));
    unsigned long long b24 = rot<2, 20, NF>(TH1(a[21]));
    unsigned long long b9 = rot<61, 21, NF>(TH2(a[22]));
    unsigned long long b19 = rot<56, 22, NF>(TH3(a[23]));
    unsigned long long b4 = rot<14, 23, NF>(TH4(a[24]));
#undef TH0
#undef TH1
#undef TH2
#undef TH3
#undef TH4
    a[0] = chi(b0, b1, b2) ^ rc; a[1] = chi(b1, b2, b3); a[2] = chi(b2, b3, b4); a[3] = chi(b3, b4, b0); a[4] = chi(b4, b0, b1);
    a[5] = chi(b5, b6, b7); a[6] = chi(b6, b7, b8); a[7] = chi(b7, b8, b9); a[8] = chi(b8, b9, b5); a[9] = chi(b9, b5, b6);
    a[10] = chi(b10, b11, b12); a[11] = chi(b11, b12, b13); a[12] = chi(b12, b13, b14); a[13] = chi(b13, b14, b10); a[14] = chi(b14, b10, b11);
    a[15] = chi(b15, b16, b17); a[16] = chi(b16, b17, b18); a[17] = chi(b17, b18, b19); a[18] = chi(b18, b19, b15); a[19] = chi(b19, b15, b16);
    a[20] = chi(b20, b21, b22); a[21] = chi(b21, b22, b23); a[22] = chi(b22, b23, b24); a[23] = chi(b23, b24, b20); a[24] = chi(b24, b20, b21);
}

template <int NF, int NT>
__global__ void __launch_bounds__(128) thread_kernel(int perms, unsigned long long* state_out, long long* cycles_out) {
    const int chain = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long a[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) a[k] = (unsigned long long)(k + 1) * 0x9E3779B97F4A7C15ULL + (unsigned long long)chain;
    const long long t0 = clock64();
    for (int p = 0; p < perms; ++p) {
#pragma unroll 1
        for (int r = 0; r < 24; ++r) round_fma<NF, NT>(a, c_keccak_rc[r]);
    }
    const long long t1 = clock64();
    if (chain == 0)
        for (int k = 0; k < 25; ++k) state_out[k] = a[k];
    if ((threadIdx.x & 31) == 0) cycles_out[chain >> 5] = t1 - t0;
}

static unsigned long long* d_state;
static long long* d_cyc;
static std::vector<unsigned long long> ref(25);

template <int NF, int NT>
void run(int sms, int perms) {
    const int tw[] = {1, 4, 8, 16};
    for (int wi = 0; wi < 4; ++wi) {
        const int w = tw[wi];
        const int bw = w < 4 ? w : 4, blocks = sms * (w / bw);
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        thread_kernel<NF, NT><<<blocks, bw * 32>>>(10, d_state, d_cyc);
        CK(cudaEventRecord(e0));
        thread_kernel<NF, NT><<<blocks, bw * 32>>>(perms, d_state, d_cyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<long long> cyc(blocks * bw);
        std::vector<unsigned long long> got(25);
        CK(cudaMemcpy(cyc.data(), d_cyc, blocks * bw * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(got.data(), d_state, 200, cudaMemcpyDeviceToHost));
        if (NF == 0 && NT == 0 && wi == 0) ref = got;
        double avg = 0;
        for (int i = 0; i < blocks * bw; ++i) avg += (double)cyc[i];
        avg /= blocks * bw;
        printf("rho on FMA %2d / 24, theta on FMA %d / 5  warps/SM %2d  cycles/round avg %7.1f  time %8.3f ms  %7.3f Gperm/s  %s\n", NF, NT, w, avg / perms / 24.0, ms,
               (double)blocks * bw * 32 * perms / (ms * 1e-3) / 1e9, got == ref ? "ok" : "MISMATCH");
    }
}

int main(int argc, char** argv) {
    const int perms = argc > 1 ? atoi(argv[1]) : 500;
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    printf("device %s, %d SMs, %d perms per chain\n", pr.name, sms, perms);
    uint32_t p2[32];
    for (int i = 0; i < 32; ++i) p2[i] = 1u << i;
    CK(cudaMemcpyToSymbol(c_pow2, p2, sizeof(p2)));
    CK(cudaMemcpyToSymbol(c_pow2b, p2, sizeof(p2)));
    CK(cudaMalloc(&d_state, 25 * 8));
    CK(cudaMalloc(&d_cyc, 1 << 20));
    run<0, 0>(sms, perms);
    run<8, 0>(sms, perms);
    run<12, 0>(sms, perms);
    run<16, 0>(sms, perms);
    run<18, 0>(sms, perms);
    run<20, 0>(sms, perms);
    run<22, 0>(sms, perms);
    run<24, 0>(sms, perms);
    run<20, 3>(sms, perms);
    run<24, 5>(sms, perms);
    return 0;
}
