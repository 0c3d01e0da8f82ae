// keccak_lab2.cu -- development microbenchmark, layout V22: TWO lanes per sponge chain, bit-interleaved.
// Lane 0 of a pair holds the even bits of every 64-bit Keccak lane, lane 1 the odd bits (25 registers each).  XOR / AND-NOT are bitwise, so theta's
// parity, the theta application and chi need no communication; a rotation by 2s is a 32-bit rotation by s of each half, a rotation by 2s + 1 swaps the
// halves (even' = rot(odd, s + 1), odd' = rot(even, s)): one SHFL.BFLY per odd rotation, 12 of rho's 24 and theta's five = 17 shuffles and 90 ALU-pipe
// instructions per lane and round (a thread per sponge: 180 ALU-pipe instructions, no shuffles; a warp per sponge: 18 shuffles in three dependent stages).
// 16 chains per warp.  Question: cycles per round of a warp alone on its sub-partition and with 2 / 4 warps on it.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/keccak_lab2 tools/keccak_lab2.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "../fusion_cryptography_b200/csrc/fz_keccak.cuh"

using namespace fz;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__constant__ uint32_t c_rc2[24][2];   // round constants, even bits / odd bits

__host__ __device__ inline uint32_t even_bits(unsigned long long w) {
    uint32_t r = 0;
    for (int i = 0; i < 32; ++i) r |= (uint32_t)((w >> (2 * i)) & 1ULL) << i;
    return r;
}
__host__ __device__ inline unsigned long long spread(uint32_t e, uint32_t o) {
    unsigned long long w = 0;
    for (int i = 0; i < 32; ++i) w |= ((unsigned long long)((e >> i) & 1u) << (2 * i)) | ((unsigned long long)((o >> i) & 1u) << (2 * i + 1));
    return w;
}

__device__ __forceinline__ uint32_t rotl32(uint32_t x, uint32_t s) { return __funnelshift_l(x, x, s); }

// rotation of the 64-bit lane by N, this thread holding half `odd` (0 even bits, 1 odd bits) in v
template <int N>
__device__ __forceinline__ uint32_t rot2(uint32_t v, uint32_t odd) {
    if constexpr (N == 0) return v;
    else if constexpr ((N & 1) == 0) return rotl32(v, N / 2);
    else {
        const uint32_t p = __shfl_xor_sync(0xffffffffu, v, 1);
        // even' = rot(odd, s + 1), odd' = rot(even, s)
        return rotl32(p, (uint32_t)(N / 2) + 1u - odd);
    }
}

__device__ __forceinline__ uint32_t x5(uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t e) { return a ^ b ^ c ^ d ^ e; }
__device__ __forceinline__ uint32_t ch(uint32_t a, uint32_t b, uint32_t c) { return a ^ (~b & c); }

__device__ __forceinline__ void round2(uint32_t (&a)[25], uint32_t odd, uint32_t rc) {
    const uint32_t c0 = x5(a[0], a[5], a[10], a[15], a[20]);
    const uint32_t c1 = x5(a[1], a[6], a[11], a[16], a[21]);
    const uint32_t c2 = x5(a[2], a[7], a[12], a[17], a[22]);
    const uint32_t c3 = x5(a[3], a[8], a[13], a[18], a[23]);
    const uint32_t c4 = x5(a[4], a[9], a[14], a[19], a[24]);
    const uint32_t r0 = rot2<1>(c0, odd), r1 = rot2<1>(c1, odd), r2 = rot2<1>(c2, odd), r3 = rot2<1>(c3, odd), r4 = rot2<1>(c4, odd);
#define TH0(v) ((v) ^ c4 ^ r1)
#define TH1(v) ((v) ^ c0 ^ r2)
#define TH2(v) ((v) ^ c1 ^ r3)
#define TH3(v) ((v) ^ c2 ^ r4)
#define TH4(v) ((v) ^ c3 ^ r0)
    const uint32_t b0 = TH0(a[0]);
    const uint32_t b10 = rot2<1>(TH1(a[1]), odd);
    const uint32_t b20 = rot2<62>(TH2(a[2]), odd);
    const uint32_t b5 = rot2<28>(TH3(a[3]), odd);
    const uint32_t b15 = rot2<27>(TH4(a[4]), odd);
    const uint32_t b16 = rot2<36>(TH0(a[5]), odd);
    const uint32_t b1 = rot2<44>(TH1(a[6]), odd);
    const uint32_t b11 = rot2<6>(TH2(a[7]), odd);
    const uint32_t b21 = rot2<55>(TH3(a[8]), odd);
    const uint32_t b6 = rot2<20>(TH4(a[9]), odd);
    const uint32_t b7 = rot2<3>(TH0(a[10]), odd);
    const uint32_t b17 = rot2<10>(TH1(a[11]), odd);
    const uint32_t b2 = rot2<43>(TH2(a[12]), odd);
    const uint32_t b12 = rot2<25>(TH3(a[13]), odd);
    const uint32_t b22 = rot2<39>(TH4(a[14]), odd);
    const uint32_t b23 = rot2<41>(TH0(a[15]), odd);
    const uint32_t b8 = rot2<45>(TH1(a[16]), odd);
    const uint32_t b18 = rot2<15>(TH2(a[17]), odd);
    const uint32_t b3 = rot2<21>(TH3(a[18]), odd);
    const uint32_t b13 = rot2<8>(TH4(a[19]), odd);
    const uint32_t b14 = rot2<18>(TH0(a[20]), odd);
    const uint32_t b24 = rot2<2>(TH1(a[21]), odd);
    const uint32_t b9 = rot2<61>(TH2(a[22]), odd);
    const uint32_t b19 = rot2<56>(TH3(a[23]), odd);
    const uint32_t b4 = rot2<14>(TH4(a[24]), odd);
#undef TH0
#undef TH1
#undef TH2
#undef TH3
#undef TH4
    a[0] = ch(b0, b1, b2) ^ rc; a[1] = ch(b1, b2, b3); a[2] = ch(b2, b3, b4); a[3] = ch(b3, b4, b0); a[4] = ch(b4, b0, b1);
    a[5] = ch(b5, b6, b7); a[6] = ch(b6, b7, b8); a[7] = ch(b7, b8, b9); a[8] = ch(b8, b9, b5); a[9] = ch(b9, b5, b6);
    a[10] = ch(b10, b11, b12); a[11] = ch(b11, b12, b13); a[12] = ch(b12, b13, b14); a[13] = ch(b13, b14, b10); a[14] = ch(b14, b10, b11);
    a[15] = ch(b15, b16, b17); a[16] = ch(b16, b17, b18); a[17] = ch(b17, b18, b19); a[18] = ch(b18, b19, b15); a[19] = ch(b19, b15, b16);
    a[20] = ch(b20, b21, b22); a[21] = ch(b21, b22, b23); a[22] = ch(b22, b23, b24); a[23] = ch(b23, b24, b20); a[24] = ch(b24, b20, b21);
}

template <int UNROLL>
__global__ void __launch_bounds__(128) pair_kernel(int perms, unsigned long long* state_out, long long* cycles_out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int chain = t >> 1;
    const uint32_t odd = (uint32_t)(t & 1);
    uint32_t a[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) {
        const unsigned long long w = (unsigned long long)(k + 1) * 0x9E3779B97F4A7C15ULL + (unsigned long long)chain;
        a[k] = even_bits(w >> odd);
    }
    const long long t0 = clock64();
    for (int p = 0; p < perms; ++p) {
#pragma unroll UNROLL
        for (int r = 0; r < 24; ++r) round2(a, odd, c_rc2[r][odd]);
    }
    const long long t1 = clock64();
    if (chain == 0) {
#pragma unroll
        for (int k = 0; k < 25; ++k) {
            const uint32_t o = __shfl_xor_sync(0x3u, a[k], 1);
            if (!odd) state_out[k] = spread(a[k], o);
        }
    }
    if ((threadIdx.x & 31) == 0) cycles_out[t >> 5] = t1 - t0;
}


// ---- V23: two lanes per chain, lane 0 holds the LOW 32 bits of every Keccak lane, lane 1 the HIGH 32 bits.  No conversion of the message or of the
// squeezed bytes (each lane loads / stores its own 4 bytes of every 8), but every rotation except by 0 and 32 needs the partner's half: 23 + 5 = 28
// shuffles per round.
template <int N>
__device__ __forceinline__ uint32_t rot3(uint32_t v, uint32_t hi) {
    if constexpr (N == 0) return v;
    const uint32_t p = __shfl_xor_sync(0xffffffffu, v, 1);
    if constexpr (N == 32) return p;
    // rot64 by n < 32: lo' = (lo << n) | (hi >> (32 - n)), hi' = (hi << n) | (lo >> (32 - n)): funnel(other, mine, n); by 32 + n the halves swap first
    constexpr int n = N & 31;
    if constexpr (N < 32) return __funnelshift_l(p, v, n);
    else return __funnelshift_l(v, p, n);
}

__device__ __forceinline__ void round3(uint32_t (&a)[25], uint32_t hi, uint32_t rc) {
    const uint32_t c0 = x5(a[0], a[5], a[10], a[15], a[20]);
    const uint32_t c1 = x5(a[1], a[6], a[11], a[16], a[21]);
    const uint32_t c2 = x5(a[2], a[7], a[12], a[17], a[22]);
    const uint32_t c3 = x5(a[3], a[8], a[13], a[18], a[23]);
    const uint32_t c4 = x5(a[4], a[9], a[14], a[19], a[24]);
    const uint32_t r0 = rot3<1>(c0, hi), r1 = rot3<1>(c1, hi), r2 = rot3<1>(c2, hi), r3 = rot3<1>(c3, hi), r4 = rot3<1>(c4, hi);
#define TH0(v) ((v) ^ c4 ^ r1)
#define TH1(v) ((v) ^ c0 ^ r2)
#define TH2(v) ((v) ^ c1 ^ r3)
#define TH3(v) ((v) ^ c2 ^ r4)
#define TH4(v) ((v) ^ c3 ^ r0)
    const uint32_t b0 = TH0(a[0]);
    const uint32_t b10 = rot3<1>(TH1(a[1]), hi);
    const uint32_t b20 = rot3<62>(TH2(a[2]), hi);
    const uint32_t b5 = rot3<28>(TH3(a[3]), hi);
    const uint32_t b15 = rot3<27>(TH4(a[4]), hi);
    const uint32_t b16 = rot3<36>(TH0(a[5]), hi);
    const uint32_t b1 = rot3<44>(TH1(a[6]), hi);
    const uint32_t b11 = rot3<6>(TH2(a[7]), hi);
    const uint32_t b21 = rot3<55>(TH3(a[8]), hi);
    const uint32_t b6 = rot3<20>(TH4(a[9]), hi);
    const uint32_t b7 = rot3<3>(TH0(a[10]), hi);
    const uint32_t b17 = rot3<10>(TH1(a[11]), hi);
    const uint32_t b2 = rot3<43>(TH2(a[12]), hi);
    const uint32_t b12 = rot3<25>(TH3(a[13]), hi);
    const uint32_t b22 = rot3<39>(TH4(a[14]), hi);
    const uint32_t b23 = rot3<41>(TH0(a[15]), hi);
    const uint32_t b8 = rot3<45>(TH1(a[16]), hi);
    const uint32_t b18 = rot3<15>(TH2(a[17]), hi);
    const uint32_t b3 = rot3<21>(TH3(a[18]), hi);
    const uint32_t b13 = rot3<8>(TH4(a[19]), hi);
    const uint32_t b14 = rot3<18>(TH0(a[20]), hi);
    const uint32_t b24 = rot3<2>(TH1(a[21]), hi);
    const uint32_t b9 = rot3<61>(TH2(a[22]), hi);
    const uint32_t b19 = rot3<56>(TH3(a[23]), hi);
    const uint32_t b4 = rot3<14>(TH4(a[24]), hi);
#undef TH0
#undef TH1
#undef TH2
#undef TH3
#undef TH4
    a[0] = ch(b0, b1, b2) ^ rc; a[1] = ch(b1, b2, b3); a[2] = ch(b2, b3, b4); a[3] = ch(b3, b4, b0); a[4] = ch(b4, b0, b1);
    a[5] = ch(b5, b6, b7); a[6] = ch(b6, b7, b8); a[7] = ch(b7, b8, b9); a[8] = ch(b8, b9, b5); a[9] = ch(b9, b5, b6);
    a[10] = ch(b10, b11, b12); a[11] = ch(b11, b12, b13); a[12] = ch(b12, b13, b14); a[13] = ch(b13, b14, b10); a[14] = ch(b14, b10, b11);
    a[15] = ch(b15, b16, b17); a[16] = ch(b16, b17, b18); a[17] = ch(b17, b18, b19); a[18] = ch(b18, b19, b15); a[19] = ch(b19, b15, b16);
    a[20] = ch(b20, b21, b22); a[21] = ch(b21, b22, b23); a[22] = ch(b22, b23, b24); a[23] = ch(b23, b24, b20); a[24] = ch(b24, b20, b21);
}

template <int UNROLL>
__global__ void __launch_bounds__(128) pair3_kernel(int perms, unsigned long long* state_out, long long* cycles_out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int chain = t >> 1;
    const uint32_t hi = (uint32_t)(t & 1);
    uint32_t a[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) {
        const unsigned long long w = (unsigned long long)(k + 1) * 0x9E3779B97F4A7C15ULL + (unsigned long long)chain;
        a[k] = (uint32_t)(w >> (32 * hi));
    }
    const uint32_t* rcw = reinterpret_cast<const uint32_t*>(c_keccak_rc);
    const long long t0 = clock64();
    for (int p = 0; p < perms; ++p) {
#pragma unroll UNROLL
        for (int r = 0; r < 24; ++r) round3(a, hi, rcw[2 * r + hi]);
    }
    const long long t1 = clock64();
    if (chain == 0) {
#pragma unroll
        for (int k = 0; k < 25; ++k) {
            const uint32_t o = __shfl_xor_sync(0x3u, a[k], 1);
            if (!hi) state_out[k] = ((unsigned long long)o << 32) | a[k];
        }
    }
    if ((threadIdx.x & 31) == 0) cycles_out[t >> 5] = t1 - t0;
}

__global__ void thread_kernel(int perms, unsigned long long* state_out) {
    unsigned long long a[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) a[k] = (unsigned long long)(k + 1) * 0x9E3779B97F4A7C15ULL;
    for (int p = 0; p < perms; ++p) keccak_f(a);
    for (int k = 0; k < 25; ++k) state_out[k] = a[k];
}

template <int UNROLL, int V>
void run(int sms, int perms, unsigned long long* d_state, long long* d_cyc, const std::vector<unsigned long long>& ref) {
    const int tw[] = {1, 2, 4, 8, 16};
    for (int wi = 0; wi < 5; ++wi) {
        const int w = tw[wi];
        const int bw = w < 4 ? w : 4, blocks = sms * (w / bw);
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        if (V == 22) pair_kernel<UNROLL><<<blocks, bw * 32>>>(10, d_state, d_cyc); else pair3_kernel<UNROLL><<<blocks, bw * 32>>>(10, d_state, d_cyc);
        CK(cudaEventRecord(e0));
        if (V == 22) pair_kernel<UNROLL><<<blocks, bw * 32>>>(perms, d_state, d_cyc); else pair3_kernel<UNROLL><<<blocks, bw * 32>>>(perms, d_state, d_cyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<long long> cyc(blocks * bw);
        std::vector<unsigned long long> got(25);
        CK(cudaMemcpy(cyc.data(), d_cyc, blocks * bw * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(got.data(), d_state, 200, cudaMemcpyDeviceToHost));
        double avg = 0, mx = 0;
        for (int i = 0; i < blocks * bw; ++i) { avg += (double)cyc[i]; if (cyc[i] > mx) mx = (double)cyc[i]; }
        avg /= blocks * bw;
        printf("V%d two lanes per chain (%s), x%-2d  warps/SM %2d (%6d chains)  cycles/round avg %7.1f max %7.1f  time %8.3f ms  %7.3f Gperm/s  %s\n", V, V == 22 ? "even / odd bits" : "low / high words", UNROLL, w, blocks * bw * 16,
               avg / perms / 24.0, mx / perms / 24.0, ms, (double)blocks * bw * 16 * perms / (ms * 1e-3) / 1e9, got == ref ? "ok" : "MISMATCH");
    }
}

int main(int argc, char** argv) {
    const int perms = argc > 1 ? atoi(argv[1]) : 500;
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    printf("device %s, %d SMs, %d perms per chain\n", pr.name, sms, perms);
    static const unsigned long long rc[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
        0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
        0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    uint32_t rc2[24][2];
    for (int r = 0; r < 24; ++r) { rc2[r][0] = even_bits(rc[r]); rc2[r][1] = even_bits(rc[r] >> 1); }
    CK(cudaMemcpyToSymbol(c_rc2, rc2, sizeof(rc2)));
    unsigned long long* d_state; long long* d_cyc;
    CK(cudaMalloc(&d_state, 25 * 8));
    CK(cudaMalloc(&d_cyc, 1 << 20));
    std::vector<unsigned long long> ref(25);
    thread_kernel<<<1, 1>>>(perms, d_state);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(ref.data(), d_state, 200, cudaMemcpyDeviceToHost));
    run<1, 22>(sms, perms, d_state, d_cyc, ref);
    run<2, 22>(sms, perms, d_state, d_cyc, ref);
    run<24, 22>(sms, perms, d_state, d_cyc, ref);
    run<1, 23>(sms, perms, d_state, d_cyc, ref);
    run<2, 23>(sms, perms, d_state, d_cyc, ref);
    run<4, 23>(sms, perms, d_state, d_cyc, ref);
    return 0;
}
