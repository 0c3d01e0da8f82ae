// keccak_lab5.cu -- development microbenchmark, second series: FIVE LANES PER CHAIN.
//
// The aggregation hash (fusion.py:573-591) is one serial SHAKE256 per aggregate.  keccak_lab.cu measured the layouts with one
// warp per chain (25 lanes hold one Keccak lane each, three dependent shuffle stages per round: 168 cycles per round alone,
// 197-247 with two or three chains on one SM sub-partition because the shuffle port saturates).  Here a chain is owned by 5
// lanes: lane x keeps COLUMN x of the state (a[x][0..4], ten registers).  Column parity is local, theta needs the parities of the
// two neighbour columns (4 shuffles), rho is local (per-lane shift amounts), pi is ONE exchange through shared memory (each lane
// scatters its five rotated lanes to their pi positions -- the 32-bit swap of rotations by 32 or more is folded into the store
// addresses -- and gathers columns x, x+1, x+2 of the new state), chi and iota are local again.  Six chains share a warp.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/keccak_lab5 tools/keccak_lab5.cu
//   tools/keccak_lab5 [perms]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "../fusion_cryptography_b200/csrc/fz_keccak.cuh"

using namespace fz;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
#define SH(v, s) __shfl_sync(0xffffffffu, (v), (s))

__constant__ int c_rot5[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

constexpr int GROUPS = 7;            // 6 chains + one dummy group for lanes 30, 31
template <int ROWL> struct Geo {
    static constexpr int ROW_WORDS = ROWL * 2;          // 32-bit words per row Y of the exchange tile
    static constexpr int BUF_WORDS = 5 * ROW_WORDS;     // one tile
    static constexpr int WARP_WORDS = 2 * BUF_WORDS;    // double buffered
};

struct Col {
    int lm1, lp1;        // lanes holding columns x-1, x+1 of the same chain
    int s[5];            // rho shift amounts & 31 of a[x][y]
    int wl[5], wh[5];    // word index where the low / high word of rot(a[x][y]) goes (pi position, halves swapped when rot >= 32)
    int r0, r1, r2;      // word index of B[.][x], B[.][x+1], B[.][x+2] in row 0
    uint32_t is0;        // all ones on the lane that owns a[0][0]
};

template <int ROWL>
__device__ __forceinline__ Col col_init(int L) {
    Col c;
    const int g = L / 5, x = L % 5;
    const int base = L - x;
    c.lm1 = base + (x + 4) % 5;
    c.lp1 = base + (x + 1) % 5;
    if (c.lm1 > 31) c.lm1 = L;
    if (c.lp1 > 31) c.lp1 = L;
#pragma unroll
    for (int y = 0; y < 5; ++y) {
        const int rot = c_rot5[x + 5 * y];
        c.s[y] = rot & 31;
        const int X = y, Y = (2 * x + 3 * y) % 5;
        const int w = (Y * ROWL + g * 5 + X) * 2;
        c.wl[y] = w + (rot >= 32 ? 1 : 0);
        c.wh[y] = w + (rot >= 32 ? 0 : 1);
    }
    c.r0 = (g * 5 + x) * 2;
    c.r1 = (g * 5 + (x + 1) % 5) * 2;
    c.r2 = (g * 5 + (x + 2) % 5) * 2;
    c.is0 = x == 0 ? 0xffffffffu : 0u;
    return c;
}

template <int UNROLL, int ROWL>
__device__ __forceinline__ void f_v21(uint32_t (&lo)[5], uint32_t (&hi)[5], const Col& c, uint32_t* __restrict__ sm) {
    constexpr unsigned long long RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
        0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
        0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    using G = Geo<ROWL>;
#pragma unroll UNROLL
    for (int r = 0; r < 24; ++r) {
        uint32_t* b = sm + (r & 1) * G::BUF_WORDS;
        // theta: parity of the own column, the neighbours' parities by shuffle
        const uint32_t cl = lo[0] ^ lo[1] ^ lo[2] ^ lo[3] ^ lo[4];
        const uint32_t ml = SH(cl, c.lm1), pl = SH(cl, c.lp1);
        const uint32_t ch = hi[0] ^ hi[1] ^ hi[2] ^ hi[3] ^ hi[4];
        const uint32_t mh = SH(ch, c.lm1), ph = SH(ch, c.lp1);
        const uint32_t dl = ml ^ __funnelshift_l(ph, pl, 1);
        const uint32_t dh = mh ^ __funnelshift_l(pl, ph, 1);
        // rho + pi: scatter the rotated lanes to their pi positions
#pragma unroll
        for (int y = 0; y < 5; ++y) {
            const uint32_t tl = lo[y] ^ dl, th = hi[y] ^ dh;
            b[c.wl[y]] = __funnelshift_l(th, tl, c.s[y]);
            b[c.wh[y]] = __funnelshift_l(tl, th, c.s[y]);
        }
        __syncwarp();
        // chi + iota on column x of the new state
#pragma unroll
        for (int y = 0; y < 5; ++y) {
            const uint2 b0 = *reinterpret_cast<const uint2*>(b + c.r0 + y * G::ROW_WORDS);
            const uint2 b1 = *reinterpret_cast<const uint2*>(b + c.r1 + y * G::ROW_WORDS);
            const uint2 b2 = *reinterpret_cast<const uint2*>(b + c.r2 + y * G::ROW_WORDS);
            lo[y] = b0.x ^ (~b1.x & b2.x);
            hi[y] = b0.y ^ (~b1.y & b2.y);
        }
        lo[0] ^= (uint32_t)RC[r] & c.is0;
        hi[0] ^= (uint32_t)(RC[r] >> 32) & c.is0;
    }
}

template <int UNROLL, int ROWL>
__global__ void col_kernel(int perms, unsigned long long* state_out, long long* cycles_out) {
    extern __shared__ uint32_t smem[];
    using G = Geo<ROWL>;
    const int L = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t* sm = smem + w * G::WARP_WORDS;
    const Col c = col_init<ROWL>(L);
    const int g = L / 5, x = L % 5;
    const int warp = blockIdx.x * (blockDim.x >> 5) + w;
    const int chain = warp * 6 + (g < 6 ? g : 0);
    uint32_t lo[5], hi[5];
#pragma unroll
    for (int y = 0; y < 5; ++y) {
        const unsigned long long v = (unsigned long long)(x + 5 * y + 1) * 0x9E3779B97F4A7C15ULL + (unsigned long long)(chain % 6 == 0 ? chain / 6 : 1000000 + chain);
        lo[y] = (uint32_t)v; hi[y] = (uint32_t)(v >> 32);
    }
    const long long t0 = clock64();
    for (int p = 0; p < perms; ++p) f_v21<UNROLL, ROWL>(lo, hi, c, sm);
    const long long t1 = clock64();
    if (warp == 0 && g == 0)
        for (int y = 0; y < 5; ++y) state_out[x + 5 * y] = ((unsigned long long)hi[y] << 32) | lo[y];
    if (L == 0) cycles_out[warp] = t1 - t0;
}

// reference: thread per sponge
__global__ void thread_kernel(int perms, unsigned long long* state_out) {
    unsigned long long a[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) a[k] = (unsigned long long)(k + 1) * 0x9E3779B97F4A7C15ULL;
    for (int p = 0; p < perms; ++p) keccak_f(a);
    if (threadIdx.x == 0 && blockIdx.x == 0)
        for (int k = 0; k < 25; ++k) state_out[k] = a[k];
}

template <int UNROLL, int ROWL>
static void run(const char* name, int sms, int perms, unsigned long long* d_state, long long* d_cyc, const std::vector<unsigned long long>& ref) {
    using G = Geo<ROWL>;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<long long> cyc(1 << 16);
    std::vector<unsigned long long> got(25);
    // {CTAs, warps per CTA}: one warp per SM .. 16 per SM, then the chain counts of the scheme: 1250 chains = 209 warps, 10^4 chains = 1667 warps
    const int shapes[][2] = {{sms, 1}, {sms, 2}, {sms, 4}, {2 * sms, 4}, {3 * sms, 4}, {4 * sms, 4}, {209, 1}, {53, 4}, {1667, 1}, {417, 4}};
    for (auto& sh : shapes) {
        const int smem = sh[1] * G::WARP_WORDS * 4;
        col_kernel<UNROLL, ROWL><<<sh[0], sh[1] * 32, smem>>>(10, d_state, d_cyc);
        CK(cudaEventRecord(e0));
        col_kernel<UNROLL, ROWL><<<sh[0], sh[1] * 32, smem>>>(perms, d_state, d_cyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        const int nw = sh[0] * sh[1];
        CK(cudaMemcpy(cyc.data(), d_cyc, nw * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(got.data(), d_state, 200, cudaMemcpyDeviceToHost));
        bool ok = true;
        for (int k = 0; k < 25; ++k) ok = ok && got[k] == ref[k];
        long long mx = 0; double avg = 0;
        for (int i = 0; i < nw; ++i) { mx = cyc[i] > mx ? cyc[i] : mx; avg += (double)cyc[i]; }
        avg /= nw;
        printf("%-22s %5d CTAs x %d warps (%5d chains, %5.2f warps/SM): cycles/round avg %6.1f max %6.1f  %8.3f ms  %6.3f Gperm/s  %s\n", name, sh[0], sh[1],
               nw * 6, (double)nw / sms, avg / perms / 24.0, (double)mx / perms / 24.0, ms, (double)nw * 6 * perms / (ms * 1e-3) / 1e9, ok ? "ok" : "MISMATCH");
    }
}

int main(int argc, char** argv) {
    const int perms = argc > 1 ? atoi(argv[1]) : 1000;
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    printf("device %s, %d SMs, %d perms per chain\n", pr.name, sms, perms);
    unsigned long long* d_state; long long* d_cyc;
    CK(cudaMalloc(&d_state, 25 * 8));
    CK(cudaMalloc(&d_cyc, 1 << 20));
    std::vector<unsigned long long> ref(25);
    thread_kernel<<<1, 32>>>(perms, d_state);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(ref.data(), d_state, 200, cudaMemcpyDeviceToHost));
    run<24, 35>("V21 x24 row 35", sms, perms, d_state, d_cyc, ref);
    run<24, 36>("V21 x24 row 36", sms, perms, d_state, d_cyc, ref);
    run<24, 37>("V21 x24 row 37", sms, perms, d_state, d_cyc, ref);
    run<8, 35>("V21 x8 row 35", sms, perms, d_state, d_cyc, ref);
    run<2, 35>("V21 x2 row 35", sms, perms, d_state, d_cyc, ref);
    return 0;
}
