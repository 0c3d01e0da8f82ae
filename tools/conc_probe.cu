// concurrency probe: does a small kernel on stream B start while a long latency-bound kernel (1250 CTAs x 32 threads) runs on stream A?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void longk(long long iters, float* out) {
    float x = threadIdx.x;
    unsigned v = threadIdx.x;
    for (long long i = 0; i < iters; ++i) { v = __shfl_sync(0xffffffffu, v, (threadIdx.x * 7 + 3) & 31) + 1; }
    if (v == 0x12345) out[0] = x;
}
__global__ void smallk(float* out, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = 1.0f; }
int main() {
    cudaStream_t A, B; cudaStreamCreateWithFlags(&A, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&B, cudaStreamNonBlocking);
    float* o; cudaMalloc(&o, 1 << 26);
    cudaEvent_t a0, a1, b0, b1; cudaEventCreate(&a0); cudaEventCreate(&a1); cudaEventCreate(&b0); cudaEventCreate(&b1);
    for (int ctas : {1250, 1184, 592, 148}) for (int big : {0, 1}) {
        longk<<<ctas, 32, 0, A>>>(1000, o); smallk<<<64, 256, 0, B>>>(o, 1); cudaDeviceSynchronize();
        cudaEventRecord(a0, A); longk<<<ctas, 32, 0, A>>>(1500000, o); cudaEventRecord(a1, A);
        cudaEventRecord(b0, B);
        if (big) smallk<<<65536, 128, 0, B>>>(o, 65536 * 128); else smallk<<<64, 256, 0, B>>>(o, 64 * 256);
        cudaEventRecord(b1, B);
        cudaDeviceSynchronize();
        float ta, tb0, tb1; cudaEventElapsedTime(&ta, a0, a1); cudaEventElapsedTime(&tb0, a0, b0); cudaEventElapsedTime(&tb1, a0, b1);
        printf("long kernel %4d CTAs: %.1f ms; %s kernel on the other stream finished at %.2f ms\n", ctas, ta, big ? "65536-CTA" : "64-CTA", tb1);
    }
    return 0;
}
