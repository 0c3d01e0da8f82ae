// fz_stream.cu -- HBM-streaming NTT-domain kernels: matrix-vector MAC (A.s, A.z), the sign
// mul-add (sk_l * c_hat + sk_r), the hash-weighted aggregation sum and the verify target.
//
// Replaces algebra/matrices.py:108-131 (GeneralMatrix.__mul__, both branches) and :74-91
// (__add__) as used by fusion/fusion.py:369-370 (keygen), :557 (sign), :670-677 (aggregate),
// :706-717 (verify).  The reference re-centres after every multiply and every add, so the
// final value is the canonical centred residue of the exact integer sum; any accumulation
// order is therefore bit-exact as long as the final store is centred.
//
// All of these are bandwidth-bound (4 B of HBM traffic per multiply-accumulate): products go
// into signed 64-bit accumulators (one IMAD.WIDE each) that are folded with 2^31 == 17919
// (mod q) every second term, and reduced once per output coefficient.
// Contract: the "matrix side" operand (A, c_hat, alpha_hat -- always produced by this
// library) is canonical (|x| <= (q-1)/2 < 2^30); the streamed operand may be any int32.
#include "fz_common.cuh"
#include "fz_internal.h"
#include "fz_mac.cuh"

namespace fz {

// mac4 / fold4 / cent4: fz_mac.cuh

// out[b][k] = cent( sum_r A[r][k] * s[b][r][k] ).  A: [ell][d] canonical, s: [batch][ell][d], out: [batch][d].
// thread <-> (b, 4 consecutive k); the ell rows of s[b] are streamed with 4 loads in flight.
// s is addressed as s[b * batch_stride4 + r * row_stride4 + c] (row_stride4 = 0 broadcasts one polynomial to every row).
__global__ void __launch_bounds__(256) matvec_kernel(const int4* __restrict__ A, const int4* __restrict__ s, size_t row_stride4, size_t batch_stride4,
                                                       int4* __restrict__ out, size_t batch, int ell, int d4) {
    const size_t total = batch * (size_t)d4;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int sh = 31 - __clz(d4);  // d4 is a power of two (16 or 64)
        const size_t b = idx >> sh;
        const int c = (int)(idx & (size_t)(d4 - 1));
        const int4* sp = s + b * batch_stride4 + c;
        const int4* ap = A + c;
        long long acc[4] = {0, 0, 0, 0};
        int r = 0;
        for (; r + 4 <= ell; r += 4) {
            int4 s0 = ld_stream(sp + (size_t)(r + 0) * row_stride4), s1 = ld_stream(sp + (size_t)(r + 1) * row_stride4);
            int4 s2 = ld_stream(sp + (size_t)(r + 2) * row_stride4), s3 = ld_stream(sp + (size_t)(r + 3) * row_stride4);
            int4 a0 = __ldg(ap + (size_t)(r + 0) * d4), a1 = __ldg(ap + (size_t)(r + 1) * d4);
            int4 a2 = __ldg(ap + (size_t)(r + 2) * d4), a3 = __ldg(ap + (size_t)(r + 3) * d4);
            mac4(acc, a0, s0); mac4(acc, a1, s1); fold4(acc);
            mac4(acc, a2, s2); mac4(acc, a3, s3); fold4(acc);
        }
        for (; r < ell; ++r) {
            mac4(acc, __ldg(ap + (size_t)r * d4), ld_stream(sp + (size_t)r * row_stride4));
            fold4(acc);
        }
        out[idx] = cent4(acc);
    }
}

// sig[b][r][k] = cent( sk_l[b][r][k] * c_hat[b][k] + sk_r[b][r][k] )   (fusion.py:557)
// grid = (ceil(ell*d4 / 256), keys per launch); 32-bit index arithmetic only (a 64-bit divide per int4 would make this
// write-bound kernel instruction-bound)
__global__ void __launch_bounds__(256) sign_mac_kernel(const int4* __restrict__ sk_l, const int4* __restrict__ sk_r, size_t row_stride4,
                                                         size_t batch_stride4, const int4* __restrict__ c_hat, int4* __restrict__ sig, size_t batch0,
                                                         int ell, int d4) {
    const int per_key = ell * d4;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= per_key) return;
    const int sh = 31 - __clz(d4);
    const int row = i >> sh, c = i & (d4 - 1);
    const size_t b = batch0 + blockIdx.y;
    const size_t si = b * batch_stride4 + (size_t)row * row_stride4 + c;
    const int4 l = ld_stream(sk_l + si), r = ld_stream(sk_r + si), ch = __ldg(c_hat + b * d4 + c);
    int4 o;
    o.x = cent_i64((long long)l.x * ch.x + r.x);
    o.y = cent_i64((long long)l.y * ch.y + r.y);
    o.z = cent_i64((long long)l.z * ch.z + r.z);
    o.w = cent_i64((long long)l.w * ch.w + r.w);
    st_stream(sig + b * (size_t)per_key + i, o);
}

// Keys generated from a seed have ell IDENTICAL rows on each side (fusion.py:159-168: every entry is sampled with the same
// seed), so their signature has ell identical rows too: one multiply-add per coefficient, ell streamed
// stores.  sk_hat: [batch][2][d] (the one distinct row of sk_l and of sk_r), sig: [batch][ell][d].  Write-bound.
__global__ void __launch_bounds__(256) sign_rows_kernel(const int4* __restrict__ sk_hat, const int4* __restrict__ c_hat, int4* __restrict__ sig,
                                                          size_t batch, int ell, int d4) {
    const size_t total = batch * (size_t)d4;
    const int sh = 31 - __clz(d4);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t b = idx >> sh;
        const int c = (int)(idx & (size_t)(d4 - 1));
        const int4 l = __ldg(sk_hat + b * 2 * d4 + c), r = __ldg(sk_hat + b * 2 * d4 + d4 + c), ch = __ldg(c_hat + idx);
        int4 o;
        o.x = cent_i64((long long)l.x * ch.x + r.x);
        o.y = cent_i64((long long)l.y * ch.y + r.y);
        o.z = cent_i64((long long)l.z * ch.z + r.z);
        o.w = cent_i64((long long)l.w * ch.w + r.w);
        int4* dst = sig + b * (size_t)ell * d4 + c;
#pragma unroll 4
        for (int row = 0; row < ell; ++row) st_stream(dst + (size_t)row * d4, o);
    }
}

// asum[k] = cent( sum_r A[r][k] ): with identical rows s, A . s = asum o s   (matrices.py:121-131 for a broadcast column)
__global__ void __launch_bounds__(64) asum_kernel(const int4* __restrict__ A, int4* __restrict__ asum, int ell, int d4) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d4) return;
    long long acc[4] = {0, 0, 0, 0};
    for (int r = 0; r < ell; ++r) {
        const int4 a = __ldg(A + (size_t)r * d4 + c);
        acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
    }
    asum[c] = cent4(acc);
}

// out[b][k] = cent( asum[k] * s[b * batch_stride + k] )
__global__ void __launch_bounds__(256) mul_bcast_kernel(const int4* __restrict__ asum, const int4* __restrict__ s, size_t batch_stride4,
                                                          int4* __restrict__ out, size_t batch, int d4) {
    const size_t total = batch * (size_t)d4;
    const int sh = 31 - __clz(d4);
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t b = idx >> sh;
        const int c = (int)(idx & (size_t)(d4 - 1));
        const int4 a = __ldg(asum + c), v = ld_stream(s + b * batch_stride4 + c);
        int4 o;
        o.x = cent_i64((long long)a.x * v.x); o.y = cent_i64((long long)a.y * v.y);
        o.z = cent_i64((long long)a.z * v.z); o.w = cent_i64((long long)a.w * v.w);
        out[idx] = o;
    }
}

static int stream_grid(const fz_ctx* ctx, size_t items, int threads) {
    size_t blocks = (items + threads - 1) / threads;
    const size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    return blocks < 1 ? 1 : (int)blocks;
}

int launch_matvec(const fz_ctx* ctx, const int32_t* A, const int32_t* s, size_t s_row_stride, size_t s_batch_stride, int32_t* out, size_t batch,
                  cudaStream_t st) {
    if (batch == 0) return 0;
    const int d4 = ctx->d / 4;
    FZ_LAUNCH(matvec_kernel, stream_grid(ctx, batch * d4, 256), 256, 0, st, (const int4*)A, (const int4*)s, s_row_stride / 4, s_batch_stride / 4, (int4*)out, batch,
                                                                     ctx->ell, d4);
    return check(cudaGetLastError(), "matvec launch");
}

int launch_sign_mac(const fz_ctx* ctx, const int32_t* sk_l, const int32_t* sk_r, size_t sk_row_stride, size_t sk_batch_stride, const int32_t* c_hat,
                    int32_t* sig, size_t batch, cudaStream_t st) {
    if (batch == 0) return 0;
    const int d4 = ctx->d / 4;
    const unsigned gx = (unsigned)((ctx->ell * d4 + 255) / 256);
    for (size_t b0 = 0; b0 < batch; b0 += 65535) {   // gridDim.y limit
        const size_t nb = batch - b0 < 65535 ? batch - b0 : 65535;
        FZ_LAUNCH(sign_mac_kernel, dim3(gx, (unsigned)nb), 256, 0, st, (const int4*)sk_l, (const int4*)sk_r, sk_row_stride / 4, sk_batch_stride / 4, (const int4*)c_hat,
                                                                 (int4*)sig, b0, ctx->ell, d4);
        if (check(cudaGetLastError(), "sign_mac launch")) return -2;
    }
    return 0;
}

int launch_sign_rows(const fz_ctx* ctx, const int32_t* sk_hat, const int32_t* c_hat, int32_t* sig, size_t batch, cudaStream_t st) {
    if (batch == 0) return 0;
    const int d4 = ctx->d / 4;
    FZ_LAUNCH(sign_rows_kernel, stream_grid(ctx, batch * d4, 256), 256, 0, st, (const int4*)sk_hat, (const int4*)c_hat, (int4*)sig, batch, ctx->ell, d4);
    return check(cudaGetLastError(), "sign_rows launch");
}

// vk[b] = A . s[b] for s with ell identical rows: asum (d int32 of scratch) = column sums of A, then one pointwise product
int launch_matvec_rows(const fz_ctx* ctx, const int32_t* A, int32_t* asum, const int32_t* s, size_t s_batch_stride, int32_t* out, size_t batch,
                       cudaStream_t st) {
    if (batch == 0) return 0;
    const int d4 = ctx->d / 4;
    FZ_LAUNCH(asum_kernel, (d4 + 63) / 64, 64, 0, st, (const int4*)A, (int4*)asum, ctx->ell, d4);
    if (check(cudaGetLastError(), "asum launch")) return -2;
    FZ_LAUNCH(mul_bcast_kernel, stream_grid(ctx, batch * d4, 256), 256, 0, st, (const int4*)asum, (const int4*)s, s_batch_stride / 4, (int4*)out, batch, d4);
    return check(cudaGetLastError(), "mul_bcast launch");
}

}  // namespace fz
