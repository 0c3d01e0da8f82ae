// fz_scheme.cu -- aggregate / verify kernels and the sort-by-str(vk) step.
//
//   * sort:      fusion.py:661-663, :693  sorted(..., key=lambda x: str(x[0])) -- Python str order of
//                the verification-key repr, reproduced WITHOUT materialising the text: two reprs
//                first differ inside one coefficient's decimal digits, and the byte after a shorter
//                number is ',' (or ']' for the last coefficient of a polynomial);
//   * aggregate: fusion.py:670-677  sum_i sig_i * alpha_i  (matrices.py:109-114, 74-91), streaming,
//                split over signers so that one capacity-sized aggregate still fills the chip;
//   * verify:    fusion.py:706-714 target, :718-721 comparison, :722-725 infinity norm.
#include "fz_common.cuh"
#include "fz_internal.h"
#include "fz_repr_order.cuh"

namespace fz {

// ---------------------------------------------------------------------------------------- sort
// repr_key / repr_cmp: fz_repr_order.cuh
// Rank sort, one CTA column per aggregate.  sel[p] = key index of position p (NULL: identity).
// perm[agg_off[j] + rank] = p  (sorted position -> original position), stable.
__global__ void __launch_bounds__(256) sort_kernel(const int32_t* __restrict__ vk_l, const int32_t* __restrict__ vk_r, const uint32_t* __restrict__ sel,
                                                    const unsigned long long* __restrict__ agg_off, size_t j0, int d, uint32_t* __restrict__ perm) {
    FZ_DYN_SHARED unsigned long long skeys[];
    const size_t j = j0 + blockIdx.y;
    const size_t lo = agg_off[j], hi = agg_off[j + 1];
    const int n = (int)(hi - lo);
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const size_t ki = sel ? sel[lo + t] : lo + t;
        skeys[t] = repr_key(vk_l[ki * d], d == 1);
    }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long mine = skeys[i];
    const size_t ki = sel ? sel[lo + i] : lo + i;
    int rank = 0;
    for (int t = 0; t < n; ++t) {
        const unsigned long long other = skeys[t];
        if (other < mine) ++rank;
        else if (other == mine && t != i) {
            const size_t kt = sel ? sel[lo + t] : lo + t;
            const int c = repr_cmp(vk_l + kt * d, vk_r + kt * d, vk_l + ki * d, vk_r + ki * d, d);
            if (c < 0 || (c == 0 && t < i)) ++rank;
        }
    }
    perm[lo + rank] = (uint32_t)(lo + i);
}

// gather per sorted position: vk_l_s[p] = vk_l[sel[perm[p]]] etc.  One warp per position.
__global__ void __launch_bounds__(128) gather_sorted_kernel(const int32_t* __restrict__ vk_l, const int32_t* __restrict__ vk_r,
                                                              const uint8_t* __restrict__ digests, const uint32_t* __restrict__ sel,
                                                              const uint32_t* __restrict__ perm, size_t n_total, int d, int32_t* __restrict__ vk_l_s,
                                                              int32_t* __restrict__ vk_r_s, uint8_t* __restrict__ dg_s) {
    const int lane = threadIdx.x & 31;
    const size_t p = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n_total) return;
    const size_t po = perm ? perm[p] : p;
    const size_t ki = sel ? sel[po] : po;
    const int4* sl = reinterpret_cast<const int4*>(vk_l + ki * d);
    const int4* sr = reinterpret_cast<const int4*>(vk_r + ki * d);
    int4* dl = reinterpret_cast<int4*>(vk_l_s + p * d);
    int4* dr = reinterpret_cast<int4*>(vk_r_s + p * d);
    for (int c = lane; c < d / 4; c += 32) { dl[c] = __ldg(sl + c); dr[c] = __ldg(sr + c); }
    if (lane < 8) reinterpret_cast<uint32_t*>(dg_s + p * 32)[lane] = reinterpret_cast<const uint32_t*>(digests + po * 32)[lane];
}

// ---------------------------------------------------------------------------------------- aggregate
// part[j][s][c] = cent( sum_{p in split s of aggregate j} sig[perm[p]][c] * alpha[p][c mod d] ),  c over ell*d
// grid = (col blocks, split, n_aggs); thread <-> 4 consecutive coefficients.
__global__ void __launch_bounds__(256) aggregate_kernel(const int4* __restrict__ sigs, const uint32_t* __restrict__ perm, const int4* __restrict__ alpha,
                                                          const unsigned long long* __restrict__ agg_off, size_t j0, int ell, int d4, int split,
                                                          int4* __restrict__ part) {
    const int cols = ell * d4;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    const int s = blockIdx.y;
    const size_t j = j0 + blockIdx.z;
    const size_t lo = agg_off[j], hi = agg_off[j + 1];
    const size_t n = hi - lo, per = (n + split - 1) / split;
    size_t p0 = lo + (size_t)s * per, p1 = p0 + per;
    if (p1 > hi) p1 = hi;
    const int ca = c & (d4 - 1);  // d4 is a power of two
    long long acc[4] = {0, 0, 0, 0};
    size_t p = p0;
    for (; p + 4 <= p1; p += 4) {
        const size_t i0 = perm ? perm[p] : p, i1 = perm ? perm[p + 1] : p + 1, i2 = perm ? perm[p + 2] : p + 2, i3 = perm ? perm[p + 3] : p + 3;
        const int4 z0 = ld_stream(sigs + i0 * cols + c), z1 = ld_stream(sigs + i1 * cols + c);
        const int4 z2 = ld_stream(sigs + i2 * cols + c), z3 = ld_stream(sigs + i3 * cols + c);
        const int4 a0 = __ldg(alpha + p * d4 + ca), a1 = __ldg(alpha + (p + 1) * d4 + ca);
        const int4 a2 = __ldg(alpha + (p + 2) * d4 + ca), a3 = __ldg(alpha + (p + 3) * d4 + ca);
        acc[0] += (long long)z0.x * a0.x + (long long)z1.x * a1.x; acc[1] += (long long)z0.y * a0.y + (long long)z1.y * a1.y;
        acc[2] += (long long)z0.z * a0.z + (long long)z1.z * a1.z; acc[3] += (long long)z0.w * a0.w + (long long)z1.w * a1.w;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = fold_i64(acc[k]);
        acc[0] += (long long)z2.x * a2.x + (long long)z3.x * a3.x; acc[1] += (long long)z2.y * a2.y + (long long)z3.y * a3.y;
        acc[2] += (long long)z2.z * a2.z + (long long)z3.z * a3.z; acc[3] += (long long)z2.w * a2.w + (long long)z3.w * a3.w;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[k] = fold_i64(acc[k]);
    }
    for (; p < p1; ++p) {
        const size_t i0 = perm ? perm[p] : p;
        const int4 z0 = ld_stream(sigs + i0 * cols + c);
        const int4 a0 = __ldg(alpha + p * d4 + ca);
        acc[0] = fold_i64(acc[0] + (long long)z0.x * a0.x); acc[1] = fold_i64(acc[1] + (long long)z0.y * a0.y);
        acc[2] = fold_i64(acc[2] + (long long)z0.z * a0.z); acc[3] = fold_i64(acc[3] + (long long)z0.w * a0.w);
    }
    part[(j * split + s) * (size_t)cols + c] = make_int4(cent_i64(acc[0]), cent_i64(acc[1]), cent_i64(acc[2]), cent_i64(acc[3]));
}

// out[j][c] = cent( sum_s part[j][s][c] ); also the fold used for partial aggregates gathered from other ranks
__global__ void __launch_bounds__(256) fold_partials_kernel(const int4* __restrict__ part, int4* __restrict__ out, size_t n_aggs, int cols, int split) {
    const size_t total = n_aggs * (size_t)cols;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t j = idx / cols;
        const int c = (int)(idx % cols);
        long long a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int s = 0; s < split; ++s) {
            const int4 v = part[(j * split + s) * (size_t)cols + c];
            a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
        }
        out[idx] = make_int4(cent_i64(a0), cent_i64(a1), cent_i64(a2), cent_i64(a3));
    }
}

// ---------------------------------------------------------------------------------------- verify
// target[j][k] = cent( sum_p cent(vk_l_s[p][k] * c_hat[p][k] + vk_r_s[p][k]) * alpha[p][k] )      (fusion.py:706-714)
// One CTA per aggregate: 4 signer groups x d/4 column threads, reduced through shared memory.
__global__ void __launch_bounds__(256) verify_target_kernel(const int4* __restrict__ vk_l_s, const int4* __restrict__ vk_r_s, const int4* __restrict__ c_hat,
                                                              const int4* __restrict__ alpha, const unsigned long long* __restrict__ agg_off, int d4,
                                                              int4* __restrict__ target) {
    __shared__ long long red[256 * 4];
    const size_t j = blockIdx.x;
    const size_t lo = agg_off[j], hi = agg_off[j + 1];
    const int groups = blockDim.x / d4;
    const int g = threadIdx.x / d4, c = threadIdx.x % d4;
    long long acc[4] = {0, 0, 0, 0};
    if (g < groups) {
        for (size_t p = lo + g; p < hi; p += groups) {
            const int4 l = __ldg(vk_l_s + p * d4 + c), r = __ldg(vk_r_s + p * d4 + c), ch = __ldg(c_hat + p * d4 + c), al = __ldg(alpha + p * d4 + c);
            const long long t0 = cent_i64((long long)l.x * ch.x + r.x), t1 = cent_i64((long long)l.y * ch.y + r.y);
            const long long t2 = cent_i64((long long)l.z * ch.z + r.z), t3 = cent_i64((long long)l.w * ch.w + r.w);
            acc[0] = fold_i64(acc[0] + t0 * al.x); acc[1] = fold_i64(acc[1] + t1 * al.y);
            acc[2] = fold_i64(acc[2] + t2 * al.z); acc[3] = fold_i64(acc[3] + t3 * al.w);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) red[threadIdx.x * 4 + k] = acc[k];
    __syncthreads();
    if (threadIdx.x < d4) {
        long long s[4] = {0, 0, 0, 0};
        for (int gg = 0; gg < groups; ++gg)
#pragma unroll
            for (int k = 0; k < 4; ++k) s[k] += red[(gg * d4 + threadIdx.x) * 4 + k];
        target[j * d4 + threadIdx.x] = make_int4(cent_i64(s[0]), cent_i64(s[1]), cent_i64(s[2]), cent_i64(s[3]));
    }
}

// codes[j] = 3 if target != observed (mod q), else 4 if max|coef| > beta_vf, else 0.   (fusion.py:718-728)
// target / observed: [n_aggs][d] canonical; coef: [n_aggs][ell][d] = INTT of the aggregate's rows.
__global__ void __launch_bounds__(256) verify_finalize_kernel(const int32_t* __restrict__ target, const int32_t* __restrict__ observed,
                                                                const int32_t* __restrict__ coef, int d, int ell, long long beta_vf,
                                                                uint8_t* __restrict__ codes) {
    __shared__ int s_mismatch;
    __shared__ unsigned int s_max;
    const size_t j = blockIdx.x;
    if (threadIdx.x == 0) { s_mismatch = 0; s_max = 0; }
    __syncthreads();
    int mism = 0;
    for (int k = threadIdx.x; k < d; k += blockDim.x) {
        long long diff = (long long)target[j * d + k] - (long long)observed[j * d + k];
        if (diff % (long long)Q != 0) mism = 1;
    }
    unsigned int mx = 0;
    const int32_t* cj = coef + j * (size_t)ell * d;
    for (int k = threadIdx.x; k < ell * d; k += blockDim.x) {
        const int32_t v = cj[k];
        const unsigned int a = v < 0 ? (unsigned int)(-(long long)v) : (unsigned int)v;
        mx = a > mx ? a : mx;
    }
    if (mism) atomicOr(&s_mismatch, 1);
    atomicMax(&s_max, mx);
    __syncthreads();
    if (threadIdx.x == 0) codes[j] = s_mismatch ? 3 : ((long long)s_max > beta_vf ? 4 : 0);
}

// ---------------------------------------------------------------------------------------- launchers
int launch_sort(const fz_ctx* ctx, const int32_t* vk_l, const int32_t* vk_r, const uint32_t* sel, const unsigned long long* agg_off, size_t n_aggs,
                size_t max_keys, uint32_t* perm, cudaStream_t st) {
    if (n_aggs == 0 || max_keys == 0) return 0;
    const size_t smem = max_keys * sizeof(unsigned long long);
    if (smem > (size_t)ctx->sort_smem_max) {   // the opt-in was made for this context's device in scheme_init
        set_error("fz_sort_keys: %zu keys in one aggregate need %zu bytes of shared memory, the device allows %d", max_keys, smem, ctx->sort_smem_max);
        return -1;
    }
    for (size_t j0 = 0; j0 < n_aggs; j0 += 65535) {   // gridDim.y is limited to 65,535
        const size_t nj = n_aggs - j0 < 65535 ? n_aggs - j0 : 65535;
        dim3 grid((unsigned)((max_keys + 255) / 256), (unsigned)nj);
        FZ_LAUNCH(sort_kernel, grid, 256, smem, st, vk_l, vk_r, sel, agg_off, j0, ctx->d, perm);
        if (check(cudaGetLastError(), "sort launch")) return -2;
    }
    return 0;
}

// per-device kernel attributes of this translation unit (called from fz_init with the context's device current)
int scheme_init(fz_ctx* ctx) {
    int optin = 0;
    if (check(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device), "shared-memory opt-in query")) return -2;
    if (check(cudaFuncSetAttribute(sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin), "sort smem attr")) return -2;
    ctx->sort_smem_max = optin;
    return 0;
}

int launch_gather_sorted(const fz_ctx* ctx, const int32_t* vk_l, const int32_t* vk_r, const uint8_t* digests, const uint32_t* sel, const uint32_t* perm,
                         size_t n_total, int32_t* vk_l_s, int32_t* vk_r_s, uint8_t* dg_s, cudaStream_t st) {
    if (n_total == 0) return 0;
    FZ_LAUNCH(gather_sorted_kernel, (int)((n_total + 3) / 4), 128, 0, st, vk_l, vk_r, digests, sel, perm, n_total, ctx->d, vk_l_s, vk_r_s, dg_s);
    return check(cudaGetLastError(), "gather_sorted launch");
}

int aggregate_split(const fz_ctx* ctx, size_t n_aggs, size_t max_keys) {
    const int cols = ctx->ell * ctx->d / 4;
    const size_t col_blocks = (cols + 255) / 256;
    size_t want = ((size_t)ctx->sm_count * 4 + col_blocks * n_aggs - 1) / (col_blocks * n_aggs);
    if (want < 1) want = 1;
    if (want > 64) want = 64;
    if (want > max_keys) want = max_keys ? max_keys : 1;
    return (int)want;
}

int launch_aggregate(const fz_ctx* ctx, const int32_t* sigs, const uint32_t* perm, const int32_t* alpha, const unsigned long long* agg_off, size_t n_aggs,
                     int split, int32_t* part, int32_t* out, cudaStream_t st) {
    if (n_aggs == 0) return 0;
    const int d4 = ctx->d / 4, cols = ctx->ell * d4;
    for (size_t j0 = 0; j0 < n_aggs; j0 += 65535) {   // gridDim.z is limited to 65,535
        const size_t nj = n_aggs - j0 < 65535 ? n_aggs - j0 : 65535;
        dim3 grid((unsigned)((cols + 255) / 256), (unsigned)split, (unsigned)nj);
        FZ_LAUNCH(aggregate_kernel, grid, 256, 0, st, (const int4*)sigs, perm, (const int4*)alpha, agg_off, j0, ctx->ell, d4, split, (int4*)(split == 1 ? out : part));
        if (check(cudaGetLastError(), "aggregate launch")) return -2;
    }
    if (split > 1) {
        size_t blocks = (n_aggs * cols + 255) / 256;
        const size_t cap = (size_t)ctx->sm_count * 8;
        if (blocks > cap) blocks = cap;
        FZ_LAUNCH(fold_partials_kernel, (int)blocks, 256, 0, st, (const int4*)part, (int4*)out, n_aggs, cols, split);
        return check(cudaGetLastError(), "fold_partials launch");
    }
    return 0;
}

int launch_fold_partials(const fz_ctx* ctx, const int32_t* part, int32_t* out, size_t n_aggs, int n_parts, cudaStream_t st) {
    if (n_aggs == 0) return 0;
    const int cols = ctx->ell * ctx->d / 4;
    size_t blocks = (n_aggs * cols + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    FZ_LAUNCH(fold_partials_kernel, (int)blocks, 256, 0, st, (const int4*)part, (int4*)out, n_aggs, cols, n_parts);
    return check(cudaGetLastError(), "fold_partials launch");
}

int launch_verify_target(const fz_ctx* ctx, const int32_t* vk_l_s, const int32_t* vk_r_s, const int32_t* c_hat, const int32_t* alpha,
                         const unsigned long long* agg_off, size_t n_aggs, int32_t* target, cudaStream_t st) {
    if (n_aggs == 0) return 0;
    FZ_LAUNCH(verify_target_kernel, (int)n_aggs, 256, 0, st, (const int4*)vk_l_s, (const int4*)vk_r_s, (const int4*)c_hat, (const int4*)alpha, agg_off, ctx->d / 4,
                                                       (int4*)target);
    return check(cudaGetLastError(), "verify_target launch");
}

int launch_verify_finalize(const fz_ctx* ctx, const int32_t* target, const int32_t* observed, const int32_t* coef, size_t n_aggs, uint8_t* codes,
                           cudaStream_t st) {
    if (n_aggs == 0) return 0;
    FZ_LAUNCH(verify_finalize_kernel, (int)n_aggs, 256, 0, st, target, observed, coef, ctx->d, ctx->ell, ctx->beta_vf, codes);
    return check(cudaGetLastError(), "verify_finalize launch");
}

}  // namespace fz
