// fz_api.cu -- the extern "C" surface declared in include/fusion_b200.h.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/fusion_b200.h"
#include "fz_internal.h"

namespace fz {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return -2;
}

static int ceil_div(int a, int b) { return (a + b - 1) / b; }
static int ilog2(int n) {
    int k = 0;
    while ((1 << k) < n) ++k;
    return k;
}

}  // namespace fz

using namespace fz;

#define FZ_REQUIRE(cond, msg)            \
    do {                                 \
        if (!(cond)) {                   \
            fz::set_error("%s", msg);    \
            return -1;                   \
        }                                \
    } while (0)

extern "C" {

int fz_version(void) { return 100; }

const char* fz_last_error(void) { return fz::g_err; }

int fz_init(int device, int secpar, fz_ctx** out) {
    FZ_REQUIRE(out != nullptr, "fz_init: out is NULL");
    FZ_REQUIRE(secpar == 128 || secpar == 256, "fz_init: secpar must be 128 or 256");
    int ndev = 0;
    if (check(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount")) return -2;
    FZ_REQUIRE(device >= 0 && device < ndev, "fz_init: no such CUDA device (libfusion_b200 has no CPU fallback)");
    if (check(cudaSetDevice(device), "cudaSetDevice")) return -2;
    fz_ctx* c = new (std::nothrow) fz_ctx();
    FZ_REQUIRE(c != nullptr, "fz_init: out of host memory");
    memset(c, 0, sizeof(*c));
    c->device = device;
    c->secpar = secpar;
    // fusion/fusion.py:16-141
    if (secpar == 128) {
        c->d = 64; c->logd = 6; c->ell = 195; c->capacity = 1796; c->omega_ch = 27; c->omega_ag = 35;
        c->root = 23584283u; c->inv_root = 540632852u;
    } else {
        c->d = 256; c->logd = 8; c->ell = 83; c->capacity = 2818; c->omega_ch = 60; c->omega_ag = 60;
        c->root = 3337519u; c->inv_root = 1978410468u;
    }
    c->bytes_coef = ceil_div(1 + secpar, 8);             // fusion.py:123-137 with beta = 1
    c->bytes_idx = ceil_div(ilog2(c->d) + secpar, 8);    // fusion.py:138-141
    c->n_ch = ceil_div(c->omega_ch, 8) + c->bytes_coef * c->omega_ch + c->d * c->bytes_idx;               // fusion.py:520-524
    c->n_ch_used = ceil_div(c->omega_ch, 8) + c->bytes_coef * c->omega_ch + (c->d - 1 - c->omega_ch) * c->bytes_idx;  // fusion.py:436-438, 473
    c->n_ag = ceil_div(c->omega_ag, 8) + (c->bytes_coef + c->bytes_idx) * c->omega_ag;                    // fusion.py:583-585
    {
        const int ch_bd = secpar == 128 ? 3 : 1, ag_bd = secpar == 128 ? 2 : 1;  // fusion.py:32-35, 57-68
        const long long inter = 52LL * (1 + (long long)(c->d < c->omega_ch ? c->d : c->omega_ch) * ch_bd);
        c->beta_vf = (long long)c->capacity * (c->d < c->omega_ag ? c->d : c->omega_ag) * ag_bd * inter;
    }
    cudaDeviceProp prop;
    if (check(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) { delete c; return -2; }
    c->sm_count = prop.multiProcessorCount;
    if (ntt_init_tables(c)) { delete c; return -2; }
    *out = c;
    return 0;
}

int fz_destroy(fz_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    if (ctx->scratch) cudaFree(ctx->scratch);
    delete ctx;
    return 0;
}

long long fz_param(const fz_ctx* c, const char* name) {
    if (!c || !name) return -1;
#define P(n) if (!strcmp(name, #n)) return (long long)c->n;
    P(secpar) P(d) P(logd) P(ell) P(capacity) P(omega_ch) P(omega_ag) P(bytes_coef) P(bytes_idx)
    P(n_ch) P(n_ch_used) P(n_ag) P(beta_vf) P(sm_count) P(root) P(inv_root) P(device)
#undef P
    if (!strcmp(name, "modulus")) return 2147465729LL;
    return -1;
}

int fz_ntt_fwd(fz_ctx* ctx, const int32_t* in, int32_t* out, size_t n_polys, void* stream) {
    FZ_REQUIRE(ctx && (n_polys == 0 || (in && out)), "fz_ntt_fwd: NULL argument");
    return launch_ntt(ctx, 0, in, out, n_polys, (cudaStream_t)stream);
}

int fz_ntt_inv(fz_ctx* ctx, const int32_t* in, int32_t* out, size_t n_polys, void* stream) {
    FZ_REQUIRE(ctx && (n_polys == 0 || (in && out)), "fz_ntt_inv: NULL argument");
    return launch_ntt(ctx, 1, in, out, n_polys, (cudaStream_t)stream);
}

int fz_pointwise_mul(fz_ctx* ctx, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, void* stream) {
    FZ_REQUIRE(ctx && (n_polys == 0 || (a && b && out)), "fz_pointwise_mul: NULL argument");
    return launch_pointwise(ctx, a, b, nullptr, out, n_polys * ctx->d, (cudaStream_t)stream);
}

int fz_pointwise_muladd(fz_ctx* ctx, const int32_t* a, const int32_t* b, const int32_t* c, int32_t* out, size_t n_polys, void* stream) {
    FZ_REQUIRE(ctx && (n_polys == 0 || (a && b && c && out)), "fz_pointwise_muladd: NULL argument");
    return launch_pointwise(ctx, a, b, c, out, n_polys * ctx->d, (cudaStream_t)stream);
}

int fz_polymul(fz_ctx* ctx, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, void* stream) {
    FZ_REQUIRE(ctx && (n_polys == 0 || (a && b && out)), "fz_polymul: NULL argument");
    return launch_polymul(ctx, a, b, out, n_polys, (cudaStream_t)stream);
}

int fz_matvec(fz_ctx* ctx, const int32_t* A, const int32_t* s, int32_t* out, size_t batch, void* stream) {
    FZ_REQUIRE(ctx && (batch == 0 || (A && s && out)), "fz_matvec: NULL argument");
    return launch_matvec(ctx, A, s, out, batch, (cudaStream_t)stream);
}

int fz_sign_mac(fz_ctx* ctx, const int32_t* sk_l, const int32_t* sk_r, const int32_t* c_hat, int32_t* sig, size_t batch, void* stream) {
    FZ_REQUIRE(ctx && (batch == 0 || (sk_l && sk_r && c_hat && sig)), "fz_sign_mac: NULL argument");
    return launch_sign_mac(ctx, sk_l, sk_r, c_hat, sig, batch, (cudaStream_t)stream);
}

// Chunked H2D -> kernels -> D2H pipeline on internal streams (both copy engines run
// concurrently with the kernels of the neighbouring chunks).  mode 0: forward NTT; 1: inverse NTT;
// 2: forward NTT -> pointwise product with the device-resident b_hat -> inverse NTT.
static int host_pipeline(fz_ctx* ctx, int mode, const int32_t* host_in, const int32_t* dev_bhat, int32_t* host_out, size_t n_polys) {
    if (n_polys == 0) return 0;
    if (check(cudaSetDevice(ctx->device), "cudaSetDevice")) return -2;
    const size_t poly_bytes = (size_t)ctx->d * sizeof(int32_t);
    const size_t chunk = ((size_t)32 << 20) / poly_bytes;  // 32 MiB per chunk
    const int NBUF = 4;
    const size_t need = NBUF * chunk * poly_bytes;
    if (ctx->scratch_bytes < need) {
        if (ctx->scratch) cudaFree(ctx->scratch);
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        if (check(cudaMalloc(&ctx->scratch, need), "host pipeline: cudaMalloc")) return -2;
        ctx->scratch_bytes = need;
    }
    cudaStream_t st[NBUF];
    for (int i = 0; i < NBUF; ++i)
        if (check(cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking), "cudaStreamCreate")) return -2;
    int rc = 0;
    size_t done = 0;
    for (int it = 0; done < n_polys && !rc; ++it) {
        const size_t n = n_polys - done < chunk ? n_polys - done : chunk;
        const int bi = it % NBUF;
        int32_t* buf = (int32_t*)((char*)ctx->scratch + (size_t)bi * chunk * poly_bytes);
        rc = check(cudaMemcpyAsync(buf, host_in + done * ctx->d, n * poly_bytes, cudaMemcpyHostToDevice, st[bi]), "H2D");
        if (!rc && mode == 2) {
            rc = launch_ntt(ctx, 0, buf, buf, n, st[bi]);
            if (!rc) rc = launch_pointwise(ctx, buf, dev_bhat + done * ctx->d, nullptr, buf, n * ctx->d, st[bi]);
            if (!rc) rc = launch_ntt(ctx, 1, buf, buf, n, st[bi]);
        } else if (!rc) {
            rc = launch_ntt(ctx, mode, buf, buf, n, st[bi]);
        }
        if (!rc) rc = check(cudaMemcpyAsync(host_out + done * ctx->d, buf, n * poly_bytes, cudaMemcpyDeviceToHost, st[bi]), "D2H");
        done += n;
    }
    for (int i = 0; i < NBUF; ++i) {
        cudaError_t e = cudaStreamSynchronize(st[i]);
        if (!rc) rc = check(e, "host pipeline: sync");
        cudaStreamDestroy(st[i]);
    }
    return rc;
}

int fz_ntt_host(fz_ctx* ctx, int inverse, const int32_t* host_in, int32_t* host_out, size_t n_polys) {
    FZ_REQUIRE(ctx && (n_polys == 0 || (host_in && host_out)), "fz_ntt_host: NULL argument");
    return host_pipeline(ctx, inverse ? 1 : 0, host_in, nullptr, host_out, n_polys);
}

int fz_ntt_pass_host(fz_ctx* ctx, const int32_t* host_a, const int32_t* dev_bhat, int32_t* host_out, size_t n_polys) {
    FZ_REQUIRE(ctx && (n_polys == 0 || (host_a && dev_bhat && host_out)), "fz_ntt_pass_host: NULL argument");
    return host_pipeline(ctx, 2, host_a, dev_bhat, host_out, n_polys);
}

}  // extern "C"
