// fz_api.cu -- the extern "C" surface declared in include/fusion_b200.h.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/fusion_b200.h"
#include <atomic>
#include "fz_internal.h"

namespace fz {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned> g_attr_done[64];

bool attr_needed(const fz_ctx* ctx, int slot) {
    const int dev = ctx && ctx->device >= 0 && ctx->device < 64 ? ctx->device : 0;
    const unsigned bit = 1u << slot;
    return (g_attr_done[dev].fetch_or(bit) & bit) == 0;
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

// workspace carving (caller-allocated device scratch; nothing is allocated per call)
namespace {
struct Carver {
    char* base; size_t cap, used; bool ok;
    Carver(void* p, size_t bytes) : base((char*)p), cap(bytes), used(0), ok(true) {}
    template <typename T> T* take(size_t count) {
        const size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        T* r = (T*)(base + used);
        used += bytes;
        if (used > cap) ok = false;
        return r;
    }
};
constexpr size_t M1_STRIDE = 7424;       // >= 3 + longest str(vk) (512 values x 13 chars + 0.5 KB) + 1 + 78, multiple of 16
constexpr size_t PIECE_STRIDE = 11264;   // >= longest "(" VK ", " INT ", " CH ")", multiple of 16
inline size_t r256(size_t b) { return (b + 255) & ~(size_t)255; }
inline size_t msg_stride_for(size_t max_keys) { return ((max_keys * (PIECE_STRIDE + 2) + 5 + 16) + 15) & ~(size_t)15; }
}  // namespace

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
    if (ntt_init_tables(c) || hash_init_tables(c) || hashstream_init_tables(c) || sample_init_tables(c) || scheme_init(c)) { delete c; return -2; }
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

// named integer option: "buffered_hash" = 1 selects the scratch-buffer hashing pipeline (fz_hash.cu) instead of the
// streaming generator (fz_hashstream.cu); both are bit-identical (tests/test_gpu_scheme.py)
int fz_set_option(fz_ctx* c, const char* name, long long value) {
    FZ_REQUIRE(c && name, "fz_set_option: NULL argument");
    if (!strcmp(name, "buffered_hash")) { c->buffered_hash = value ? 1 : 0; return 0; }
    if (!strcmp(name, "hash_ag_impl")) { c->hash_ag_impl = (int)value; return 0; }
    if (!strcmp(name, "sample_impl")) { fz::set_sample_impl((int)value); return 0; }
    fz::set_error("fz_set_option: unknown option %s", name);
    return -1;
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

int fz_pointwise_add(fz_ctx* ctx, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, void* stream) {
    FZ_REQUIRE(ctx && (n_polys == 0 || (a && b && out)), "fz_pointwise_add: NULL argument");
    return launch_pointwise_add(ctx, a, b, out, n_polys * ctx->d, (cudaStream_t)stream);
}

int fz_pointwise_muladd(fz_ctx* ctx, const int32_t* a, const int32_t* b, const int32_t* c, int32_t* out, size_t n_polys, void* stream) {
    FZ_REQUIRE(ctx && (n_polys == 0 || (a && b && c && out)), "fz_pointwise_muladd: NULL argument");
    return launch_pointwise(ctx, a, b, c, out, n_polys * ctx->d, (cudaStream_t)stream);
}

int fz_polymul(fz_ctx* ctx, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, void* stream) {
    FZ_REQUIRE(ctx && (n_polys == 0 || (a && b && out)), "fz_polymul: NULL argument");
    return launch_polymul(ctx, a, b, out, n_polys, (cudaStream_t)stream);
}

int fz_ntt_pass(fz_ctx* ctx, const int32_t* a, const int32_t* b_hat, int32_t* out, size_t n_polys, void* stream) {
    FZ_REQUIRE(ctx && (n_polys == 0 || (a && b_hat && out)), "fz_ntt_pass: NULL argument");
    return launch_ntt_pass(ctx, a, b_hat, out, n_polys, (cudaStream_t)stream);
}

int fz_matvec(fz_ctx* ctx, const int32_t* A, const int32_t* s, int32_t* out, size_t batch, void* stream) {
    FZ_REQUIRE(ctx && (batch == 0 || (A && s && out)), "fz_matvec: NULL argument");
    return launch_matvec(ctx, A, s, (size_t)ctx->d, (size_t)ctx->ell * ctx->d, out, batch, (cudaStream_t)stream);
}

int fz_sign_mac(fz_ctx* ctx, const int32_t* sk_l, const int32_t* sk_r, const int32_t* c_hat, int32_t* sig, size_t batch, void* stream) {
    FZ_REQUIRE(ctx && (batch == 0 || (sk_l && sk_r && c_hat && sig)), "fz_sign_mac: NULL argument");
    return launch_sign_mac(ctx, sk_l, sk_r, (size_t)ctx->d, (size_t)ctx->ell * ctx->d, c_hat, sig, batch, (cudaStream_t)stream);
}

// Chunked H2D -> kernels -> D2H pipeline on internal streams (both copy engines run
// concurrently with the kernels of the neighbouring chunks).  mode 0: forward NTT; 1: inverse NTT;
// 2: forward NTT -> pointwise product with the device-resident b_hat -> inverse NTT.
static int host_pipeline(fz_ctx* ctx, int mode, const int32_t* host_in, const int32_t* dev_bhat, int32_t* host_out, size_t n_polys) {
    if (n_polys == 0) return 0;
    if (check(cudaSetDevice(ctx->device), "cudaSetDevice")) return -2;
    const size_t poly_bytes = (size_t)ctx->d * sizeof(int32_t);
    size_t chunk_mb = 32;   // MiB per chunk; FZ_HOST_CHUNK_MB overrides (tuning knob of the host pipeline)
    if (const char* e = getenv("FZ_HOST_CHUNK_MB")) { const long v = atol(e); if (v >= 1 && v <= 1024) chunk_mb = (size_t)v; }
    const size_t chunk = (chunk_mb << 20) / poly_bytes;
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
            rc = launch_ntt_pass(ctx, buf, dev_bhat + done * ctx->d, buf, n, st[bi]);   // in place: every lane reads its tile before it writes it
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


// ------------------------------------------------------------------------------------------

// SHA3-256 / SHAKE256 over n byte strings data[offs[i]..offs[i+1]) (device pointers); generic, byte granular.
int fz_sha3_256(fz_ctx* ctx, const uint8_t* data, const uint64_t* offs, size_t n, uint8_t* out32, void* stream) {
    FZ_REQUIRE(ctx && (n == 0 || (data && offs && out32)), "fz_sha3_256: NULL argument");
    return launch_sponge_bytes(ctx, data, (const unsigned long long*)offs, n, 0, 0, 0x06, out32, 32, 32, (cudaStream_t)stream);
}

int fz_shake256(fz_ctx* ctx, const uint8_t* data, const uint64_t* offs, size_t n, uint8_t* out, size_t out_len, void* stream) {
    FZ_REQUIRE(ctx && (n == 0 || (data && offs && out)), "fz_shake256: NULL argument");
    FZ_REQUIRE(out_len < ((size_t)1 << 32), "fz_shake256: out_len too large");
    return launch_sponge_bytes(ctx, data, (const unsigned long long*)offs, n, 0, 0, 0x1F, out, out_len, (uint32_t)out_len, (cudaStream_t)stream);
}

int fz_hash_msg(fz_ctx* ctx, const uint8_t* msgs, const uint64_t* offs, size_t n, uint8_t* digests, void* stream) {
    FZ_REQUIRE(ctx && (n == 0 || (msgs && offs && digests)), "fz_hash_msg: NULL argument");
    const uint32_t prefix = (uint32_t)(ctx->secpar == 256 ? 3 : 1) | (0u << 8) | ((uint32_t)',' << 16);  // fusion.py:38-43, 407
    return launch_sponge_bytes(ctx, msgs, (const unsigned long long*)offs, n, prefix, 3, 0x06, digests, 32, 32, (cudaStream_t)stream);
}

size_t fz_text_stride(const fz_ctx* ctx, int mode) { (void)ctx; return mode == 0 ? M1_STRIDE : PIECE_STRIDE; }

int fz_format_text(fz_ctx* ctx, int mode, const int32_t* vk_l, const int32_t* vk_r, const uint8_t* digests, const int32_t* c_hat, size_t n,
                   uint8_t* out, uint32_t* out_len, void* stream) {
    FZ_REQUIRE(ctx && (n == 0 || (vk_l && vk_r && digests && out && out_len)), "fz_format_text: NULL argument");
    FZ_REQUIRE(mode == 0 || (mode == 1 && (n == 0 || c_hat)), "fz_format_text: mode must be 0 (hash_ch message) or 1 (hash_ag piece, needs c_hat)");
    return launch_format(ctx, mode, vk_l, vk_r, digests, c_hat, n, out, mode == 0 ? M1_STRIDE : PIECE_STRIDE, out_len, (cudaStream_t)stream);
}

int fz_decode_poly(fz_ctx* ctx, const uint8_t* bytes, size_t stride, size_t blen, size_t n, long long log2_bias, long long norm_bound, int weight_bound,
                   int32_t* coef_out, void* stream) {
    FZ_REQUIRE(ctx && (n == 0 || (bytes && coef_out)), "fz_decode_poly: NULL argument");
    FZ_REQUIRE(weight_bound >= 0 && weight_bound <= 256, "fz_decode_poly: weight_bound must be in [0, 256]");
    FZ_REQUIRE(log2_bias > -(1LL << 20) && log2_bias < (1LL << 20), "fz_decode_poly: log2_bias out of range");
    long long bound = norm_bound < 1073732864LL ? norm_bound : 1073732864LL;   // fusion.py:432
    if (bound < 1) bound = 1;
    int lb = 0;
    while ((1LL << (lb + 1)) <= bound) ++lb;
    const int frac = ((1LL << lb) != bound);
    const int bpc = ceil_div(lb + frac + 1 + (int)log2_bias, 8), bpi = ceil_div(ilog2(ctx->d) + (int)log2_bias, 8);   // fusion.py:433-434
    const size_t total = (size_t)ceil_div(weight_bound, 8) + (size_t)(bpc + bpi) * weight_bound;             // fusion.py:435-438
    if (blen < total) {
        set_error("Too few bytes to decode polynomial. Expected %zu but got %zu", total, blen);  // fusion.py:439-442
        return -3;
    }
    return launch_decode(ctx, bytes, stride, (uint32_t)blen, n, weight_bound, (uint32_t)bound, bpc, bpi, coef_out, (cudaStream_t)stream);
}

size_t fz_hash_ch_workspace(const fz_ctx* ctx, size_t n) { return ctx->buffered_hash ? r256(n * M1_STRIDE) + r256(n * 4) + 512 : 256; }

// fusion.py:511-531 hash_ch for n (vk, prehash digest) pairs: c_hat_out[i] = NTT(decode(SHAKE256(dst "," str(vk_i) "," str(int_i))))
int fz_hash_ch(fz_ctx* ctx, const int32_t* vk_l, const int32_t* vk_r, const uint8_t* digests, size_t n, int32_t* c_hat_out, void* ws, size_t ws_bytes,
               void* stream) {
    FZ_REQUIRE(ctx && (n == 0 || (vk_l && vk_r && digests && c_hat_out)), "fz_hash_ch: NULL argument");
    if (!ctx->buffered_hash) {
        int rc = launch_hash_ch_stream(ctx, vk_l, vk_r, digests, n, c_hat_out, (cudaStream_t)stream);
        if (!rc) rc = launch_ntt(ctx, 0, c_hat_out, c_hat_out, n, (cudaStream_t)stream);
        return rc;
    }
    FZ_REQUIRE(n == 0 || ws, "fz_hash_ch: workspace required in buffered mode");
    Carver cv(ws, ws_bytes);
    uint8_t* m1 = cv.take<uint8_t>(n * M1_STRIDE);
    uint32_t* len = cv.take<uint32_t>(n);
    FZ_REQUIRE(cv.ok, "fz_hash_ch: workspace too small (see fz_hash_ch_workspace)");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_format(ctx, 0, vk_l, vk_r, digests, nullptr, n, m1, M1_STRIDE, len, st);
    if (!rc) rc = launch_shake_decode(ctx, m1, M1_STRIDE, len, n, c_hat_out, st);
    if (!rc) rc = launch_ntt(ctx, 0, c_hat_out, c_hat_out, n, st);
    return rc;
}

size_t fz_hash_ag_workspace(const fz_ctx* ctx, size_t n_total, size_t n_aggs, size_t max_keys) {
    if (!ctx->buffered_hash) return r256(n_aggs * ag_state_bytes()) + 256;
    return r256(n_total * PIECE_STRIDE) + r256(n_total * 4) + r256(n_total * 8) + r256(n_aggs * msg_stride_for(max_keys)) + r256(n_aggs * 4) +
           r256(n_total * (size_t)ctx->n_ag + 16) + 2048;
}

// fusion.py:632-652 hash_ag for n_aggs aggregates whose (already sorted) signers are the slices
// [agg_off[j], agg_off[j+1]) of vk_l_s / vk_r_s / dg_s / c_hat_s.  alpha_hat_out: [n_total][d] in the same order.
int fz_hash_ag(fz_ctx* ctx, const int32_t* vk_l_s, const int32_t* vk_r_s, const uint8_t* dg_s, const int32_t* c_hat_s, const uint64_t* agg_off,
               size_t n_aggs, size_t n_total, size_t max_keys, int32_t* alpha_hat_out, void* ws, size_t ws_bytes, void* stream) {
    FZ_REQUIRE(ctx && (n_total == 0 || (vk_l_s && vk_r_s && dg_s && c_hat_s && agg_off && alpha_hat_out)), "fz_hash_ag: NULL argument");
    if (n_aggs == 0 || n_total == 0) return 0;
    if (!ctx->buffered_hash) {
        // one window covering every signer: begin -> absorb -> squeeze on a state record per aggregate carved from ws
        FZ_REQUIRE(ws, "fz_hash_ag: workspace required (see fz_hash_ag_workspace)");
        Carver cv(ws, ws_bytes);
        void* state = cv.take<uint8_t>(n_aggs * ag_state_bytes());
        FZ_REQUIRE(cv.ok, "fz_hash_ag: workspace too small (see fz_hash_ag_workspace)");
        cudaStream_t st = (cudaStream_t)stream;
        int rc = launch_ag_state_init(ctx, state, (const unsigned long long*)agg_off, n_aggs, st);
        if (!rc) rc = launch_hash_ag_absorb(ctx, state, vk_l_s, vk_r_s, dg_s, c_hat_s, (const unsigned long long*)agg_off, n_aggs, st);
        if (!rc) rc = launch_hash_ag_squeeze(ctx, state, (const unsigned long long*)agg_off, n_aggs, alpha_hat_out, st);
        if (!rc) rc = launch_ntt(ctx, 0, alpha_hat_out, alpha_hat_out, n_total, st);
        return rc;
    }
    FZ_REQUIRE(ws, "fz_hash_ag: workspace required in buffered mode");
    Carver cv(ws, ws_bytes);
    const size_t mstride = msg_stride_for(max_keys);
    uint8_t* pieces = cv.take<uint8_t>(n_total * PIECE_STRIDE);
    uint32_t* plen = cv.take<uint32_t>(n_total);
    unsigned long long* ppos = cv.take<unsigned long long>(n_total);
    uint8_t* msgs = cv.take<uint8_t>(n_aggs * mstride);
    uint32_t* mlen = cv.take<uint32_t>(n_aggs);
    uint8_t* xof = cv.take<uint8_t>(n_total * (size_t)ctx->n_ag + 16);
    FZ_REQUIRE(cv.ok, "fz_hash_ag: workspace too small (see fz_hash_ag_workspace)");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_format(ctx, 1, vk_l_s, vk_r_s, dg_s, c_hat_s, n_total, pieces, PIECE_STRIDE, plen, st);
    if (!rc) rc = launch_ag_message(ctx, pieces, PIECE_STRIDE, plen, nullptr, (const unsigned long long*)agg_off, n_aggs, n_total, ppos, msgs, mstride, mlen, st);
    if (!rc) rc = launch_shake_ag(ctx, msgs, mstride, mlen, (const unsigned long long*)agg_off, xof, n_aggs, st);
    if (!rc) rc = launch_decode(ctx, xof, (size_t)ctx->n_ag, (uint32_t)ctx->n_ag, n_total, ctx->omega_ag, 1u, ctx->bytes_coef, ctx->bytes_idx, alpha_hat_out, st);
    if (!rc) rc = launch_ntt(ctx, 0, alpha_hat_out, alpha_hat_out, n_total, st);
    return rc;
}

// ---- windowed hash_ag: the same serial sponge per aggregate, consumed a window of sorted signer positions at a time.
// state: n_aggs records of fz_hash_ag_state_bytes() bytes (device memory owned by the caller).
size_t fz_hash_ag_state_bytes(const fz_ctx* ctx) { (void)ctx; return ag_state_bytes(); }

// agg_off: n_aggs + 1 offsets; only the signer COUNTS agg_off[j+1] - agg_off[j] are recorded
int fz_hash_ag_begin(fz_ctx* ctx, void* state, const uint64_t* agg_off, size_t n_aggs, void* stream) {
    FZ_REQUIRE(ctx && (n_aggs == 0 || (state && agg_off)), "fz_hash_ag_begin: NULL argument");
    return launch_ag_state_init(ctx, state, (const unsigned long long*)agg_off, n_aggs, (cudaStream_t)stream);
}

// absorbs the next win_off[j+1] - win_off[j] sorted signers of every aggregate j; the window arrays hold them contiguously.
// The call that delivers an aggregate's last signer pads and closes its message (an aggregate without signers closes at once).
int fz_hash_ag_absorb(fz_ctx* ctx, void* state, const int32_t* vk_l_w, const int32_t* vk_r_w, const uint8_t* dg_w, const int32_t* c_hat_w,
                      const uint64_t* win_off, size_t n_aggs, void* stream) {
    FZ_REQUIRE(ctx && (n_aggs == 0 || (state && win_off)), "fz_hash_ag_absorb: NULL argument");
    return launch_hash_ag_absorb(ctx, state, vk_l_w, vk_r_w, dg_w, c_hat_w, (const unsigned long long*)win_off, n_aggs, (cudaStream_t)stream);
}

// squeezes, decodes and transforms the aggregation coefficients of the next window of signers: alpha_hat_w: [n_window][d]
int fz_hash_ag_squeeze(fz_ctx* ctx, void* state, const uint64_t* win_off, size_t n_aggs, size_t n_window, int32_t* alpha_hat_w, void* stream) {
    FZ_REQUIRE(ctx && (n_aggs == 0 || (state && win_off)) && (n_window == 0 || alpha_hat_w), "fz_hash_ag_squeeze: NULL argument");
    int rc = launch_hash_ag_squeeze(ctx, state, (const unsigned long long*)win_off, n_aggs, alpha_hat_w, (cudaStream_t)stream);
    if (!rc) rc = launch_ntt(ctx, 0, alpha_hat_w, alpha_hat_w, n_window, (cudaStream_t)stream);
    return rc;
}

// ---- windowed hash_ag with the text generated off the chain: fz_hash_ag_text writes the bytes a window adds to every
// aggregate's message into per-aggregate byte rings (parallel work, any stream), fz_hash_ag_absorb_ring advances the sponges
// over them (the serial chain), fz_hash_ag_squeeze_xof writes whole squeezed blocks per aggregate and fz_hash_ag_decode turns
// them into coefficients (parallel work again).  Windows are uniform: window w0 holds sorted positions w0 .. w0 + P - 1.
size_t fz_hash_ag_ring_bytes(const fz_ctx* ctx, size_t max_window) {   // per aggregate; holds two windows of text in flight
    (void)ctx;
    const size_t need = 2 * max_window * agw_max_piece() + 2 * 272;
    return (need + 271) / 272 * 272;
}
size_t fz_hash_ag_xof_bytes(const fz_ctx* ctx, size_t max_window) { return ((max_window * (size_t)ctx->n_ag + 2 * 136 + 32) + 15) & ~(size_t)15; }
size_t fz_hash_ag_text_workspace(const fz_ctx* ctx, size_t n_window) {
    (void)ctx;
    return r256(n_window * 4) + r256(n_window * 8) + r256(n_window * agw_int_text_stride()) + r256(n_window * 4) + 256;
}

int fz_hash_ag_text(fz_ctx* ctx, const int32_t* vk_l_w, const int32_t* vk_r_w, const uint8_t* dg_w, const int32_t* c_hat_w, const uint64_t* win_off,
                    const uint32_t* agg_of, const uint64_t* agg_off, size_t n_aggs, size_t n_window, uint32_t w0, uint64_t* text_pos, uint8_t* ring,
                    size_t ring_cap, const void* state, uint64_t* wend_out, void* ws, size_t ws_bytes, void* stream) {
    FZ_REQUIRE(ctx && (n_aggs == 0 || (win_off && agg_off && text_pos && ring && wend_out)), "fz_hash_ag_text: NULL argument");
    FZ_REQUIRE(n_window == 0 || (vk_l_w && vk_r_w && dg_w && c_hat_w && agg_of && ws), "fz_hash_ag_text: NULL window argument");
    FZ_REQUIRE(ring_cap % 272 == 0 && ring_cap >= 544, "fz_hash_ag_text: ring_cap must be a multiple of 272 (see fz_hash_ag_ring_bytes)");
    Carver cv(ws, ws_bytes);
    uint32_t* piece_len = cv.take<uint32_t>(n_window);
    unsigned long long* piece_off = cv.take<unsigned long long>(n_window);
    uint8_t* int_text = cv.take<uint8_t>(n_window * agw_int_text_stride());   // the decimal prehash of every signer of the window, 80-byte slots
    uint32_t* int_len = cv.take<uint32_t>(n_window);
    FZ_REQUIRE(cv.ok, "fz_hash_ag_text: workspace too small (see fz_hash_ag_text_workspace)");
    return launch_agw_text(ctx, vk_l_w, vk_r_w, dg_w, c_hat_w, (const unsigned long long*)win_off, agg_of, (const unsigned long long*)agg_off, n_aggs,
                           n_window, w0, (unsigned long long*)text_pos, piece_len, piece_off, int_text, int_len, (unsigned long long*)wend_out, ring, ring_cap,
                           state, (cudaStream_t)stream);
}

int fz_hash_ag_absorb_ring(fz_ctx* ctx, void* state, const uint8_t* ring, size_t ring_cap, const uint64_t* wend, const uint64_t* win_off, size_t n_aggs,
                           void* stream) {
    FZ_REQUIRE(ctx && (n_aggs == 0 || (state && ring && wend && win_off)), "fz_hash_ag_absorb_ring: NULL argument");
    return launch_agw_absorb(ctx, state, ring, ring_cap, (const unsigned long long*)wend, (const unsigned long long*)win_off, n_aggs, (cudaStream_t)stream);
}

// Time-sliced absorb for warp-per-aggregate batches slightly above two chains per SM sub-partition: one launch advances the n_active
// aggregates first_active .. (mod n_aggs) by at most budget_blocks 136-byte blocks each; wend as written by fz_hash_ag_text (bit 63 =
// text complete).  The caller rotates first_active so that every aggregate rests equally often, keeps the rings large enough for the
// backlog of the resting aggregates, and finishes with launches over all aggregates until every sponge is closed.
int fz_hash_ag_absorb_sliced(fz_ctx* ctx, void* state, const uint8_t* ring, size_t ring_cap, const uint64_t* wend, size_t n_aggs, size_t first_active,
                             size_t n_active, uint64_t budget_blocks, void* stream) {
    FZ_REQUIRE(ctx && (n_aggs == 0 || (state && ring && wend)), "fz_hash_ag_absorb_sliced: NULL argument");
    FZ_REQUIRE(n_active <= n_aggs && (n_aggs == 0 || first_active < n_aggs), "fz_hash_ag_absorb_sliced: bad active range");
    return launch_agw_absorb_sliced(ctx, state, ring, ring_cap, (const unsigned long long*)wend, n_aggs, first_active, n_active, budget_blocks,
                                    (cudaStream_t)stream);
}

int fz_hash_ag_squeeze_xof(fz_ctx* ctx, void* state, const uint64_t* win_off, size_t n_aggs, uint8_t* xof, size_t xof_cap, uint32_t* bpos_out,
                           void* stream) {
    FZ_REQUIRE(ctx && (n_aggs == 0 || (state && win_off && xof && bpos_out)), "fz_hash_ag_squeeze_xof: NULL argument");
    FZ_REQUIRE(xof_cap % 16 == 0, "fz_hash_ag_squeeze_xof: xof_cap must be a multiple of 16 (see fz_hash_ag_xof_bytes)");
    return launch_agw_squeeze(ctx, state, (const unsigned long long*)win_off, n_aggs, xof, xof_cap, bpos_out, (cudaStream_t)stream);
}

int fz_hash_ag_decode(fz_ctx* ctx, const uint8_t* xof, size_t xof_cap, const uint32_t* bpos, const uint64_t* win_off, const uint32_t* agg_of,
                      size_t n_window, int32_t* alpha_hat_w, void* stream) {
    FZ_REQUIRE(ctx && (n_window == 0 || (xof && bpos && win_off && agg_of && alpha_hat_w)), "fz_hash_ag_decode: NULL argument");
    int rc = launch_agw_decode(ctx, xof, xof_cap, bpos, (const unsigned long long*)win_off, agg_of, n_window, alpha_hat_w, (cudaStream_t)stream);
    if (!rc) rc = launch_ntt(ctx, 0, alpha_hat_w, alpha_hat_w, n_window, (cudaStream_t)stream);
    return rc;
}

// sorted(..., key=str(vk)) (fusion.py:661-663, 693): perm[agg_off[j] + rank] = original position; sel (optional) maps positions to key rows
int fz_sort_keys(fz_ctx* ctx, const int32_t* vk_l, const int32_t* vk_r, const uint32_t* sel, const uint64_t* agg_off, size_t n_aggs, size_t max_keys,
                 uint32_t* perm, void* stream) {
    FZ_REQUIRE(ctx && (n_aggs == 0 || (vk_l && vk_r && agg_off && perm)), "fz_sort_keys: NULL argument");
    FZ_REQUIRE(max_keys <= 28000, "fz_sort_keys: more than 28000 keys per aggregate is not supported");
    return launch_sort(ctx, vk_l, vk_r, sel, (const unsigned long long*)agg_off, n_aggs, max_keys, perm, (cudaStream_t)stream);
}

int fz_gather_sorted(fz_ctx* ctx, const int32_t* vk_l, const int32_t* vk_r, const uint8_t* digests, const uint32_t* sel, const uint32_t* perm,
                     size_t n_total, int32_t* vk_l_s, int32_t* vk_r_s, uint8_t* dg_s, void* stream) {
    FZ_REQUIRE(ctx && (n_total == 0 || (vk_l && vk_r && digests && vk_l_s && vk_r_s && dg_s)), "fz_gather_sorted: NULL argument");
    return launch_gather_sorted(ctx, vk_l, vk_r, digests, sel, perm, n_total, vk_l_s, vk_r_s, dg_s, (cudaStream_t)stream);
}

size_t fz_aggregate_sum_workspace(const fz_ctx* ctx, size_t n_aggs, size_t max_keys) {
    const int split = aggregate_split(ctx, n_aggs, max_keys);
    return split == 1 ? 256 : r256((size_t)split * n_aggs * ctx->ell * ctx->d * 4) + 256;
}

// fusion.py:670-677: out[j] = sum_p sigs[perm[p]] * alpha[p] over the sorted positions p of aggregate j
int fz_aggregate_sum(fz_ctx* ctx, const int32_t* sigs, const uint32_t* perm, const int32_t* alpha, const uint64_t* agg_off, size_t n_aggs, size_t max_keys,
                     int32_t* out, void* ws, size_t ws_bytes, void* stream) {
    FZ_REQUIRE(ctx && (n_aggs == 0 || (sigs && alpha && agg_off && out)), "fz_aggregate_sum: NULL argument");
    const int split = aggregate_split(ctx, n_aggs, max_keys);
    Carver cv(ws, ws_bytes);
    int32_t* part = split == 1 ? nullptr : cv.take<int32_t>((size_t)split * n_aggs * ctx->ell * ctx->d);
    FZ_REQUIRE(cv.ok && (split == 1 || ws), "fz_aggregate_sum: workspace too small (see fz_aggregate_sum_workspace)");
    return launch_aggregate(ctx, sigs, perm, alpha, (const unsigned long long*)agg_off, n_aggs, split, part, out, (cudaStream_t)stream);
}

// modular sum of n_parts partial aggregates laid out [n_aggs][n_parts][ell][d] (e.g. all-gathered from other ranks)
int fz_fold_partials(fz_ctx* ctx, const int32_t* parts, int n_parts, size_t n_aggs, int32_t* out, void* stream) {
    FZ_REQUIRE(ctx && (n_aggs == 0 || (parts && out)) && n_parts >= 1 && n_parts <= 4096, "fz_fold_partials: bad argument");
    return launch_fold_partials(ctx, parts, out, n_aggs, n_parts, (cudaStream_t)stream);
}

size_t fz_verify_check_workspace(const fz_ctx* ctx, size_t n_aggs) {
    return 2 * r256(n_aggs * ctx->d * 4) + r256(n_aggs * (size_t)ctx->ell * ctx->d * 4) + 1024;
}

// fusion.py:690-692 + 706-728 on sorted inputs: codes[j] = 0 ok | 3 target mismatch | 4 norm too large
int fz_verify_check(fz_ctx* ctx, const int32_t* A, const int32_t* vk_l_s, const int32_t* vk_r_s, const int32_t* c_hat_s, const int32_t* alpha_s,
                    const int32_t* agg, const uint64_t* agg_off, size_t n_aggs, uint8_t* codes, void* ws, size_t ws_bytes, void* stream) {
    FZ_REQUIRE(ctx && (n_aggs == 0 || (A && vk_l_s && vk_r_s && c_hat_s && alpha_s && agg && agg_off && codes && ws)), "fz_verify_check: NULL argument");
    Carver cv(ws, ws_bytes);
    int32_t* target = cv.take<int32_t>(n_aggs * ctx->d);
    int32_t* observed = cv.take<int32_t>(n_aggs * ctx->d);
    int32_t* coef = cv.take<int32_t>(n_aggs * (size_t)ctx->ell * ctx->d);
    FZ_REQUIRE(cv.ok, "fz_verify_check: workspace too small (see fz_verify_check_workspace)");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_verify_target(ctx, vk_l_s, vk_r_s, c_hat_s, alpha_s, (const unsigned long long*)agg_off, n_aggs, target, st);
    if (!rc) rc = launch_matvec(ctx, A, agg, (size_t)ctx->d, (size_t)ctx->ell * ctx->d, observed, n_aggs, st);
    if (!rc) rc = launch_ntt(ctx, 1, agg, coef, n_aggs * ctx->ell, st);
    if (!rc) rc = launch_verify_finalize(ctx, target, observed, coef, n_aggs, codes, st);
    return rc;
}

// fusion.py:706-714 over one window of sorted signers: target_w[j] = sum_p (vk_l[p] * c_hat[p] + vk_r[p]) * alpha[p]   ([n_aggs][d]);
// window partials add up with fz_pointwise_add
int fz_verify_target(fz_ctx* ctx, const int32_t* vk_l_w, const int32_t* vk_r_w, const int32_t* c_hat_w, const int32_t* alpha_w, const uint64_t* win_off,
                     size_t n_aggs, int32_t* target_w, void* stream) {
    FZ_REQUIRE(ctx && (n_aggs == 0 || (win_off && target_w)), "fz_verify_target: NULL argument");
    return launch_verify_target(ctx, vk_l_w, vk_r_w, c_hat_w, alpha_w, (const unsigned long long*)win_off, n_aggs, target_w, (cudaStream_t)stream);
}

size_t fz_verify_finish_workspace(const fz_ctx* ctx, size_t n_aggs) { return r256(n_aggs * ctx->d * 4) + r256(n_aggs * (size_t)ctx->ell * ctx->d * 4) + 512; }

// fusion.py:690-692, 715-728 given the complete target: observed = A . agg, comparison, infinity norm of INTT(agg)
int fz_verify_finish(fz_ctx* ctx, const int32_t* A, const int32_t* target, const int32_t* agg, size_t n_aggs, uint8_t* codes, void* ws, size_t ws_bytes,
                     void* stream) {
    FZ_REQUIRE(ctx && (n_aggs == 0 || (A && target && agg && codes && ws)), "fz_verify_finish: NULL argument");
    Carver cv(ws, ws_bytes);
    int32_t* observed = cv.take<int32_t>(n_aggs * ctx->d);
    int32_t* coef = cv.take<int32_t>(n_aggs * (size_t)ctx->ell * ctx->d);
    FZ_REQUIRE(cv.ok, "fz_verify_finish: workspace too small (see fz_verify_finish_workspace)");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_matvec(ctx, A, agg, (size_t)ctx->d, (size_t)ctx->ell * ctx->d, observed, n_aggs, st);
    if (!rc) rc = launch_ntt(ctx, 1, agg, coef, n_aggs * ctx->ell, st);
    if (!rc) rc = launch_verify_finalize(ctx, target, observed, coef, n_aggs, codes, st);
    return rc;
}

// ------------------------------------------------------------------------------------------ samplers / keygen
// keys: [n][key_stride] little-endian 32-bit limbs of abs(seed) (CPython random.seed), klen[n] limbs used (1..8)
int fz_sample_coef(fz_ctx* ctx, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, int norm_bound, int weight_bound, int32_t* out,
                   void* stream) {
    FZ_REQUIRE(ctx && (n == 0 || (keys && klen && out)) && key_stride >= 1, "fz_sample_coef: bad argument");
    long long bound = norm_bound < 1073732864LL ? norm_bound : 1073732864LL;  // polynomials.py:450
    FZ_REQUIRE(bound >= 1, "fz_sample_coef: norm_bound must be >= 1");
    return launch_sample(ctx, 0, keys, key_stride, klen, n, (uint32_t)bound, weight_bound, out, (cudaStream_t)stream);
}

int fz_sample_ntt(fz_ctx* ctx, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, int32_t* out, void* stream) {
    FZ_REQUIRE(ctx && (n == 0 || (keys && klen && out)) && key_stride >= 1, "fz_sample_ntt: bad argument");
    return launch_sample(ctx, 1, keys, key_stride, klen, n, 0, 0, out, (cudaStream_t)stream);
}

size_t fz_keygen_workspace(const fz_ctx* ctx, size_t n) { return r256(2 * n * 4 * 4) + r256(2 * n * 4) + r256(2 * n * ctx->d * 4) + r256(ctx->d * 4) + 1024; }

// fusion.py:338-373 for n u64 seeds (seed_i for the left key, seed_i + 1 for the right key).
// sk_l / sk_r ([n][ell][d]) may be NULL (keys are then not materialised); sk_hat ([n][2][d], optional) receives the
// one distinct row of each side (quirk Q1: all ell rows of a seeded key are the same polynomial).
int fz_keygen(fz_ctx* ctx, const int32_t* A, const uint64_t* seeds, size_t n, int32_t* sk_l, int32_t* sk_r, int32_t* sk_hat, int32_t* vk_l, int32_t* vk_r,
              void* ws, size_t ws_bytes, void* stream) {
    FZ_REQUIRE(ctx && (n == 0 || (A && seeds && vk_l && vk_r && ws)), "fz_keygen: NULL argument");
    Carver cv(ws, ws_bytes);
    uint32_t* keys = cv.take<uint32_t>(2 * n * 4);
    int32_t* klen = cv.take<int32_t>(2 * n);
    int32_t* tmp = sk_hat ? sk_hat : cv.take<int32_t>(2 * n * ctx->d);
    int32_t* asum = cv.take<int32_t>(ctx->d);
    FZ_REQUIRE(cv.ok, "fz_keygen: workspace too small (see fz_keygen_workspace)");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t d = ctx->d;
    int rc = launch_seeds_to_keys(ctx, (const unsigned long long*)seeds, n, keys, klen, st);
    if (!rc) rc = launch_sample(ctx, 0, keys, 4, klen, 2 * n, 52u, ctx->d, tmp, st);   // beta_sk = 52, omega_sk = d (fusion.py:87, 90)
    if (!rc) rc = launch_ntt(ctx, 0, tmp, tmp, 2 * n, st);
    // vk = A . sk with every row of sk equal to the sampled polynomial: (sum of the rows of A) o sk   (quirk Q1; the general
    // ell-row product is fz_matvec)
    if (!rc) rc = launch_matvec_rows(ctx, A, asum, tmp, 2 * d, vk_l, n, st);
    if (!rc) rc = launch_matvec_rows(ctx, A, asum, tmp + d, 2 * d, vk_r, n, st);
    if (!rc && sk_l) rc = launch_replicate_rows(ctx, tmp, sk_l, n, ctx->ell, 2 * d, st);
    if (!rc && sk_r) rc = launch_replicate_rows(ctx, tmp + d, sk_r, n, ctx->ell, 2 * d, st);
    return rc;
}

// The general ell-row form of fusion.py:338-373: row k of sk_l[i] is sample_polynomial_coefficient_representation(seed = row_seeds[i][0][k])
// (algebra/polynomials.py:436-467), row k of sk_r[i] comes from row_seeds[i][1][k]; then the 2*ell transforms of fusion.py:364-368 and the two
// 1 x ell by ell x 1 products of :369-370.  With row_seeds[i][0][k] = seed_i and row_seeds[i][1][k] = seed_i + 1 for every k this IS
// keygen(params, seed_i) computed the long way (2*ell generators, 2*ell transforms, ell-term products) -- no use of quirk Q1.
size_t fz_keygen_rows_workspace(const fz_ctx* ctx, size_t n) { return r256(2 * n * ctx->ell * 4 * 4) + r256(2 * n * ctx->ell * 4) + 1024; }

int fz_keygen_rows(fz_ctx* ctx, const int32_t* A, const uint64_t* row_seeds, size_t n, int32_t* sk_l, int32_t* sk_r, int32_t* vk_l, int32_t* vk_r, void* ws,
                   size_t ws_bytes, void* stream) {
    FZ_REQUIRE(ctx && (n == 0 || (A && row_seeds && sk_l && sk_r && vk_l && vk_r && ws)), "fz_keygen_rows: NULL argument");
    if (n == 0) return 0;
    const size_t ell = ctx->ell, d = ctx->d;
    Carver cv(ws, ws_bytes);
    uint32_t* keys = cv.take<uint32_t>(2 * n * ell * 4);
    int32_t* klen = cv.take<int32_t>(2 * n * ell);
    FZ_REQUIRE(cv.ok, "fz_keygen_rows: workspace too small (see fz_keygen_rows_workspace)");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_row_seeds_to_keys(ctx, (const unsigned long long*)row_seeds, n, (int)ell, keys, klen, st);   // side-major keys
    // one generator per row, written straight into the secret-key matrices (beta_sk = 52, omega_sk = d: fusion.py:87, 90)
    if (!rc) rc = launch_sample(ctx, 0, keys, 4, klen, n * ell, 52u, ctx->d, sk_l, st);
    if (!rc) rc = launch_sample(ctx, 0, keys + n * ell * 4, 4, klen + n * ell, n * ell, 52u, ctx->d, sk_r, st);
    if (!rc) rc = launch_ntt(ctx, 0, sk_l, sk_l, n * ell, st);
    if (!rc) rc = launch_ntt(ctx, 0, sk_r, sk_r, n * ell, st);
    if (!rc) rc = launch_matvec(ctx, A, sk_l, d, ell * d, vk_l, n, st);
    if (!rc) rc = launch_matvec(ctx, A, sk_r, d, ell * d, vk_r, n, st);
    return rc;
}

// fusion.py:338-373 + 534-557 in one call: n one-time keys from u64 seeds, each signing message i (msgs[offs[i]..offs[i+1]) on
// the device).  The secret keys never leave the workspace; vk_l / vk_r: [n][d], sig: [n][ell][d].
size_t fz_keygen_sign_workspace(const fz_ctx* ctx, size_t n) {
    return r256(fz_keygen_workspace(ctx, n)) + r256(fz_hash_ch_workspace(ctx, n)) + r256(2 * n * ctx->d * 4) + r256(n * 32) + r256(n * ctx->d * 4) + 1024;
}

int fz_keygen_sign(fz_ctx* ctx, const int32_t* A, const uint64_t* seeds, const uint8_t* msgs, const uint64_t* offs, size_t n, int32_t* vk_l, int32_t* vk_r,
                   int32_t* sig, void* ws, size_t ws_bytes, void* stream) {
    FZ_REQUIRE(ctx && (n == 0 || (A && seeds && msgs && offs && vk_l && vk_r && sig && ws)), "fz_keygen_sign: NULL argument");
    if (n == 0) return 0;
    Carver cv(ws, ws_bytes);
    int32_t* sk_hat = cv.take<int32_t>(2 * n * ctx->d);
    uint8_t* digests = cv.take<uint8_t>(n * 32);
    int32_t* c_hat = cv.take<int32_t>(n * ctx->d);
    const size_t kws = fz_keygen_workspace(ctx, n), hws = fz_hash_ch_workspace(ctx, n);
    void* kbuf = cv.take<uint8_t>(kws);
    void* hbuf = cv.take<uint8_t>(hws);
    FZ_REQUIRE(cv.ok, "fz_keygen_sign: workspace too small (see fz_keygen_sign_workspace)");
    int rc = fz_keygen(ctx, A, seeds, n, nullptr, nullptr, sk_hat, vk_l, vk_r, kbuf, kws, stream);
    if (!rc) rc = fz_hash_msg(ctx, msgs, offs, n, digests, stream);
    if (!rc) rc = fz_hash_ch(ctx, vk_l, vk_r, digests, n, c_hat, hbuf, hws, stream);
    if (!rc) rc = launch_sign_rows(ctx, sk_hat, c_hat, sig, n, (cudaStream_t)stream);
    return rc;
}

// fusion.py:557 with broadcast rows: sig[b][r] = sk_hat[b][0] * c_hat[b] + sk_hat[b][1] for every r (keys from fz_keygen's sk_hat)
int fz_sign_mac_rows(fz_ctx* ctx, const int32_t* sk_hat, const int32_t* c_hat, int32_t* sig, size_t batch, void* stream) {
    FZ_REQUIRE(ctx && (batch == 0 || (sk_hat && c_hat && sig)), "fz_sign_mac_rows: NULL argument");
    return launch_sign_rows(ctx, sk_hat, c_hat, sig, batch, (cudaStream_t)stream);
}

}  // extern "C"
