// Host build of csrc/fz_ntt_core.cuh: whole transforms run through the product's own butterflies, stage functions, twiddle indexing and table
// construction, one emulated lane after the other; only the data movement of the kernels (loads, the warp-private transpose, stores) is restated
// here, with the header's row_off().  Test infrastructure only (tests/test_ntt_host.py); the library never sees this file.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>

static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((unsigned long long)a * b) >> 32); }
static inline int32_t __mulhi(int32_t a, int32_t b) { return (int32_t)(((long long)a * b) >> 32); }
using std::min;
#undef __constant__
#define __constant__

#include "../../fusion_cryptography_b200/csrc/fz_ntt_core.cuh"

namespace {

bool g_ready = false;
void init_tables() {
    if (g_ready) return;
    fz::ntt_build_tables(8, 3337519ULL, fz::c_fwd256, fz::c_inv256, fz::c_last256);
    fz::ntt_build_tables(6, 23584283ULL, fz::c_fwd64, fz::c_inv64, fz::c_last64);
    g_ready = true;
}

// ntt_fwd_kernel's body for one polynomial; y_out receives the lazy residues ([0, 2q]) of the contiguous ownership when not null
template <int LOGD>
void forward(const int32_t* in, int32_t* out, uint32_t* y_out) {
    using C = fz::Cfg<LOGD>;
    uint32_t tile[C::TPP * 20];
    for (int tau = 0; tau < C::TPP; ++tau) {
        uint32_t x[16];
        for (int k = 0; k < 16; ++k) x[k] = fz::to_lazy(in[tau + C::TPP * k]);
        fz::fwd_strided<LOGD>(x);
        for (int k = 0; k < 16; ++k) tile[fz::row_off(tau + C::TPP * k)] = x[k];
    }
    for (int tau = 0; tau < C::TPP; ++tau) {
        uint32_t y[16];
        for (int j = 0; j < 16; ++j) y[j] = tile[tau * 20 + j];
        fz::Sh tw2[15];
        fz::load_tw2<LOGD>(C::fwd(), tau, tw2);
        fz::fwd_contig<LOGD>(y, tw2);
        for (int j = 0; j < 16; ++j) {
            if (out) out[tau * 16 + j] = fz::cent_u(y[j]);
            if (y_out) y_out[tau * 16 + j] = y[j];
        }
    }
}

// ntt_inv_kernel's body from reduced residues y_in ([0, q]); LAST = 0 plain, 2 with the factor 2^32 that repays a Montgomery product
template <int LOGD, int LAST>
void inverse_from(const uint32_t* y_in, int32_t* out) {
    using C = fz::Cfg<LOGD>;
    uint32_t tile[C::TPP * 20];
    for (int tau = 0; tau < C::TPP; ++tau) {
        uint32_t y[16];
        for (int j = 0; j < 16; ++j) y[j] = y_in[tau * 16 + j];
        fz::Sh tw2[15];
        fz::load_tw2<LOGD>(C::inv(), tau, tw2);
        fz::inv_contig<LOGD>(y, tw2);
        for (int j = 0; j < 16; ++j) tile[tau * 20 + j] = y[j];
    }
    for (int tau = 0; tau < C::TPP; ++tau) {
        uint32_t x[16];
        int32_t o[16];
        for (int k = 0; k < 16; ++k) x[k] = tile[fz::row_off(tau + C::TPP * k)];
        fz::inv_strided<LOGD, LAST>(x, o);
        for (int k = 0; k < 16; ++k) out[tau + C::TPP * k] = o[k];
    }
}

template <int LOGD>
void run(int mode, const int32_t* a, const int32_t* b, int32_t* out, size_t n) {
    constexpr int D = fz::Cfg<LOGD>::D;
    for (size_t p = 0; p < n; ++p) {
        const int32_t* in = a + p * D;
        int32_t* o = out + p * D;
        if (mode == 0) {
            forward<LOGD>(in, o, nullptr);
        } else if (mode == 1) {
            uint32_t y[D];
            for (int e = 0; e < D; ++e) y[e] = fz::to_red(in[e]);
            inverse_from<LOGD, 0>(y, o);
        } else {   // ntt_pass_kernel: INTT(NTT(a) o b_hat) with the Montgomery pointwise product
            uint32_t y[D];
            forward<LOGD>(in, nullptr, y);
            for (int e = 0; e < D; ++e) y[e] = fz::mont_mul(fz::red_q(y[e]), b[p * D + e]);
            inverse_from<LOGD, 2>(y, o);
        }
    }
}

}  // namespace

extern "C" {

// mode 0: forward transform, 1: inverse transform, 2: fused pass INTT(NTT(a) o b_hat); logd 6 or 8
int h_ntt(int logd, int mode, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys) {
    init_tables();
    if (logd == 8) run<8>(mode, a, b, out, n_polys);
    else if (logd == 6) run<6>(mode, a, b, out, n_polys);
    else return -1;
    return 0;
}

// the tables the library uploads: w and wq of the forward / inverse tables
void h_tables(int logd, uint32_t* fw, uint32_t* fwq, uint32_t* iw, uint32_t* iwq) {
    init_tables();
    const int d = 1 << logd;
    const fz::Sh* f = logd == 8 ? fz::c_fwd256 : fz::c_fwd64;
    const fz::Sh* iv = logd == 8 ? fz::c_inv256 : fz::c_inv64;
    for (int k = 0; k < d; ++k) { fw[k] = f[k].w; fwq[k] = f[k].wq; iw[k] = iv[k].w; iwq[k] = iv[k].wq; }
}

}  // extern "C"
