// fz_internal.h -- context and launcher prototypes shared by the translation units of
// libfusion_b200.  Not part of the public C ABI (see include/fusion_b200.h).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

struct fz_ctx {
    int device;
    int secpar;
    int d, logd, ell, capacity;
    int omega_ch, omega_ag;
    int bytes_coef, bytes_idx;   // fusion.py:123-141 (33/33 @256, 17/17 @128)
    int n_ch, n_ch_used, n_ag;   // XOF bytes per challenge (requested / consumed) and per agg. coefficient
    long long beta_vf;
    int sm_count;
    uint32_t root, inv_root;
    // scratch owned by the context (grown on demand, never per call when large enough)
    void* scratch;
    size_t scratch_bytes;
};

namespace fz {

void set_error(const char* fmt, ...);
int check(cudaError_t e, const char* what);

// fz_ntt.cu
int ntt_init_tables(fz_ctx* ctx);
int launch_ntt(const fz_ctx* ctx, int inverse, const int32_t* in, int32_t* out, size_t n_polys, cudaStream_t st);
int launch_polymul(const fz_ctx* ctx, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, cudaStream_t st);
int launch_pointwise(const fz_ctx* ctx, const int32_t* a, const int32_t* b, const int32_t* c, int32_t* out, size_t n_coef, cudaStream_t st);

// fz_stream.cu
int launch_matvec(const fz_ctx* ctx, const int32_t* A, const int32_t* s, int32_t* out, size_t batch, cudaStream_t st);
int launch_sign_mac(const fz_ctx* ctx, const int32_t* sk_l, const int32_t* sk_r, const int32_t* c_hat, int32_t* sig, size_t batch, cudaStream_t st);

}  // namespace fz
