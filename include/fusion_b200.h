/*
 * fusion_b200.h -- C ABI of libfusion_b200.so: the batched Fusion lattice-signature hot path
 * on NVIDIA B200 (sm_100a).
 *
 * The reference (geometry-labs/fusion-cryptography) is pure Python and has no FFI of its own;
 * its boundary for this path is the module fusion.fusion.  Each entry point below names the
 * reference function(s) it replaces (paths relative to the reference root).  The Python
 * drop-in (fusion_cryptography_b200/fusion/fusion.py) binds these symbols with ctypes; the
 * binding a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void* (NULL = default
 *     stream); every call is asynchronous on that stream unless stated otherwise;
 *   - device buffers hold int32 canonical centred coefficients, row-major [batch][rows][d]
 *     (d = 256, rows = ell = 83 at secpar 256; d = 64, ell = 195 at secpar 128);
 *   - "hat" (NTT-domain) buffers are in the reference's bit-reversed order
 *     (algebra/ntt.py:216-291);
 *   - return value 0 = success; negative = error, text via fz_last_error().
 *   - there is no CPU fallback: every function needs a CUDA device.
 */
#ifndef FUSION_B200_H
#define FUSION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fz_ctx fz_ctx;

/* Context: one per (process, device, security level).  Uploads the per-(q, d) twiddle tables
 * (replaces the per-call rebuild at algebra/polynomials.py:396-397, 414-417) and owns scratch.
 * secpar is 128 or 256 (fusion/fusion.py:71-121). */
int fz_init(int device, int secpar, fz_ctx** out);
int fz_destroy(fz_ctx* ctx);
const char* fz_last_error(void);
int fz_version(void);
/* named integer parameter of the level ("d", "ell", "capacity", "n_ch", "n_ag", ...); -1 if unknown */
long long fz_param(const fz_ctx* ctx, const char* name);

/* algebra/polynomials.py:394-411 transform(coefficient -> NTT) == algebra/ntt.py:216-291
 * cooley_tukey_ntt, on n_polys polynomials; in may equal out.  Input: any int32. */
int fz_ntt_fwd(fz_ctx* ctx, const int32_t* in, int32_t* out, size_t n_polys, void* stream);
/* algebra/polynomials.py:412-431 transform(NTT -> coefficient) == algebra/ntt.py:294-377
 * gentleman_sande_intt (incl. the n^-1 scaling). */
int fz_ntt_inv(fz_ctx* ctx, const int32_t* in, int32_t* out, size_t n_polys, void* stream);
/* PolynomialNTTRepresentation.__mul__ (algebra/polynomials.py:341-385) on n_polys polynomials */
int fz_pointwise_mul(fz_ctx* ctx, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, void* stream);
/* __mul__ then __add__ (algebra/polynomials.py:341-385, 282-318): out = a*b + c */
int fz_pointwise_muladd(fz_ctx* ctx, const int32_t* a, const int32_t* b, const int32_t* c, int32_t* out, size_t n_polys, void* stream);
/* algebra/ntt.py:380-484 ntt_poly_mult (result only): out = INTT(NTT(a) o NTT(b)), fused */
int fz_polymul(fz_ctx* ctx, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, void* stream);

/* GeneralMatrix.__mul__ matmul branch (algebra/matrices.py:121-131) for the scheme's shapes:
 * out[b] = A(1 x ell) . s[b](ell x 1);  A: [ell][d], s: [batch][ell][d], out: [batch][d].
 * Used by keygen (fusion.py:369-370) and verify (fusion.py:715-717). */
int fz_matvec(fz_ctx* ctx, const int32_t* A, const int32_t* s, int32_t* out, size_t batch, void* stream);
/* fusion.py:557: sig[b] = sk_l[b] * c_hat[b] + sk_r[b]  (matrices.py:109-114 then :74-91) */
int fz_sign_mac(fz_ctx* ctx, const int32_t* sk_l, const int32_t* sk_r, const int32_t* c_hat, int32_t* sig, size_t batch, void* stream);

/* Host-buffer convenience used by the end-to-end measurement: copies n_polys polynomials from
 * pinned or pageable host memory, transforms them (inverse = 0/1) and copies the result back,
 * chunked and double-buffered on internal streams.  Synchronous. */
int fz_ntt_host(fz_ctx* ctx, int inverse, const int32_t* host_in, int32_t* host_out, size_t n_polys);
/* Same pipeline for the bench pass: host a -> NTT -> pointwise product with the DEVICE-resident
 * NTT-domain operand b_hat -> INTT -> host out (transform(transform(a) * b_hat)).  Synchronous. */
int fz_ntt_pass_host(fz_ctx* ctx, const int32_t* host_a, const int32_t* dev_bhat, int32_t* host_out, size_t n_polys);

#ifdef __cplusplus
}
#endif
#endif /* FUSION_B200_H */
