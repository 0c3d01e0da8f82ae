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
/* named integer option (every choice gives the same results; the options exist for the cross-checking tests and the measurements):
 *   "buffered_hash" = 1   the HBM-scratch hashing pipeline instead of the streaming one
 *   "hash_ag_impl"        layout of the serial SHAKE256 chains of hash_ag: 0 chosen by the number of aggregates (default), 2 a thread,
 *                         3 a warp, 4 a lane pair per aggregate
 *   "sample_impl"         kernel serving the dense secret-key sampler: 1 byte-tile kernel (default), 2 16-byte-store kernel, 0 the general kernel */
int fz_set_option(fz_ctx* ctx, const char* name, long long value);

/* algebra/polynomials.py:394-411 transform(coefficient -> NTT) == algebra/ntt.py:216-291
 * cooley_tukey_ntt, on n_polys polynomials; in may equal out.  Input: any int32. */
int fz_ntt_fwd(fz_ctx* ctx, const int32_t* in, int32_t* out, size_t n_polys, void* stream);
/* algebra/polynomials.py:412-431 transform(NTT -> coefficient) == algebra/ntt.py:294-377
 * gentleman_sande_intt (incl. the n^-1 scaling). */
int fz_ntt_inv(fz_ctx* ctx, const int32_t* in, int32_t* out, size_t n_polys, void* stream);
/* PolynomialNTTRepresentation.__mul__ (algebra/polynomials.py:341-385) on n_polys polynomials */
int fz_pointwise_mul(fz_ctx* ctx, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, void* stream);
/* __add__ of either representation (algebra/polynomials.py:114-148, 282-318): out = a + b */
int fz_pointwise_add(fz_ctx* ctx, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, void* stream);
/* __mul__ then __add__ (algebra/polynomials.py:341-385, 282-318): out = a*b + c */
int fz_pointwise_muladd(fz_ctx* ctx, const int32_t* a, const int32_t* b, const int32_t* c, int32_t* out, size_t n_polys, void* stream);
/* algebra/ntt.py:380-484 ntt_poly_mult (result only): out = INTT(NTT(a) o NTT(b)), fused */
int fz_polymul(fz_ctx* ctx, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, void* stream);
/* transform(transform(a) * b_hat) (algebra/polynomials.py:391-433 around :341-385) in one pass over HBM: a, out in the coefficient
 * domain, b_hat in the NTT domain; bit-identical to fz_ntt_fwd -> fz_pointwise_mul -> fz_ntt_inv */
int fz_ntt_pass(fz_ctx* ctx, const int32_t* a, const int32_t* b_hat, int32_t* out, size_t n_polys, void* stream);

/* GeneralMatrix.__mul__ matmul branch (algebra/matrices.py:121-131) for the scheme's shapes:
 * out[b] = A(1 x ell) . s[b](ell x 1);  A: [ell][d], s: [batch][ell][d], out: [batch][d].
 * Used by keygen (fusion.py:369-370) and verify (fusion.py:715-717). */
int fz_matvec(fz_ctx* ctx, const int32_t* A, const int32_t* s, int32_t* out, size_t batch, void* stream);
/* fusion.py:557: sig[b] = sk_l[b] * c_hat[b] + sk_r[b]  (matrices.py:109-114 then :74-91) */
int fz_sign_mac(fz_ctx* ctx, const int32_t* sk_l, const int32_t* sk_r, const int32_t* c_hat, int32_t* sig, size_t batch, void* stream);

/* ---- hashing: SHA3-256 / SHAKE256 (hashlib call sites fusion/fusion.py:408, 419, 591) ------------------- */
/* generic batched hashes over device byte strings data[offs[i]..offs[i+1]); out32: [n][32]; out: [n][out_len] */
int fz_sha3_256(fz_ctx* ctx, const uint8_t* data, const uint64_t* offs, size_t n, uint8_t* out32, void* stream);
int fz_shake256(fz_ctx* ctx, const uint8_t* data, const uint64_t* offs, size_t n, uint8_t* out, size_t out_len, void* stream);
/* fusion.py:405-409 hash_message_to_int: digests[i] = SHA3-256(dst "," msg_i); the int is the little-endian value of the 32 bytes */
int fz_hash_msg(fz_ctx* ctx, const uint8_t* msgs, const uint64_t* offs, size_t n, uint8_t* digests, void* stream);
/* str(OneTimeVerificationKey) / str(SignatureChallenge) as hashed at fusion.py:416-418 (mode 0: dst "," VK "," INT) and
 * fusion.py:589 (mode 1: one "(VK, INT, CH)" list element).  out: [n][fz_text_stride(mode)] bytes, out_len[n]. */
size_t fz_text_stride(const fz_ctx* ctx, int mode);
int fz_format_text(fz_ctx* ctx, int mode, const int32_t* vk_l, const int32_t* vk_r, const uint8_t* digests, const int32_t* c_hat, size_t n,
                   uint8_t* out, uint32_t* out_len, void* stream);
/* fusion.py:422-481 decode_bytes_to_polynomial_coefficients on n byte strings bytes[i*stride .. +blen); returns -3
 * ("Too few bytes ...", the reference's ValueError) when blen is too small.  coef_out: [n][d]. */
int fz_decode_poly(fz_ctx* ctx, const uint8_t* bytes, size_t stride, size_t blen, size_t n, long long log2_bias, long long norm_bound, int weight_bound,
                   int32_t* coef_out, void* stream);
/* fusion.py:511-531 hash_ch (prehash digests from fz_hash_msg): c_hat_out: [n][d] */
size_t fz_hash_ch_workspace(const fz_ctx* ctx, size_t n);
int fz_hash_ch(fz_ctx* ctx, const int32_t* vk_l, const int32_t* vk_r, const uint8_t* digests, size_t n, int32_t* c_hat_out, void* ws, size_t ws_bytes,
               void* stream);
/* fusion.py:632-652 hash_ag for n_aggs aggregates; inputs already sorted (fz_sort_keys + fz_gather_sorted); aggregate j
 * owns positions [agg_off[j], agg_off[j+1]) (device array of n_aggs + 1 offsets).  alpha_hat_out: [n_total][d]. */
size_t fz_hash_ag_workspace(const fz_ctx* ctx, size_t n_total, size_t n_aggs, size_t max_keys);
int fz_hash_ag(fz_ctx* ctx, const int32_t* vk_l_s, const int32_t* vk_r_s, const uint8_t* dg_s, const int32_t* c_hat_s, const uint64_t* agg_off,
               size_t n_aggs, size_t n_total, size_t max_keys, int32_t* alpha_hat_out, void* ws, size_t ws_bytes, void* stream);

/* Windowed form of the same hash (one serial SHAKE256 per aggregate, fusion.py:591): the sponge of every aggregate is a
 * resumable record in `state` (n_aggs * fz_hash_ag_state_bytes() bytes of device memory), fed a window of the next sorted
 * signers of every aggregate per call -- win_off[j]..win_off[j+1] index the window arrays -- so that all aggregates of a
 * batch advance together however many signers fit in memory at once.  begin -> absorb* -> squeeze*. */
size_t fz_hash_ag_state_bytes(const fz_ctx* ctx);
int fz_hash_ag_begin(fz_ctx* ctx, void* state, const uint64_t* agg_off, size_t n_aggs, void* stream);
int fz_hash_ag_absorb(fz_ctx* ctx, void* state, const int32_t* vk_l_w, const int32_t* vk_r_w, const uint8_t* dg_w, const int32_t* c_hat_w,
                      const uint64_t* win_off, size_t n_aggs, void* stream);
int fz_hash_ag_squeeze(fz_ctx* ctx, void* state, const uint64_t* win_off, size_t n_aggs, size_t n_window, int32_t* alpha_hat_w, void* stream);
/* The same windowed hash with everything but the permutations taken off the serial chain: the text a window adds to each
 * message (str() of fusion.py:586-590) is written, a warp per signer, into a byte ring per aggregate (n_aggs * ring_cap
 * bytes, ring_cap = fz_hash_ag_ring_bytes(P) for windows of <= P signers per aggregate; text_pos: n_aggs running byte
 * counts, zero-initialised; wend_out: n_aggs message lengths after this window); the chain absorbs ring blocks up to wend;
 * squeezed blocks go to a per-aggregate XOF buffer (xof_cap = fz_hash_ag_xof_bytes(P)) which fz_hash_ag_decode turns into
 * the window's coefficients (fusion.py:594-629 + NTT).  Window w0 holds the sorted positions w0.. of every aggregate;
 * agg_of[i] is the aggregate of window entry i.  Text of window w+2 may only be written once window w has been absorbed.
 * `state` (optional, the sponge records of fz_hash_ag_begin) lets the writer check that discipline itself: if the new text would reach
 * back over bytes an aggregate's sponge has not absorbed yet, bit 62 of wend_out is set and the absorb calls poison that sponge
 * (it never closes; fz_hash_ag_squeeze_xof produces nothing for it) instead of hashing overwritten bytes. */
size_t fz_hash_ag_ring_bytes(const fz_ctx* ctx, size_t max_window);
size_t fz_hash_ag_xof_bytes(const fz_ctx* ctx, size_t max_window);
size_t fz_hash_ag_text_workspace(const fz_ctx* ctx, size_t n_window);
int fz_hash_ag_text(fz_ctx* ctx, const int32_t* vk_l_w, const int32_t* vk_r_w, const uint8_t* dg_w, const int32_t* c_hat_w, const uint64_t* win_off,
                    const uint32_t* agg_of, const uint64_t* agg_off, size_t n_aggs, size_t n_window, uint32_t w0, uint64_t* text_pos, uint8_t* ring,
                    size_t ring_cap, const void* state, uint64_t* wend_out, void* ws, size_t ws_bytes, void* stream);
int fz_hash_ag_absorb_ring(fz_ctx* ctx, void* state, const uint8_t* ring, size_t ring_cap, const uint64_t* wend, const uint64_t* win_off, size_t n_aggs,
                           void* stream);
/* time-sliced absorb (same sponge, fusion.py:591): one launch advances n_active aggregates, first_active.. modulo n_aggs, by at
 * most budget_blocks blocks each; wend from fz_hash_ag_text (bit 63 set once an aggregate's text is complete) */
int fz_hash_ag_absorb_sliced(fz_ctx* ctx, void* state, const uint8_t* ring, size_t ring_cap, const uint64_t* wend, size_t n_aggs, size_t first_active,
                             size_t n_active, uint64_t budget_blocks, void* stream);
int fz_hash_ag_squeeze_xof(fz_ctx* ctx, void* state, const uint64_t* win_off, size_t n_aggs, uint8_t* xof, size_t xof_cap, uint32_t* bpos_out,
                           void* stream);
int fz_hash_ag_decode(fz_ctx* ctx, const uint8_t* xof, size_t xof_cap, const uint32_t* bpos, const uint64_t* win_off, const uint32_t* agg_of,
                      size_t n_window, int32_t* alpha_hat_w, void* stream);

/* ---- aggregate / verify ------------------------------------------------------------------------------- */
/* sorted(zip(...), key=lambda x: str(x[0])) at fusion.py:661-663 and :693.  sel (optional) maps a position to a row of
 * vk_l / vk_r; perm[agg_off[j] + rank] = position.  Stable. */
int fz_sort_keys(fz_ctx* ctx, const int32_t* vk_l, const int32_t* vk_r, const uint32_t* sel, const uint64_t* agg_off, size_t n_aggs, size_t max_keys,
                 uint32_t* perm, void* stream);
int fz_gather_sorted(fz_ctx* ctx, const int32_t* vk_l, const int32_t* vk_r, const uint8_t* digests, const uint32_t* sel, const uint32_t* perm,
                     size_t n_total, int32_t* vk_l_s, int32_t* vk_r_s, uint8_t* dg_s, void* stream);
/* fusion.py:670-677: out[j] = sum_p sigs[perm[p]] * alpha[p]; sigs: [n_sigs][ell][d], out: [n_aggs][ell][d] */
size_t fz_aggregate_sum_workspace(const fz_ctx* ctx, size_t n_aggs, size_t max_keys);
int fz_aggregate_sum(fz_ctx* ctx, const int32_t* sigs, const uint32_t* perm, const int32_t* alpha, const uint64_t* agg_off, size_t n_aggs, size_t max_keys,
                     int32_t* out, void* ws, size_t ws_bytes, void* stream);
/* modular sum of partial aggregates [n_aggs][n_parts][ell][d] (the multi-GPU reduce of partial sums) */
int fz_fold_partials(fz_ctx* ctx, const int32_t* parts, int n_parts, size_t n_aggs, int32_t* out, void* stream);
/* fusion.py:690-692, 706-728 on sorted inputs: codes[j] = 0 (True, "") | 3 "Target doesn't match image of aggregate
 * signature." | 4 "Norm of aggregate signature too large."  (reasons 1, 2 are argument checks done by the caller) */
size_t fz_verify_check_workspace(const fz_ctx* ctx, size_t n_aggs);
int fz_verify_check(fz_ctx* ctx, const int32_t* A, const int32_t* vk_l_s, const int32_t* vk_r_s, const int32_t* c_hat_s, const int32_t* alpha_s,
                    const int32_t* agg, const uint64_t* agg_off, size_t n_aggs, uint8_t* codes, void* ws, size_t ws_bytes, void* stream);

/* The two halves of fz_verify_check for windowed batches: the target sum of fusion.py:706-714 over one window of sorted
 * signers (partials add with fz_pointwise_add), then fusion.py:690-692, 715-728 on the complete target. */
int fz_verify_target(fz_ctx* ctx, const int32_t* vk_l_w, const int32_t* vk_r_w, const int32_t* c_hat_w, const int32_t* alpha_w, const uint64_t* win_off,
                     size_t n_aggs, int32_t* target_w, void* stream);
size_t fz_verify_finish_workspace(const fz_ctx* ctx, size_t n_aggs);
int fz_verify_finish(fz_ctx* ctx, const int32_t* A, const int32_t* target, const int32_t* agg, size_t n_aggs, uint8_t* codes, void* ws, size_t ws_bytes,
                     void* stream);

/* ---- seeded samplers and keygen (CPython random: MT19937 + randrange; algebra/polynomials.py:436-488) ----- */
/* keys: [n][key_stride] little-endian 32-bit limbs of abs(seed), klen[n] in 1..8 */
int fz_sample_coef(fz_ctx* ctx, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, int norm_bound, int weight_bound, int32_t* out,
                   void* stream);
int fz_sample_ntt(fz_ctx* ctx, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, int32_t* out, void* stream);
/* fusion.py:338-373 keygen for n u64 seeds; sk_l / sk_r ([n][ell][d]) and sk_hat ([n][2][d]) are optional outputs */
size_t fz_keygen_workspace(const fz_ctx* ctx, size_t n);
int fz_keygen(fz_ctx* ctx, const int32_t* A, const uint64_t* seeds, size_t n, int32_t* sk_l, int32_t* sk_r, int32_t* sk_hat, int32_t* vk_l, int32_t* vk_r,
              void* ws, size_t ws_bytes, void* stream);
/* fusion.py:338-373 in its general ell-row form: row k of sk_l[i] / sk_r[i] is sample_polynomial_coefficient_representation
 * (algebra/polynomials.py:436-467) seeded with row_seeds[i][0][k] / row_seeds[i][1][k] (u64, [n][2][ell]), then the 2*ell transforms of
 * fusion.py:364-368 and the two ell-term products of :369-370 (fz_matvec).  row_seeds[i][0][*] = s, row_seeds[i][1][*] = s + 1 reproduces
 * keygen(params, s) without using the fact that a seeded key has identical rows.  sk_l, sk_r: [n][ell][d] (required), vk_l, vk_r: [n][d]. */
size_t fz_keygen_rows_workspace(const fz_ctx* ctx, size_t n);
int fz_keygen_rows(fz_ctx* ctx, const int32_t* A, const uint64_t* row_seeds, size_t n, int32_t* sk_l, int32_t* sk_r, int32_t* vk_l, int32_t* vk_r, void* ws,
                   size_t ws_bytes, void* stream);
/* keygen (fusion.py:338-373) + sign (fusion.py:534-557) of n one-time keys in one call: key i comes from seeds[i] and signs
 * the message bytes msgs[offs[i]..offs[i+1]) (device memory).  vk_l / vk_r: [n][d], sig: [n][ell][d]. */
size_t fz_keygen_sign_workspace(const fz_ctx* ctx, size_t n);
int fz_keygen_sign(fz_ctx* ctx, const int32_t* A, const uint64_t* seeds, const uint8_t* msgs, const uint64_t* offs, size_t n, int32_t* vk_l, int32_t* vk_r,
                   int32_t* sig, void* ws, size_t ws_bytes, void* stream);
/* fusion.py:557 for keys kept as their one distinct row (sk_hat from fz_keygen): sig: [batch][ell][d] */
int fz_sign_mac_rows(fz_ctx* ctx, const int32_t* sk_hat, const int32_t* c_hat, int32_t* sig, size_t batch, void* stream);

/* Host-buffer convenience used by the end-to-end measurement: copies n_polys polynomials from
 * pinned or pageable host memory, transforms them (inverse = 0/1) and copies the result back,
 * chunked and double-buffered on internal streams.  Synchronous. */
int fz_ntt_host(fz_ctx* ctx, int inverse, const int32_t* host_in, int32_t* host_out, size_t n_polys);
/* Same pipeline for the bench pass: host a -> NTT -> pointwise product with the DEVICE-resident
 * NTT-domain operand b_hat -> INTT -> host out (transform(transform(a) * b_hat)).  Synchronous. */
int fz_ntt_pass_host(fz_ctx* ctx, const int32_t* host_a, const int32_t* dev_bhat, int32_t* host_out, size_t n_polys);

/* ---- any ring Z_q[X]/(X^d + 1) of the reference's algebra layer ------------------------------------------------------
 * The scheme's two parameter sets have compile-time kernels (fz_ntt_* above).  A fz_ring computes the same transforms and
 * pointwise operations for ANY odd prime modulus q < 2^31 and degree d = 2^logd (1 <= logd <= 14) with CALLER-SUPPLIED
 * bit-reversed power tables, which is the actual contract of cooley_tukey_ntt / gentleman_sande_intt: they take the table
 * as an argument (algebra/ntt.py:216-219, :294-297) and the reference's own unit tests run them on moduli below 2^17 and
 * degrees 4..64 (tests/test_ntt.py:18-50).  Tables are HOST arrays of d values (entry 0 is never read by the butterflies);
 * either (or both, for pointwise operations and sampling only) may be NULL.  Results are canonical centred residues (algebra/ntt.py:93-123). */
typedef struct fz_ring fz_ring;
int fz_ring_create(int device, uint32_t q, int logd, const uint32_t* fwd_table, const uint32_t* inv_table, fz_ring** out);
int fz_ring_destroy(fz_ring* ring);
/* algebra/ntt.py:271-291 (cooley_tukey_ntt): natural order in, bit-reversed order out; in == out allowed */
int fz_ring_ntt_fwd(fz_ring* ring, const int32_t* in, int32_t* out, size_t n_polys, void* stream);
/* algebra/ntt.py:350-377 (gentleman_sande_intt) including the n^-1 = pow(d, q - 2, q) scaling of :353, :373-376 */
int fz_ring_ntt_inv(fz_ring* ring, const int32_t* in, int32_t* out, size_t n_polys, void* stream);
/* algebra/polynomials.py:341-385 / :282-318 (and :114-148): cent(x * y), cent(x + y) coefficient by coefficient */
int fz_ring_pointwise_mul(fz_ring* ring, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, void* stream);
int fz_ring_pointwise_add(fz_ring* ring, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, void* stream);
/* algebra/polynomials.py:436-467 / :470-488 with a seed, for this ring (same key layout as fz_sample_coef / fz_sample_ntt) */
int fz_ring_sample_coef(fz_ring* ring, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, int norm_bound, int weight_bound, int32_t* out,
                        void* stream);
int fz_ring_sample_ntt(fz_ring* ring, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, int32_t* out, void* stream);
/* fusion/fusion.py:422-481 (decode_bytes_to_polynomial_coefficients) for this ring: bytes [n][stride] (device), the first blen of each row decoded;
 * bytes_per_coefficient / bytes_per_index as :433-434 computes them; -3 = the reference's "Too few bytes to decode polynomial" ValueError */
int fz_ring_decode_poly(fz_ring* ring, const uint8_t* bytes, size_t stride, size_t blen, size_t n, long long norm_bound, int weight_bound,
                        int bytes_per_coefficient, int bytes_per_index, int32_t* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FUSION_B200_H */
