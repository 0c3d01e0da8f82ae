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
    int sort_smem_max;  // dynamic shared memory sort_kernel may use on this device (opt-in made in scheme_init)
    int hash_ag_impl;   // 0 auto (ring-fed chains: a warp per aggregate up to 1,776 aggregates, two lanes per aggregate up to 40,960, a thread per aggregate above), 2 thread, 3 warp, 4 pair
    int buffered_hash;  // 1: format text into HBM scratch first (fz_hash.cu); 0 (default): streaming generator (fz_hashstream.cu)
};

// A kernel launch of the host code.  The CPU test suite compiles the kernels and this launch code with g++ on a small CUDA-on-CPU emulator
// (tests/host/cuda_emu.h), which defines FZ_LAUNCH before this header is read; in the library it is the ordinary launch.
#ifndef FZ_LAUNCH
#define FZ_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

namespace fz {

void set_error(const char* fmt, ...);
int check(cudaError_t e, const char* what);
// Kernel attributes (dynamic shared-memory opt-in, carve-out) are PER DEVICE: true the first time `slot` is asked about on the context's device.
// Setting an attribute twice is harmless, so two host threads racing here only repeat the call.
enum AttrSlot { ATTR_FORMAT = 0, ATTR_AGW_FORMAT = 1, ATTR_WARP_HASH = 2, ATTR_CHAIN_CARVEOUT = 3 };
bool attr_needed(const fz_ctx* ctx, int slot);

// fz_ntt.cu
int ntt_init_tables(fz_ctx* ctx);
int launch_ntt(const fz_ctx* ctx, int inverse, const int32_t* in, int32_t* out, size_t n_polys, cudaStream_t st);
int launch_polymul(const fz_ctx* ctx, const int32_t* a, const int32_t* b, int32_t* out, size_t n_polys, cudaStream_t st);
int launch_ntt_pass(const fz_ctx* ctx, const int32_t* a, const int32_t* bhat, int32_t* out, size_t n_polys, cudaStream_t st);
int launch_pointwise_add(const fz_ctx* ctx, const int32_t* a, const int32_t* b, int32_t* out, size_t n_coef, cudaStream_t st);
int launch_pointwise(const fz_ctx* ctx, const int32_t* a, const int32_t* b, const int32_t* c, int32_t* out, size_t n_coef, cudaStream_t st);

// fz_stream.cu
int launch_matvec(const fz_ctx* ctx, const int32_t* A, const int32_t* s, size_t s_row_stride, size_t s_batch_stride, int32_t* out, size_t batch,
                  cudaStream_t st);
int launch_sign_mac(const fz_ctx* ctx, const int32_t* sk_l, const int32_t* sk_r, size_t sk_row_stride, size_t sk_batch_stride, const int32_t* c_hat,
                    int32_t* sig, size_t batch, cudaStream_t st);

int launch_sign_rows(const fz_ctx* ctx, const int32_t* sk_hat, const int32_t* c_hat, int32_t* sig, size_t batch, cudaStream_t st);
int launch_matvec_rows(const fz_ctx* ctx, const int32_t* A, int32_t* asum, const int32_t* s, size_t s_batch_stride, int32_t* out, size_t batch,
                       cudaStream_t st);

// fz_hash.cu
int hash_init_tables(fz_ctx* ctx);
int launch_sponge_bytes(const fz_ctx* ctx, const uint8_t* data, const unsigned long long* offs, size_t n, uint32_t prefix, int prefix_len,
                        uint8_t suffix, uint8_t* out, size_t out_stride, uint32_t out_len, cudaStream_t st);
int launch_format(const fz_ctx* ctx, int mode, const int32_t* vk_l, const int32_t* vk_r, const uint8_t* digests, const int32_t* c_hat, size_t n,
                  uint8_t* out, size_t out_stride, uint32_t* out_len, cudaStream_t st);
int launch_shake_decode(const fz_ctx* ctx, const uint8_t* msg, size_t stride, const uint32_t* len, size_t n, int32_t* coef_out, cudaStream_t st);
int launch_decode(const fz_ctx* ctx, const uint8_t* bytes, size_t stride, uint32_t blen, size_t n, int weight, uint32_t bound, int bpc, int bpi,
                  int32_t* coef_out, cudaStream_t st);
int launch_ag_message(const fz_ctx* ctx, const uint8_t* pieces, size_t piece_stride, const uint32_t* piece_len, const uint32_t* order,
                      const unsigned long long* agg_off, size_t n_aggs, size_t n_total, unsigned long long* piece_pos, uint8_t* msgs, size_t msg_stride,
                      uint32_t* msg_len, cudaStream_t st);
int launch_shake_ag(const fz_ctx* ctx, const uint8_t* msgs, size_t msg_stride, const uint32_t* msg_len, const unsigned long long* agg_off, uint8_t* xof,
                    size_t n_aggs, cudaStream_t st);

// fz_hashstream.cu
int hashstream_init_tables(fz_ctx* ctx);
int launch_hash_ch_stream(const fz_ctx* ctx, const int32_t* vk_l, const int32_t* vk_r, const uint8_t* digests, size_t n, int32_t* coef_out,
                          cudaStream_t st);
size_t ag_state_bytes();
int launch_ag_state_init(const fz_ctx* ctx, void* state, const unsigned long long* agg_off, size_t n_aggs, cudaStream_t st);
int launch_hash_ag_absorb(const fz_ctx* ctx, void* state, const int32_t* vk_l_s, const int32_t* vk_r_s, const uint8_t* dg_s, const int32_t* c_hat_s,
                          const unsigned long long* win_off, size_t n_aggs, cudaStream_t st);
int launch_hash_ag_squeeze(const fz_ctx* ctx, void* state, const unsigned long long* win_off, size_t n_aggs, int32_t* coef_out, cudaStream_t st);

int launch_agw_absorb(const fz_ctx* ctx, void* state, const uint8_t* ring, size_t ring_cap, const unsigned long long* wend,
                      const unsigned long long* win_off, size_t n_aggs, cudaStream_t st);
int launch_agw_absorb_sliced(const fz_ctx* ctx, void* state, const uint8_t* ring, size_t ring_cap, const unsigned long long* wend, size_t n_aggs,
                             size_t first_active, size_t n_active, unsigned long long budget_blocks, cudaStream_t st);
int launch_agw_squeeze(const fz_ctx* ctx, void* state, const unsigned long long* win_off, size_t n_aggs, uint8_t* xof, size_t xof_cap, uint32_t* bpos_out,
                       cudaStream_t st);

// bit 63 of the per-aggregate message length written by the text stage: the text is complete
constexpr unsigned long long AGW_TEXT_COMPLETE = 1ULL << 63;
// bit 62: the text writer found that this window's text would overwrite ring bytes the aggregate's sponge has not absorbed yet (ring too small
// for the backlog): the absorb kernels poison the sponge (phase 0xDEAD) instead of hashing overwritten bytes
constexpr unsigned long long AGW_TEXT_OVERRUN = 1ULL << 62;
constexpr unsigned long long AGW_TEXT_FLAGS = AGW_TEXT_COMPLETE | AGW_TEXT_OVERRUN;
size_t ag_state_rpos_offset();
// fz_hash.cu (windowed text / decode)
size_t agw_max_piece();
int launch_agw_text(const fz_ctx* ctx, const int32_t* vk_l, const int32_t* vk_r, const uint8_t* digests, const int32_t* c_hat, const unsigned long long* win_off,
                    const uint32_t* agg_of, const unsigned long long* agg_off, size_t n_aggs, size_t n_win, uint32_t w0, unsigned long long* text_pos,
                    uint32_t* piece_len, unsigned long long* piece_off, uint8_t* int_text, uint32_t* int_len, unsigned long long* wend, uint8_t* ring,
                    size_t ring_cap, const void* state, cudaStream_t st);
size_t agw_int_text_stride();
int launch_agw_decode(const fz_ctx* ctx, const uint8_t* xof, size_t xof_cap, const uint32_t* bpos, const unsigned long long* win_off, const uint32_t* agg_of,
                      size_t n_win, int32_t* coef_out, cudaStream_t st);

// fz_sample.cu
int sample_init_tables(fz_ctx* ctx);
void set_sample_impl(int v);   // 0 general kernel, 1 dense kernel with the byte tile (default), 2 dense kernel with 16-byte stores
int launch_sample(const fz_ctx* ctx, int mode, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, uint32_t bound, int weight,
                  int32_t* out, cudaStream_t st);
int launch_sample_ring(int d, uint32_t q, int mode, const uint32_t* keys, int key_stride, const int32_t* klen, size_t n, uint32_t bound, int weight,
                       int32_t* out, cudaStream_t st);
int launch_seeds_to_keys(const fz_ctx* ctx, const unsigned long long* seeds, size_t n, uint32_t* keys, int32_t* klen, cudaStream_t st);
int launch_row_seeds_to_keys(const fz_ctx* ctx, const unsigned long long* seeds, size_t n, int rows, uint32_t* keys, int32_t* klen, cudaStream_t st);
int launch_replicate_rows(const fz_ctx* ctx, const int32_t* src, int32_t* dst, size_t batch, int rows, size_t src_batch_stride, cudaStream_t st);

// fz_scheme.cu
int scheme_init(fz_ctx* ctx);
int launch_sort(const fz_ctx* ctx, const int32_t* vk_l, const int32_t* vk_r, const uint32_t* sel, const unsigned long long* agg_off, size_t n_aggs,
                size_t max_keys, uint32_t* perm, cudaStream_t st);
int launch_gather_sorted(const fz_ctx* ctx, const int32_t* vk_l, const int32_t* vk_r, const uint8_t* digests, const uint32_t* sel, const uint32_t* perm,
                         size_t n_total, int32_t* vk_l_s, int32_t* vk_r_s, uint8_t* dg_s, cudaStream_t st);
int aggregate_split(const fz_ctx* ctx, size_t n_aggs, size_t max_keys);
int launch_aggregate(const fz_ctx* ctx, const int32_t* sigs, const uint32_t* perm, const int32_t* alpha, const unsigned long long* agg_off, size_t n_aggs,
                     int split, int32_t* part, int32_t* out, cudaStream_t st);
int launch_fold_partials(const fz_ctx* ctx, const int32_t* part, int32_t* out, size_t n_aggs, int n_parts, cudaStream_t st);
int launch_verify_target(const fz_ctx* ctx, const int32_t* vk_l_s, const int32_t* vk_r_s, const int32_t* c_hat, const int32_t* alpha,
                         const unsigned long long* agg_off, size_t n_aggs, int32_t* target, cudaStream_t st);
int launch_verify_finalize(const fz_ctx* ctx, const int32_t* target, const int32_t* observed, const int32_t* coef, size_t n_aggs, uint8_t* codes,
                           cudaStream_t st);

}  // namespace fz
