// Drives the product's ring-fed aggregation-hash pipeline on the CUDA-on-CPU emulator exactly as Engine._hash_windows / _squeeze_windows
// (fusion_cryptography_b200/batch.py) drive it on the GPU: per window of sorted signer positions fz's own launch_agw_text (prehash text, piece
// lengths, per-aggregate layout, format into the byte rings) and launch_agw_absorb (the sponge chains, in the layout asked for), then per window
// launch_agw_squeeze and launch_agw_decode.  The kernels and the launch code are the library's source files, compiled by g++ (hash_emu.cpp,
// hashstream_emu.cpp).  Test infrastructure only (tests/test_pipeline_emu.py).
#include "../cuda_emu.h"

#include <cstdarg>
#include <cstdio>
#include <vector>

#include "../../../include/fusion_b200.h"
#include "../../../fusion_cryptography_b200/csrc/fz_internal.h"

static int ceil_div(int a, int b) { return (a + b - 1) / b; }

extern "C" {

const char* h_last_error() { return fz_last_error(); }

// Inputs are per SORTED position (agg_off[n_aggs] of them): vk_l_s, vk_r_s, c_hat_s [total][d]; digests_s [total][32].  positions_per_window = P.
// layout: 2 a thread, 3 a warp, 4 a lane pair per aggregate.  alpha_coef_out: [total][d], the aggregation coefficients before their transform.
int h_agw_pipeline(int secpar, int layout, size_t n_aggs, const unsigned long long* agg_off, const int32_t* vk_l_s, const int32_t* vk_r_s,
                   const uint8_t* digests_s, const int32_t* c_hat_s, int P, int32_t* alpha_coef_out) {
    using namespace fz;
    fz_ctx c;
    memset(&c, 0, sizeof(c));
    c.secpar = secpar;
    if (secpar == 128) { c.d = 64; c.logd = 6; c.ell = 195; c.capacity = 1796; c.omega_ch = 27; c.omega_ag = 35; }
    else { c.d = 256; c.logd = 8; c.ell = 83; c.capacity = 2818; c.omega_ch = 60; c.omega_ag = 60; }
    c.bytes_coef = ceil_div(1 + secpar, 8);
    c.bytes_idx = ceil_div(c.logd + secpar, 8);
    c.n_ag = ceil_div(c.omega_ag, 8) + (c.bytes_coef + c.bytes_idx) * c.omega_ag;      // as fz_init derives them (fusion.py:123-141, 583-585)
    c.sm_count = 2;
    c.hash_ag_impl = layout;
    if (hash_init_tables(&c) || hashstream_init_tables(&c)) return -1;
    const int d = c.d;
    const size_t total = (size_t)agg_off[n_aggs];
    size_t max_count = 0;
    for (size_t j = 0; j < n_aggs; ++j) max_count = std::max(max_count, (size_t)(agg_off[j + 1] - agg_off[j]));

    std::vector<unsigned long long> state((n_aggs + 1) * ag_state_bytes() / 8, 0xA5A5A5A5A5A5A5A5ULL);
    if (launch_ag_state_init(&c, state.data(), agg_off, n_aggs, nullptr)) return -2;
    const size_t need = 2 * (size_t)P * agw_max_piece() + 2 * 272;                      // fz_hash_ag_ring_bytes
    const size_t ring_cap = (need + 271) / 272 * 272;
    std::vector<uint8_t> ring(n_aggs * ring_cap, 0x5A);
    std::vector<unsigned long long> text_pos(n_aggs, 0);

    struct Win { std::vector<unsigned long long> win_off; std::vector<uint32_t> agg_of; std::vector<size_t> pos; };
    std::vector<Win> wins;
    for (size_t w = 0; w < std::max<size_t>(max_count, 1); w += (size_t)P) {
        Win W;
        W.win_off.assign(n_aggs + 1, 0);
        for (size_t j = 0; j < n_aggs; ++j) {
            const size_t cnt = (size_t)(agg_off[j + 1] - agg_off[j]);
            const size_t take = cnt > w ? std::min<size_t>((size_t)P, cnt - w) : 0;
            for (size_t r = 0; r < take; ++r) { W.agg_of.push_back((uint32_t)j); W.pos.push_back((size_t)agg_off[j] + w + r); }
            W.win_off[j + 1] = W.win_off[j] + take;
        }
        const size_t n_win = W.pos.size();
        std::vector<int32_t> vl(n_win * d + 4), vr(n_win * d + 4), ch(n_win * d + 4);
        std::vector<uint8_t> dg(n_win * 32 + 32);
        for (size_t i = 0; i < n_win; ++i) {
            memcpy(&vl[i * d], vk_l_s + W.pos[i] * d, (size_t)d * 4);
            memcpy(&vr[i * d], vk_r_s + W.pos[i] * d, (size_t)d * 4);
            memcpy(&ch[i * d], c_hat_s + W.pos[i] * d, (size_t)d * 4);
            memcpy(&dg[i * 32], digests_s + W.pos[i] * 32, 32);
        }
        std::vector<uint32_t> piece_len(n_win + 1), int_len(n_win + 1);
        std::vector<unsigned long long> piece_off(n_win + 1), wend(n_aggs + 1);
        std::vector<uint8_t> int_text((n_win + 1) * agw_int_text_stride());
        if (launch_agw_text(&c, vl.data(), vr.data(), dg.data(), ch.data(), W.win_off.data(), W.agg_of.data(), agg_off, n_aggs, n_win, (uint32_t)w,
                            text_pos.data(), piece_len.data(), piece_off.data(), int_text.data(), int_len.data(), wend.data(), ring.data(), ring_cap,
                            state.data(), nullptr))
            return -3;
        if (launch_agw_absorb(&c, state.data(), ring.data(), ring_cap, wend.data(), W.win_off.data(), n_aggs, nullptr)) return -4;
        wins.push_back(std::move(W));
    }
    // every sponge must have closed its message (phase word 1), as Engine._hash_windows checks
    for (size_t j = 0; j < n_aggs; ++j)
        if (reinterpret_cast<const uint32_t*>(state.data() + j * ag_state_bytes() / 8)[107] != 1) { set_error("sponge %zu did not close", j); return -5; }

    const size_t xof_cap = (((size_t)P * (size_t)c.n_ag + 2 * 136 + 32) + 15) & ~(size_t)15;   // fz_hash_ag_xof_bytes
    for (const Win& W : wins) {
        const size_t n_win = W.pos.size();
        std::vector<uint8_t> xof(n_aggs * xof_cap + 64, 0xC3);
        std::vector<uint32_t> bpos(n_aggs + 1, 0);
        if (launch_agw_squeeze(&c, state.data(), W.win_off.data(), n_aggs, xof.data(), xof_cap, bpos.data(), nullptr)) return -6;
        std::vector<int32_t> coef(n_win * d + 4, 12345);
        if (launch_agw_decode(&c, xof.data(), xof_cap, bpos.data(), W.win_off.data(), W.agg_of.data(), n_win, coef.data(), nullptr)) return -7;
        for (size_t i = 0; i < n_win; ++i) memcpy(alpha_coef_out + W.pos[i] * d, &coef[i * d], (size_t)d * 4);
    }
    (void)total;
    return 0;
}

}  // extern "C"
