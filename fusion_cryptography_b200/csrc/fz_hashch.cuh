// fz_hashch.cuh -- the challenge hash as ONE sponge thread runs it (fusion.py:412-531: hash_vk_and_int_to_bytes, decode_bytes_to_polynomial_coefficients
// with bound = 1, the coefficient part of hash_ch): the fixed-text tables, the token generator over a verification key and the prehash, and
// hash_ch_thread -- absorb the generated text, squeeze, decode the sparse ternary polynomial.  fz_hashstream.cu wraps it in hash_ch_stream_kernel
// (and reuses the generator for the aggregation hash); the CPU suite compiles the same functions for the host and compares the coefficients with
// the reference's (tests/test_hashch_host.py).  The device tables are DEFINED here: include this header in exactly one translation unit.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "fz_common.cuh"
#include "fz_keccak.cuh"
#include "fz_text.cuh"

namespace fz {

// decoder (shared with fz_hash.cu through this small duplicate of the constant table)
__constant__ uint32_t c_magic2[257];

struct SegTable {   // byte offsets / lengths into the fixed-text pool of one level; every segment starts 16-byte aligned
    uint32_t off[8], len[8];
};
// segment ids
enum { SEG_CH_OPEN = 0, SEG_MID = 1, SEG_CH_CLOSE = 2, SEG_AG_OPEN_FIRST = 3, SEG_AG_OPEN_NEXT = 4, SEG_AG_AFTER_VK = 5, SEG_AG_BEFORE_C = 6, SEG_AG_CLOSE = 7 };

__device__ __align__(16) uint8_t g_text[2][4096];
__constant__ SegTable c_seg[2];

// ---------------------------------------------------------------------------------- token generator
// A text is a list of items; items are fixed segments, value lists or the prehash decimal.
struct Gen {
    int item;        // index in the program
    int idx;         // token index inside the item
    int int_len;     // length of the decimal prehash of the current signer
    uint8_t intbuf[96];
};

__device__ __forceinline__ bool gen_segment(Gen& g, Queue& q, int level, int seg) {  // returns true when the item is finished
    const uint32_t len = c_seg[level].len[seg];
    const uint8_t* p = g_text[level] + c_seg[level].off[seg] + 16 * g.idx;
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const int n = (int)len - 16 * g.idx < 16 ? (int)len - 16 * g.idx : 16;
    q_push(q, ((unsigned long long)t.y << 32) | t.x, ((unsigned long long)t.w << 32) | t.z, n);
    ++g.idx;
    return 16 * g.idx >= (int)len;
}

// four values per step (one 16-byte load); d is a multiple of 4
__device__ __forceinline__ bool gen_values(Gen& g, Queue& q, const int32_t* __restrict__ vals, int d) {
    const int4 v = __ldg(reinterpret_cast<const int4*>(vals) + g.idx);
    unsigned long long lo, hi;
    int n = fmt_value(v.x, true, lo, hi); q_push(q, lo, hi, n);
    n = fmt_value(v.y, true, lo, hi); q_push(q, lo, hi, n);
    n = fmt_value(v.z, true, lo, hi); q_push(q, lo, hi, n);
    n = fmt_value(v.w, 4 * g.idx + 4 != d, lo, hi); q_push(q, lo, hi, n);
    ++g.idx;
    return 4 * g.idx >= d;
}

__device__ __forceinline__ bool gen_int(Gen& g, Queue& q, const uint8_t* __restrict__ digest) {
    if (g.idx == 0) g.int_len = fmt_le256(digest, g.intbuf);
    unsigned long long lo = 0, hi = 0;
    const int base = 16 * g.idx;
    const int n = g.int_len - base < 16 ? g.int_len - base : 16;
    for (int k = 0; k < 8; ++k) lo |= (unsigned long long)g.intbuf[base + k] << (8 * k);
    for (int k = 0; k < 8; ++k) hi |= (unsigned long long)g.intbuf[base + 8 + k] << (8 * k);
    q_push(q, lo, hi, n);
    ++g.idx;
    return 16 * g.idx >= g.int_len;
}

// ---------------------------------------------------------------------------------- hash_ch
// one thread per signer: text = dst "," str(vk) "," str(prehash)   (fusion.py:416-418), then squeeze + decode
constexpr int HS_THREADS = 64;

// what one thread of hash_ch_stream_kernel does for its signer; `my` is the thread's column of the CTA's staging tile (word w of the thread at my[w * HS_THREADS])
__device__ __forceinline__ void hash_ch_thread(const int32_t* __restrict__ vl, const int32_t* __restrict__ vr, const uint8_t* __restrict__ digest, int level, int d,
                                               int weight, int bpc, int bpi, uint32_t n_used, int32_t* __restrict__ o, uint32_t* my) {
    unsigned long long a[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) a[k] = 0;
    Queue q;
    q_reset(q);
    Gen g;
    g.item = 0; g.idx = 0; g.int_len = 0;
    // program: 0 SEG_CH_OPEN, 1 values(vk_l), 2 SEG_MID, 3 values(vk_r), 4 SEG_CH_CLOSE, 5 INT, 6 done
    for (;;) {
        while (q_bytes(q) < KECCAK_RATE && g.item < 6) {
            bool fin;
            switch (g.item) {
                case 0: fin = gen_segment(g, q, level, SEG_CH_OPEN); break;
                case 1: fin = gen_values(g, q, vl, d); break;
                case 2: fin = gen_segment(g, q, level, SEG_MID); break;
                case 3: fin = gen_values(g, q, vr, d); break;
                case 4: fin = gen_segment(g, q, level, SEG_CH_CLOSE); break;
                default: fin = gen_int(g, q, digest); break;
            }
            if (fin) { ++g.item; g.idx = 0; }
        }
        if (q_bytes(q) >= KECCAK_RATE) {
            q_absorb_full(q, a);
            keccak_f(a);
        } else {
            q_absorb_last(q, a);
            keccak_f(a);
            break;
        }
    }
    // ---- squeeze + decode (fusion.py:422-481 with bound = 1: every magnitude is 1, its bytes are skipped) ----
    // The squeezed stream has the same layout for every signer: bfs signum bytes, weight * bpc magnitude bytes, then one
    // bpi-byte big-endian integer per Fisher-Yates draw (i = d-1 down to weight+1), so all threads walk it in lock step.
    // Each block is staged in shared memory ([word][thread]: conflict-free) and every draw is reduced mod (i + 1) two
    // bytes per step: acc < 256, so acc * 2^16 + 2 bytes < 2^24 and the reciprocal multiplication is exact.
    int8_t c[256];
    for (int k = 0; k < d; ++k) c[k] = 0;
    const int bfs = (weight + 7) / 8;
    const int first = bfs + bpc * weight;                 // first byte of the index draws
    const int n_swaps = weight < d ? d - 1 - weight : 0;
    const uint8_t* myb = reinterpret_cast<const uint8_t*>(my);
    int t = 0, rem = bpi;                                 // current draw and its bytes still to come
    uint32_t acc = 0, m = (uint32_t)d, mg = c_magic2[m <= 256 ? m : 0];
    for (uint32_t base = 0; base < n_used; base += KECCAK_RATE) {
        if (base) keccak_f(a);
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) {
            my[(2 * k) * HS_THREADS] = (uint32_t)a[k];
            my[(2 * k + 1) * HS_THREADS] = (uint32_t)(a[k] >> 32);
        }
        // byte o of the block is at myb[(o >> 2) * 4 * HS_THREADS + (o & 3)]
#define BLKB(o) ((uint32_t)myb[((o) >> 2) * (4 * HS_THREADS) + ((o) & 3)])
        if (base == 0) {
            // signums: bit i (LSB first) of the big-endian integer of the first bfs bytes (bfs <= 32 < 136)
            for (int k = 0; k < weight && k < d; ++k) c[k] = (int8_t)(2 * (int)((BLKB(bfs - 1 - (k >> 3)) >> (k & 7)) & 1u) - 1);
        }
        int pos = (int)base < first ? first - (int)base : 0;   // skip signums / magnitudes
        const int lim = (int)(n_used - base) < KECCAK_RATE ? (int)(n_used - base) : KECCAK_RATE;
        while (pos < lim && t < n_swaps) {
            int nb = lim - pos < rem ? lim - pos : rem;
            rem -= nb;
            for (; nb >= 2; nb -= 2, pos += 2) {
                const uint32_t v = (acc << 16) | (BLKB(pos) << 8) | BLKB(pos + 1);
                acc = v - __umulhi(v, mg) * m;
            }
            if (nb) {
                const uint32_t v = (acc << 8) | BLKB(pos);
                acc = v - __umulhi(v, mg) * m;
                ++pos;
            }
            if (rem == 0) {
                const int i2 = d - 1 - t, j2 = (int)acc;
                const int8_t tmp = c[i2]; c[i2] = c[j2]; c[j2] = tmp;
                ++t; rem = bpi; acc = 0;
                m = (uint32_t)(d - t); mg = c_magic2[m];
            }
        }
#undef BLKB
        if (t >= n_swaps) break;
    }
    // a stream shorter than the draws need (never for hash_ch: n_used covers them) would leave truncated / empty draws
    for (; t < n_swaps; ++t) {
        const int i2 = d - 1 - t, j2 = rem == bpi ? 0 : (int)acc;
        const int8_t tmp = c[i2]; c[i2] = c[j2]; c[j2] = tmp;
        rem = bpi; acc = 0;
    }
    for (int k = 0; k < d; k += 4) *reinterpret_cast<int4*>(o + k) = make_int4(c[k], c[k + 1], c[k + 2], c[k + 3]);
}

// ---------------------------------------------------------------------------------- table construction (host)
static inline void hs_add_seg(std::vector<uint8_t>& pool, SegTable& t, int id, const std::string& s) {
    while (pool.size() % 16) pool.push_back(0);
    t.off[id] = (uint32_t)pool.size();
    t.len[id] = (uint32_t)s.size();
    pool.insert(pool.end(), s.begin(), s.end());
    while (pool.size() % 16) pool.push_back(0);
}

// magic[257]: reciprocals of the decoder's moduli; tabs[2] / pools[2][4096]: the fixed text of both levels ([0] secpar 128, [1] secpar 256).
// Returns false when a pool overflows.
static inline bool hashstream_build_tables(uint32_t* magic, SegTable* tabs, uint8_t* pools) {
    magic[0] = 0; magic[1] = 0;
    for (uint32_t m = 2; m <= 256; ++m) magic[m] = 0xFFFFFFFFu / m + 1u;
    const int degs[2] = {64, 256};
    const unsigned roots[2] = {23584283u, 3337519u}, iroots[2] = {540632852u, 1978410468u};
    for (int L = 0; L < 2; ++L) {
        char ph[192];
        snprintf(ph, sizeof(ph), "PolynomialNTTRepresentation(modulus=2147465729, degree=%d, root=%u, inv_root=%u, root_order=%d, values=[", degs[L], roots[L],
                 iroots[L], 2 * degs[L]);
        const std::string poly_head(ph);
        const std::string gm_head = "GeneralMatrix(elem_class=<class 'algebra.polynomials.PolynomialNTTRepresentation'>, matrix=[[";
        const std::string vk_head = "OneTimeVerificationKey(left_vk_hat=" + gm_head + poly_head;
        const std::string mid = "])]]), right_vk_hat=" + gm_head + poly_head;
        const std::string dst(1, (char)(L == 1 ? 3 : 1));
        std::vector<uint8_t> pool;
        memset(&tabs[L], 0, sizeof(SegTable));
        hs_add_seg(pool, tabs[L], SEG_CH_OPEN, dst + std::string(1, (char)1) + "," + vk_head);                          // fusion.py:416-418
        hs_add_seg(pool, tabs[L], SEG_MID, mid);
        hs_add_seg(pool, tabs[L], SEG_CH_CLOSE, "])]])),");
        hs_add_seg(pool, tabs[L], SEG_AG_OPEN_FIRST, dst + std::string(1, (char)2) + ",[(" + vk_head);                 // fusion.py:586-590
        hs_add_seg(pool, tabs[L], SEG_AG_OPEN_NEXT, ", (" + vk_head);
        hs_add_seg(pool, tabs[L], SEG_AG_AFTER_VK, "])]])), ");
        hs_add_seg(pool, tabs[L], SEG_AG_BEFORE_C, ", SignatureChallenge(c_hat=" + poly_head);
        hs_add_seg(pool, tabs[L], SEG_AG_CLOSE, "])))");
        if (pool.size() > 4096) return false;
        pool.resize(4096, 0);
        memcpy(pools + (size_t)L * 4096, pool.data(), 4096);
    }
    return true;
}

}  // namespace fz
