// fz_hashstream.cu -- scratch-free challenge / aggregation-coefficient derivation.
//
// hash_ch (fusion.py:511-531) and hash_ag (fusion.py:632-652) absorb the *text* of Python reprs
// (str(OneTimeVerificationKey), str(SignatureChallenge), decimal prehash).  The
// first implementation (fz_hash.cu) materialises that text in HBM (7-11 KB per signer) before hashing
// it, which caps how many aggregates can be in flight.  Here every sponge thread GENERATES the text it
// absorbs, 16 bytes (one value or one chunk of fixed text) at a time, straight from the int32
// coefficients: no intermediate buffers, no extra HBM traffic, and the squeezed bytes go straight into
// the sparse-ternary decoder.  hash_ag stays one serial sponge per aggregate (it is one SHAKE256 call
// in the reference); parallelism comes from many aggregates / signers in flight.
//
// Work per thread is a resumable token generator (phase, index) so that the threads of a warp can be
// at different places of their texts while the Keccak-f calls stay convergent:
//     loop: { fill: emit tokens until >= 136 bytes are buffered;  absorb 17 lanes;  permute }
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "fz_common.cuh"
#include "fz_internal.h"
#include "fz_keccak.cuh"
#include "fz_text.cuh"   // byte queue + decimal formatter of the sponge threads
#include "fz_hashch.cuh" // fixed-text tables, token generator, the challenge hash of one sponge thread, table construction
#include "fz_keccak_lanes.cuh"   // Keccak-f over a lane pair / over a warp
#include "fz_dec_stream.cuh"     // the decoder of the one-shot aggregation-hash kernels

namespace fz {

// ---------------------------------------------------------------------------------- hash_ch
// one thread per signer (hash_ch_thread, fz_hashch.cuh)
__global__ void __launch_bounds__(HS_THREADS) hash_ch_stream_kernel(const int32_t* __restrict__ vk_l, const int32_t* __restrict__ vk_r,
                                                                     const uint8_t* __restrict__ digests, size_t n, int level, int d, int weight, int bpc,
                                                                     int bpi, uint32_t n_used, int32_t* __restrict__ coef_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    __shared__ uint32_t blk_sm[2 * KECCAK_RATE_LANES * HS_THREADS];
    hash_ch_thread(vk_l + i * d, vk_r + i * d, digests + i * 32, level, d, weight, bpc, bpi, n_used, coef_out + i * d, blk_sm + threadIdx.x);
}

// ---------------------------------------------------------------------------------- hash_ag
// The aggregation hash is ONE serial SHAKE256 per aggregate (fusion.py:573-591) over
//     dst "," "[" "(" VK ", " INT ", " CH ")" ", " ... "]"
// followed by total * n_ag squeezed bytes, decoded signer by signer (fusion.py:594-629).  The sponge is RESUMABLE: its
// state (25 lanes, the text bytes not yet absorbed, the squeeze position) lives in an AgState record per aggregate, so a
// batch of aggregates can be hashed WINDOW BY WINDOW over the sorted signer positions: every call consumes the next few
// signers of every aggregate.  All chains of the batch are then in flight at once, independent of how many signers fit
// in HBM at a time, and the per-signer inputs of the next window (gather, hash_ch) can be computed on another stream
// while this window is absorbed.
struct __align__(16) AgState {
    unsigned long long a[25];   // sponge lanes
    unsigned long long q[26];   // absorb phase: text bytes generated but not yet absorbed (little endian, nq of them)
    uint32_t nq;
    uint32_t absorbed;          // signers whose text has been generated
    uint32_t total;             // signers of this aggregate
    uint32_t produced;          // signers whose coefficient polynomial has been squeezed
    uint32_t bpos;              // squeeze phase: bytes of the current block (lanes 0..16 of a) already handed out
    uint32_t phase;             // 0 absorbing, 1 squeezing
    unsigned long long rpos;    // ring-fed absorb (fz_hash.cu agw_*): message bytes absorbed so far (a multiple of 136 until closing)
    uint32_t pad[18];
};
static_assert(sizeof(AgState) == 512, "AgState layout");

__global__ void ag_state_init_kernel(AgState* __restrict__ st, const unsigned long long* __restrict__ agg_off, size_t n_aggs) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_aggs) return;
    AgState* s = st + j;
    for (int k = 0; k < 25; ++k) s->a[k] = 0;
    for (int k = 0; k < 26; ++k) s->q[k] = 0;
    s->nq = 0; s->absorbed = 0; s->total = (uint32_t)(agg_off[j + 1] - agg_off[j]); s->produced = 0; s->bpos = 0; s->phase = 0;
    s->rpos = 0;
}

// one thread per aggregate: absorbs the signers [win_off[j], win_off[j+1]) of this window (the next ones in sorted order)
__global__ void __launch_bounds__(32) hash_ag_absorb_thread_kernel(AgState* __restrict__ st, const int32_t* __restrict__ vk_l_s,
                                                                     const int32_t* __restrict__ vk_r_s, const uint8_t* __restrict__ dg_s,
                                                                     const int32_t* __restrict__ c_hat_s, const unsigned long long* __restrict__ win_off,
                                                                     size_t n_aggs, int level, int d) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_aggs) return;
    AgState* s = st + j;
    if (s->phase != 0) return;
    const size_t lo = win_off[j], hi = win_off[j + 1];
    const uint32_t total = s->total;
    uint32_t absorbed = s->absorbed;
    if (hi == lo && total != 0) return;
    unsigned long long a[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) a[k] = s->a[k];
    Queue q;
    {
        const int nq = (int)s->nq;
        q.wi = nq >> 3; q.nb = nq & 7;
        for (int k = 0; k < q.wi; ++k) q.blk[k] = s->q[k];
        q.cur = q.nb ? s->q[q.wi] : 0ULL;
    }
    Gen g;
    g.item = 0; g.idx = 0; g.int_len = 0;
    size_t p = lo;
    bool more = hi > lo, closing = false;
    if (total == 0) {  // "[]": dst "," "[" "]"
        q_push(q, (unsigned long long)(level == 1 ? 3 : 1) | (2ULL << 8) | ((unsigned long long)',' << 16) | ((unsigned long long)'[' << 24) |
                      ((unsigned long long)']' << 32), 0, 5);
        more = false;
        closing = true;
    }
    // per-signer program: 0 OPEN (first / next), 1 values(vk_l), 2 MID, 3 values(vk_r), 4 AFTER_VK, 5 INT, 6 BEFORE_C, 7 values(c_hat), 8 CLOSE
    for (;;) {
        while (q_bytes(q) < KECCAK_RATE && more) {
            bool fin;
            switch (g.item) {
                case 0: fin = gen_segment(g, q, level, absorbed == 0 ? SEG_AG_OPEN_FIRST : SEG_AG_OPEN_NEXT); break;
                case 1: fin = gen_values(g, q, vk_l_s + p * d, d); break;
                case 2: fin = gen_segment(g, q, level, SEG_MID); break;
                case 3: fin = gen_values(g, q, vk_r_s + p * d, d); break;
                case 4: fin = gen_segment(g, q, level, SEG_AG_AFTER_VK); break;
                case 5: fin = gen_int(g, q, dg_s + p * 32); break;
                case 6: fin = gen_segment(g, q, level, SEG_AG_BEFORE_C); break;
                case 7: fin = gen_values(g, q, c_hat_s + p * d, d); break;
                default: fin = gen_segment(g, q, level, SEG_AG_CLOSE); break;
            }
            if (fin) {
                g.idx = 0;
                if (++g.item == 9) {
                    g.item = 0;
                    ++absorbed;
                    if (++p == hi) more = false;
                    if (absorbed == total) {
                        q_push(q, (unsigned long long)']', 0, 1);
                        closing = true;
                    }
                }
            }
        }
        if (q_bytes(q) >= KECCAK_RATE) {
            q_absorb_full(q, a);
            keccak_f(a);
        } else if (closing) {
            q_absorb_last(q, a);
            keccak_f(a);
            q_reset(q);
            s->phase = 1;
            s->bpos = 0;
            break;
        } else {
            break;   // window exhausted: the leftover text waits in the state
        }
    }
#pragma unroll
    for (int k = 0; k < 25; ++k) s->a[k] = a[k];
    for (int k = 0; k < q.wi; ++k) s->q[k] = q.blk[k];
    s->q[q.wi] = q.cur;
    s->nq = (uint32_t)q_bytes(q);
    s->absorbed = absorbed;
}

// one thread per aggregate: squeezes and decodes the coefficients of the next hi - lo signers, n_ag bytes each
__global__ void __launch_bounds__(32) hash_ag_squeeze_thread_kernel(AgState* __restrict__ st, const unsigned long long* __restrict__ win_off, size_t n_aggs,
                                                                      int d, int weight, int bpc, int bpi, uint32_t n_ag, int32_t* __restrict__ coef_out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_aggs) return;
    AgState* s = st + j;
    const size_t lo = win_off[j], hi = win_off[j + 1];
    if (hi == lo || s->phase != 1) return;
    unsigned long long a[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) a[k] = s->a[k];
    int32_t c[256];
    unsigned long long blk[KECCAK_RATE_LANES];
#pragma unroll
    for (int k = 0; k < KECCAK_RATE_LANES; ++k) blk[k] = a[k];
    int bpos = (int)s->bpos;  // bytes consumed from the current block
    for (size_t sp = lo; sp < hi; ++sp) {
        for (int k = 0; k < d; ++k) c[k] = 0;
        Dec ds;
        dinit(ds, c, d, weight, 1u, bpc, bpi);
        uint32_t used = 0;
        while (used < n_ag) {
            if (bpos == KECCAK_RATE) {
                keccak_f(a);
#pragma unroll
                for (int k = 0; k < KECCAK_RATE_LANES; ++k) blk[k] = a[k];
                bpos = 0;
            }
            const uint32_t avail = KECCAK_RATE - bpos, need = n_ag - used;
            const uint32_t take = avail < need ? avail : need;
            if (ds.phase != 3)
                for (uint32_t t = 0; t < take && ds.phase != 3; ++t) {
                    const int jb = bpos + (int)t;
                    dbyte(ds, c, (uint32_t)(blk[jb >> 3] >> (8 * (jb & 7))) & 0xffu);
                }
            bpos += (int)take;
            used += take;
        }
        while (ds.phase != 3) dfinish(ds, c);  // the slice is exhausted: remaining draws see empty bytes (quirk Q4)
        int32_t* o = coef_out + sp * d;
        for (int k = 0; k < d; ++k) o[k] = c[k];
    }
#pragma unroll
    for (int k = 0; k < 25; ++k) s->a[k] = a[k];
    s->bpos = (uint32_t)bpos;
    s->produced += (uint32_t)(hi - lo);
}

// ---------------------------------------------------------------------------------- ring-fed chain (thread per aggregate)
// The text of the window is already in the aggregate's byte ring (fz_hash.cu agw_format_kernel): the chain only loads the
// 17 lanes of each block (the next block's loads are issued before the permutation of the current one) and permutes.
__global__ void __launch_bounds__(32) agw_absorb_thread_kernel(AgState* __restrict__ st, const uint8_t* __restrict__ ring, size_t ring_cap,
                                                                const unsigned long long* __restrict__ wend, const unsigned long long* __restrict__ win_off,
                                                                size_t n_aggs) {
    // The next block travels HBM -> shared memory with cp.async while the current one is permuted.  (Plain loads into registers are
    // sunk by the compiler to just before their use, behind the permutation, which exposed the whole memory latency: 14 % of the
    // stall samples.)  Word k of thread t sits at stage[buf][k][t]: conflict-free.
    __shared__ __align__(8) unsigned long long stage[2][KECCAK_RATE_LANES][32];
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_aggs) return;
    AgState* s = st + j;
    if (s->phase != 0) return;
    const uint32_t absorbed = s->absorbed + (uint32_t)(win_off[j + 1] - win_off[j]);
    const bool closing = absorbed == s->total;
    unsigned long long rpos = s->rpos;
    const unsigned long long end = wend[j] & ~AGW_TEXT_FLAGS;
    if (wend[j] & AGW_TEXT_OVERRUN) { s->phase = 0xDEAD; return; }   // the writer overwrote bytes this sponge had not absorbed: never hash them
    s->absorbed = absorbed;
    if (rpos + KECCAK_RATE > end && !closing) return;
    unsigned long long a[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) a[k] = s->a[k];
    const uint8_t* base = ring + j * ring_cap;
    const int t = threadIdx.x;
    auto fetch = [&](unsigned long long pos, int buf) {   // 17 asynchronous 8-byte copies of the block at message offset pos
        const uint8_t* src = base + (pos % ring_cap);
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) {
#ifdef __CUDA_ARCH__
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&stage[buf][k][t]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src + 8 * k) : "memory");
#else   // CPU test suite (tests/host/cuda_emu.h): the copy is synchronous
            memcpy(&stage[buf][k][t], src + 8 * k, 8);
#endif
        }
#ifdef __CUDA_ARCH__
        asm volatile("cp.async.commit_group;" ::: "memory");
#endif
    };
    bool have = rpos + KECCAK_RATE <= end;
    int buf = 0;
    if (have) fetch(rpos, 0);
    while (have) {
#ifdef __CUDA_ARCH__
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) a[k] ^= stage[buf][k][t];
        rpos += KECCAK_RATE;
        have = rpos + KECCAK_RATE <= end;
        buf ^= 1;
        if (have) fetch(rpos, buf);
        keccak_f(a);
    }
    if (closing) {   // last (partial) block with SHAKE padding; rem < 136
        const uint32_t rem = (uint32_t)(end - rpos);
        const unsigned long long* b = reinterpret_cast<const unsigned long long*>(base + (rpos % ring_cap));
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) {
            unsigned long long w = 0;
            if ((uint32_t)(8 * k) < rem) {
                w = b[k];
                const uint32_t valid = rem - 8 * k;
                if (valid < 8) w &= (1ULL << (8 * valid)) - 1ULL;
            }
            if ((rem >> 3) == (uint32_t)k) w ^= 0x1FULL << (8 * (rem & 7));
            if (k == 16) w ^= 0x8000000000000000ULL;
            a[k] ^= w;
        }
        keccak_f(a);
        rpos = end;
        s->phase = 1;
        s->bpos = 0;
    }
#pragma unroll
    for (int k = 0; k < 25; ++k) s->a[k] = a[k];
    s->rpos = rpos;
}

// squeezes whole blocks into the aggregate's XOF buffer: offset 0 is the start of the sponge's current block, the window's
// bytes begin bpos_out[j] bytes into it (signer r of the window at bpos_out[j] + r * n_ag)
__global__ void __launch_bounds__(32) agw_squeeze_thread_kernel(AgState* __restrict__ st, const unsigned long long* __restrict__ win_off, size_t n_aggs,
                                                                 uint32_t n_ag, uint8_t* __restrict__ xof, size_t xof_cap, uint32_t* __restrict__ bpos_out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_aggs) return;
    AgState* s = st + j;
    const unsigned long long cnt = win_off[j + 1] - win_off[j];
    if (cnt == 0 || s->phase != 1) return;
    unsigned long long a[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) a[k] = s->a[k];
    const uint32_t bpos = s->bpos;
    bpos_out[j] = bpos;
    unsigned long long* out = reinterpret_cast<unsigned long long*>(xof + j * xof_cap);
    const unsigned long long need = bpos + cnt * n_ag;   // bytes counted from the start of the current block
    unsigned long long o = 0;
#pragma unroll
    for (int k = 0; k < KECCAK_RATE_LANES; ++k) __stcs(out + k, a[k]);
    while (o + KECCAK_RATE < need) {
        keccak_f(a);
        o += KECCAK_RATE;
        unsigned long long* ob = out + o / 8;
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) __stcs(ob + k, a[k]);
    }
#pragma unroll
    for (int k = 0; k < 25; ++k) s->a[k] = a[k];
    s->bpos = (uint32_t)(need - o);
    s->produced += (uint32_t)cnt;
}

// ---------------------------------------------------------------------------------- host
// ---------------------------------------------------------------------------------- ring-fed chain (TWO lanes per aggregate)
// Between about 1,800 and 40,000 aggregates per GPU neither of the two layouts above fits: a warp per chain puts three and more warps on a
// sub-partition and their 18 shuffles per round saturate its shuffle port (247 cycles per round at three chains, 350 at four), a thread per chain
// leaves half of the sub-partitions without a warp below 19,000 chains while each warp is bound by its own ALU pipe (180 instructions x 2 cycles
// per round: 370-400 cycles).  Here the even lane of a pair holds the LOW 32 bits of every Keccak lane and the odd lane the HIGH 32 bits: XOR and
// AND-NOT need no communication, a 64-bit rotation is one SHFL.BFLY for the partner's half and one funnel shift, i.e. 90 ALU instructions and 28
// shuffles per thread and round and 16 chains per warp.  tools/keccak_lab2.cu (profiles/r02_keccak_lab2_v23.txt): 246-255 cycles per round up to
// 9,472 chains (a warp per sub-partition), 351 at 18,944, 545 at 37,888 (3.94 G Keccak-f/s against 4.27 G of the thread layout at saturation); the
// even / odd bit-interleaved split (V22, 17 shuffles) measures 230 but needs every message word interleaved and every squeezed word
// de-interleaved (about 40 instructions per 64-bit word), which costs more than the 8 % it saves.  Message bytes and squeezed bytes need no
// conversion here: each lane loads / stores its own 4 bytes of every 8.
// Shuffles with a mask narrower than the warp compile to WARPSYNC.COLLECTIVE sequences (4x the instructions), so every shuffle is full-mask and the
// control flow is warp-uniform: all 16 pairs run as many permutations as the busiest pair of the warp and a pair snapshots its state after its own
// last one (25 predicated moves per block, about 1 %).
namespace pairag {

// rot / round / keccak_f_pair: fz_keccak_lanes.cuh

__device__ __forceinline__ uint32_t warp_max(uint32_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// CTAs of four warps (64 aggregates): the four warps of a CTA sit on the four sub-partitions of its SM whatever else is resident, single-warp CTAs
// do not spread evenly (measured: 592 single-warp CTAs ran like two warps per sub-partition).  No CTA-wide barrier anywhere.
constexpr int PAIR_WARPS = 4;
__global__ void __launch_bounds__(PAIR_WARPS * 32) agw_absorb_pair_kernel(AgState* __restrict__ st, const uint8_t* __restrict__ ring, size_t ring_cap,
                                                                          const unsigned long long* __restrict__ wend,
                                                                          const unsigned long long* __restrict__ win_off, size_t n_aggs) {
    // word k of lane t of the next block at stage[buf][k][t] (cp.async while the current block is permuted, as in the thread kernel)
    __shared__ __align__(4) uint32_t stage_all[PAIR_WARPS][2][KECCAK_RATE_LANES][32];
    uint32_t (*stage)[KECCAK_RATE_LANES][32] = stage_all[threadIdx.x >> 5];
    const int t = threadIdx.x & 31;
    const uint32_t hi = (uint32_t)(t & 1);
    const size_t grp = (size_t)blockIdx.x * PAIR_WARPS + (threadIdx.x >> 5);
    if (grp * 16 >= n_aggs) return;   // warp-uniform
    const size_t j = grp * 16 + (size_t)(t >> 1);
    const bool valid = j < n_aggs;
    AgState* s = st + (valid ? j : 0);
    bool act = valid && s->phase == 0;
    unsigned long long rpos = 0, end = 0;
    bool closing = false;
    if (act) {
        const unsigned long long we = wend[j];
        if (we & AGW_TEXT_OVERRUN) {   // the writer overwrote bytes this sponge had not absorbed: never hash them
            act = false;
        } else {
            const uint32_t absorbed = s->absorbed + (uint32_t)(win_off[j + 1] - win_off[j]);
            closing = absorbed == s->total;
            rpos = s->rpos;
            end = we & ~AGW_TEXT_FLAGS;
        }
    }
    const uint32_t nblk = act ? (uint32_t)((end - rpos) / KECCAK_RATE) : 0u;
    const bool work = act && (nblk > 0 || closing);
    uint32_t a[25], sv[25];
    const uint32_t* sa = reinterpret_cast<const uint32_t*>(s->a) + hi;
#pragma unroll
    for (int k = 0; k < 25; ++k) { a[k] = work ? sa[2 * k] : 0u; sv[k] = a[k]; }
    const uint8_t* base = ring + (valid ? j : 0) * ring_cap + 4 * hi;
    auto fetch = [&](unsigned long long pos, int buf) {   // this lane's 4 bytes of each of the 17 words of the block at message offset pos
        const uint8_t* src = base + (pos % ring_cap);
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) {
#ifdef __CUDA_ARCH__
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&stage[buf][k][t]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + 8 * k) : "memory");
#else
            memcpy(&stage[buf][k][t], src + 8 * k, 4);
#endif
        }
    };
    const uint32_t maxblk = warp_max(nblk);
    int buf = 0;
    if (nblk > 0) fetch(rpos, 0);
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
    for (uint32_t b = 0; b < maxblk; ++b) {
#ifdef __CUDA_ARCH__
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
        const bool mine = b < nblk;
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) a[k] ^= mine ? stage[buf][k][t] : 0u;
        buf ^= 1;
        if (b + 1 < nblk) fetch(rpos + (unsigned long long)(b + 1) * KECCAK_RATE, buf);
#ifdef __CUDA_ARCH__
        asm volatile("cp.async.commit_group;" ::: "memory");
#endif
        keccak_f_pair(a, hi);
        if (b + 1 == nblk) {
#pragma unroll
            for (int k = 0; k < 25; ++k) sv[k] = a[k];
        }
    }
#pragma unroll
    for (int k = 0; k < 25; ++k) a[k] = sv[k];
    rpos += (unsigned long long)nblk * KECCAK_RATE;
    if (__any_sync(0xffffffffu, closing)) {   // last (partial) block with SHAKE padding; rem < 136
        if (closing) {
            const uint32_t rem = (uint32_t)(end - rpos);
            const uint32_t* bw = reinterpret_cast<const uint32_t*>(base + (rpos % ring_cap));
#pragma unroll
            for (int k = 0; k < KECCAK_RATE_LANES; ++k) {
                const uint32_t off = 8u * k + 4u * hi;   // byte offset of this lane's half of word k
                uint32_t w = 0;
                if (off < rem) {
                    w = bw[2 * k];
                    const uint32_t vb = rem - off;
                    if (vb < 4) w &= (1u << (8 * vb)) - 1u;
                }
                if (rem >= off && rem < off + 4) w ^= 0x1Fu << (8 * (rem - off));
                if (k == 16 && hi) w ^= 0x80000000u;
                a[k] ^= w;
            }
        }
        keccak_f_pair(a, hi);
        if (!closing) {
#pragma unroll
            for (int k = 0; k < 25; ++k) a[k] = sv[k];
        }
    }
    if (valid && s->phase == 0 && !act) { if (!hi) s->phase = 0xDEAD; return; }   // overrun (act was cleared above)
    if (!act) return;
    if (work) {
        uint32_t* da = reinterpret_cast<uint32_t*>(s->a) + hi;
#pragma unroll
        for (int k = 0; k < 25; ++k) da[2 * k] = a[k];
    }
    if (!hi) {
        s->absorbed += (uint32_t)(win_off[j + 1] - win_off[j]);
        if (closing) { s->rpos = end; s->phase = 1; s->bpos = 0; }
        else s->rpos = rpos;
    }
}

// squeezes whole blocks into the aggregate's XOF buffer (same layout as agw_squeeze_thread_kernel).
// A time-sliced form (launches of at most 592 warps with a block budget, the resting groups rotating, for the 625 warps of 10^4 aggregates) was built
// and measured: 458 ms against 465 ms of squeeze kernels over the 109 windows of the configs[4] batch, i.e. nothing -- the squeeze phase is not paced by
// the 33 sub-partitions that carry two warps but by the decode / gather / target kernels of the previous window sharing the issue slots (4.2 ms per
// window = 2.7 ms of bare chain + 1.5 ms of side-stream work; 2.7 ms per window at 2,500 aggregates, where that work is four times smaller).  Not kept.
__global__ void __launch_bounds__(PAIR_WARPS * 32) agw_squeeze_pair_kernel(AgState* __restrict__ st, const unsigned long long* __restrict__ win_off,
                                                                           size_t n_aggs, uint32_t n_ag, uint8_t* __restrict__ xof, size_t xof_cap,
                                                                           uint32_t* __restrict__ bpos_out) {
    const int t = threadIdx.x & 31;
    const uint32_t hi = (uint32_t)(t & 1);
    const size_t grp = (size_t)blockIdx.x * PAIR_WARPS + (threadIdx.x >> 5);
    if (grp * 16 >= n_aggs) return;   // warp-uniform
    const size_t j = grp * 16 + (size_t)(t >> 1);
    const bool valid = j < n_aggs;
    AgState* s = st + (valid ? j : 0);
    const unsigned long long cnt = valid ? win_off[j + 1] - win_off[j] : 0ULL;
    const bool act = valid && cnt != 0 && s->phase == 1;
    const uint32_t bpos = act ? s->bpos : 0u;
    const unsigned long long need = (unsigned long long)bpos + cnt * n_ag;   // bytes counted from the start of the current block
    const uint32_t nperm = act && need > 0 ? (uint32_t)((need - 1) / KECCAK_RATE) : 0u;   // trips of `while (o + 136 < need)`
    uint32_t a[25], sv[25];
    const uint32_t* sa = reinterpret_cast<const uint32_t*>(s->a) + hi;
#pragma unroll
    for (int k = 0; k < 25; ++k) { a[k] = act ? sa[2 * k] : 0u; sv[k] = a[k]; }
    uint32_t* out = reinterpret_cast<uint32_t*>(xof + (valid ? j : 0) * xof_cap) + hi;
    if (act) {
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) __stcs(out + 2 * k, a[k]);
    }
    const uint32_t maxperm = warp_max(nperm);
    for (uint32_t it = 0; it < maxperm; ++it) {
        keccak_f_pair(a, hi);
        if (it < nperm) {
            uint32_t* ob = out + (size_t)(it + 1) * (KECCAK_RATE / 4);
#pragma unroll
            for (int k = 0; k < KECCAK_RATE_LANES; ++k) __stcs(ob + 2 * k, a[k]);
        }
        if (it + 1 == nperm) {
#pragma unroll
            for (int k = 0; k < 25; ++k) sv[k] = a[k];
        }
    }
    if (!act) return;
    uint32_t* da = reinterpret_cast<uint32_t*>(s->a) + hi;
#pragma unroll
    for (int k = 0; k < 25; ++k) da[2 * k] = sv[k];
    if (!hi) {
        bpos_out[j] = bpos;
        s->bpos = (uint32_t)(need - (unsigned long long)nperm * KECCAK_RATE);
        s->produced += (uint32_t)cnt;
    }
}

}  // namespace pairag

int hashstream_init_tables(fz_ctx* ctx) {
    uint32_t magic[257];
    SegTable tabs[2];
    std::vector<uint8_t> pools(2 * 4096);
    if (!hashstream_build_tables(magic, tabs, pools.data())) { set_error("fixed-text pool overflow"); return -1; }
    if (check(cudaMemcpyToSymbol(c_magic2, magic, sizeof(magic)), "decoder constants upload")) return -2;
    for (int L = 0; L < 2; ++L)
        if (check(cudaMemcpyToSymbol(g_text, pools.data() + (size_t)L * 4096, 4096, (size_t)L * 4096), "fixed text upload")) return -2;
    (void)ctx;
    return check(cudaMemcpyToSymbol(c_seg, tabs, sizeof(tabs)), "segment table upload");
}

int launch_hash_ch_stream(const fz_ctx* ctx, const int32_t* vk_l, const int32_t* vk_r, const uint8_t* digests, size_t n, int32_t* coef_out,
                          cudaStream_t st) {
    if (n == 0) return 0;
    FZ_LAUNCH(hash_ch_stream_kernel, (int)((n + HS_THREADS - 1) / HS_THREADS), HS_THREADS, 0, st, vk_l, vk_r, digests, n, ctx->secpar == 256 ? 1 : 0, ctx->d,
                                                                                            ctx->omega_ch, ctx->bytes_coef, ctx->bytes_idx,
                                                                                            (uint32_t)ctx->n_ch_used, coef_out);
    return check(cudaGetLastError(), "hash_ch_stream launch");
}


// ==========================================================================================
// Warp-cooperative variant of hash_ag: ONE WARP per aggregate.
//
// The aggregation hash is a single serial sponge per aggregate (fusion.py:573-591), ~279 K Keccak-f for a capacity
// aggregate at secpar 256.  With one thread per sponge a permutation takes ~14 K cycles of one warp, which bounds the
// latency of a batch of a few thousand aggregates.  Here 25 lanes hold one 64-bit Keccak lane each and exchange
// through warp shuffles (theta: 4 + 2, pi: 1, chi: 2 sixty-four-bit shuffles per round), the 32 lanes format 32
// coefficients of the repr text at a time into a block-aligned shared-memory ring, and the squeezed bytes are decoded
// cooperatively (the 60 / 35 index draws of one coefficient polynomial are independent Horner reductions).
namespace warpag {

constexpr int RING_BLOCKS = 16;
constexpr int RING_BYTES = RING_BLOCKS * KECCAK_RATE;   // 2176, every block start is 8-byte aligned
constexpr int XOF_BYTES = 4096;                         // >= n_ag (3968 @256, 1195 @128) rounded up
constexpr int WARPS = 4;
constexpr int SMEM_PER_WARP = RING_BYTES + XOF_BYTES + 1024 + 64;

// c_rot / Lane / lane_init / keccak_f_warp / keccak_f_warp_u: fz_keccak_lanes.cuh


struct Ring {
    uint8_t* buf;   // RING_BYTES, block aligned
    int wr;         // total bytes written
    int rd;         // total bytes absorbed (multiple of 136)
};

// all lanes: append `n` (<= 16) bytes of this lane's token after the tokens of the lower lanes
__device__ __forceinline__ void ring_put_tokens(Ring& rg, unsigned long long lo, unsigned long long hi, int n, int L) {
    int x = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, x, o);
        if (L >= o) x += y;
    }
    const int total = __shfl_sync(0xffffffffu, x, 31);
    int off = (rg.wr + x - n) % RING_BYTES;
    for (int k = 0; k < n; ++k) {
        rg.buf[off] = (uint8_t)((k < 8 ? lo >> (8 * k) : hi >> (8 * (k - 8))) & 0xff);
        off = off + 1 == RING_BYTES ? 0 : off + 1;
    }
    rg.wr += total;
}

// absorb every complete block currently in the ring
__device__ __forceinline__ void ring_drain(Ring& rg, uint32_t& lo, uint32_t& hi, const Lane& c, int L) {
    __syncwarp();
    while (rg.wr - rg.rd >= KECCAK_RATE) {
        if (L < KECCAK_RATE_LANES) {
            const uint2 w = *reinterpret_cast<const uint2*>(rg.buf + (rg.rd % RING_BYTES) + 8 * L);
            lo ^= w.x;
            hi ^= w.y;
        }
        keccak_f_warp(lo, hi, c, L);
        rg.rd += KECCAK_RATE;
    }
    __syncwarp();
}

__device__ __forceinline__ void ring_segment(Ring& rg, int level, int seg, uint32_t& lo, uint32_t& hi, const Lane& c, int L) {
    const int len = (int)c_seg[level].len[seg];
    const uint8_t* p = g_text[level] + c_seg[level].off[seg];
    for (int base = 0; base < len; base += 32 * 16) {   // 32 lanes x 16-byte tokens per round
        const int t = base + 16 * L;
        int n = len - t;
        n = n < 0 ? 0 : (n > 16 ? 16 : n);
        unsigned long long a = 0, b = 0;
        if (n > 0) {
            const uint4 v = *reinterpret_cast<const uint4*>(p + t);
            a = ((unsigned long long)v.y << 32) | v.x;
            b = ((unsigned long long)v.w << 32) | v.z;
        }
        ring_put_tokens(rg, a, b, n, L);
        ring_drain(rg, lo, hi, c, L);
    }
}

__device__ __forceinline__ void ring_values(Ring& rg, const int32_t* __restrict__ vals, int d, uint32_t& lo, uint32_t& hi, const Lane& c, int L) {
    for (int base = 0; base < d; base += 32) {
        unsigned long long a, b;
        const int n = fmt_value(vals[base + L], base + L != d - 1, a, b);
        ring_put_tokens(rg, a, b, n, L);
        ring_drain(rg, lo, hi, c, L);
    }
}

// absorb the signers [win_off[j], win_off[j+1]) of this window into the sponge of aggregate j (one warp per aggregate)
__global__ void __launch_bounds__(WARPS * 32) hash_ag_absorb_warp_kernel(AgState* __restrict__ st, const int32_t* __restrict__ vk_l_s,
                                                                          const int32_t* __restrict__ vk_r_s, const uint8_t* __restrict__ dg_s,
                                                                          const int32_t* __restrict__ c_hat_s, const unsigned long long* __restrict__ win_off,
                                                                          size_t n_aggs, int level, int d) {
    FZ_DYN_SHARED __align__(16) uint8_t wsm[];
    const int L = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t j = (size_t)blockIdx.x * WARPS + warp;
    if (j >= n_aggs) return;
    AgState* s = st + j;
    if (s->phase != 0) return;
    const size_t p0 = win_off[j], p1 = win_off[j + 1];
    const uint32_t total = s->total;
    uint32_t absorbed = s->absorbed;
    if (p1 == p0 && total != 0) return;
    uint8_t* base = wsm + (size_t)warp * SMEM_PER_WARP;
    Ring rg;
    rg.buf = base; rg.wr = 0; rg.rd = 0;
    uint8_t* intbuf = base + RING_BYTES + XOF_BYTES + 512;                   // 96 bytes decimal prehash
    const Lane c = lane_init(L);
    uint32_t lo = 0, hi = 0;
    if (L < 25) { const unsigned long long v = s->a[L]; lo = (uint32_t)v; hi = (uint32_t)(v >> 32); }
    {   // text left over from the previous window
        const int nq = (int)s->nq;
        const uint8_t* qb = reinterpret_cast<const uint8_t*>(s->q);
        for (int k = L; k < nq; k += 32) rg.buf[k] = qb[k];
        rg.wr = nq;
        __syncwarp();
    }
    bool closing = false;
    if (total == 0) {
        const unsigned long long t = (unsigned long long)(level == 1 ? 3 : 1) | (2ULL << 8) | ((unsigned long long)',' << 16) |
                                     ((unsigned long long)'[' << 24) | ((unsigned long long)']' << 32);
        ring_put_tokens(rg, L == 0 ? t : 0ULL, 0ULL, L == 0 ? 5 : 0, L);
        closing = true;
    }
    for (size_t p = p0; p < p1; ++p) {
        ring_segment(rg, level, absorbed == 0 ? SEG_AG_OPEN_FIRST : SEG_AG_OPEN_NEXT, lo, hi, c, L);
        ring_values(rg, vk_l_s + p * d, d, lo, hi, c, L);
        ring_segment(rg, level, SEG_MID, lo, hi, c, L);
        ring_values(rg, vk_r_s + p * d, d, lo, hi, c, L);
        ring_segment(rg, level, SEG_AG_AFTER_VK, lo, hi, c, L);
        int ilen = 0;
        if (L == 0) ilen = fmt_le256(dg_s + p * 32, intbuf);
        ilen = __shfl_sync(0xffffffffu, ilen, 0);
        __syncwarp();
        {
            const int t = 16 * L;
            int n = ilen - t;
            n = n < 0 ? 0 : (n > 16 ? 16 : n);
            unsigned long long a = 0, b = 0;
            if (n > 0) {
                for (int k = 0; k < 8; ++k) a |= (unsigned long long)intbuf[t + k] << (8 * k);
                for (int k = 0; k < 8; ++k) b |= (unsigned long long)intbuf[t + 8 + k] << (8 * k);
            }
            ring_put_tokens(rg, a, b, n, L);
            ring_drain(rg, lo, hi, c, L);
        }
        ring_segment(rg, level, SEG_AG_BEFORE_C, lo, hi, c, L);
        ring_values(rg, c_hat_s + p * d, d, lo, hi, c, L);
        ring_segment(rg, level, SEG_AG_CLOSE, lo, hi, c, L);
        if (++absorbed == total) {
            ring_put_tokens(rg, L == 0 ? (unsigned long long)']' : 0ULL, 0ULL, L == 0 ? 1 : 0, L);
            ring_drain(rg, lo, hi, c, L);
            closing = true;
        }
    }
    __syncwarp();
    const int rem = rg.wr - rg.rd;   // < 136
    unsigned long long w = 0;        // lane L < 17: bytes [8L, 8L + 8) of the pending text
    if (L < KECCAK_RATE_LANES) {
        const int b0 = 8 * L;
        for (int k = 0; k < 8; ++k)
            if (b0 + k < rem) w |= (unsigned long long)rg.buf[(rg.rd + b0 + k) % RING_BYTES] << (8 * k);
    }
    if (closing) {   // last block with padding
        if (L < KECCAK_RATE_LANES) {
            const int b0 = 8 * L;
            if (rem >= b0 && rem < b0 + 8) w ^= 0x1FULL << (8 * (rem - b0));
            if (L == 16) w ^= 0x8000000000000000ULL;
            lo ^= (uint32_t)w;
            hi ^= (uint32_t)(w >> 32);
        }
        keccak_f_warp(lo, hi, c, L);
        if (L == 0) { s->phase = 1; s->bpos = 0; s->nq = 0; }
    } else {
        if (L < KECCAK_RATE_LANES) s->q[L] = w;
        if (L == 0) s->nq = (uint32_t)rem;
    }
    if (L < 25) s->a[L] = ((unsigned long long)hi << 32) | lo;
    if (L == 0) s->absorbed = absorbed;
}

// squeeze + decode the coefficients of the next p1 - p0 signers of aggregate j, n_ag bytes per signer
__global__ void __launch_bounds__(WARPS * 32) hash_ag_squeeze_warp_kernel(AgState* __restrict__ st, const unsigned long long* __restrict__ win_off,
                                                                           size_t n_aggs, int d, int weight, int bpc, int bpi, uint32_t n_ag,
                                                                           int32_t* __restrict__ coef_out) {
    FZ_DYN_SHARED __align__(16) uint8_t wsm[];
    const int L = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t j = (size_t)blockIdx.x * WARPS + warp;
    if (j >= n_aggs) return;
    AgState* s = st + j;
    const size_t p0 = win_off[j], p1 = win_off[j + 1];
    if (p1 == p0 || s->phase != 1) return;
    uint8_t* base = wsm + (size_t)warp * SMEM_PER_WARP;
    uint8_t* xof = base + RING_BYTES;            // squeezed bytes of the current signer
    int8_t* cf = reinterpret_cast<int8_t*>(base + RING_BYTES + XOF_BYTES);   // d coefficients
    int* jdraw = reinterpret_cast<int*>(base + RING_BYTES + XOF_BYTES + 640);  // up to 64 index draws
    const Lane c = lane_init(L);
    uint32_t lo = 0, hi = 0;
    if (L < 25) { const unsigned long long v = s->a[L]; lo = (uint32_t)v; hi = (uint32_t)(v >> 32); }
    int have = 0;                     // bytes buffered in xof[] for the current signer
    int blk_used = (int)s->bpos;      // bytes of the current squeezed block already handed out
    const int bfs = (weight + 7) / 8;
    for (size_t sp = p0; sp < p1; ++sp) {
        // collect n_ag bytes
        while (have < (int)n_ag) {
            if (blk_used == KECCAK_RATE) {
                keccak_f_warp(lo, hi, c, L);
                blk_used = 0;
            }
            // the current block lives in registers: lanes 0..16 publish the needed part
            const int take = min(KECCAK_RATE - blk_used, (int)n_ag - have);
            if (L < KECCAK_RATE_LANES) {
                const unsigned long long w = ((unsigned long long)hi << 32) | lo;
                for (int k = 0; k < 8; ++k) {
                    const int bpos = 8 * L + k;
                    if (bpos >= blk_used && bpos < blk_used + take) xof[have + bpos - blk_used] = (uint8_t)(w >> (8 * k));
                }
            }
            have += take;
            blk_used += take;
        }
        __syncwarp();
        // decode (fusion.py:422-481) with bound = 1: signums, `weight` ignored magnitudes, then the partial Fisher-Yates
        for (int k = L; k < d; k += 32) {
            int v = 0;
            if (k < weight) v = 2 * ((xof[bfs - 1 - k / 8] >> (k % 8)) & 1) - 1;
            cf[k] = (int8_t)v;
        }
        const int n_swaps = weight < d ? d - 1 - weight : 0;
        const int first = bfs + bpc * weight;
        // index draws whose bytes exist (possibly truncated); the others see b"" -> 0 (quirk Q4)
        const int n_full = (int)n_ag > first ? ((int)n_ag - first + bpi - 1) / bpi : 0;
        const int n_draw = n_full < n_swaps ? n_full : n_swaps;
        for (int t = L; t < n_draw; t += 32) {
            const int i = d - 1 - t;
            const uint32_t m = (uint32_t)(i + 1), mg = c_magic2[m];
            uint32_t acc = 0;
            const int s0 = first + t * bpi;
            const int e0 = min(s0 + bpi, (int)n_ag);
            for (int q = s0; q < e0; ++q) {
                const uint32_t v = (acc << 8) | xof[q];
                acc = v - __umulhi(v, mg) * m;
            }
            jdraw[t] = (int)acc;
        }
        __syncwarp();
        if (L == 0) {
            for (int t = 0; t < n_swaps; ++t) {
                const int i = d - 1 - t;
                const int jj = t < n_draw ? jdraw[t] : 0;
                const int8_t tmp = cf[i]; cf[i] = cf[jj]; cf[jj] = tmp;
            }
        }
        __syncwarp();
        int32_t* o = coef_out + sp * d;
        for (int k = L; k < d; k += 32) o[k] = (int32_t)cf[k];
        __syncwarp();
        have = 0;
    }
    if (L < 25) s->a[L] = ((unsigned long long)hi << 32) | lo;
    if (L == 0) { s->bpos = (uint32_t)blk_used; s->produced += (uint32_t)(p1 - p0); }
}

// ring-fed chain, one warp per aggregate: lanes 0..16 load the block (prefetched one block ahead)
template <int UNROLL>
__global__ void __launch_bounds__(32) agw_absorb_warp_kernel(AgState* __restrict__ st, const uint8_t* __restrict__ ring, size_t ring_cap,
                                                                      const unsigned long long* __restrict__ wend,
                                                                      const unsigned long long* __restrict__ win_off, size_t n_aggs) {
    const int L = threadIdx.x & 31;
    const size_t j = blockIdx.x;   // one warp per CTA: the chains spread evenly over the SMs
    if (j >= n_aggs) return;
    AgState* s = st + j;
    if (s->phase != 0) return;
    const uint32_t absorbed = s->absorbed + (uint32_t)(win_off[j + 1] - win_off[j]);
    const bool closing = absorbed == s->total;
    unsigned long long rpos = s->rpos;
    const unsigned long long end = wend[j] & ~AGW_TEXT_FLAGS;
    __syncwarp();
    if (wend[j] & AGW_TEXT_OVERRUN) {   // the writer overwrote bytes this sponge had not absorbed: never hash them
        if (L == 0) s->phase = 0xDEAD;
        return;
    }
    if (L == 0) s->absorbed = absorbed;
    if (rpos + KECCAK_RATE > end && !closing) return;
    const Lane c = lane_init(L);
    uint32_t lo = 0, hi = 0;
    if (L < 25) { const unsigned long long v = s->a[L]; lo = (uint32_t)v; hi = (uint32_t)(v >> 32); }
    const uint8_t* rj = ring + j * ring_cap;
    bool have = rpos + KECCAK_RATE <= end;
    uint2 nxt = make_uint2(0u, 0u);
    if (have && L < KECCAK_RATE_LANES) nxt = __ldcs(reinterpret_cast<const uint2*>(rj + (rpos % ring_cap)) + L);
    while (have) {
        lo ^= nxt.x;
        hi ^= nxt.y;
        rpos += KECCAK_RATE;
        have = rpos + KECCAK_RATE <= end;
        nxt = make_uint2(0u, 0u);
        if (have && L < KECCAK_RATE_LANES) nxt = __ldcs(reinterpret_cast<const uint2*>(rj + (rpos % ring_cap)) + L);
        keccak_f_warp_u<UNROLL>(lo, hi, c, L);
    }
    if (closing) {
        const uint32_t rem = (uint32_t)(end - rpos);
        if (L < KECCAK_RATE_LANES) {
            unsigned long long w = 0;
            if ((uint32_t)(8 * L) < rem) {
                w = *(reinterpret_cast<const unsigned long long*>(rj + (rpos % ring_cap)) + L);
                const uint32_t valid = rem - 8 * L;
                if (valid < 8) w &= (1ULL << (8 * valid)) - 1ULL;
            }
            if ((rem >> 3) == (uint32_t)L) w ^= 0x1FULL << (8 * (rem & 7));
            if (L == 16) w ^= 0x8000000000000000ULL;
            lo ^= (uint32_t)w;
            hi ^= (uint32_t)(w >> 32);
        }
        keccak_f_warp_u<UNROLL>(lo, hi, c, L);
        rpos = end;
        if (L == 0) { s->phase = 1; s->bpos = 0; }
    }
    if (L < 25) s->a[L] = ((unsigned long long)hi << 32) | lo;
    if (L == 0) s->rpos = rpos;
}

// Time-sliced form of the ring-fed absorb for batches a little larger than two chains per SM sub-partition (8 per SM): three chains
// on one sub-partition saturate its shuffle port (3 x 81 issue cycles per round against a 197-cycle round) and run at 247 cycles per
// round, which would set the time of the whole launch.  Here a launch serves exactly n_active chains -- the aggregates
// first_active .. first_active + n_active - 1 (mod n_aggs) -- and every chain absorbs at most budget_blocks blocks, so all CTAs of
// a launch finish together; the caller rotates the set that rests.  A chain closes its message when the text is complete
// (bit 63 of wend) and fewer than 136 bytes are left.
template <int UNROLL>
__global__ void __launch_bounds__(128) agw_absorb_sliced_kernel(AgState* __restrict__ st, const uint8_t* __restrict__ ring, size_t ring_cap,
                                                                const unsigned long long* __restrict__ wend, size_t n_aggs, size_t first_active,
                                                                size_t n_active, unsigned long long budget_blocks) {
    // four chains per CTA, one per SM sub-partition: the even spread over the sub-partitions does not depend on what else is resident
    const int L = threadIdx.x & 31;
    const size_t slot = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (slot >= n_active) return;
    size_t j = first_active + slot;
    if (j >= n_aggs) j -= n_aggs;
    AgState* s = st + j;
    if (s->phase != 0) return;
    const unsigned long long we = wend[j];
    const unsigned long long end = we & ~AGW_TEXT_FLAGS;
    const bool complete = (we & AGW_TEXT_COMPLETE) != 0;
    unsigned long long rpos = s->rpos;
    if ((we & AGW_TEXT_OVERRUN) || end - rpos > ring_cap) {   // the text writer lapped this chain (budget too small for the text produced): fail loudly, never hash garbage
        if (L == 0) s->phase = 0xDEAD;
        return;
    }
    unsigned long long nb = (end - rpos) / KECCAK_RATE;
    if (nb > budget_blocks) nb = budget_blocks;
    const bool closing = complete && rpos + nb * KECCAK_RATE + KECCAK_RATE > end;
    if (nb == 0 && !closing) return;
    const Lane c = lane_init(L);
    uint32_t lo = 0, hi = 0;
    if (L < 25) { const unsigned long long v = s->a[L]; lo = (uint32_t)v; hi = (uint32_t)(v >> 32); }
    const uint8_t* rj = ring + j * ring_cap;
    uint2 nxt = make_uint2(0u, 0u);
    if (nb && L < KECCAK_RATE_LANES) nxt = __ldcs(reinterpret_cast<const uint2*>(rj + (rpos % ring_cap)) + L);
    for (unsigned long long b = 0; b < nb; ++b) {
        lo ^= nxt.x;
        hi ^= nxt.y;
        rpos += KECCAK_RATE;
        nxt = make_uint2(0u, 0u);
        if (b + 1 < nb && L < KECCAK_RATE_LANES) nxt = __ldcs(reinterpret_cast<const uint2*>(rj + (rpos % ring_cap)) + L);
        keccak_f_warp_u<UNROLL>(lo, hi, c, L);
    }
    if (closing) {
        const uint32_t rem = (uint32_t)(end - rpos);
        if (L < KECCAK_RATE_LANES) {
            unsigned long long w = 0;
            if ((uint32_t)(8 * L) < rem) {
                w = *(reinterpret_cast<const unsigned long long*>(rj + (rpos % ring_cap)) + L);
                const uint32_t valid = rem - 8 * L;
                if (valid < 8) w &= (1ULL << (8 * valid)) - 1ULL;
            }
            if ((rem >> 3) == (uint32_t)L) w ^= 0x1FULL << (8 * (rem & 7));
            if (L == 16) w ^= 0x8000000000000000ULL;
            lo ^= (uint32_t)w;
            hi ^= (uint32_t)(w >> 32);
        }
        keccak_f_warp_u<UNROLL>(lo, hi, c, L);
        rpos = end;
        if (L == 0) { s->phase = 1; s->bpos = 0; s->absorbed = s->total; }
    }
    if (L < 25) s->a[L] = ((unsigned long long)hi << 32) | lo;
    if (L == 0) s->rpos = rpos;
}

template <int UNROLL>
__global__ void __launch_bounds__(32) agw_squeeze_warp_kernel(AgState* __restrict__ st, const unsigned long long* __restrict__ win_off, size_t n_aggs,
                                                                       uint32_t n_ag, uint8_t* __restrict__ xof, size_t xof_cap,
                                                                       uint32_t* __restrict__ bpos_out) {
    const int L = threadIdx.x & 31;
    const size_t j = blockIdx.x;
    if (j >= n_aggs) return;
    AgState* s = st + j;
    const unsigned long long cnt = win_off[j + 1] - win_off[j];
    if (cnt == 0 || s->phase != 1) return;
    const Lane c = lane_init(L);
    uint32_t lo = 0, hi = 0;
    if (L < 25) { const unsigned long long v = s->a[L]; lo = (uint32_t)v; hi = (uint32_t)(v >> 32); }
    const uint32_t bpos = s->bpos;
    __syncwarp();
    if (L == 0) bpos_out[j] = bpos;
    uint2* out = reinterpret_cast<uint2*>(xof + j * xof_cap);
    const unsigned long long need = bpos + cnt * n_ag;
    unsigned long long o = 0;
    if (L < KECCAK_RATE_LANES) __stcs(out + L, make_uint2(lo, hi));
    while (o + KECCAK_RATE < need) {
        keccak_f_warp_u<UNROLL>(lo, hi, c, L);
        o += KECCAK_RATE;
        if (L < KECCAK_RATE_LANES) __stcs(out + o / 8 + L, make_uint2(lo, hi));
    }
    if (L < 25) s->a[L] = ((unsigned long long)hi << 32) | lo;
    if (L == 0) { s->bpos = (uint32_t)(need - o); s->produced += (uint32_t)cnt; }
}

}  // namespace warpag

size_t ag_state_bytes() { return sizeof(AgState); }
size_t ag_state_rpos_offset() { return offsetof(AgState, rpos); }

int launch_ag_state_init(const fz_ctx* ctx, void* state, const unsigned long long* agg_off, size_t n_aggs, cudaStream_t st) {
    if (n_aggs == 0) return 0;
    (void)ctx;
    FZ_LAUNCH(ag_state_init_kernel, (int)((n_aggs + 127) / 128), 128, 0, st, (AgState*)state, agg_off, n_aggs);
    return check(cudaGetLastError(), "hash_ag state init launch");
}

static int warp_smem_attr(const fz_ctx* ctx) {
    const int smem = warpag::WARPS * warpag::SMEM_PER_WARP;
    if (attr_needed(ctx, ATTR_WARP_HASH)) {
        if (check(cudaFuncSetAttribute(warpag::hash_ag_absorb_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "hash_ag absorb smem attr"))
            return -2;
        if (check(cudaFuncSetAttribute(warpag::hash_ag_squeeze_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "hash_ag squeeze smem attr"))
            return -2;
    }
    return 0;
}

// one serial sponge per aggregate: a warp per aggregate minimises the chain latency, a thread per aggregate maximises
// throughput once there are enough aggregates to fill the machine
// ... and two lanes per aggregate in between (ring-fed kernels only; the one-shot hash_ag_* kernels keep the two-way choice)
static bool use_warp(const fz_ctx* ctx, size_t n_aggs) { return ctx->hash_ag_impl == 3 || ((ctx->hash_ag_impl == 0 || ctx->hash_ag_impl == 4) && n_aggs < 2816); }
// measured crossovers (profiles/r02_keccak_lab2_v23.txt, r01_keccak_lab_v3.txt): a warp per chain runs 197 cycles per round at two chains per
// sub-partition (1,184 chains; the time-sliced absorb keeps that up to 1,776) and 247 at three; the pair layout 250 up to 9,472 chains, 351 at
// 18,944, 545 at 37,888; a thread per chain 370-400 up to 18,944 chains, 554 at 37,888 and wins from there on (4.27 against 3.94 G Keccak-f/s)
static int chain_layout(const fz_ctx* ctx, size_t n_aggs) {   // 2 thread, 3 warp, 4 pair
    if (ctx->hash_ag_impl == 2 || ctx->hash_ag_impl == 3 || ctx->hash_ag_impl == 4) return ctx->hash_ag_impl;
    return n_aggs <= 1776 ? 3 : n_aggs < 40960 ? 4 : 2;
}

int launch_hash_ag_absorb(const fz_ctx* ctx, void* state, const int32_t* vk_l_s, const int32_t* vk_r_s, const uint8_t* dg_s, const int32_t* c_hat_s,
                          const unsigned long long* win_off, size_t n_aggs, cudaStream_t st) {
    if (n_aggs == 0) return 0;
    const int level = ctx->secpar == 256 ? 1 : 0;
    if (use_warp(ctx, n_aggs)) {
        if (warp_smem_attr(ctx)) return -2;
        FZ_LAUNCH(warpag::hash_ag_absorb_warp_kernel, (int)((n_aggs + warpag::WARPS - 1) / warpag::WARPS), warpag::WARPS * 32, warpag::WARPS * warpag::SMEM_PER_WARP, st, (AgState*)state, vk_l_s, vk_r_s, dg_s, c_hat_s, win_off, n_aggs, level, ctx->d);
    } else {
        FZ_LAUNCH(hash_ag_absorb_thread_kernel, (int)((n_aggs + 31) / 32), 32, 0, st, (AgState*)state, vk_l_s, vk_r_s, dg_s, c_hat_s, win_off, n_aggs, level,
                                                                                ctx->d);
    }
    return check(cudaGetLastError(), "hash_ag absorb launch");
}

int launch_hash_ag_squeeze(const fz_ctx* ctx, void* state, const unsigned long long* win_off, size_t n_aggs, int32_t* coef_out, cudaStream_t st) {
    if (n_aggs == 0) return 0;
    if (use_warp(ctx, n_aggs)) {
        if (warp_smem_attr(ctx)) return -2;
        FZ_LAUNCH(warpag::hash_ag_squeeze_warp_kernel, (int)((n_aggs + warpag::WARPS - 1) / warpag::WARPS), warpag::WARPS * 32, warpag::WARPS * warpag::SMEM_PER_WARP, st, (AgState*)state, win_off, n_aggs, ctx->d, ctx->omega_ag, ctx->bytes_coef, ctx->bytes_idx,
                                                    (uint32_t)ctx->n_ag, coef_out);
    } else {
        FZ_LAUNCH(hash_ag_squeeze_thread_kernel, (int)((n_aggs + 31) / 32), 32, 0, st, (AgState*)state, win_off, n_aggs, ctx->d, ctx->omega_ag, ctx->bytes_coef,
                                                                                 ctx->bytes_idx, (uint32_t)ctx->n_ag, coef_out);
    }
    return check(cudaGetLastError(), "hash_ag squeeze launch");
}

// The chain kernels use no shared memory, so by default their CTAs run on SMs configured with the smallest shared-memory
// carve-out -- and an SM cannot change its carve-out while CTAs are resident: every kernel of the side stream that needs shared
// memory (torch's scans, hash_ch, the text formatter) then waits until a whole SM has drained, i.e. almost to the end of the
// chain kernel (44 of 52 ms, found with tools/agg_host_probe.py).  Asking for the largest carve-out keeps the SMs hospitable.
static int chain_carveout(const fz_ctx* ctx) {
    if (!attr_needed(ctx, ATTR_CHAIN_CARVEOUT)) return 0;
    const void* ks[] = {(const void*)agw_absorb_thread_kernel,          (const void*)agw_squeeze_thread_kernel,
                        (const void*)warpag::agw_absorb_warp_kernel<24>, (const void*)warpag::agw_squeeze_warp_kernel<24>,
                        (const void*)warpag::agw_absorb_sliced_kernel<24>, (const void*)pairag::agw_absorb_pair_kernel,
                        (const void*)pairag::agw_squeeze_pair_kernel};
    for (const void* k : ks)
        if (check(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared), "chain kernel carve-out")) return -2;
    return 0;
}

int launch_agw_absorb(const fz_ctx* ctx, void* state, const uint8_t* ring, size_t ring_cap, const unsigned long long* wend,
                      const unsigned long long* win_off, size_t n_aggs, cudaStream_t st) {
    if (n_aggs == 0) return 0;
    if (chain_carveout(ctx)) return -2;
    const int layout = chain_layout(ctx, n_aggs);
    if (layout == 3)
        FZ_LAUNCH(warpag::agw_absorb_warp_kernel<24>, (unsigned)n_aggs, 32, 0, st, (AgState*)state, ring, ring_cap, wend, win_off, n_aggs);
    else if (layout == 4)
        FZ_LAUNCH(pairag::agw_absorb_pair_kernel, (unsigned)((n_aggs + 63) / 64), pairag::PAIR_WARPS * 32, 0, st, (AgState*)state, ring, ring_cap, wend, win_off, n_aggs);
    else
        FZ_LAUNCH(agw_absorb_thread_kernel, (int)((n_aggs + 31) / 32), 32, 0, st, (AgState*)state, ring, ring_cap, wend, win_off, n_aggs);
    return check(cudaGetLastError(), "agw_absorb launch");
}

int launch_agw_absorb_sliced(const fz_ctx* ctx, void* state, const uint8_t* ring, size_t ring_cap, const unsigned long long* wend, size_t n_aggs,
                             size_t first_active, size_t n_active, unsigned long long budget_blocks, cudaStream_t st) {
    if (n_aggs == 0 || n_active == 0) return 0;
    if (chain_carveout(ctx)) return -2;
    FZ_LAUNCH(warpag::agw_absorb_sliced_kernel<24>, (unsigned)((n_active + 3) / 4), 128, 0, st, (AgState*)state, ring, ring_cap, wend, n_aggs, first_active, n_active,
                                                                                        budget_blocks);
    return check(cudaGetLastError(), "agw_absorb_sliced launch");
}

int launch_agw_squeeze(const fz_ctx* ctx, void* state, const unsigned long long* win_off, size_t n_aggs, uint8_t* xof, size_t xof_cap, uint32_t* bpos_out,
                       cudaStream_t st) {
    if (n_aggs == 0) return 0;
    if (chain_carveout(ctx)) return -2;
    const int layout = chain_layout(ctx, n_aggs);
    if (layout == 3)
        FZ_LAUNCH(warpag::agw_squeeze_warp_kernel<24>, (unsigned)n_aggs, 32, 0, st, (AgState*)state, win_off, n_aggs, (uint32_t)ctx->n_ag, xof, xof_cap, bpos_out);
    else if (layout == 4)
        FZ_LAUNCH(pairag::agw_squeeze_pair_kernel, (unsigned)((n_aggs + 63) / 64), pairag::PAIR_WARPS * 32, 0, st, (AgState*)state, win_off, n_aggs, (uint32_t)ctx->n_ag,
                                                                                                          xof, xof_cap, bpos_out);
    else
        FZ_LAUNCH(agw_squeeze_thread_kernel, (int)((n_aggs + 31) / 32), 32, 0, st, (AgState*)state, win_off, n_aggs, (uint32_t)ctx->n_ag, xof, xof_cap, bpos_out);
    return check(cudaGetLastError(), "agw_squeeze launch");
}

}  // namespace fz
