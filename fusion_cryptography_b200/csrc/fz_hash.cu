// fz_hash.cu -- device-side derivation of challenges and aggregation coefficients, bit-exact
// with the reference's hashlib / str() pipeline:
//   * generic batched SHA3-256 / SHAKE256 over byte strings (fusion.py:405-409 prehash)
//   * repr() formatter: int32 coefficients -> the exact ASCII of str(OneTimeVerificationKey) and
//     str(SignatureChallenge) (fusion.py:328-329, 382-383; matrices.py:40-41; polynomials.py:257-258)
//   * hash_ch: SHAKE256(dst "," str(vk) "," str(prehash)) squeezed straight into the
//     sparse-ternary decoder (fusion.py:412-419, 422-508)
//   * hash_ag: one serial SHAKE256 per aggregate over str(list(zip(keys, prehashes, challenges)))
//     and slice decoding (fusion.py:573-629)
#include <cstdio>
#include "fz_common.cuh"
#include "fz_internal.h"
#include "fz_keccak.cuh"

namespace fz {

// ------------------------------------------------------------------------------------------
// generic sponge: one thread per message, bytes absorbed from global memory
// out[i*out_stride .. +out_len) = XOF(prefix || data[offs[i]..offs[i+1]))
__global__ void __launch_bounds__(128) sponge_bytes_kernel(const uint8_t* __restrict__ data, const unsigned long long* __restrict__ offs, size_t n,
                                                             uint32_t prefix, int prefix_len, uint8_t suffix, uint8_t* __restrict__ out,
                                                             size_t out_stride, uint32_t out_len) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* msg = data + offs[i];
    const size_t mlen = (size_t)(offs[i + 1] - offs[i]);
    const size_t total = mlen + prefix_len;
    unsigned long long a[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) a[k] = 0;
    unsigned long long blk[KECCAK_RATE_LANES];
    size_t pos = 0;
    for (;;) {
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) blk[k] = 0;
        const size_t take = total - pos < (size_t)KECCAK_RATE ? total - pos : (size_t)KECCAK_RATE;
        for (size_t j = 0; j < take; ++j) {
            const size_t g = pos + j;
            const uint8_t b = g < (size_t)prefix_len ? (uint8_t)(prefix >> (8 * g)) : msg[g - prefix_len];
            blk[j >> 3] |= (unsigned long long)b << (8 * (j & 7));
        }
        pos += take;
        const bool last = take < (size_t)KECCAK_RATE;
        if (last) {
            blk[take >> 3] ^= (unsigned long long)suffix << (8 * (take & 7));
            blk[16] ^= 0x8000000000000000ULL;
        }
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) a[k] ^= blk[k];
        keccak_f(a);
        if (last) break;
    }
    uint8_t* o = out + i * out_stride;
    uint32_t produced = 0;
    while (produced < out_len) {
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) blk[k] = a[k];
        const uint32_t take = out_len - produced < (uint32_t)KECCAK_RATE ? out_len - produced : (uint32_t)KECCAK_RATE;
        for (uint32_t j = 0; j < take; ++j) o[produced + j] = (uint8_t)(blk[j >> 3] >> (8 * (j & 7)));
        produced += take;
        if (produced < out_len) keccak_f(a);
    }
}

}  // namespace fz
#include "fz_format.cuh"    // repr formatter of one warp (its __constant__ table is defined at this point of the translation unit)
namespace fz {

constexpr int FMT_WARPS = 4;
constexpr int FMT_SMEM_PER_WARP = 11264;  // >= longest piece "(" VK ", " INT ", " CH ")": 768 values x 13 chars + 0.7 KB fixed text

__device__ __forceinline__ void copy_out(uint8_t* __restrict__ dst, const uint8_t* src, int len, int lane) {
    // dst is 16-byte aligned, src is 16-byte aligned shared memory; pad bytes beyond len are don't-care
    const int n16 = (len + 15) >> 4;
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (int i = lane; i < n16; i += 32) d4[i] = s4[i];
}

// mode 0: hash_ch message   dst "," VK "," INT                                  (fusion.py:416-418)
// mode 1: hash_ag piece     "(" VK ", " INT ", " "SignatureChallenge(c_hat=" POLY ")" ")"   (fusion.py:589)
// One warp per signer; text is assembled in shared memory and copied out with 16-byte stores.
__global__ void __launch_bounds__(FMT_WARPS * 32) format_kernel(int mode, int level, int d, uint8_t dst0, const int32_t* __restrict__ vk_l,
                                                                 const int32_t* __restrict__ vk_r, const uint8_t* __restrict__ digests,
                                                                 const int32_t* __restrict__ c_hat, size_t n, uint8_t* __restrict__ out, size_t out_stride,
                                                                 uint32_t* __restrict__ out_len) {
    FZ_DYN_SHARED __align__(16) uint8_t fsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* buf = fsm + (size_t)warp * FMT_SMEM_PER_WARP;
    const FmtConst& fc = c_fmt[level];
    for (size_t i = (size_t)blockIdx.x * FMT_WARPS + warp; i < n; i += (size_t)gridDim.x * FMT_WARPS) {
        int o = 0;
        if (mode == 0) {
            if (lane == 0) { buf[0] = dst0; buf[1] = 1; buf[2] = ','; }
            o = 3;
        } else {
            if (lane == 0) buf[0] = '(';
            o = 1;
        }
        const char h1[] = "OneTimeVerificationKey(left_vk_hat=";
        const char h2[] = ", right_vk_hat=";
        o += put_str(buf + o, h1, 35, lane);
        o += warp_format_gm1(buf + o, fc, vk_l + i * d, d, lane);
        o += put_str(buf + o, h2, 15, lane);
        o += warp_format_gm1(buf + o, fc, vk_r + i * d, d, lane);
        if (lane == 0) buf[o] = ')';
        o += 1;
        if (mode == 0) {
            if (lane == 0) buf[o] = ',';
            o += 1;
        } else {
            if (lane == 0) { buf[o] = ','; buf[o + 1] = ' '; }
            o += 2;
        }
        __syncwarp();
        int li = 0;
        if (lane == 0) li = format_le256(buf + o, digests + i * 32);
        li = __shfl_sync(0xffffffffu, li, 0);
        o += li;
        if (mode == 1) {
            const char h3[] = ", SignatureChallenge(c_hat=";
            o += put_str(buf + o, h3, 27, lane);
            o += warp_format_poly(buf + o, fc, c_hat + i * d, d, lane);
            if (lane == 0) { buf[o] = ')'; buf[o + 1] = ')'; }
            o += 2;
        }
        __syncwarp();
        // zero the tail of the last 16-byte group so the copy never publishes stale bytes
        for (int k = o + lane; k < ((o + 15) & ~15); k += 32) buf[k] = 0;
        __syncwarp();
        copy_out(out + i * out_stride, buf, o, lane);
        if (lane == 0) out_len[i] = (uint32_t)o;
        __syncwarp();
    }
}

}  // namespace fz
#include "fz_decoder.cuh"   // the byte-driven sparse ternary decoder (its __constant__ table is defined at this point of the translation unit)
namespace fz {

constexpr int HC_THREADS = 64;

// hash_ch core: one thread per signer.  msg: [n][stride] (8-byte aligned), 17 x 8-byte loads per block.
// Squeezes n_used bytes and feeds them to the decoder; writes d int32 coefficients (NOT yet transformed).
__global__ void __launch_bounds__(HC_THREADS) shake_decode_kernel(const uint8_t* __restrict__ msg, size_t stride, const uint32_t* __restrict__ len,
                                                                    size_t n, int d, int weight, uint32_t bound, int bpc, int bpi, uint32_t n_used,
                                                                    int32_t* __restrict__ coef_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long* m = reinterpret_cast<const unsigned long long*>(msg + i * stride);
    const uint32_t L = len[i];
    unsigned long long a[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) a[k] = 0;
    const uint32_t nfull = L / KECCAK_RATE;
    for (uint32_t b = 0; b < nfull; ++b) {
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) a[k] ^= m[(size_t)b * KECCAK_RATE_LANES + k];
        keccak_f(a);
    }
    {
        unsigned long long blk[KECCAK_RATE_LANES];
        const uint32_t rem = L - nfull * KECCAK_RATE;
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) {
            unsigned long long w = 0;
            if ((uint32_t)(8 * k) < rem) {
                w = m[(size_t)nfull * KECCAK_RATE_LANES + k];
                const uint32_t valid = rem - 8 * k;
                if (valid < 8) w &= (1ULL << (8 * valid)) - 1ULL;
            }
            blk[k] = w;
        }
        blk[rem >> 3] ^= 0x1FULL << (8 * (rem & 7));
        blk[16] ^= 0x8000000000000000ULL;
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) a[k] ^= blk[k];
        keccak_f(a);
    }
    int32_t c[256];
    for (int k = 0; k < d; ++k) c[k] = 0;
    Decoder s;
    dec_init(s, c, d, weight, bound, bpc, bpi);
    uint32_t produced = 0;
    while (produced < n_used && s.phase != 3) {
        unsigned long long blk[KECCAK_RATE_LANES];
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) blk[k] = a[k];
        const uint32_t take = n_used - produced < (uint32_t)KECCAK_RATE ? n_used - produced : (uint32_t)KECCAK_RATE;
        for (uint32_t j = 0; j < take && s.phase != 3; ++j) dec_byte(s, c, (uint32_t)(blk[j >> 3] >> (8 * (j & 7))) & 0xffu);
        produced += take;
        if (produced < n_used && s.phase != 3) keccak_f(a);
    }
    dec_drain(s, c);
    int32_t* o = coef_out + i * d;
    for (int k = 0; k < d; ++k) o[k] = c[k];
}

// standalone decoder over caller-supplied bytes (one thread per polynomial): bytes[i*stride .. +blen)
__global__ void __launch_bounds__(HC_THREADS) decode_kernel(const uint8_t* __restrict__ bytes, size_t stride, uint32_t blen, size_t n, int d, int weight,
                                                             uint32_t bound, int bpc, int bpi, int32_t* __restrict__ coef_out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* b = bytes + i * stride;
    int32_t c[256];
    for (int k = 0; k < d; ++k) c[k] = 0;
    Decoder s;
    dec_init(s, c, d, weight, bound, bpc, bpi);
    for (uint32_t j = 0; j < blen && s.phase != 3; ++j) dec_byte(s, c, b[j]);
    dec_drain(s, c);
    int32_t* o = coef_out + i * d;
    for (int k = 0; k < d; ++k) o[k] = c[k];
}

// ------------------------------------------------------------------------------------------
// hash_ag message layout: per aggregate j with sorted signer list order[agg_off[j] .. agg_off[j+1]):
//   message = dst(2) "," "[" piece_0 ", " piece_1 ... "]"
// One CTA per aggregate computes the exclusive scan of piece lengths (+2 separators).
__global__ void __launch_bounds__(256) ag_layout_kernel(const uint32_t* __restrict__ piece_len, const uint32_t* __restrict__ order,
                                                          const unsigned long long* __restrict__ agg_off, size_t msg_stride,
                                                          unsigned long long* __restrict__ piece_pos, uint32_t* __restrict__ msg_len) {
    __shared__ unsigned long long warp_sum[8];
    __shared__ unsigned long long carry;
    const size_t j = blockIdx.x;
    const size_t lo = agg_off[j], hi = agg_off[j + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 4;  // dst0 dst1 ',' '['
    __syncthreads();
    for (size_t base = lo; base < hi; base += 256) {
        const size_t p = base + threadIdx.x;
        unsigned long long v = 0;
        if (p < hi) v = (unsigned long long)piece_len[order ? order[p] : (uint32_t)p] + (p > lo ? 2 : 0);
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        unsigned long long wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += warp_sum[w];
        const unsigned long long c0 = carry;
        if (p < hi) piece_pos[p] = j * msg_stride + c0 + wbase + x - v + (p > lo ? 2 : 0);  // start of the piece text itself
        __syncthreads();
        if (threadIdx.x == 255) carry = c0 + wbase + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) msg_len[j] = (uint32_t)(carry + 1);  // closing ']'
}

// copies piece text to its place in the aggregate's message; one CTA per sorted position
__global__ void __launch_bounds__(128) ag_pack_kernel(const uint8_t* __restrict__ pieces, size_t piece_stride, const uint32_t* __restrict__ piece_len,
                                                        const uint32_t* __restrict__ order, const unsigned long long* __restrict__ agg_off,
                                                        const unsigned long long* __restrict__ piece_pos, size_t msg_stride, uint8_t dst0,
                                                        uint8_t* __restrict__ msgs, const uint32_t* __restrict__ msg_len, size_t n_aggs, size_t n_total) {
    const size_t p = blockIdx.x;
    if (p < n_total) {
        const uint32_t sidx = order ? order[p] : (uint32_t)p;
        const uint8_t* src = pieces + (size_t)sidx * piece_stride;
        const uint32_t L = piece_len[sidx];
        uint8_t* dst = msgs + piece_pos[p];
        for (uint32_t k = threadIdx.x; k < L; k += blockDim.x) dst[k] = src[k];
        // separator before every piece but the first of its aggregate: detect via position
        if (threadIdx.x == 0) {
            // find whether p is the first of its aggregate: piece_pos[p] - (agg base) == 4
            const unsigned long long rel = piece_pos[p] % msg_stride;
            if (rel != 4) { dst[-2] = ','; dst[-1] = ' '; }
        }
    } else {
        // header / footer blocks: one per aggregate
        const size_t j = p - n_total;
        if (j < n_aggs && threadIdx.x == 0) {
            uint8_t* m = msgs + j * msg_stride;
            m[0] = dst0; m[1] = 2; m[2] = ','; m[3] = '[';
            m[msg_len[j] - 1] = ']';
        }
    }
}

// one serial sponge per aggregate: absorb msgs[j], squeeze out_len[j] bytes to xof + xof_off[j]
__global__ void __launch_bounds__(32) shake_ag_kernel(const uint8_t* __restrict__ msgs, size_t msg_stride, const uint32_t* __restrict__ msg_len,
                                                        const unsigned long long* __restrict__ agg_off, uint32_t n_ag, uint8_t* __restrict__ xof,
                                                        size_t n_aggs) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_aggs) return;
    const unsigned long long* m = reinterpret_cast<const unsigned long long*>(msgs + j * msg_stride);
    const uint32_t L = msg_len[j];
    unsigned long long a[25];
#pragma unroll
    for (int k = 0; k < 25; ++k) a[k] = 0;
    const uint32_t nfull = L / KECCAK_RATE;
    for (uint32_t b = 0; b < nfull; ++b) {
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) a[k] ^= m[(size_t)b * KECCAK_RATE_LANES + k];
        keccak_f(a);
    }
    {
        unsigned long long blk[KECCAK_RATE_LANES];
        const uint32_t rem = L - nfull * KECCAK_RATE;
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) {
            unsigned long long w = 0;
            if ((uint32_t)(8 * k) < rem) {
                w = m[(size_t)nfull * KECCAK_RATE_LANES + k];
                const uint32_t valid = rem - 8 * k;
                if (valid < 8) w &= (1ULL << (8 * valid)) - 1ULL;
            }
            blk[k] = w;
        }
        blk[rem >> 3] ^= 0x1FULL << (8 * (rem & 7));
        blk[16] ^= 0x8000000000000000ULL;
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) a[k] ^= blk[k];
        keccak_f(a);
    }
    // squeeze (hi - lo) * n_ag bytes, contiguous at xof + lo * n_ag  (fusion.py:583-591, 599-605)
    const unsigned long long lo = agg_off[j], hi = agg_off[j + 1];
    unsigned long long total = (hi - lo) * (unsigned long long)n_ag, produced = 0;
    uint8_t* o = xof + lo * (unsigned long long)n_ag;
    while (produced < total) {
        unsigned long long blk[KECCAK_RATE_LANES];
#pragma unroll
        for (int k = 0; k < KECCAK_RATE_LANES; ++k) blk[k] = a[k];
        const uint32_t take = total - produced < (unsigned long long)KECCAK_RATE ? (uint32_t)(total - produced) : (uint32_t)KECCAK_RATE;
        for (uint32_t k = 0; k < take; ++k) o[produced + k] = (uint8_t)(blk[k >> 3] >> (8 * (k & 7)));
        produced += take;
        if (produced < total) keccak_f(a);
    }
}

// ------------------------------------------------------------------------------------------
// Windowed aggregation hash, text side (the sponge side is in fz_hashstream.cu).  The text a window of sorted signers adds
// to each aggregate's message is generated in parallel, a warp per signer, straight into a per-aggregate byte RING in HBM,
// so the serial sponge chain only loads 17 lanes per block.  The message of aggregate j is
//     dst 2 "," "[" piece_0 ", " piece_1 ... "]",    piece = "(" VK ", " INT ", " "SignatureChallenge(c_hat=" POLY ")" ")"
// and the bytes signer r contributes are [opener] piece [closer]: opener = dst 2 ",[" for r = 0 else ", "; closer = "]" for
// the aggregate's last signer.  Three passes: lengths -> per-aggregate scan of stream offsets -> format + write.
constexpr int AGW_SMEM_PER_WARP = FMT_SMEM_PER_WARP + 64;   // piece + opener / closer + 16 bytes of alignment slack

__device__ __forceinline__ int values_len(const int32_t* __restrict__ vals, int d, int lane) {
    int my = 0;
    for (int k = lane; k < d; k += 32) my += dec_len(vals[k]);
#pragma unroll
    for (int o = 16; o; o >>= 1) my += __shfl_xor_sync(0xffffffffu, my, o);
    return my + 2 * (d - 1);
}

// The decimal text of the prehash (str(int.from_bytes(digest, "little")), up to 78 digits) is a serial big-number conversion of about 1,500 instructions.
// The length and format kernels below used to run it on lane 0 of the signer's warp -- 1,500 WARP instructions each, a third of what the two kernels
// executed per signer (ncu: 6.9 K + 1.7 K warp instructions).  It is now computed once, by a thread per signer (47 warp instructions per signer), into
// 80-byte slots the two kernels read.
constexpr int INT_TEXT_STRIDE = 80;
__global__ void __launch_bounds__(128) agw_int_text_kernel(const uint8_t* __restrict__ digests, size_t n, uint8_t* __restrict__ int_text,
                                                            uint32_t* __restrict__ int_len) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    __align__(16) uint8_t t[96];
#pragma unroll
    for (int k = 0; k < 96; ++k) t[k] = 0;
    const int li = format_le256(t, digests + i * 32);
    uint4* dst = reinterpret_cast<uint4*>(int_text + i * INT_TEXT_STRIDE);
    const uint4* src = reinterpret_cast<const uint4*>(t);
#pragma unroll
    for (int k = 0; k < INT_TEXT_STRIDE / 16; ++k) dst[k] = src[k];
    int_len[i] = (uint32_t)li;
}

__global__ void __launch_bounds__(128) agw_len_kernel(int level, int d, const int32_t* __restrict__ vk_l, const int32_t* __restrict__ vk_r,
                                                       const uint32_t* __restrict__ int_len, const int32_t* __restrict__ c_hat,
                                                       const unsigned long long* __restrict__ win_off, const uint32_t* __restrict__ agg_of,
                                                       const unsigned long long* __restrict__ agg_off, size_t n_win, uint32_t w0,
                                                       uint32_t* __restrict__ piece_len) {
    const int lane = threadIdx.x & 31;
    const size_t i = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (i >= n_win) return;
    const FmtConst& fc = c_fmt[level];
    const uint32_t j = agg_of[i];
    const unsigned long long rank = w0 + (i - win_off[j]), total = agg_off[j + 1] - agg_off[j];
    const int poly_fixed = fc.poly_head_len + 2;                 // head ... "])"
    const int gm_fixed = fc.gm_head_len + poly_fixed + 3;        // head POLY "]])"
    int len = 1 + 35 + gm_fixed + 15 + gm_fixed + 1 + 2 + 27 + poly_fixed + 2;
    len += values_len(vk_l + i * d, d, lane) + values_len(vk_r + i * d, d, lane) + values_len(c_hat + i * d, d, lane);
    const int li = (int)int_len[i];
    len += li + (rank == 0 ? 4 : 2) + (rank + 1 == total ? 1 : 0);
    if (lane == 0) piece_len[i] = (uint32_t)len;
}

// one warp per aggregate: stream offsets of this window's pieces, new end of the generated text
__global__ void __launch_bounds__(128) agw_layout_kernel(const uint32_t* __restrict__ piece_len, const unsigned long long* __restrict__ win_off,
                                                          const unsigned long long* __restrict__ agg_off, size_t n_aggs, uint8_t dst0,
                                                          unsigned long long* __restrict__ text_pos, unsigned long long* __restrict__ piece_off,
                                                          unsigned long long* __restrict__ wend, uint8_t* __restrict__ ring, size_t ring_cap, uint32_t w0,
                                                          const unsigned long long* __restrict__ rpos0, size_t rpos_stride) {
    const int lane = threadIdx.x & 31;
    const size_t j = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (j >= n_aggs) return;
    unsigned long long pos = text_pos[j];
    if (agg_off[j + 1] == agg_off[j] && pos == 0) {   // "[]"
        if (lane == 0) { uint8_t* r = ring + j * ring_cap; r[0] = dst0; r[1] = 2; r[2] = ','; r[3] = '['; r[4] = ']'; }
        pos = 5;
    }
    const size_t lo = win_off[j], hi = win_off[j + 1];
    for (size_t base = lo; base < hi; base += 32) {
        const size_t p = base + lane;
        const unsigned long long v = p < hi ? piece_len[p] : 0;
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (p < hi) piece_off[p] = pos + x - v;
        pos += __shfl_sync(0xffffffffu, x, 31);
    }
    // bit 63 of wend: the message text of this aggregate is complete (its last signer is in this window or an earlier one)
    const bool complete = (unsigned long long)w0 + (hi - lo) >= agg_off[j + 1] - agg_off[j];
    // ring discipline: everything from the sponge's read position (never ahead of the value read here: it only grows) to the new end must fit
    const bool overrun = rpos0 != nullptr && pos - rpos0[j * rpos_stride] > ring_cap;
    if (lane == 0) { text_pos[j] = pos; wend[j] = pos | (complete ? AGW_TEXT_COMPLETE : 0ULL) | (overrun ? AGW_TEXT_OVERRUN : 0ULL); }
}

// dst and src are congruent mod 16; bytes outside [dst, dst + len) are never written (neighbouring pieces are written concurrently)
__device__ __forceinline__ void copy_congruent(uint8_t* __restrict__ dst, const uint8_t* src, int len, int lane) {
    int head = (int)((16 - ((size_t)dst & 15)) & 15);
    if (head > len) head = len;
    if (lane < head) dst[lane] = src[lane];
    const int n16 = (len - head) >> 4;
    const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    for (int k = lane; k < n16; k += 32) d4[k] = s4[k];
    const int done = head + 16 * n16;
    if (lane < len - done) dst[done + lane] = src[done + lane];
}

__global__ void __launch_bounds__(FMT_WARPS * 32) agw_format_kernel(int level, int d, uint8_t dst0, const int32_t* __restrict__ vk_l,
                                                                     const int32_t* __restrict__ vk_r, const uint8_t* __restrict__ int_text,
                                                                     const uint32_t* __restrict__ int_len,
                                                                     const int32_t* __restrict__ c_hat, const unsigned long long* __restrict__ win_off,
                                                                     const uint32_t* __restrict__ agg_of, const unsigned long long* __restrict__ agg_off,
                                                                     size_t n_win, uint32_t w0, const unsigned long long* __restrict__ piece_off,
                                                                     const uint32_t* __restrict__ piece_len, uint8_t* __restrict__ ring, size_t ring_cap) {
    FZ_DYN_SHARED __align__(16) uint8_t fsm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const FmtConst& fc = c_fmt[level];
    for (size_t i = (size_t)blockIdx.x * FMT_WARPS + warp; i < n_win; i += (size_t)gridDim.x * FMT_WARPS) {
        const uint32_t j = agg_of[i];
        const unsigned long long rank = w0 + (i - win_off[j]), total = agg_off[j + 1] - agg_off[j];
        const unsigned long long off = piece_off[i];
        uint8_t* buf = fsm + (size_t)warp * AGW_SMEM_PER_WARP + (off & 15);   // congruent to the ring position mod 16
        int o = 0;
        if (rank == 0) {
            if (lane == 0) { buf[0] = dst0; buf[1] = 2; buf[2] = ','; buf[3] = '['; buf[4] = '('; }
            o = 5;
        } else {
            if (lane == 0) { buf[0] = ','; buf[1] = ' '; buf[2] = '('; }
            o = 3;
        }
        const char h1[] = "OneTimeVerificationKey(left_vk_hat=";
        const char h2[] = ", right_vk_hat=";
        o += put_str(buf + o, h1, 35, lane);
        o += warp_format_gm1(buf + o, fc, vk_l + i * d, d, lane);
        o += put_str(buf + o, h2, 15, lane);
        o += warp_format_gm1(buf + o, fc, vk_r + i * d, d, lane);
        if (lane == 0) { buf[o] = ')'; buf[o + 1] = ','; buf[o + 2] = ' '; }
        o += 3;
        __syncwarp();
        const int li = (int)int_len[i];   // the decimal prehash, computed by agw_int_text_kernel
        for (int k = lane; k < li; k += 32) buf[o + k] = int_text[i * INT_TEXT_STRIDE + k];
        o += li;
        const char h3[] = ", SignatureChallenge(c_hat=";
        o += put_str(buf + o, h3, 27, lane);
        o += warp_format_poly(buf + o, fc, c_hat + i * d, d, lane);
        if (lane == 0) { buf[o] = ')'; buf[o + 1] = ')'; buf[o + 2] = ']'; }
        o += 2 + (rank + 1 == total ? 1 : 0);
        __syncwarp();
        // o == piece_len[i] by construction (checked in debug builds through the tests' parity with the streaming sponge)
        uint8_t* rj = ring + (size_t)j * ring_cap;
        const size_t start = (size_t)(off % ring_cap);
        const int seg1 = (size_t)o <= ring_cap - start ? o : (int)(ring_cap - start);
        copy_congruent(rj + start, buf, seg1, lane);
        if (o > seg1) copy_congruent(rj, buf + seg1, o - seg1, lane);   // ring_cap is a multiple of 16: still congruent
        (void)piece_len;
        __syncwarp();
    }
}

// decode_bytes_to_agg_coefs (fusion.py:594-629) for a window: one warp per signer, n_ag squeezed bytes each, taken from the
// per-aggregate XOF buffer the sponge chain wrote (signer r of the window starts bpos[j] + r * n_ag bytes into it)
__global__ void __launch_bounds__(128) agw_decode_kernel(const uint8_t* __restrict__ xof, size_t xof_cap, const uint32_t* __restrict__ bpos,
                                                          const unsigned long long* __restrict__ win_off, const uint32_t* __restrict__ agg_of, size_t n_win,
                                                          int d, int weight, int bpc, int bpi, uint32_t n_ag, int32_t* __restrict__ coef_out) {
    FZ_DYN_SHARED __align__(16) uint8_t dsm[];
    const int L = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t i = (size_t)blockIdx.x * 4 + warp;
    if (i >= n_win) return;
    const int per_warp = (((int)n_ag + 32 + 15) & ~15) + 256 + 256;
    uint8_t* xs = dsm + (size_t)warp * per_warp;
    int8_t* cf = reinterpret_cast<int8_t*>(xs + (((int)n_ag + 32 + 15) & ~15));
    int* jdraw = reinterpret_cast<int*>(cf + 256);
    const uint32_t j = agg_of[i];
    const uint8_t* src = xof + (size_t)j * xof_cap + bpos[j] + (size_t)(i - win_off[j]) * n_ag;
    const int shift = (int)((size_t)src & 15);
    {   // aligned 16-byte loads covering [src, src + n_ag)
        const uint4* s4 = reinterpret_cast<const uint4*>(src - shift);
        const int n16 = (shift + (int)n_ag + 15) >> 4;
        uint4* x4 = reinterpret_cast<uint4*>(xs);
        for (int k = L; k < n16; k += 32) x4[k] = __ldg(s4 + k);
    }
    __syncwarp();
    const uint8_t* x = xs + shift;
    const int bfs = (weight + 7) / 8;
    for (int k = L; k < d; k += 32) {
        int v = 0;
        if (k < weight) v = 2 * ((x[bfs - 1 - k / 8] >> (k % 8)) & 1) - 1;
        cf[k] = (int8_t)v;
    }
    const int n_swaps = weight < d ? d - 1 - weight : 0;
    const int first = bfs + bpc * weight;
    // index draws whose bytes exist (possibly truncated); the others see b"" -> 0 (quirk Q4)
    const int n_full = (int)n_ag > first ? ((int)n_ag - first + bpi - 1) / bpi : 0;
    const int n_draw = n_full < n_swaps ? n_full : n_swaps;
    for (int t = L; t < n_draw; t += 32) {
        const int idx = d - 1 - t;
        const uint32_t m = (uint32_t)(idx + 1), mg = c_magic[m];
        uint32_t acc = 0;
        const int s0 = first + t * bpi;
        const int e0 = min(s0 + bpi, (int)n_ag);
        for (int q = s0; q < e0; ++q) {
            const uint32_t v = (acc << 8) | x[q];
            acc = v - __umulhi(v, mg) * m;
        }
        jdraw[t] = (int)acc;
    }
    __syncwarp();
    if (L == 0) {   // the draws that have bytes: a data-dependent sequence of swaps, one lane
        for (int t = 0; t < n_draw; ++t) {
            const int idx = d - 1 - t, jj = jdraw[t];
            const int8_t tmp = cf[idx]; cf[idx] = cf[jj]; cf[jj] = tmp;
        }
    }
    __syncwarp();
    // the draws past the end of the squeezed bytes all see b"" -> 0 (quirk Q4): swap(cf[idx], cf[0]) for idx = a, a - 1, ..., b is the rotation
    // cf[a] <- cf[0], cf[k] <- cf[k + 1] (b <= k < a), cf[0] <- cf[b]: done by the whole warp (about as many steps as there are real draws at secpar 256)
    const int a = d - 1 - n_draw, b = d - n_swaps;
    if (n_draw < n_swaps && a >= b) {
        int8_t v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int k = b + L + 32 * r;
            v[r] = k < a ? cf[k + 1] : (k == a ? cf[0] : (int8_t)0);
        }
        const int8_t vb = cf[b];
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int k = b + L + 32 * r;
            if (k <= a) cf[k] = v[r];
        }
        if (L == 0) cf[0] = vb;
        __syncwarp();
    }
    int32_t* o = coef_out + i * d;
    for (int k = L; k < d; k += 32) o[k] = (int32_t)cf[k];
}

// ------------------------------------------------------------------------------------------ host
int hash_init_tables(fz_ctx* ctx) {
    FmtConst h[2];
    format_build_tables(h);
    (void)ctx;
    uint32_t magic[257];
    decoder_build_magic(magic);
    if (check(cudaMemcpyToSymbol(c_magic, magic, sizeof(magic)), "decoder constants upload")) return -2;
    return check(cudaMemcpyToSymbol(c_fmt, h, sizeof(h)), "format constants upload");
}

static int grid_for(size_t n, int per_block) { return (int)((n + per_block - 1) / per_block); }

int launch_sponge_bytes(const fz_ctx* ctx, const uint8_t* data, const unsigned long long* offs, size_t n, uint32_t prefix, int prefix_len,
                        uint8_t suffix, uint8_t* out, size_t out_stride, uint32_t out_len, cudaStream_t st) {
    if (n == 0) return 0;
    FZ_LAUNCH(sponge_bytes_kernel, grid_for(n, 128), 128, 0, st, data, offs, n, prefix, prefix_len, suffix, out, out_stride, out_len);
    (void)ctx;
    return check(cudaGetLastError(), "sponge_bytes launch");
}

int launch_format(const fz_ctx* ctx, int mode, const int32_t* vk_l, const int32_t* vk_r, const uint8_t* digests, const int32_t* c_hat, size_t n,
                  uint8_t* out, size_t out_stride, uint32_t* out_len, cudaStream_t st) {
    if (n == 0) return 0;
    const int smem = FMT_WARPS * FMT_SMEM_PER_WARP;
    if (attr_needed(ctx, ATTR_FORMAT) &&
        check(cudaFuncSetAttribute(format_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "format smem attr")) return -2;
    size_t blocks = (n + FMT_WARPS - 1) / FMT_WARPS;
    const size_t cap = (size_t)ctx->sm_count * 5;
    if (blocks > cap) blocks = cap;
    FZ_LAUNCH(format_kernel, (int)blocks, FMT_WARPS * 32, smem, st, mode, ctx->secpar == 256 ? 1 : 0, ctx->d, (uint8_t)(ctx->secpar == 256 ? 3 : 1), vk_l, vk_r,
                                                              digests, c_hat, n, out, out_stride, out_len);
    return check(cudaGetLastError(), "format launch");
}

int launch_shake_decode(const fz_ctx* ctx, const uint8_t* msg, size_t stride, const uint32_t* len, size_t n, int32_t* coef_out, cudaStream_t st) {
    if (n == 0) return 0;
    FZ_LAUNCH(shake_decode_kernel, grid_for(n, HC_THREADS), HC_THREADS, 0, st, msg, stride, len, n, ctx->d, ctx->omega_ch, 1u, ctx->bytes_coef, ctx->bytes_idx,
                                                                           (uint32_t)ctx->n_ch_used, coef_out);
    return check(cudaGetLastError(), "shake_decode launch");
}

int launch_decode(const fz_ctx* ctx, const uint8_t* bytes, size_t stride, uint32_t blen, size_t n, int weight, uint32_t bound, int bpc, int bpi,
                  int32_t* coef_out, cudaStream_t st) {
    if (n == 0) return 0;
    FZ_LAUNCH(decode_kernel, grid_for(n, HC_THREADS), HC_THREADS, 0, st, bytes, stride, blen, n, ctx->d, weight, bound, bpc, bpi, coef_out);
    return check(cudaGetLastError(), "decode launch");
}

int launch_ag_message(const fz_ctx* ctx, const uint8_t* pieces, size_t piece_stride, const uint32_t* piece_len, const uint32_t* order,
                      const unsigned long long* agg_off, size_t n_aggs, size_t n_total, unsigned long long* piece_pos, uint8_t* msgs, size_t msg_stride,
                      uint32_t* msg_len, cudaStream_t st) {
    if (n_aggs == 0) return 0;
    FZ_LAUNCH(ag_layout_kernel, (int)n_aggs, 256, 0, st, piece_len, order, agg_off, msg_stride, piece_pos, msg_len);
    if (check(cudaGetLastError(), "ag_layout launch")) return -2;
    FZ_LAUNCH(ag_pack_kernel, (int)(n_total + n_aggs), 128, 0, st, pieces, piece_stride, piece_len, order, agg_off, piece_pos, msg_stride,
                                                            (uint8_t)(ctx->secpar == 256 ? 3 : 1), msgs, msg_len, n_aggs, n_total);
    return check(cudaGetLastError(), "ag_pack launch");
}

int launch_shake_ag(const fz_ctx* ctx, const uint8_t* msgs, size_t msg_stride, const uint32_t* msg_len, const unsigned long long* agg_off, uint8_t* xof,
                    size_t n_aggs, cudaStream_t st) {
    if (n_aggs == 0) return 0;
    FZ_LAUNCH(shake_ag_kernel, grid_for(n_aggs, 32), 32, 0, st, msgs, msg_stride, msg_len, agg_off, (uint32_t)ctx->n_ag, xof, n_aggs);
    return check(cudaGetLastError(), "shake_ag launch");
}

size_t agw_max_piece() { return (size_t)FMT_SMEM_PER_WARP + 8; }
size_t agw_int_text_stride() { return INT_TEXT_STRIDE; }

int launch_agw_text(const fz_ctx* ctx, const int32_t* vk_l, const int32_t* vk_r, const uint8_t* digests, const int32_t* c_hat, const unsigned long long* win_off,
                    const uint32_t* agg_of, const unsigned long long* agg_off, size_t n_aggs, size_t n_win, uint32_t w0, unsigned long long* text_pos,
                    uint32_t* piece_len, unsigned long long* piece_off, uint8_t* int_text, uint32_t* int_len, unsigned long long* wend, uint8_t* ring,
                    size_t ring_cap, const void* state, cudaStream_t st) {
    if (n_aggs == 0) return 0;
    const int level = ctx->secpar == 256 ? 1 : 0;
    const uint8_t dst0 = (uint8_t)(ctx->secpar == 256 ? 3 : 1);
    if (n_win) {
        FZ_LAUNCH(agw_int_text_kernel, (unsigned)((n_win + 127) / 128), 128, 0, st, digests, n_win, int_text, int_len);
        if (check(cudaGetLastError(), "agw_int_text launch")) return -2;
        FZ_LAUNCH(agw_len_kernel, (unsigned)((n_win + 3) / 4), 128, 0, st, level, ctx->d, vk_l, vk_r, int_len, c_hat, win_off, agg_of, agg_off, n_win, w0, piece_len);
        if (check(cudaGetLastError(), "agw_len launch")) return -2;
    }
    FZ_LAUNCH(agw_layout_kernel, (unsigned)((n_aggs + 3) / 4), 128, 0, st, piece_len, win_off, agg_off, n_aggs, dst0, text_pos, piece_off, wend, ring, ring_cap, w0,
                                                                    state ? (const unsigned long long*)((const uint8_t*)state + ag_state_rpos_offset()) : nullptr,
                                                                    ag_state_bytes() / 8);
    if (check(cudaGetLastError(), "agw_layout launch")) return -2;
    if (n_win) {
        const int smem = FMT_WARPS * AGW_SMEM_PER_WARP;
        if (attr_needed(ctx, ATTR_AGW_FORMAT) &&
            check(cudaFuncSetAttribute(agw_format_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), "agw_format smem attr")) return -2;
        size_t blocks = (n_win + FMT_WARPS - 1) / FMT_WARPS;
        const size_t cap = (size_t)ctx->sm_count * 16;
        if (blocks > cap) blocks = cap;
        FZ_LAUNCH(agw_format_kernel, (unsigned)blocks, FMT_WARPS * 32, smem, st, level, ctx->d, dst0, vk_l, vk_r, int_text, int_len, c_hat, win_off, agg_of, agg_off, n_win, w0,
                                                                          piece_off, piece_len, ring, ring_cap);
        if (check(cudaGetLastError(), "agw_format launch")) return -2;
    }
    return 0;
}

int launch_agw_decode(const fz_ctx* ctx, const uint8_t* xof, size_t xof_cap, const uint32_t* bpos, const unsigned long long* win_off, const uint32_t* agg_of,
                      size_t n_win, int32_t* coef_out, cudaStream_t st) {
    if (n_win == 0) return 0;
    const int per_warp = ((ctx->n_ag + 32 + 15) & ~15) + 256 + 256;
    const int smem = 4 * per_warp;   // <= 18 KiB at either level: below the default dynamic shared memory limit
    FZ_LAUNCH(agw_decode_kernel, (unsigned)((n_win + 3) / 4), 128, smem, st, xof, xof_cap, bpos, win_off, agg_of, n_win, ctx->d, ctx->omega_ag, ctx->bytes_coef,
                                                                      ctx->bytes_idx, (uint32_t)ctx->n_ag, coef_out);
    return check(cudaGetLastError(), "agw_decode launch");
}

}  // namespace fz
