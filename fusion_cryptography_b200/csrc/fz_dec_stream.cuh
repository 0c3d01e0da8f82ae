// fz_dec_stream.cuh -- the byte-driven sparse ternary decoder (fusion.py:422-481) of the one-shot aggregation-hash kernels in fz_hashstream.cu: the same
// state machine as fz_decoder.cuh, reading the reciprocal table of ITS translation unit (c_magic2, fz_hashch.cuh -- __constant__ data is per
// translation unit).  Replayed on the host with the same cases as fz_decoder.cuh (tests/test_decoder_host.py).
#pragma once
#include "fz_hashch.cuh"

namespace fz {

struct Dec {
    int d, weight, bpc, bpi;
    uint32_t bound;
    int phase, item, left;
    unsigned long long acc;
    uint32_t acc32, m, magic;
    unsigned long long sg[4];
};
__device__ __forceinline__ int dsign(const Dec& s, int i) { return 2 * (int)((s.sg[(i >> 6) & 3] >> (i & 63)) & 1ULL) - 1; }
__device__ __forceinline__ void dset(Dec& s, int i) {
    s.item = i; s.left = s.bpi; s.acc32 = 0; s.m = (uint32_t)(i + 1); s.magic = c_magic2[s.m <= 256 ? s.m : 0];
}
__device__ __forceinline__ void dfinish(Dec& s, int32_t* c) {
    if (s.phase == 0) {
        s.phase = 1; s.item = 0; s.left = s.bpc; s.acc = 0;
        if (s.weight == 0) { s.phase = 2; dset(s, s.d - 1); }
    } else if (s.phase == 1) {
        const int i = s.item;
        const long long mag = (long long)(s.bound == 1 ? 0ULL : s.acc % s.bound) + 1;
        if (i < s.d) c[i] = (int32_t)(mag * dsign(s, i));
        s.item++; s.left = s.bpc; s.acc = 0;
        if (s.item >= s.weight) { s.phase = 2; dset(s, s.d - 1); }
    } else if (s.phase == 2) {
        const int i = s.item, j = (int)s.acc32;
        const int32_t t = c[i]; c[i] = c[j]; c[j] = t;
        dset(s, i - 1);
    }
    if (s.phase == 2 && (s.item <= s.weight || s.weight >= s.d)) s.phase = 3;
}
__device__ __forceinline__ void dinit(Dec& s, int32_t* c, int d, int weight, uint32_t bound, int bpc, int bpi) {
    s.d = d; s.weight = weight; s.bound = bound; s.bpc = bpc; s.bpi = bpi;
    s.phase = 0; s.item = 0; s.left = (weight + 7) / 8; s.acc = 0; s.acc32 = 0; s.m = 1; s.magic = 0;
    s.sg[0] = s.sg[1] = s.sg[2] = s.sg[3] = 0;
    while (s.left == 0 && s.phase != 3) dfinish(s, c);
}
__device__ __forceinline__ void dbyte(Dec& s, int32_t* c, uint32_t b) {
    if (s.phase == 2) {
        const uint32_t v = (s.acc32 << 8) | b;
        s.acc32 = v - __umulhi(v, s.magic) * s.m;
    } else if (s.phase == 1) {
        if (s.bound != 1) s.acc = ((s.acc << 8) | b) % s.bound;
    } else {
        s.sg[3] = (s.sg[3] << 8) | (s.sg[2] >> 56);
        s.sg[2] = (s.sg[2] << 8) | (s.sg[1] >> 56);
        s.sg[1] = (s.sg[1] << 8) | (s.sg[0] >> 56);
        s.sg[0] = (s.sg[0] << 8) | b;
    }
    if (--s.left == 0) dfinish(s, c);
}

}  // namespace fz
