// Host build of csrc/fz_dec_stream.cuh (the decoder of the one-shot aggregation-hash kernels) fed byte by byte as those kernels feed it.
// Test infrastructure only (tests/test_decoder_host.py); the library never sees this file.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((unsigned long long)a * b) >> 32); }
static inline int32_t __mulhi(int32_t a, int32_t b) { return (int32_t)(((long long)a * b) >> 32); }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t n) {
    n &= 31;
    return n ? (hi << n) | (lo >> (32 - n)) : hi;
}
template <typename T> static inline T __ldg(const T* p) { return *p; }
using std::min;
#undef __constant__
#define __constant__

#include "../../fusion_cryptography_b200/csrc/fz_dec_stream.cuh"

extern "C" {

// same contract as decoder_host.cpp's h_decode
void h_decode(const uint8_t* b, uint32_t blen, int d, int weight, uint32_t bound, int bpc, int bpi, int32_t* out) {
    static bool ready = false;
    if (!ready) {
        if (!fz::hashstream_build_tables(fz::c_magic2, fz::c_seg, &fz::g_text[0][0])) return;
        ready = true;
    }
    int32_t c[256];
    for (int k = 0; k < d; ++k) c[k] = 0;
    fz::Dec s;
    fz::dinit(s, c, d, weight, bound, bpc, bpi);
    for (uint32_t j = 0; j < blen && s.phase != 3; ++j) fz::dbyte(s, c, b[j]);
    while (s.phase != 3) fz::dfinish(s, c);     // the byte string is exhausted: the remaining items see empty / truncated slices (quirk Q4)
    for (int k = 0; k < d; ++k) out[k] = c[k];
}

}  // extern "C"
