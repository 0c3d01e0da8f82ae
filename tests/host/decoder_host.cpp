// Host build of csrc/fz_decoder.cuh with decode_kernel's per-thread body around it.  Test infrastructure only (tests/test_decoder_host.py);
// the library never sees this file.
#include <cstddef>
#include <cstdint>

#define __device__
#define __forceinline__ inline
#define __constant__
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((unsigned long long)a * b) >> 32); }
#include "../../fusion_cryptography_b200/csrc/fz_decoder.cuh"

extern "C" {

// one polynomial from blen caller-supplied bytes, as one thread of decode_kernel (fz_hash.cu) does it; d <= 256
void h_decode(const uint8_t* b, uint32_t blen, int d, int weight, uint32_t bound, int bpc, int bpi, int32_t* out) {
    static bool ready = false;
    if (!ready) { fz::decoder_build_magic(fz::c_magic); ready = true; }
    int32_t c[256];
    for (int k = 0; k < d; ++k) c[k] = 0;
    fz::Decoder s;
    fz::dec_init(s, c, d, weight, bound, bpc, bpi);
    for (uint32_t j = 0; j < blen && s.phase != 3; ++j) fz::dec_byte(s, c, b[j]);
    fz::dec_drain(s, c);
    for (int k = 0; k < d; ++k) out[k] = c[k];
}

}  // extern "C"
