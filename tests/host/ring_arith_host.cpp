// Host build of csrc/fz_ring_arith.cuh (the general-ring arithmetic: any odd prime modulus below 2^31).  Test infrastructure only
// (tests/test_ring_arith_host.py); the library never sees this file.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((unsigned long long)a * b) >> 32); }
static inline int32_t __mulhi(int32_t a, int32_t b) { return (int32_t)(((long long)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
using std::min;

#include "../../fusion_cryptography_b200/csrc/fz_ring_arith.cuh"

namespace {
fz::RingK ring(uint32_t q) {          // the constants fz_ring.cu's ring_k() derives from the modulus
    fz::RingK k;
    k.q = q; k.half = q / 2; k.logd = 0; k.ninv.w = 0; k.ninv.wq = 0;
    k.mu = (unsigned long long)((((unsigned __int128)1) << 64) / q);
    k.mu32 = (uint32_t)((1ULL << 32) / q);
    k.c31 = (uint32_t)((1ULL << 31) % q);
    return k;
}
}  // namespace

extern "C" {

// op 0: rq_from(a) -> [0, q)   1: cent(a * b)   2: cent(a + b)   3: rq_mul(a, Shoup pair of w = b) for a any 32-bit value   4: rq_cent(a) for a in [0, q)
void h_ring_op(int op, uint32_t q, const int32_t* a, const int32_t* b, int64_t* out, size_t n) {
    const fz::RingK k = ring(q);
    for (size_t i = 0; i < n; ++i) {
        switch (op) {
            case 0: out[i] = fz::rq_from(a[i], k); break;
            case 1: out[i] = fz::ring_pw1<true>(k, a[i], b[i]); break;
            case 2: out[i] = fz::ring_pw1<false>(k, a[i], b[i]); break;
            case 3: {
                fz::Sh w;
                w.w = (uint32_t)b[i] % q;
                w.wq = (uint32_t)(((unsigned long long)w.w << 32) / q);     // ring_table_kernel
                out[i] = fz::rq_mul((uint32_t)a[i], w, q);
                break;
            }
            default: out[i] = fz::rq_cent((uint32_t)a[i], q, k.half);
        }
    }
}

}  // extern "C"
