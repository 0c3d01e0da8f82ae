// fz_mac.cuh -- the multiply-accumulate step of the streaming NTT-domain kernels (fz_stream.cu, fz_scheme.cu): four coefficients per thread, products in
// signed 64-bit accumulators, one pseudo-Mersenne fold after every SECOND product (two products of a canonical matrix-side operand, |a| < 2^30, and
// any int32 are below 2^62 together, the folded rest below 2^47: no overflow), one centring per output.  The reference re-centres after every
// multiply and add (algebra/matrices.py:108-131, 74-91), so the result is cent() of the exact sum.  Checked on the host at the extreme operands
// (tests/test_mac_host.py).
#pragma once
#include "fz_common.cuh"

namespace fz {

__device__ __forceinline__ void mac4(long long (&acc)[4], const int4& a, const int4& b) {
    acc[0] += (long long)a.x * b.x;
    acc[1] += (long long)a.y * b.y;
    acc[2] += (long long)a.z * b.z;
    acc[3] += (long long)a.w * b.w;
}
__device__ __forceinline__ void fold4(long long (&acc)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = fold_i64(acc[i]);
}
__device__ __forceinline__ int4 cent4(const long long (&acc)[4]) {
    return make_int4(cent_i64(acc[0]), cent_i64(acc[1]), cent_i64(acc[2]), cent_i64(acc[3]));
}

}  // namespace fz
