// fz_repr_order.cuh -- Python's str() order of verification keys without the text (fusion.py:661-663, 693: sorted(..., key=lambda x: str(x[0]))).
// Two reprs of the same level share every byte up to the first coefficient that differs, so the order of two keys is the order of that
// coefficient's decimal text followed by its terminator (',' inside a value list, ']' after the last value).  fz_scheme.cu's sort_kernel ranks with
// these two functions; the CPU suite compiles them for the host and compares the order with Python's sorted() on the real repr strings
// (tests/test_repr_order_host.py).
#pragma once
#include <cstdint>

namespace fz {

// 4-bit codes preserve ASCII order of the alphabet { ',' < '-' < '0'..'9' < ']' }; 0 pads.
__device__ __forceinline__ unsigned long long repr_key(int32_t v, bool last) {
    uint32_t u = v < 0 ? (uint32_t)(-(long long)v) : (uint32_t)v;
    uint32_t digs[10];
    int n = 0;
    do { digs[n++] = u % 10u; u /= 10u; } while (u);
    unsigned long long key = 0;
    int pos = 11;
    if (v < 0) { key |= 2ULL << (4 * pos); --pos; }
    for (int k = n - 1; k >= 0; --k) { key |= (unsigned long long)(3u + digs[k]) << (4 * pos); --pos; }
    key |= (last ? 13ULL : 1ULL) << (4 * pos);
    return key;
}

// full comparison of str(vk_a) vs str(vk_b): -1, 0, +1
__device__ int repr_cmp(const int32_t* al, const int32_t* ar, const int32_t* bl, const int32_t* br, int d) {
    for (int k = 0; k < d; ++k)
        if (al[k] != bl[k]) {
            const unsigned long long ka = repr_key(al[k], k == d - 1), kb = repr_key(bl[k], k == d - 1);
            return ka < kb ? -1 : 1;
        }
    for (int k = 0; k < d; ++k)
        if (ar[k] != br[k]) {
            const unsigned long long ka = repr_key(ar[k], k == d - 1), kb = repr_key(br[k], k == d - 1);
            return ka < kb ? -1 : 1;
        }
    return 0;
}

}  // namespace fz
