// Host build of csrc/fz_mac.cuh (+ the folds of fz_common.cuh) with the accumulation cadences of matvec_kernel / sign_mac (mac, mac, fold) and of
// aggregate_kernel (two products, add, fold) restated around them, for one thread's four coefficients.  Test infrastructure only
// (tests/test_mac_host.py); the library never sees this file.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((unsigned long long)a * b) >> 32); }
static inline int32_t __mulhi(int32_t a, int32_t b) { return (int32_t)(((long long)a * b) >> 32); }
using std::min;

#include "../../fusion_cryptography_b200/csrc/fz_mac.cuh"

extern "C" {

// matvec_kernel's loop for one thread: out[k] = cent(sum_r A[r][k] * s[r][k]), k < 4; A canonical, s any int32
void h_matvec4(const int32_t* A, const int32_t* s, int ell, int32_t* out) {
    const int4* ap = reinterpret_cast<const int4*>(A);
    const int4* sp = reinterpret_cast<const int4*>(s);
    long long acc[4] = {0, 0, 0, 0};
    int r = 0;
    for (; r + 4 <= ell; r += 4) {
        fz::mac4(acc, ap[r + 0], sp[r + 0]); fz::mac4(acc, ap[r + 1], sp[r + 1]); fz::fold4(acc);
        fz::mac4(acc, ap[r + 2], sp[r + 2]); fz::mac4(acc, ap[r + 3], sp[r + 3]); fz::fold4(acc);
    }
    for (; r < ell; ++r) { fz::mac4(acc, ap[r], sp[r]); fz::fold4(acc); }
    const int4 o = fz::cent4(acc);
    out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = o.w;
}

// aggregate_kernel's loop for one thread and one coefficient: out = cent(sum_p z[p] * a[p]); a canonical, z any int32
int32_t h_aggregate1(const int32_t* z, const int32_t* a, int n) {
    long long acc = 0;
    int p = 0;
    for (; p + 4 <= n; p += 4) {
        acc += (long long)z[p] * a[p] + (long long)z[p + 1] * a[p + 1];
        acc = fz::fold_i64(acc);
        acc += (long long)z[p + 2] * a[p + 2] + (long long)z[p + 3] * a[p + 3];
        acc = fz::fold_i64(acc);
    }
    for (; p < n; ++p) acc = fz::fold_i64(acc + (long long)z[p] * a[p]);
    return fz::cent_i64(acc);
}

}  // extern "C"
