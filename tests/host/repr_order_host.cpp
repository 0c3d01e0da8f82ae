// Host build of csrc/fz_repr_order.cuh (the str(vk) order of keys without the text) with sort_kernel's rank loop restated around it.
// Test infrastructure only (tests/test_repr_order_host.py); the library never sees this file.
#include <cstddef>
#include <cstdint>
#include <vector>

#define __device__
#define __forceinline__ inline
#include "../../fusion_cryptography_b200/csrc/fz_repr_order.cuh"

extern "C" {

int h_repr_cmp(const int32_t* al, const int32_t* ar, const int32_t* bl, const int32_t* br, int d) { return fz::repr_cmp(al, ar, bl, br, d); }

// sort_kernel for one aggregate of n keys: perm[rank] = original position, stable
void h_rank_sort(const int32_t* vk_l, const int32_t* vk_r, int n, int d, uint32_t* perm) {
    std::vector<unsigned long long> skeys(n);
    for (int t = 0; t < n; ++t) skeys[t] = fz::repr_key(vk_l[(size_t)t * d], d == 1);
    for (int i = 0; i < n; ++i) {
        const unsigned long long mine = skeys[i];
        int rank = 0;
        for (int t = 0; t < n; ++t) {
            const unsigned long long other = skeys[t];
            if (other < mine) ++rank;
            else if (other == mine && t != i) {
                const int c = fz::repr_cmp(vk_l + (size_t)t * d, vk_r + (size_t)t * d, vk_l + (size_t)i * d, vk_r + (size_t)i * d, d);
                if (c < 0 || (c == 0 && t < i)) ++rank;
            }
        }
        perm[rank] = (uint32_t)i;
    }
}

}  // extern "C"
