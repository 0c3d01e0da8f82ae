// The dynamic shared memory of the library's kernels (FZ_DYN_SHARED arrays) for the CUDA-on-CPU emulator: static storage, one CTA runs at a time.
// Test infrastructure only.
#include <cstdint>

namespace fz {
alignas(16) uint8_t fsm[256 * 1024];                 // fz_hash.cu: format_kernel, agw_format_kernel
alignas(16) uint8_t dsm[64 * 1024];                  // fz_hash.cu: agw_decode_kernel
alignas(16) uint32_t ring_smem[64 * 1024];           // fz_ring.cu: ring_ntt_kernel
alignas(16) uint32_t smp_sm[64 * 1024];              // fz_sample.cu: sample_kernel, sample_dense_kernel
alignas(16) unsigned long long skeys[32 * 1024];     // fz_scheme.cu: sort_kernel
namespace warpag { alignas(16) uint8_t wsm[256 * 1024]; }   // fz_hashstream.cu: hash_ag_*_warp_kernel
}  // namespace fz
