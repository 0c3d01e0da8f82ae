// fz_ntt.cu (kernels AND host code) compiled for the CPU test suite on the CUDA-on-CPU emulator.  Test infrastructure only.
#include "../cuda_emu.h"
#undef __constant__
#define __constant__ static      // __constant__ data is per translation unit on the device; keep it so here
#include "../../../fusion_cryptography_b200/csrc/fz_ntt.cu"
