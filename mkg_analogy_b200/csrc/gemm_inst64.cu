// Instantiations of the persistent tcgen05 GEMM for the 64-column N tile (one translation unit per tile width so
// that nvcc compiles them in parallel).
#include "gemm_kernel.cuh"

gemmk::GemmKernel gemm_pick_kernel_64(int a_mn, int b_mn, int epi) { return gemmk::pick_kernel_bn<64>(a_mn, b_mn, epi); }
