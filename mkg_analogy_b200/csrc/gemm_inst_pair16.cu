// Instantiations of the CTA-pair tcgen05 GEMM with 16 epilogue warps on 32-column passes (PairWide16 in gemm_kernel.cuh).
#include "gemm_kernel.cuh"

gemmk::GemmKernel gemm_pick_kernel_pair16(int a_mn, int b_mn, int epi) { return gemmk::pick_kernel_pair16(a_mn, b_mn, epi); }
