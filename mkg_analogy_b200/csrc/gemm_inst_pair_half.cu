// Instantiations of the half-footprint CTA-pair tcgen05 GEMM (two CTAs per SM: 3 stages, one accumulator buffer,
// 4 epilogue warps, 32-column epilogue passes; PairHalf in gemm_kernel.cuh).
#include "gemm_kernel.cuh"

gemmk::GemmKernel gemm_pick_kernel_pair_half(int a_mn, int b_mn, int epi) { return gemmk::pick_kernel_pair_half(a_mn, b_mn, epi); }
