// Instantiations of the CTA-pair (cta_group::2, 256 x 256 tile) tcgen05 GEMM.
#include "gemm_kernel.cuh"

gemmk::GemmKernel gemm_pick_kernel_pair(int a_mn, int b_mn, int epi) { return gemmk::pick_kernel_pair(a_mn, b_mn, epi); }
