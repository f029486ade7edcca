// Host-side description of one tcgen05 GEMM launch (see gemm_tcgen05.cu).
#pragma once
#include "common.cuh"

// D[M,N] = epilogue( sum_k A[m,k] * B[n,k] ).  Both operands are bf16 in HBM; accumulation is fp32
// in TMEM.  "mn" flags give the storage order of each operand:
//   a_mn = 0 : A stored [M][K] (K contiguous, e.g. activations X for Y = X W^T)
//   a_mn = 1 : A stored [K][M] (M contiguous, e.g. dY^T for dW = dY^T X)
//   b_mn = 0 : B stored [N][K] (nn.Linear weight layout for the forward product)
//   b_mn = 1 : B stored [K][N] (nn.Linear weight for dX = dY W, activations X for dW)
// An optional second operand pair (A2,B2) extends the contraction (K = K1 + K2): this is how
// gelu(X Wi^T + G Wf^T) (modeling_unimo.py:458-464) runs as ONE GEMM.
struct GemmOp {
    int M = 0, N = 0, K1 = 0, K2 = 0;
    int a_mn = 0, b_mn = 0;
    const bf16* A = nullptr; long lda = 0;
    const bf16* B = nullptr; long ldb = 0;
    const bf16* A2 = nullptr; long lda2 = 0;
    const bf16* B2 = nullptr; long ldb2 = 0;
    // epilogue, applied in this order
    float alpha = 1.0f;
    const float* bias = nullptr;          // + bias[n]
    const float* bias2 = nullptr;         // + bias2[n] (second operand pair's bias)
    bf16* out_preact = nullptr; long ldp = 0;   // store value before activation (saved for backward)
    int act = 0;                          // 0 none, 1 gelu(erf), 2 quick_gelu
    int save_grad = 0;                    // out_preact receives act'(pre-activation) instead of the pre-activation
    const bf16* dact_src = nullptr; long ldx = 0; int dact = 0;  // *= act'(dact_src[m,n]) (backward); dact = 3: dact_src already holds act'
    DropCfg drop = DropCfg{0u, 0u, 1.f};  // dropout applied to the value before the residual add (element index m*N+n)
    const float* residual = nullptr; long ldr = 0;   // + residual[m,n] (fp32)
    float* out_f32 = nullptr; long ldo32 = 0;
    bf16* out_bf16 = nullptr; long ldo16 = 0;
    float* colsum_out = nullptr;          // with dact = 3 and a bf16 output: += column sums of the output (fp32, atomic)
    float* colsum_out2 = nullptr;
    int accumulate = 0;                   // out_f32 += value (atomic; required when split_k > 1)
    int split_k = 1;
    int bn = 0;                           // N tile (0 = auto: 128)
    int max_ctas = 0;                     // bound of the persistent grid (0 = one CTA per SM): leaves SMs to concurrent streams
    int b_static = 0;                     // B (and B2) is not written by the preceding kernel of the stream (model weights): its first
                                          // tiles are requested BEFORE griddepcontrol.wait, under the predecessor's tail
};

struct GemmPlan;  // opaque: encoded tensor maps + launch geometry

// Encode tensor maps for `op`; the returned plan can be launched any number of times while the
// operand pointers stay valid.
int gemm_plan_create(const GemmOp& op, GemmPlan** out);
void gemm_plan_destroy(GemmPlan* p);
// drop: optional override of the plan's dropout configuration (same on/off state, fresh seed)
int gemm_plan_launch(const GemmPlan* p, cudaStream_t stream, const DropCfg* drop = nullptr);
// flops of one launch (2*M*N*K)
double gemm_plan_flops(const GemmPlan* p);
