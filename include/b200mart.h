/* b200mart — C ABI of the B200-native MarT (MKGformer) training hot path.
 *
 * This is the drop-in boundary of SURVEY.md §8(b): plain C, raw device pointers and sizes, no torch
 * types.  The reference (zjunlp/MKG_Analogy) has no native code on this path; what these entry
 * points replace is the torch-op graph that the reference's model layer issues, i.e. the body of
 *   MarT/models/modeling_unimo.py:848-893   UnimoForMaskedLM.forward           (model forward)
 *   MarT/lit_models/transformer.py:59-113   TransformerLitModel.training_step  (score head + loss)
 *   MarT/lit_models/transformer.py:115-166  TransformerLitModel._eval          (ranks)
 *   MarT/lit_models/utils.py:42-66          LabelSmoothSoftmaxCEV1             (loss + its backward)
 *   MarT/lit_models/transformer.py:224-241  configure_optimizers (AdamW)       (parameter update)
 * and torch autograd's backward of the same graph.  The Python side binds them with ctypes
 * (mkg_analogy_b200/_lib.py) — the same convention the reference itself uses for its only native
 * artefact (M-KGE/IKRL_TransAE/DATA_/TrainDataLoader.py:41-55, ctypes.cdll.LoadLibrary on Base.so).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; b200mart_last_error() gives the
 *     thread-local message.  Nothing throws, aborts or synchronises the stream (unless stated).
 *   - all device buffers are allocated and owned by the caller (PyTorch); the library borrows them.
 *   - `stream` is a cudaStream_t passed as void*; work is stream-ordered.
 *   - one handle per device and process; calls on one handle must not overlap from two threads.
 */
#ifndef B200MART_H
#define B200MART_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MART_VERSION 1

#define B200MART_OK 0
#define B200MART_ERR_ARG 1
#define B200MART_ERR_UNSUPPORTED 2
#define B200MART_ERR_CUDA 3

const char* b200mart_last_error(void);
int32_t b200mart_version(void);

/* ------------------------------------------------------------------------------------------------
 * Stand-alone tensor-core product (tcgen05 + TMA), used by the model functions below and exposed
 * for kernel-level parity tests.  D[M,N] = epilogue( A[M,K] * B[N,K]^T ), bf16 in, fp32 accumulate.
 * Replaces torch.nn.functional.linear / its autograd products as called from
 * modeling_unimo.py:221-226,:283-287 (CLIP), :327-334,:388,:458-462,:475 (BERT), :972-976,:958 (head).
 *   a_mn/b_mn: 0 = operand stored [rows][K] (K contiguous); 1 = stored [K][rows].
 *   (A2,B2,K2): optional second operand pair extending the contraction (K must be a multiple of 64).
 *   epilogue order: *alpha, +bias[n], store out_preact, act (1 gelu-erf, 2 quick-gelu),
 *                   *act'(dact_src) (dact 1/2), +residual, store out_f32 (or atomically accumulate) / out_bf16.
 */
typedef struct b200mart_gemm_desc {
    int32_t M, N, K, K2;
    int32_t a_mn, b_mn;
    const void* A;  int64_t lda;   /* bf16 */
    const void* B;  int64_t ldb;   /* bf16 */
    const void* A2; int64_t lda2;
    const void* B2; int64_t ldb2;
    float alpha;
    const float* bias;
    void* out_preact; int64_t ldp; /* bf16 */
    int32_t act;
    const void* dact_src; int64_t ldx; int32_t dact;
    const float* residual; int64_t ldr;
    float* out_f32; int64_t ldo32;
    void* out_bf16; int64_t ldo16;
    int32_t accumulate;
    int32_t split_k;
    int32_t bn;                    /* N tile: 0 (auto), 64, 128, 256 */
} b200mart_gemm_desc;

int32_t b200mart_gemm_bf16(const b200mart_gemm_desc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200MART_H */
