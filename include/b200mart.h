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
/* kernels launched by this library in this process so far (bench.py reports the per-step difference) */
int64_t b200mart_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Stand-alone tensor-core product (tcgen05 + TMA), used by the model functions below and exposed
 * for kernel-level parity tests.  D[M,N] = epilogue( A[M,K] * B[N,K]^T ), bf16 in, fp32 accumulate.
 * Replaces torch.nn.functional.linear / its autograd products as called from
 * modeling_unimo.py:221-226,:283-287 (CLIP), :327-334,:388,:458-462,:475 (BERT), :972-976,:958 (head).
 *   a_mn/b_mn: 0 = operand stored [rows][K] (K contiguous); 1 = stored [K][rows].
 *   (A2,B2,K2): optional second operand pair extending the contraction (K must be a multiple of 64).
 *   epilogue order: *alpha, +bias[n] (+bias2[n]), store out_preact, act (1 gelu-erf, 2 quick-gelu),
 *                   *act'(dact_src) (dact 1/2) or *dact_src (dact 3), +residual, store out_f32 (or atomically accumulate) / out_bf16.
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
    const float* bias2;            /* optional second bias (bias of the second operand pair) */
    void* out_preact; int64_t ldp; /* bf16 */
    int32_t act;
    int32_t save_grad;             /* out_preact receives act'(pre-activation) (bf16) instead of the pre-activation */
    const void* dact_src; int64_t ldx; int32_t dact;   /* dact 1/2: *= act'(dact_src); 3: *= dact_src (saved derivative) */
    const float* residual; int64_t ldr;
    float* out_f32; int64_t ldo32;
    void* out_bf16; int64_t ldo16;
    int32_t accumulate;
    int32_t split_k;
    int32_t bn;                    /* N tile: 0 (auto), 64, 128, 256 */
} b200mart_gemm_desc;

int32_t b200mart_gemm_bf16(const b200mart_gemm_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MKGformer model handle.  Shapes follow the reference's two HF configs (BertConfig for the text
 * stream, CLIPVisionConfig for the vision stream; MarT/main.py:76-86); fuse_from is the layer index 8
 * hard-coded at MarT/models/modeling_unimo.py:616,627.
 */
typedef struct b200mart_model_desc {
    int32_t hidden, layers, heads, intermediate;
    int32_t vocab, max_position, type_vocab;
    int32_t image_size, patch;
    int32_t fuse_from;
    float text_eps, vision_eps;
} b200mart_model_desc;

typedef struct b200mart_handle b200mart_handle;

int32_t b200mart_create(const b200mart_model_desc* desc, b200mart_handle** out);
void b200mart_destroy(b200mart_handle* h);

/* Parameter storage.  All parameters of the reference module (state_dict contract of SURVEY.md §8a,
 * modeling_unimo.py:100-186,:189-287,:290-478,:661-673,:942-976) live in ONE flat fp32 buffer owned by the
 * caller (so that torch nn.Parameters can be views of it), mirrored by a flat bf16 buffer of the same
 * layout (tensor-core operands) and a flat fp32 gradient buffer.  The library fixes the layout
 * (q/k/v projection weights and biases of a layer are adjacent so that they form one [3H, H] operand). */
int32_t b200mart_param_count(const b200mart_handle* h);
int64_t b200mart_param_total(const b200mart_handle* h);              /* elements of the flat buffers */
int32_t b200mart_param_info(const b200mart_handle* h, int32_t index, const char** name, int64_t* offset,
                            int32_t* ndim, int64_t shape[4], int32_t* decay, int32_t* unused_in_finetune);
/* Flat extent [begin, end) holding every parameter of layer index `layer` (vision layer + text layer): the unit of the
 * overlapped data-parallel gradient exchange (see layer_events of b200mart_backward). */
int32_t b200mart_layer_extent(const b200mart_handle* h, int32_t layer, int64_t* begin, int64_t* end);
/* Runtime options: "streams" (default 1) runs the vision and text chains on two internal CUDA streams;
 * "training" (default 0) enables dropout in b200mart_forward (nn.Module.train()/eval() of the reference). */
int32_t b200mart_set_option(b200mart_handle* h, const char* name, int32_t value);
/* Dropout of the BERT stream (modeling_unimo.py:185 embeddings, :362 attention probabilities, :389 / :476 sub-layer
 * outputs; the CLIP stream has none): hidden_dropout_prob, attention_probs_dropout_prob (reference: 0.1 / 0.1) and a seed.
 * Masks are a stateless function of (seed, forward-call counter, site, element index): backward regenerates them. */
int32_t b200mart_set_dropout(b200mart_handle* h, float p_hidden, float p_attn, uint32_t seed);
/* per-forward seed of the most recent b200mart_forward (tests rebuild the masks from it) */
uint32_t b200mart_last_dropout_seed(const b200mart_handle* h);
/* params16[i] = bf16(params32[i]) over the whole flat buffer */
int32_t b200mart_cast_params(const b200mart_handle* h, const float* params32, void* params16, void* stream);
/* the same over the flat range [begin, end): refreshes the mirror of parameters another rank's optimizer shard just delivered
 * (reduce-scatter -> AdamW on the local shard -> all-gather of the fp32 masters) */
int32_t b200mart_cast_params_range(const b200mart_handle* h, const float* params32, void* params16, int64_t begin, int64_t end,
                                   void* stream);

/* Workspace: activations saved for backward + scratch, for a batch of B samples x L text tokens. */
int64_t b200mart_workspace_bytes(const b200mart_handle* h, int32_t B, int32_t L);

typedef struct b200mart_buffers {
    const float* params32;     /* flat fp32 parameters (biases, LayerNorm, embeddings are read from here) */
    const void* params16;      /* flat bf16 mirror (GEMM operands) */
    float* grads32;            /* flat fp32 gradients, ACCUMULATED into (caller zeroes); may be NULL for inference */
    void* workspace;
    int64_t workspace_bytes;
} b200mart_buffers;

/* One batch, as produced by the reference collator (MarT/data/data_module.py:48-182) and passed as
 * `self.model(**batch)` at MarT/lit_models/transformer.py:68.  Device pointers. */
typedef struct b200mart_batch {
    int32_t B, L;
    const int64_t* input_ids;       /* (B, L) */
    const int64_t* attention_mask;  /* (B, L) */
    const int64_t* token_type_ids;  /* (B, L) or NULL (zeros) */
    const float* pixel_values;      /* (B, 2, 3, S, S) fp32 */
    const int64_t* sep_idx;         /* (B, 6) or NULL: no adaptive scaling (pre-training) */
} b200mart_batch;

/* UnimoForMaskedLM.forward (modeling_unimo.py:848-893); dropout only with the "training" option (see above).
 *   z_out      (B, L, hidden) fp32: trans_hidden_states = model_output[1]              (may be NULL)
 *   logits_out (B, L, vocab)  fp32: model_output[0].logits, the dense "compat" head    (may be NULL)
 * Activations needed by b200mart_backward stay in the workspace until the next forward. */
int32_t b200mart_forward(b200mart_handle* h, const b200mart_buffers* bufs, const b200mart_batch* batch,
                         float* z_out, float* logits_out, void* stream);

/* Score head + loss on the [MASK] rows only (TransformerLitModel.training_step, transformer.py:59-113).
 *   mode 0 (fine-tune, :92-109): logits over `col_ids` (n_cols analogy entity ids), loss =
 *       LS-CE(label) + alpha * relaxation(rel_idx, q_head_idx, a_head_idx).
 *   mode 1 (pre-train, :72-90): rows with pre_type != 2 are scored over [ent_st, ent_ed), rows with
 *       pre_type == 2 over [rel_st, rel_ed); loss = CE_entity + CE_relation.
 * label_smoothing = 0 selects plain cross-entropy (transformer.py:22-25).
 *   loss_out        device float[1]
 *   mask_logits_out fp32, mode 0: (B, n_cols); mode 1: (B, n_ent + n_rel) = [entity | relation] scores  (may be NULL)
 *   ranks_out       int64, mode 0: (B); mode 1: (B) rank within the row's own group (_eval, :115-166) (may be NULL)
 * With compute_grad != 0 the gradient wrt trans_hidden_states is left in the workspace for
 * b200mart_backward and the tied-embedding / decoder-bias gradients are accumulated into grads32. */
typedef struct b200mart_head {
    int32_t mode;
    const int64_t* col_ids; int32_t n_cols;
    int32_t ent_st, ent_ed, rel_st, rel_ed;
    const int64_t* label;
    const int64_t* pre_type;
    const int64_t* rel_idx;      /* (B, 2) */
    const int64_t* q_head_idx;   /* (B) */
    const int64_t* a_head_idx;   /* (B) */
    float label_smoothing, alpha;
    int32_t mask_token_id;
    const float* grad_scale;     /* device scalar multiplying every gradient this call seeds (NULL = 1): the incoming
                                  * d(loss) of `loss.backward()` under loss scaling / gradient accumulation, read on the
                                  * device so that no host synchronisation is needed; the returned loss is unscaled */
} b200mart_head;

int32_t b200mart_head_loss(b200mart_handle* h, const b200mart_buffers* bufs, const b200mart_batch* batch,
                           const b200mart_head* head, float* loss_out, float* mask_logits_out, int64_t* ranks_out,
                           int32_t compute_grad, void* stream);

/* Backward of b200mart_forward (what torch autograd does for the reference graph).  The gradient wrt
 * trans_hidden_states is the one left by b200mart_head_loss, plus dz_extra (B, L, hidden) fp32 when given
 * (use_head_grad = 0: dz_extra alone).  dlogits (B, L, vocab) fp32, when given, is the gradient of the dense
 * compat logits; dlogits_scratch must then hold B*L*roundup(vocab, 8) bf16.  Accumulates into grads32.
 * layer_events: NULL, or 2*layers cudaEvent_t handles; events [2i] and [2i+1] are recorded (on the library's internal
 * streams) when layer i's parameter gradients are final, so that the caller can start that layer's gradient all-reduce
 * on another stream while the layers below are still running (what DDP's bucket hooks do for the reference). */
int32_t b200mart_backward(b200mart_handle* h, const b200mart_buffers* bufs, const b200mart_batch* batch,
                          int32_t use_head_grad, const float* dz_extra, const float* dlogits, void* dlogits_scratch,
                          void* const* layer_events, void* stream);

/* torch.optim.AdamW step as configured at transformer.py:224-241 over the flat buffers, refreshing the bf16
 * mirror.  `step` is 1-based.  Parameters the reference never gives a gradient (text_pooler, vision_post_layernorm,
 * fusion_dense of text layers < fuse_from; in pre-training also adaptive_weight) are left untouched, as torch skips
 * parameters whose .grad is None.  [begin, end) restricts the update to a flat range (end <= 0: whole buffer) so that a
 * layer can be updated as soon as its gradients are final (b200mart_layer_extent + layer_events); zero_grads clears the
 * gradient range after use (the next step accumulates into it). */
int32_t b200mart_adamw(b200mart_handle* h, float* params32, void* params16, float* grads32, float* exp_avg,
                       float* exp_avg_sq, int32_t step, float lr, float beta1, float beta2, float eps,
                       float weight_decay, int32_t pretrain, int64_t begin, int64_t end, int32_t zero_grads, void* stream);

/* Per-step scalars in DEVICE memory, for CUDA-graph replay: a captured graph freezes kernel arguments, so what
 * changes from step to step must be read from memory.  `device_block` = 16 bytes {uint32 salt, float lr, float 1 - beta1^t,
 * float 1 - beta2^t}, written by the caller (stream-ordered) before each step.  While a block is set, every dropout mask of
 * b200mart_forward / b200mart_backward is drawn from site_seed ^ salt (site_seed = the one of the forward that was captured) and
 * b200mart_adamw takes lr and the bias corrections from the block instead of its by-value arguments; the same code path runs eagerly
 * and under capture.  NULL restores by-value behaviour.  The reference has no counterpart (PyTorch eager re-issues every launch;
 * MarT/lit_models/base.py:86-91 builds a stock torch.optim.AdamW). */
int32_t b200mart_set_step_block(b200mart_handle* h, void* device_block);

/* The optimizer step with the data-parallel exchange fused in, over NVSwitch multicast (one kernel = reduce-scatter + AdamW +
 * all-gather): for this rank's shard [begin, end) the mean gradient is pulled with multimem.ld_reduce from grads_mc (multicast
 * address of the symmetric gradient buffers, every rank's local gradients complete), AdamW runs on the rank's own masters /
 * moments, and the new fp32 parameters and their bf16 mirror are stored through multimem.st to params32_mc / params16_mc
 * (multicast addresses of the symmetric parameter buffers): every replica receives the same bits.  The caller runs a cross-rank
 * barrier before the call (every rank's local gradients of the range are complete) and one after it (every shard has been read:
 * each rank may now clear its own gradient replica; every new parameter has been delivered).  max_ctas bounds the grid (0 =
 * the default of 32 CTAs, which leaves the SMs to the backward pass the call overlaps; extents exchanged after backward take more). */
int32_t b200mart_adamw_multicast(b200mart_handle* h, const float* params32, float* exp_avg, float* exp_avg_sq,
                                 float* grads_mc, float* params32_mc, void* params16_mc, int32_t step, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, int32_t pretrain, int64_t begin,
                                 int64_t end, int32_t world, int32_t max_ctas, void* stream);

/* Device-side status flags of the last calls on this workspace (synchronises the stream):
 * bit0 token id out of range, bit1 token type out of range, bit2 a row without exactly one [MASK], bit3 a label or a
 * relaxation-loss position out of range (the reference raises IndexError / asserts in these cases). */
int32_t b200mart_status(b200mart_handle* h, const b200mart_buffers* bufs, void* stream, int32_t* flags_out);

/* NVSwitch multicast reduction (data-parallel gradient exchange without a collective library): mc_dst is the MULTICAST address
 * of a buffer replicated on every GPU of the group (CUDA multicast object / torch symmetric memory); every replica receives
 * mc_dst[i] += scale * src[i] (multimem.red.add.v4.f32: one send per GPU, reduced in the switch).  n % 4 == 0, 16-byte aligned. */
int32_t b200mart_multimem_add(const float* src, float* mc_dst, int64_t n, float scale, void* stream);

/* Layer-by-layer parity hook (what `output_hidden_states=True` gives for the reference, modeling_unimo.py:603-647): copies the
 * fp32 residual stream saved by the last b200mart_forward on this workspace into `out`.  stream_kind 0 = text hidden state
 * entering text layer `layer` (layer = layers: the encoder output), (B, L, hidden); 1 = vision hidden state entering vision
 * layer `layer` (layer = layers: the vision output), (B, 1 + 2 * patches, hidden). */
int32_t b200mart_read_hidden(b200mart_handle* h, const b200mart_buffers* bufs, int32_t stream_kind, int32_t layer, float* out,
                             void* stream);

/* Optional device timing by kernel family, for roofline reporting: between begin and end every tensor-core GEMM
 * (index 0), fused attention (1) and fusion (2) launch is bracketed by CUDA events on the launching stream.
 * end synchronises the stream; gemm_flops = sum of 2*M*N*K over the GEMM launches. Never used in a timed region. */
int32_t b200mart_profile_begin(b200mart_handle* h);
int32_t b200mart_profile_end(b200mart_handle* h, void* stream, double ms_by_cat[3], int64_t launches_by_cat[3],
                             double* gemm_flops);
/* algorithmic bytes (operands + epilogue inputs/outputs, each once) of the GEMM launches of the last profiled window */
double b200mart_profile_gemm_bytes(const b200mart_handle* h);
/* Intra-kernel timeline of the tensor-core product (profiling runs only, never inside a timed region): while a device buffer
 * of `records` x 148 x 16 uint64 is registered, the i-th GEMM launch writes, per CTA, clock64 stamps of its phases (entry,
 * prologue done, first TMA issued, first operands landed, first tile's MMAs issued / retired, first / last tile's epilogue
 * done, exit) plus globaltimer at entry/exit, tile count, SM id and the problem shape (gemm_kernel.cuh: GemmStamp).
 * Passing NULL switches it off. Returns the number of launches recorded since the previous call. */
int64_t b200mart_gemm_timeline(void* device_buffer, int64_t records);

/* ------------------------------------------------------------------------------------------------
 * Pre-LayerNorm encoder stack: the block FLAVA and ViLT are built from
 *   FlavaLayer / FlavaSelfAttention  MarT/models/modeling_flava.py:620-665, :441-521  (FlavaEncoder: image, text and
 *                                    multimodal encoders; adaptive scaling on rows 1:s of the text encoder)
 *   ViltLayer / ViltSelfAttention    MarT/models/modeling_vilt.py:491-528, :329-400   (rows img_length+1 : s')
 * h = x + drop(Wo attn(LN x) + bo); y = h + drop(W2 act(W1 LN h + b1) + b2); head_dim 64; act 1 = erf-GELU, 2 = quick-GELU.
 * Parameters live in flat buffers laid out by the library (names follow the HF modules: layer.<i>.attention.attention.query.weight,
 * layer.<i>.attention.output.dense.*, layer.<i>.intermediate.dense.*, layer.<i>.output.dense.*, layer.<i>.layernorm_before/after.*,
 * layer.<i>.attention.attention.adaptive_weight.{0,1}); buffers/workspace conventions as for the MKGformer handle.
 *
 * post_ln = 1 selects the post-LayerNorm block instead:
 *   VisualBertLayer / VisualBertSelfAttention  MarT/models/modeling_visual_bert.py:381-421, :203-286 (adaptive scaling :255-260 on
 *                                              rows img_length+1 : s' of the [text ; 72 region features] sequence)
 * h = LN(x + drop(Wo attn(x) + bo)); y = LN(h + drop(W2 act(W1 h + b1) + b2)), names layer.<i>.attention.self.{query,key,value}.*,
 * layer.<i>.attention.self.adaptive_weight.{0,1}, layer.<i>.attention.output.{dense,LayerNorm}.*, layer.<i>.intermediate.dense.*,
 * layer.<i>.output.{dense,LayerNorm}.*. */
typedef struct b200mart_stack_desc {
    int32_t hidden, layers, heads, intermediate;
    int32_t act;
    float eps;
    int32_t post_ln;           /* 0: the pre-LayerNorm block above; 1: the BERT (post-LayerNorm) block of VisualBERT, see below */
} b200mart_stack_desc;
typedef struct b200mart_stack b200mart_stack;
typedef struct b200mart_stack_io {
    int32_t B, T;
    const float* x;            /* (B*T, hidden) fp32 input hidden states */
    const int64_t* key_mask;   /* (B, T) attention_mask (1 = keep) -> additive -1e4, or NULL */
    const int64_t* split;      /* (B) adaptive split position s per sample (sep_idx[:, 2]), or NULL: no scaling */
    int32_t row_lo;            /* scores of rows [row_lo, s) x cols >= s are scaled by clamp(w0,0,.5); rows >= s by clamp(w1,.5,1) */
    float* y;                  /* (B*T, hidden) fp32 output hidden states (forward) */
    float p_hidden, p_attn;    /* dropout probabilities (0 = eval) and seed; sites as in b200mart_set_dropout */
    uint32_t seed;
} b200mart_stack_io;
int32_t b200mart_stack_create(const b200mart_stack_desc* desc, b200mart_stack** out);
void b200mart_stack_destroy(b200mart_stack* st);
int32_t b200mart_stack_param_count(const b200mart_stack* st);
int64_t b200mart_stack_param_total(const b200mart_stack* st);
int32_t b200mart_stack_param_info(const b200mart_stack* st, int32_t index, const char** name, int64_t* offset, int32_t* ndim,
                                  int64_t shape[4]);
int64_t b200mart_stack_workspace_bytes(const b200mart_stack* st, int32_t B, int32_t T);
int32_t b200mart_stack_forward(b200mart_stack* st, const b200mart_buffers* bufs, const b200mart_stack_io* io, void* stream);
/* dy (B*T, hidden) fp32 in; dx (B*T, hidden) fp32 out (may be NULL); parameter gradients accumulate into bufs->grads32 */
int32_t b200mart_stack_backward(b200mart_stack* st, const b200mart_buffers* bufs, const b200mart_stack_io* io, const float* dy,
                                float* dx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Stand-alone operators (kernel-level parity tests; also the building blocks of the functions above).
 */
/* Fused attention, head_dim 64 (BertSelfAttention :317-377 with adaptive scaling :342-349 and -1e4 mask;
 * CLIPAttention :212-272 with an optional K/V prefix).  q/k/v bf16 row-major with leading dimensions in
 * elements, head h at columns [64h, 64h+64).  prefix_k/v: (B*n_prefix rows) or NULL. */
typedef struct b200mart_attn_desc {
    int32_t B, Lq, heads, n_prefix, n_keys;
    float scale;
    const void* q; int64_t ldq;
    const void* prefix_k; const void* prefix_v; int64_t ld_prefix;
    const void* k; const void* v; int64_t ldkv;
    const int64_t* key_mask;     /* (B, n_keys) 1 = keep, NULL = no mask */
    const int64_t* sep_idx;      /* (B, 6) or NULL */
    const float* w0; const float* w1;
    void* out; int64_t ldo;      /* bf16 context (B*Lq, heads*64): written by fwd, read by bwd (delta = rowsum(dO*O)) */
    float* out_f32;              /* optional fp32 copy, same layout */
    float* lse;                  /* (B, heads, Lq) */
    /* backward only */
    const float* dout; int64_t lddo;
    void* dq; void* dk; void* dv; int64_t ldd;             /* bf16 */
    const float* dk_add; const float* dv_add; int64_t ldadd;
    float* dprefix_k; float* dprefix_v; int64_t ldd_prefix; /* fp32 */
    float* dw0; float* dw1;
    float* delta;                /* scratch (B, heads, Lq) */
} b200mart_attn_desc;
int32_t b200mart_attention_fwd(const b200mart_attn_desc* d, void* stream);
int32_t b200mart_attention_bwd(const b200mart_attn_desc* d, void* stream);

/* BertFusion (modeling_unimo.py:400-414), fp32: out = softmax(ctx vh^T) vh. */
int32_t b200mart_fusion_fwd(const float* ctx, const float* vh, void* out_bf16, float* lse, int32_t B, int32_t L,
                            int32_t Nv, int32_t H, void* stream);
int32_t b200mart_fusion_bwd(const float* dg, const float* ctx, const float* vh, const float* lse, float* pbuf,
                            float* dsbuf, float* dctx_accum, const float* dvh_in, float* dvh_out, int32_t B, int32_t L,
                            int32_t Nv, int32_t H, void* stream);

/* LayerNorm forward / backward over (M, H) fp32 rows, H % 128 == 0, H <= 1024. */
int32_t b200mart_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, float* y32, void* y16,
                               float* mean, float* rstd, int32_t M, int32_t H, void* stream);
int32_t b200mart_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                               const float* add, float* dx32, void* dx16, float* dgamma, float* dbeta, int32_t M,
                               int32_t H, void* stream);

/* LabelSmoothSoftmaxCEV1 (lit_models/utils.py:42-66) fused with its backward: loss[0] += mean loss,
 * dlogits (B, C) fp32. */
int32_t b200mart_ls_ce(const float* logits, int32_t B, int32_t C, const int64_t* label, float eps, float* loss,
                       float* dlogits, void* stream);

/* Input path.  Replaces the image half of the reference collator, MarT/data/data_module.py:127-160
 * (`entity2visual_features[entities.index(name)]` per sample on the host, then a 38.5 MB host -> device copy per 32-example
 * batch): the preprocessed entity images live on the device once (`table`: (n_entities, image_elems) fp32 or bf16) and a batch's
 * pixel_values (n_images = 2 B images of image_elems = 3 S S floats) are gathered by entity index; idx < 0 = the all-zero
 * "[PAD] image" (:141-142, :152-153).  idx >= n_entities sets bit 16 of *status (may be NULL) and yields zeros. */
int32_t b200mart_gather_images(const void* table, int32_t table_is_bf16, const int64_t* idx, float* out, int32_t n_images,
                               int64_t image_elems, int64_t n_entities, int32_t* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200MART_H */
