// Launchers of the non-GEMM kernels of the MKGformer step. Every launcher is stream-ordered, allocates
// nothing and returns 0 / error code (message via b200mart_last_error()).
#pragma once
#include "common.cuh"

// ---- elementwise.cu
int launch_cast_bf16(const float* src, bf16* dst, long n, cudaStream_t s);
int launch_multimem_add(const float* src, float* mc_dst, long n, float scale, cudaStream_t s);
int launch_fill_f32(float* dst, float v, long n, cudaStream_t s);
int launch_axpy_f32(const float* src, float* dst, long n, cudaStream_t s);   // dst += src
int launch_cast_pad_bf16(const float* src, long ld_src, bf16* dst, long ld_dst, int rows, int cols, cudaStream_t s);
// pixel_values (B,2,3,S,S) fp32 -> patch rows bf16 [(b*2+img)*P + py*g + px][c*p*p + iy*p + ix]
int launch_im2col(const float* px, bf16* out, int B, int S, int p, long ldo, cudaStream_t s);
// CLIPVisionEmbeddings + vision_pre_layrnorm (modeling_unimo.py:119-132, :711)
int launch_vision_embed_ln(const float* patch_out, const float* cls, const float* pos, const float* gamma,
                           const float* beta, float eps, float* x0, float* out_f32, float* mean, float* rstd, int B,
                           int P, int H, cudaStream_t s);
// BertEmbeddings (modeling_unimo.py:152-186), dropout p = 0
int launch_text_embed_ln(const int64_t* ids, const int64_t* tt, const float* E, const float* pos, const float* type,
                         const float* gamma, const float* beta, float eps, float* pre, float* out_f32, bf16* out_bf16,
                         float* mean, float* rstd, int B, int L, int H, int V, int TV, int* status, DropCfg drop, cudaStream_t s);
int launch_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, float* y32, bf16* y16,
                         float* mean, float* rstd, int M, int H, cudaStream_t s);
// dx = LN'(dy) [* act'(dact_src)] [+ add];  dgamma += ..., dbeta += ...; dxsum[c] += sum_rows dx[:, c] when given
// (the bias gradient of the linear layer whose output this LayerNorm normalised)
int launch_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                         const float* add, const bf16* dact_src, int dact, float* dx32, bf16* dx16, float* dgamma,
                         float* dbeta, float* dxsum, int M, int H, cudaStream_t s, DropCfg dy_drop = DropCfg{0u, 0u, 1.f},
                         DropCfg dx16_drop = DropCfg{0u, 0u, 1.f});
// dy_drop: dy is the gradient wrt a DROPPED copy of this LayerNorm's output (element index row*H+col);
// dx16_drop: dx16 (and dxsum) additionally carry the dropout mask of the dense branch that fed the LayerNorm input
int launch_colsum_bf16(const bf16* x, long ld, float* out, int M, int N, cudaStream_t s);
// dst16 = bf16(src * dropout mask), colsum += column sums of the masked values (src contiguous [M, N])
int launch_mask_cast_colsum(const float* src, bf16* dst, float* colsum, int M, int N, DropCfg drop, cudaStream_t s);  // out[n] += sum_m x[m,n]
int launch_colsum_f32(const float* x, long ld, float* out, int M, int N, cudaStream_t s);
// text embedding backward: dE[ids] += d, dtype[tt] += d, dpos[l] += d  (d = gradient wrt the pre-LN sum)
int launch_text_embed_bwd(const float* d, const int64_t* ids, const int64_t* tt, float* dE, float* dpos, float* dtype,
                          int B, int L, int H, int V, int TV, cudaStream_t s);   // ids / types clamped as in the forward kernel
// vision embedding backward: dcls, dpos accumulate; dpatch_out bf16 [(B*2*P), H] (A^T operand of the dWp GEMM)
int launch_vision_embed_bwd(const float* dx0, float* dcls, float* dpos, bf16* dpatch, int B, int P, int H,
                            cudaStream_t s);
// gather rows: out[r, :] = src[idx[r], :]
// idx values outside [0, V): status bit 0 is raised and row 0 is read (gathers), the element is skipped (scatters)
int launch_gather_rows_bf16(const bf16* src, const int64_t* idx, bf16* out, int R, int H, long V, int* status, cudaStream_t s);
int launch_gather_f32(const float* src, const int64_t* idx, float* out, int R, long V, cudaStream_t s);
int launch_scatter_add_rows_f32(const float* src, const int64_t* idx, float* dst, int R, int H, long V, cudaStream_t s);
int launch_scatter_add_f32(const float* src, const int64_t* idx, float* dst, int R, long V, cudaStream_t s);
// AdamW over a flat buffer with per-segment flags (lit_models/transformer.py:224-241)
struct AdamSeg { long begin, end; int decay; int skip; };
// over the flat range [begin, end); zero_grads: clear g after use
int launch_adamw(float* p, float* g, float* m, float* v, bf16* p16, const AdamSeg* segs_dev, int nseg, long begin, long end,
                 int zero_grads, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2, cudaStream_t s,
                 const float* dyn = nullptr);   // dyn: device {salt, lr, bc1, bc2} overriding the by-value scalars (graph replay)

// the same on this rank's shard with the exchange fused in (multimem.ld_reduce of the gradients, multimem.st of the result)
int launch_adamw_mc(const float* p, float* m, float* v, float* g_mc, float* p_mc, bf16* p16_mc, const AdamSeg* segs_dev,
                    int nseg, long begin, long end, float inv_world, float lr, float b1, float b2, float eps, float wd, float bc1,
                    float bc2, cudaStream_t s, int max_ctas = 0);

// entity-image table gather (collator replacement, MarT/data/data_module.py:127-160): out[i] = table[idx[i]], zeros for idx < 0
int launch_gather_images(const void* table, int table_is_bf16, const int64_t* idx, float* out, int n_images, long image_elems,
                         long n_entities, int* status, cudaStream_t s);

// ---- attention_mma.cu (self-attention) / fusion.cu (BertFusion)
struct AttnFwd {
    DropCfg drop;                 // dropout on the attention probabilities (thresh 0 = off); element ((b*nh+h)*Lq+r)*n+c
    // Q rows: [B*Lq] rows of q (ldq), head h at column h*64. Keys: optional prefix source (n1 keys per sample) then own (n2).
    const bf16* q; long ldq;
    const bf16* k1; const bf16* v1; long ld1; int n1;   // prefix (text K/V for the vision stream), may be null
    const bf16* k2; const bf16* v2; long ld2; int n2;
    int B, Lq, nh;
    float scale;
    const int64_t* key_mask;      // (B, n2) attention_mask (1 = keep) or null; adds -10000 where 0
    const int64_t* sep_idx;       // adaptive split column s per sample: s = sep_idx[b * sep_stride]; null = no scaling
    int sep_stride;               //   (MKGformer passes sep_idx + 2 with stride 6, i.e. sep_idx[b][2])
    int row_lo;                   // rows [row_lo, s) x cols >= s are scaled by clamp(w0,0,.5); rows >= s by clamp(w1,.5,1)
    const float* w0; const float* w1;   // adaptive weights (device scalars), used when sep_idx != null
    bf16* out; long ldo;          // context [B*Lq, nh*64]
    float* lse;                   // [B, nh, Lq] log-sum-exp of the final scores
};
// out32: optional fp32 copy of the context (consumed by BertFusion)
int launch_attention_fwd(const AttnFwd& a, float* out32, cudaStream_t s);
struct AttnBwd {
    AttnFwd f;                    // same inputs as forward (out unused)
    const float* dout; long lddo; // fp32 gradient wrt context [B*Lq, nh*64]
    const bf16* dout16 = nullptr; long lddo16 = 0;   // optional bf16 copy of dout (same layout): lets the tiled keys pass stream it with cp.async
    bf16* dq; bf16* dk2; bf16* dv2; long ldd;      // bf16 grads (same column layout as q/k2/v2)
    const float* dk2_add; const float* dv2_add; long ldadd;  // optional fp32 extra gradient added into dk2/dv2
    float* dk1; float* dv1; long ldd1;             // fp32 grads wrt the prefix keys/values (plain store), when n1 > 0
    float* dw0; float* dw1;       // adaptive-weight gradients (atomic accumulate), when sep_idx != null
};
// ctx: the forward output (bf16, [B*Lq, nh*64], leading dimension ldctx); delta: scratch [B, nh, Lq] fp32
int launch_attention_bwd(const AttnBwd& a, const bf16* ctx, long ldctx, float* delta, cudaStream_t s);
// BertFusion (modeling_unimo.py:400-414): out = softmax(ctx vh^T) vh, no scale, no mask
int launch_fusion_fwd(const float* ctx, const float* vh, bf16* out, float* lse, int B, int L, int Nv, int H,
                      cudaStream_t s);
// fusion backward, phase A (per query row): P, dS stored to pbuf/dsbuf [B, L, Nv]; dctx (fp32) += dS vh
int launch_fusion_bwd_rows(const float* dg, const float* ctx, const float* vh, const float* lse, float* pbuf,
                           float* dsbuf, float* dctx, int B, int L, int Nv, int H, cudaStream_t s);
// phase B (per key row): dvh_out = dvh_in (optional) + dS^T ctx + P^T dg ; writes fp32 + bf16
int launch_fusion_bwd_keys(const float* dg, const float* ctx, const float* pbuf, const float* dsbuf,
                           const float* dvh_in, float* dvh32, bf16* dvh16, int B, int L, int Nv, int H, cudaStream_t s);

// ---- loss.cu
int launch_find_mask(const int64_t* ids, int* mask_pos, int B, int L, int mask_id, int* status, cudaStream_t s);
// zm[b, :] = z16[b*L + mask_pos[b], :]
int launch_gather_mask_rows(const bf16* z16, const int* mask_pos, bf16* zm, int B, int L, int H, cudaStream_t s);
struct LossGroup {
    const float* logits; long ld; int C;   // (B, C) fp32
    float* dlogits32;                      // (B, ldd) fp32 gradient (zero rows for unselected samples)
    bf16* dlogits16; long ldd;
    int select;                            // 0 = all rows, 1 = rows with pre_type != 2, 2 = rows with pre_type == 2
};
// label-smoothed CE (lit_models/utils.py:42-66) or plain CE (eps = 0), mean over the group's rows; loss[0] += CE
int launch_ls_ce(const LossGroup& g, const int64_t* label, const int64_t* pre_type, float eps, float* loss, int B,
                 int* status, cudaStream_t s, const float* grad_scale = nullptr);   // status (may be null): bit 8 = label out of range
// relaxation loss (lit_models/transformer.py:103-108): loss[0] += alpha * mean(...); dz rows += gradient
int launch_relaxation(const float* z, const int64_t* rel_idx, const int64_t* q_head, const int64_t* a_head, float alpha,
                      float* loss, float* dz, int B, int L, int H, int* status, cudaStream_t s, const float* grad_scale = nullptr);
// dz[b*L + mask_pos[b], :] += dzm[b, :]
int launch_scatter_mask_rows(const float* dzm, const int* mask_pos, float* dz, int B, int L, int H, cudaStream_t s);
// rank of gold = 1 + #{c : x[b,c] > x[b,label]}  (lit_models/transformer.py:162-164 without the double sort)
int launch_ranks(const float* logits, long ld, int C, const int64_t* label, const int64_t* pre_type, int select,
                 int64_t* ranks, int B, cudaStream_t s);
