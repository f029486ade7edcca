// MKGformer ("Unimo") forward / score-head loss / backward / AdamW as a stream of sm_100a kernels behind the
// C ABI of include/b200mart.h.  What it computes is specified in SURVEY.md Appendix A; the reference graph is
//   UnimoForMaskedLM.forward            MarT/models/modeling_unimo.py:848-893
//   UnimoEncoder.forward (interleaving) :589-658      CLIPEncoderLayer :490-527      BertLayer :540-577
//   TransformerLitModel.training_step   MarT/lit_models/transformer.py:59-113
// Design: the host side is a fixed, allocation-free launch sequence (CUDA-graph capturable); every matrix
// product is one tcgen05 GEMM with a fused epilogue (gemm_tcgen05.cu); the fp32 residual streams, LayerNorm
// statistics, softmax and the loss stay fp32; GEMM operands are bf16 (activations are written in bf16 by the
// producing kernel, weights come from the flat bf16 mirror of the parameters).
#include "gemm.cuh"
#include "kernels.cuh"
#include "../../include/b200mart.h"

#include <stdlib.h>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct PInfo {
    std::string name;
    long offset = 0;
    int ndim = 0;
    long shape[4] = {0, 0, 0, 0};
    int decay = 1;
    int unused_ft = 0;   // never receives a gradient in the reference (fine-tune)
    long numel() const { long n = 1; for (int i = 0; i < ndim; ++i) n *= shape[i]; return n; }
};

struct VisLayerP { long ln1_w, ln1_b, qkv_w, qkv_b, o_w, o_b, ln2_w, ln2_b, fc1_w, fc1_b, fc2_w, fc2_b; };
struct TxtLayerP { long qkv_w, qkv_b, aw0, aw1, ao_w, ao_b, aln_w, aln_b, i_w, i_b, f_w, f_b, o_w, o_b, oln_w, oln_b; };
struct ParamLayout {
    long cls, patch_w, vpos, preln_w, preln_b, postln_w, postln_b;
    long word, tpos, ttype, eln_w, eln_b, pool_w, pool_b;
    long dec_b, tr_w, tr_b, trln_w, trln_b;
    std::vector<VisLayerP> vis;
    std::vector<TxtLayerP> txt;
    long total = 0;
};

// ---- workspace carving
struct Carver {
    uint8_t* base; size_t off = 0;
    template <typename T> T* take(size_t n) {
        off = (off + 255) & ~size_t(255);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};

// pre16 / ipre16 hold the activation DERIVATIVE act'(pre-activation) (saved by the forward GEMM epilogue)
struct VisLayerA { bf16 *u16, *qkv16, *ctx16, *u2_16, *pre16, *h16; float *mean1, *rstd1, *lse, *x1, *mean2, *rstd2; };
struct TxtLayerA {
    bf16 *qkv16, *ctx16, *g16, *a16, *ipre16, *i16;
    float *lse, *ctx32, *flse, *apre, *amean, *arstd, *a32, *opre, *omean, *orstd;
};
struct Workspace {
    int B = 0, L = 0;
    int* status; int* mask_pos; float* loss;
    bf16* patches16; bf16* wp16; float* patch_out; float* vx0; float *vmean0, *vrstd0;
    std::vector<float*> V32;            // D+1 vision residual streams
    std::vector<VisLayerA> vl;
    float* tpre; float *emean, *erstd;
    std::vector<float*> T32; std::vector<bf16*> T16;
    std::vector<TxtLayerA> tl;
    // head
    bf16* hpre16; float* hact; float *hmean, *hrstd; float* z32; bf16* z16; bf16* zm16;
    bf16* esel16; float* bias_sel; float* logits; bf16* dlogits16; float* dzm; float* desel; float* dbias_sel;
    float* dz;                          // gradient wrt trans_hidden_states (B*L, H)
    // backward scratch
    // dY scratch read by the weight-gradient side streams alternates by layer parity ([2])
    float *t32a, *t32b, *t32c, *dT; bf16 *t16a[2], *t16b[2], *tF16[2], *tdqkv16[2];
    float *v32a, *v32b, *dV[2]; bf16 *v16a[2], *dV16[2], *vF16[2], *vdqkv16[2];
    float *dk1, *dv1; float *pbuf, *dsbuf; float *delta, *delta_v;   // delta: text chain, delta_v: vision chain
    bf16* dpatch16;
    size_t bytes = 0;
};

inline long align_up(long x, long a) { return (x + a - 1) / a * a; }

}  // namespace

struct b200mart_handle {
    b200mart_model_desc d;
    int P = 0, Nv = 0, patch_dim = 0, patch_ld = 0;   // patch_ld: patch_dim rounded up to 8 (16-byte bf16 rows for TMA)
    std::vector<PInfo> params;
    ParamLayout lay;
    std::vector<long> layer_begin, layer_end;   // flat extent of layer i's parameters (vision + text)
    std::unordered_map<std::string, GemmPlan*> plans;
    AdamSeg* segs_dev[2] = {nullptr, nullptr};   // [0] fine-tune, [1] pre-train (uploaded on first use)
    std::vector<AdamSeg> segs_host[2];
    int nseg = 0;
    // cached workspace carving
    Workspace ws; void* ws_base = nullptr;
    // dropout (BERT stream; reference p = 0.1 in training): masks are regenerated from (seed, site, element) in backward
    bool training = false;
    float p_hidden = 0.f, p_attn = 0.f;
    uint32_t seed0 = 0, cur_seed = 0;
    uint32_t* step_block = nullptr;            // device block {salt, lr, 1 - beta1^t, 1 - beta2^t} (b200mart_set_step_block) or null
    unsigned long long fwd_count = 0;
    float cur_ph = 0.f, cur_pa = 0.f;          // probabilities in effect for the forward whose activations are saved
    // two-chain execution: the vision and text streams of the model are independent below the fusion depth
    // (modeling_unimo.py:616-640), so each runs on its own CUDA stream, ordered by events where they meet
    bool split_streams = true;
    cudaStream_t sV = nullptr, sT = nullptr, sVw = nullptr, sTw = nullptr;
    std::vector<cudaEvent_t> evpool;
    size_t ev_used = 0;
    // optional per-category device timing (CUDA events around each launch group; see b200mart_profile_begin/end)
    bool profiling = false;
    struct ProfRec { cudaEvent_t e0, e1; int cat; };
    std::vector<ProfRec> prof;
    size_t prof_used = 0;
    double prof_flops = 0.0, prof_bytes = 0.0;   // GEMM FLOPs / algorithmic operand+result bytes of the profiled launches
    ~b200mart_handle() {
        for (auto& kv : plans) gemm_plan_destroy(kv.second);
        for (auto* p : segs_dev) if (p) cudaFree(p);
        for (auto& r : prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
        for (auto& e : evpool) cudaEventDestroy(e);
        if (sV) cudaStreamDestroy(sV);
        if (sT) cudaStreamDestroy(sT);
        if (sVw) cudaStreamDestroy(sVw);
        if (sTw) cudaStreamDestroy(sTw);
    }
};

enum { CAT_GEMM = 0, CAT_ATTN = 1, CAT_FUSION = 2, CAT_COUNT = 3 };

// RAII: when profiling, brackets the launches issued in its scope with two events on the launching stream
struct ProfScope {
    b200mart_handle* h; cudaStream_t s; b200mart_handle::ProfRec* rec = nullptr;
    ProfScope(b200mart_handle* h_, cudaStream_t s_, int cat) : h(h_), s(s_) {
        if (!h->profiling) return;
        if (h->prof_used == h->prof.size()) {
            b200mart_handle::ProfRec r;
            if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
            h->prof.push_back(r);
        }
        rec = &h->prof[h->prof_used++];
        rec->cat = cat;
        cudaEventRecord(rec->e0, s);
    }
    ~ProfScope() { if (rec) cudaEventRecord(rec->e1, s); }
};

namespace {

// ------------------------------------------------------------------ parameter layout
long add_param(b200mart_handle* h, const std::string& name, std::initializer_list<long> shape) {
    PInfo p;
    p.name = name;
    p.ndim = (int)shape.size();
    int i = 0;
    for (long s : shape) p.shape[i++] = s;
    p.offset = h->lay.total;
    // reference rule (lit_models/transformer.py:225-230): no decay for names containing "bias" or "LayerNorm.weight"
    p.decay = (name.find("bias") != std::string::npos || name.find("LayerNorm.weight") != std::string::npos) ? 0 : 1;
    h->lay.total += align_up(p.numel(), 64);
    h->params.push_back(p);
    return p.offset;
}

void build_layout(b200mart_handle* h) {
    const b200mart_model_desc& d = h->d;
    const long H = d.hidden, F = d.intermediate, V = d.vocab;
    ParamLayout& l = h->lay;
    l.cls = add_param(h, "unimo.vision_embeddings.class_embedding", {H});
    l.patch_w = add_param(h, "unimo.vision_embeddings.patch_embedding.weight", {H, 3, d.patch, d.patch});
    l.vpos = add_param(h, "unimo.vision_embeddings.position_embedding.weight", {h->P + 1, H});
    l.preln_w = add_param(h, "unimo.vision_pre_layrnorm.weight", {H});
    l.preln_b = add_param(h, "unimo.vision_pre_layrnorm.bias", {H});
    l.postln_w = add_param(h, "unimo.vision_post_layernorm.weight", {H});
    l.postln_b = add_param(h, "unimo.vision_post_layernorm.bias", {H});
    h->params[h->params.size() - 1].unused_ft = h->params[h->params.size() - 2].unused_ft = 1;
    l.word = add_param(h, "unimo.text_embeddings.word_embeddings.weight", {V, H});
    l.tpos = add_param(h, "unimo.text_embeddings.position_embeddings.weight", {d.max_position, H});
    l.ttype = add_param(h, "unimo.text_embeddings.token_type_embeddings.weight", {d.type_vocab, H});
    l.eln_w = add_param(h, "unimo.text_embeddings.LayerNorm.weight", {H});
    l.eln_b = add_param(h, "unimo.text_embeddings.LayerNorm.bias", {H});
    char buf[160];
    // one contiguous extent per layer index (vision layer i, then text layer i): the gradient exchange of a layer
    // can start as soon as that layer's backward has run on both chains
    for (int i = 0; i < d.layers; ++i) {
        h->layer_begin.push_back(l.total);
        {
        VisLayerP v;
        auto nm = [&](const char* s) { snprintf(buf, sizeof(buf), "unimo.encoder.vision_layers.%d.%s", i, s); return std::string(buf); };
        v.qkv_w = add_param(h, nm("self_attn.q_proj.weight"), {H, H});
        add_param(h, nm("self_attn.k_proj.weight"), {H, H});
        add_param(h, nm("self_attn.v_proj.weight"), {H, H});
        v.qkv_b = add_param(h, nm("self_attn.q_proj.bias"), {H});
        add_param(h, nm("self_attn.k_proj.bias"), {H});
        add_param(h, nm("self_attn.v_proj.bias"), {H});
        v.o_w = add_param(h, nm("self_attn.out_proj.weight"), {H, H});
        v.o_b = add_param(h, nm("self_attn.out_proj.bias"), {H});
        v.ln1_w = add_param(h, nm("layer_norm1.weight"), {H});
        v.ln1_b = add_param(h, nm("layer_norm1.bias"), {H});
        v.fc1_w = add_param(h, nm("mlp.fc1.weight"), {F, H});
        v.fc1_b = add_param(h, nm("mlp.fc1.bias"), {F});
        v.fc2_w = add_param(h, nm("mlp.fc2.weight"), {H, F});
        v.fc2_b = add_param(h, nm("mlp.fc2.bias"), {H});
        v.ln2_w = add_param(h, nm("layer_norm2.weight"), {H});
        v.ln2_b = add_param(h, nm("layer_norm2.bias"), {H});
        l.vis.push_back(v);
        }
        {
        TxtLayerP t;
        auto nm = [&](const char* s) { snprintf(buf, sizeof(buf), "unimo.encoder.text_layer.%d.%s", i, s); return std::string(buf); };
        t.qkv_w = add_param(h, nm("attention.self.query.weight"), {H, H});
        add_param(h, nm("attention.self.key.weight"), {H, H});
        add_param(h, nm("attention.self.value.weight"), {H, H});
        t.qkv_b = add_param(h, nm("attention.self.query.bias"), {H});
        add_param(h, nm("attention.self.key.bias"), {H});
        add_param(h, nm("attention.self.value.bias"), {H});
        t.aw0 = add_param(h, nm("attention.self.adaptive_weight.0"), {1});
        t.aw1 = add_param(h, nm("attention.self.adaptive_weight.1"), {1});
        t.ao_w = add_param(h, nm("attention.output.dense.weight"), {H, H});
        t.ao_b = add_param(h, nm("attention.output.dense.bias"), {H});
        t.aln_w = add_param(h, nm("attention.output.LayerNorm.weight"), {H});
        t.aln_b = add_param(h, nm("attention.output.LayerNorm.bias"), {H});
        t.i_w = add_param(h, nm("intermediate.dense.weight"), {F, H});
        t.i_b = add_param(h, nm("intermediate.dense.bias"), {F});
        t.f_w = add_param(h, nm("intermediate.fusion_dense.weight"), {F, H});
        t.f_b = add_param(h, nm("intermediate.fusion_dense.bias"), {F});
        if (i < d.fuse_from) h->params[h->params.size() - 1].unused_ft = h->params[h->params.size() - 2].unused_ft = 1;
        t.o_w = add_param(h, nm("output.dense.weight"), {H, F});
        t.o_b = add_param(h, nm("output.dense.bias"), {H});
        t.oln_w = add_param(h, nm("output.LayerNorm.weight"), {H});
        t.oln_b = add_param(h, nm("output.LayerNorm.bias"), {H});
        l.txt.push_back(t);
        }
        h->layer_end.push_back(l.total);
    }
    l.pool_w = add_param(h, "unimo.text_pooler.dense.weight", {H, H});
    l.pool_b = add_param(h, "unimo.text_pooler.dense.bias", {H});
    h->params[h->params.size() - 1].unused_ft = h->params[h->params.size() - 2].unused_ft = 1;
    l.dec_b = add_param(h, "cls.predictions.bias", {V});
    l.tr_w = add_param(h, "cls.predictions.transform.dense.weight", {H, H});
    l.tr_b = add_param(h, "cls.predictions.transform.dense.bias", {H});
    l.trln_w = add_param(h, "cls.predictions.transform.LayerNorm.weight", {H});
    l.trln_b = add_param(h, "cls.predictions.transform.LayerNorm.bias", {H});
}

// ------------------------------------------------------------------ workspace
void carve(const b200mart_handle* h, int B, int L, void* base, Workspace& w) {
    const b200mart_model_desc& d = h->d;
    const size_t H = d.hidden, F = d.intermediate, nh = d.heads, Nv = h->Nv, P = h->P;
    const size_t Mt = (size_t)B * L, Mv = (size_t)B * Nv;
    const size_t Vp = align_up(d.vocab, 8);
    Carver c{reinterpret_cast<uint8_t*>(base)};
    w.B = B; w.L = L;
    w.status = c.take<int>(8);
    w.mask_pos = c.take<int>(B);
    w.loss = c.take<float>(8);
    w.patches16 = c.take<bf16>((size_t)B * 2 * P * h->patch_ld);
    w.wp16 = c.take<bf16>(H * h->patch_ld);
    w.patch_out = c.take<float>((size_t)B * 2 * P * H);
    w.vx0 = c.take<float>(Mv * H);
    w.vmean0 = c.take<float>(Mv); w.vrstd0 = c.take<float>(Mv);
    w.V32.assign(d.layers + 1, nullptr);
    for (auto& p : w.V32) p = c.take<float>(Mv * H);
    w.vl.assign(d.layers, VisLayerA());
    for (auto& a : w.vl) {
        a.u16 = c.take<bf16>(Mv * H); a.mean1 = c.take<float>(Mv); a.rstd1 = c.take<float>(Mv);
        a.qkv16 = c.take<bf16>(Mv * 3 * H); a.lse = c.take<float>((size_t)B * nh * Nv);
        a.ctx16 = c.take<bf16>(Mv * H); a.x1 = c.take<float>(Mv * H);
        a.u2_16 = c.take<bf16>(Mv * H); a.mean2 = c.take<float>(Mv); a.rstd2 = c.take<float>(Mv);
        a.pre16 = c.take<bf16>(Mv * F); a.h16 = c.take<bf16>(Mv * F);
    }
    w.tpre = c.take<float>(Mt * H); w.emean = c.take<float>(Mt); w.erstd = c.take<float>(Mt);
    w.T32.assign(d.layers + 1, nullptr); w.T16.assign(d.layers + 1, nullptr);
    for (int i = 0; i <= d.layers; ++i) { w.T32[i] = c.take<float>(Mt * H); w.T16[i] = c.take<bf16>(Mt * H); }
    w.tl.assign(d.layers, TxtLayerA());
    for (int i = 0; i < d.layers; ++i) {
        TxtLayerA& a = w.tl[i];
        a.qkv16 = c.take<bf16>(Mt * 3 * H); a.lse = c.take<float>((size_t)B * nh * L);
        a.ctx16 = c.take<bf16>(Mt * H);
        if (i >= d.fuse_from) { a.ctx32 = c.take<float>(Mt * H); a.g16 = c.take<bf16>(Mt * H); a.flse = c.take<float>(Mt); }
        a.apre = c.take<float>(Mt * H); a.amean = c.take<float>(Mt); a.arstd = c.take<float>(Mt);
        a.a32 = c.take<float>(Mt * H); a.a16 = c.take<bf16>(Mt * H);
        a.ipre16 = c.take<bf16>(Mt * F); a.i16 = c.take<bf16>(Mt * F);
        a.opre = c.take<float>(Mt * H); a.omean = c.take<float>(Mt); a.orstd = c.take<float>(Mt);
    }
    w.hpre16 = c.take<bf16>(Mt * H); w.hact = c.take<float>(Mt * H); w.hmean = c.take<float>(Mt); w.hrstd = c.take<float>(Mt);
    w.z32 = c.take<float>(Mt * H); w.z16 = c.take<bf16>(Mt * H);
    w.zm16 = c.take<bf16>((size_t)align_up(B, 128) * H);
    w.esel16 = c.take<bf16>(Vp * H); w.bias_sel = c.take<float>(Vp);
    w.logits = c.take<float>((size_t)B * Vp); w.dlogits16 = c.take<bf16>((size_t)align_up(B, 64) * Vp);
    w.dzm = c.take<float>((size_t)B * H); w.desel = c.take<float>(Vp * H); w.dbias_sel = c.take<float>(Vp);
    w.dz = c.take<float>(Mt * H);
    w.t32a = c.take<float>(Mt * H); w.t32b = c.take<float>(Mt * H); w.t32c = c.take<float>(Mt * H); w.dT = c.take<float>(Mt * H);
    w.v32a = c.take<float>(Mv * H); w.v32b = c.take<float>(Mv * H);
    for (int k = 0; k < 2; ++k) {
        w.t16a[k] = c.take<bf16>(Mt * H); w.t16b[k] = c.take<bf16>(Mt * H); w.tF16[k] = c.take<bf16>(Mt * F); w.tdqkv16[k] = c.take<bf16>(Mt * 3 * H);
        w.dV[k] = c.take<float>(Mv * H); w.dV16[k] = c.take<bf16>(Mv * H);
        w.v16a[k] = c.take<bf16>(Mv * H); w.vF16[k] = c.take<bf16>(Mv * F); w.vdqkv16[k] = c.take<bf16>(Mv * 3 * H);
    }
    w.dk1 = c.take<float>(Mt * H); w.dv1 = c.take<float>(Mt * H);
    w.pbuf = c.take<float>(Mt * Nv); w.dsbuf = c.take<float>(Mt * Nv);
    w.delta = c.take<float>((size_t)B * nh * L);
    w.delta_v = c.take<float>((size_t)B * nh * Nv);
    w.dpatch16 = c.take<bf16>((size_t)B * 2 * P * H);
    w.bytes = c.off + 256;
}

int bind_workspace(b200mart_handle* h, const b200mart_buffers* bufs, int B, int L) {
    B2_CHECK(bufs && bufs->workspace, "workspace pointer is null");
    if (h->ws_base != bufs->workspace || h->ws.B != B || h->ws.L != L) {
        carve(h, B, L, bufs->workspace, h->ws);
        h->ws_base = bufs->workspace;
    }
    B2_CHECK((long)h->ws.bytes <= bufs->workspace_bytes, "workspace too small: need %zu bytes, have %ld", h->ws.bytes,
             (long)bufs->workspace_bytes);
    return 0;
}

// ------------------------------------------------------------------ GEMM through the plan cache
int run_gemm(b200mart_handle* h, const GemmOp& op, cudaStream_t s) {
    // key = every field that shapes the plan (pointers included: tensor maps embed them)
    long key[40];
    int k = 0;
    key[k++] = op.M; key[k++] = op.N; key[k++] = op.K1; key[k++] = op.K2; key[k++] = op.a_mn; key[k++] = op.b_mn;
    key[k++] = (long)op.A; key[k++] = op.lda; key[k++] = (long)op.B; key[k++] = op.ldb;
    key[k++] = (long)op.A2; key[k++] = op.lda2; key[k++] = (long)op.B2; key[k++] = op.ldb2;
    float al = op.alpha; int ali; memcpy(&ali, &al, 4); key[k++] = ali;
    key[k++] = (long)op.bias; key[k++] = (long)op.bias2; key[k++] = (long)op.out_preact; key[k++] = op.ldp; key[k++] = op.act;
    key[k++] = (long)op.dact_src; key[k++] = op.ldx; key[k++] = op.dact; key[k++] = (long)op.residual; key[k++] = op.ldr;
    key[k++] = (long)op.out_f32; key[k++] = op.ldo32; key[k++] = (long)op.out_bf16; key[k++] = op.ldo16;
    key[k++] = op.accumulate; key[k++] = op.split_k; key[k++] = op.bn; key[k++] = op.save_grad;
    key[k++] = (long)op.colsum_out; key[k++] = (long)op.colsum_out2; key[k++] = op.drop.thresh; key[k++] = op.max_ctas;
    key[k++] = op.b_static;
    std::string ks(reinterpret_cast<const char*>(key), k * sizeof(long));
    auto it = h->plans.find(ks);
    GemmPlan* plan = nullptr;
    if (it == h->plans.end()) {
        if (h->plans.size() > 8192) {   // pointer churn (e.g. a caller re-allocating buffers every step): start over
            for (auto& kv : h->plans) gemm_plan_destroy(kv.second);
            h->plans.clear();
        }
        B2_TRY(gemm_plan_create(op, &plan));
        h->plans.emplace(std::move(ks), plan);
    } else {
        plan = it->second;
    }
    ProfScope ps(h, s, CAT_GEMM);
    if (h->profiling) {
        h->prof_flops += gemm_plan_flops(plan);
        // algorithmic bytes of this product: both operands once, every epilogue input and output once
        double b = 2.0 * ((double)op.M * (op.K1 + op.K2) + (double)op.N * (op.K1 + op.K2));
        const double mn = (double)op.M * op.N;
        if (op.out_f32) b += (op.accumulate ? 8.0 : 4.0) * mn;
        if (op.out_bf16) b += 2.0 * mn;
        if (op.out_preact) b += 2.0 * mn;
        if (op.residual) b += 4.0 * mn;
        if (op.dact_src) b += 2.0 * mn;
        h->prof_bytes += b;
    }
    return gemm_plan_launch(plan, s, op.drop.thresh ? &op.drop : nullptr);
}

// bound of the persistent grid of the weight-gradient GEMMs (B200MART_DW_MAXCTAS; 0 = one CTA per SM): they run on side
// streams next to the two dependent chains, and a resident persistent CTA is never preempted whatever the stream priority
int dw_max_ctas() {
    static int cap = -1;
    // round 1: 40 of 148 SMs (7.87 -> 7.67 ms per step against an unbounded grid); round 2, with the lighter LayerNorm / column-sum
    // kernels next to them: 32 / 40 / 48 / 56 / 64 CTAs = 7.10 / 7.04 / 7.03 / 7.00 / 7.00 ms (gpurun_out/ab_r2o.txt) -> 56
    if (cap < 0) { const char* e = getenv("B200MART_DW_MAXCTAS"); cap = e ? atoi(e) : 56; if (cap < 0) cap = 0; }
    return cap;
}

// bounds of the persistent GEMM grids of the vision / text chain while both chains run (B200MART_CAP_V / _T; 0 = none)
int chain_cap(int which) {
    static int cap[2] = {-1, -1};
    if (cap[0] < 0) {
        const char* v = getenv("B200MART_CAP_V"); const char* t = getenv("B200MART_CAP_T");
        cap[0] = v ? atoi(v) : 0; cap[1] = t ? atoi(t) : 0;
        if (cap[0] < 0) cap[0] = 0;
        if (cap[1] < 0) cap[1] = 0;
    }
    return cap[which];
}

// split-K factor for weight-gradient GEMMs so that the grid covers the CTAs it may use (148 SMs, or the side-stream bound)
int pick_split(int M, int N, int K, int ctas = 148) {
    const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
    const int kb = (K + 63) / 64;
    const int want = ctas * 120 / 148;
    int split = 1;
    while (tiles * split < want && split * 2 <= kb && split < 16) split *= 2;
    return split;
}

struct Run {
    b200mart_handle* h;
    const b200mart_buffers* bufs;
    cudaStream_t s;
    int cap = 0;                                // bound of the persistent GEMM grids of this chain (0 = one CTA per SM)
    DropCfg drop_next = DropCfg{0u, 0u, 1.f};   // consumed by the next linear(): dropout before its residual add
    const float* p32(long off) const { return bufs->params32 + off; }
    const bf16* p16(long off) const { return reinterpret_cast<const bf16*>(bufs->params16) + off; }
    float* g32(long off) const { return bufs->grads32 + off; }

    // Y = act(X W^T + b) with the usual optional outputs.  X (M, K) bf16, W (N, K) bf16 from the mirror.
    int linear(int M, int N, int K, const bf16* X, long w_off, long b_off, int act, bf16* preact, const float* residual,
               float* out32, bf16* out16, const bf16* X2 = nullptr, long w2_off = -1, long b2_off = -1) {
        GemmOp op;
        op.M = M; op.N = N; op.K1 = K;
        op.A = X; op.lda = K; op.B = p16(w_off); op.ldb = K;
        if (X2) { op.K2 = K; op.A2 = X2; op.lda2 = K; op.B2 = p16(w2_off); op.ldb2 = K; op.bias2 = p32(b2_off); }
        op.bias = p32(b_off);
        op.act = act; op.out_preact = preact; op.ldp = N;
        // bf16-output activations save act'(pre) for backward (cheap multiply there); the fp32-output head saves pre itself
        op.save_grad = (preact != nullptr && out32 == nullptr && act != 0) ? 1 : 0;
        op.residual = residual; op.ldr = N;
        op.drop = drop_next;
        drop_next = DropCfg{0u, 0u, 1.f};
        op.out_f32 = out32; op.ldo32 = N; op.out_bf16 = out16; op.ldo16 = N;
        op.max_ctas = cap;
        op.b_static = 1;     // B = weights from the bf16 mirror: last written by the previous step's optimizer
        return run_gemm(h, op, s);
    }
    // dX = dY W  [* act'(dact_src)] [+ residual].  dY (M, N) bf16, W (N, K) -> B operand stored [K=N][N=K] (MN-major)
    int grad_input(int M, int N, int K, const bf16* dY, long w_off, const bf16* dact_src, int dact, const float* residual,
                   float* out32, bf16* out16, long colsum_off = -1, long colsum_off2 = -1) {
        GemmOp op;
        op.M = M; op.N = K; op.K1 = N;
        op.A = dY; op.lda = N; op.B = p16(w_off); op.ldb = K; op.b_mn = 1;
        op.dact_src = dact_src; op.ldx = K; op.dact = dact;
        if (colsum_off >= 0) op.colsum_out = g32(colsum_off);      // bias gradient of the layer that produced dact_src
        if (colsum_off2 >= 0) op.colsum_out2 = g32(colsum_off2);
        op.residual = residual; op.ldr = K;
        op.out_f32 = out32; op.ldo32 = K; op.out_bf16 = out16; op.ldo16 = K;
        op.max_ctas = cap;
        op.b_static = 1;     // this layer's weights are only updated after both chains have left the layer
        return run_gemm(h, op, s);
    }
    // dW (N, K) += dY^T X ; db (N) += colsum(dY).   dY (M, N) bf16 with leading dimension ldy, X (M, K) bf16.
    int grad_weight(int M, int N, int K, const bf16* dY, long ldy, const bf16* X, long ldx, long w_off, long b_off) {
        GemmOp op;
        op.M = N; op.N = K; op.K1 = M;
        op.A = dY; op.lda = ldy; op.a_mn = 1; op.B = X; op.ldb = ldx; op.b_mn = 1;
        op.out_f32 = g32(w_off); op.ldo32 = K; op.accumulate = 1;
        // default: automatic tile choice (CTA-pair kernel where it exists) + split-K up to about one CTA per SM.
        // B200MART_DW_WIDE=1 forces 128x256 single-CTA tiles instead (A/B runs: pair is ~2 % better on the whole step)
        static int wide = -1;
        if (wide < 0) { const char* e = getenv("B200MART_DW_WIDE"); wide = e ? atoi(e) : 0; }
        op.bn = (wide && (long)N * K >= 1500000L) ? 256 : 128;
        // the bound only makes sense while other streams want the SMs: serial execution ("streams" = 0) uses the whole machine
        const int dwcap = h->split_streams ? dw_max_ctas() : 0;
        if (!wide) { op.bn = 0; op.split_k = pick_split(N, K, M, dwcap > 0 ? dwcap : 148); }
        else {
            const int tiles = ((N + 127) / 128) * ((K + op.bn - 1) / op.bn);
            const int kb = (M + 63) / 64;
            int split = 148 / tiles;
            if (split < 1) split = 1;
            if (split > 16) split = 16;
            if (split > kb) split = kb;
            op.split_k = split;
        }
        // weight-gradient products run on side streams next to the two dependent chains: a bounded grid keeps SMs free for
        // the chains' kernels (a resident persistent CTA is never preempted, whatever the stream priorities)
        op.max_ctas = dwcap;
        B2_TRY(run_gemm(h, op, s));
        if (b_off >= 0) B2_TRY(launch_colsum_bf16(dY, ldy, g32(b_off), M, N, s));
        return 0;
    }
};

// ------------------------------------------------------------------ two-chain stream plumbing
struct Chains {
    b200mart_handle* h; cudaStream_t main, v, t, vw, tw; bool split;   // vw / tw: weight-gradient side streams
    int rc = 0;
    // event recorded on `on` (nullptr when the chains share one stream: ordering is implicit)
    cudaEvent_t mark(cudaStream_t on) {
        if (!split) return nullptr;
        if (h->ev_used == h->evpool.size()) {
            cudaEvent_t e;
            if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { rc = 3; return nullptr; }
            h->evpool.push_back(e);
        }
        cudaEvent_t e = h->evpool[h->ev_used++];
        if (cudaEventRecord(e, on) != cudaSuccess) rc = 3;
        return e;
    }
    void wait(cudaStream_t who, cudaEvent_t e) {
        if (split && e && cudaStreamWaitEvent(who, e, 0) != cudaSuccess) rc = 3;
    }
};

int chains_begin(b200mart_handle* h, cudaStream_t s, Chains& c) {
    c.h = h; c.main = s; c.v = s; c.t = s; c.vw = s; c.tw = s; c.split = false;
    h->ev_used = 0;
    if (!h->split_streams) return 0;
    if (!h->sV) {
        // Priorities: the two dependent chains are the critical path, so their CTAs are scheduled first; the weight-gradient
        // products fill the SMs they leave idle; everything the caller runs at default priority next to them (the pipelined
        // AdamW / gradient exchange of the trainer) only gets what is left.  B200MART_STREAM_PRIO=0: all default (A/B runs).
        int least = 0, greatest = 0;
        CUDA_TRY(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        { const char* e = getenv("B200MART_STREAM_PRIO"); if (e && atoi(e) == 0) greatest = least; }
        const int mid = (greatest < least - 1) ? greatest + (least - greatest) / 2 : greatest;
        CUDA_TRY(cudaStreamCreateWithPriority(&h->sV, cudaStreamNonBlocking, greatest));
        CUDA_TRY(cudaStreamCreateWithPriority(&h->sT, cudaStreamNonBlocking, greatest));
        CUDA_TRY(cudaStreamCreateWithPriority(&h->sVw, cudaStreamNonBlocking, mid));
        CUDA_TRY(cudaStreamCreateWithPriority(&h->sTw, cudaStreamNonBlocking, mid));
    }
    c.v = h->sV; c.t = h->sT; c.vw = h->sVw; c.tw = h->sTw; c.split = true;
    cudaEvent_t e = c.mark(s);
    c.wait(c.v, e);
    c.wait(c.t, e);
    c.wait(c.vw, e);
    c.wait(c.tw, e);
    return c.rc ? b200mart_set_err(3, "stream fork failed: %s", cudaGetErrorString(cudaGetLastError())) : 0;
}

int chains_end(Chains& c) {
    if (c.split) {
        c.wait(c.main, c.mark(c.v));
        c.wait(c.main, c.mark(c.t));
        c.wait(c.main, c.mark(c.vw));
        c.wait(c.main, c.mark(c.tw));
    }
    return c.rc ? b200mart_set_err(3, "stream join failed: %s", cudaGetErrorString(cudaGetLastError())) : 0;
}

int check_batch(const b200mart_handle* h, const b200mart_batch* b) {
    B2_CHECK(b && b->B > 0 && b->L > 0, "bad batch");
    B2_CHECK(b->L <= h->d.max_position, "L=%d exceeds max_position %d", b->L, h->d.max_position);
    B2_CHECK(b->input_ids && b->attention_mask && b->pixel_values, "batch needs input_ids, attention_mask, pixel_values");
    return 0;
}

}  // namespace

// ======================================================================================= C ABI
extern "C" int32_t b200mart_create(const b200mart_model_desc* desc, b200mart_handle** out) {
    B2_CHECK(desc && out, "create: null argument");
    *out = nullptr;
    const b200mart_model_desc& d = *desc;
    B2_CHECK(d.hidden % 128 == 0 && d.hidden <= 1024, "hidden %d must be a multiple of 128 and <= 1024", d.hidden);
    B2_CHECK(d.heads > 0 && d.hidden == d.heads * 64, "head dim must be 64 (hidden %d, heads %d)", d.hidden, d.heads);
    B2_CHECK(d.intermediate % 64 == 0, "intermediate %d must be a multiple of 64", d.intermediate);
    B2_CHECK(d.layers > 0 && d.fuse_from >= 1, "bad layer count / fuse_from");
    B2_CHECK(d.patch > 0 && d.image_size % d.patch == 0, "patch %d must divide image size %d", d.patch, d.image_size);
    b200mart_handle* h = new b200mart_handle();
    h->d = d;
    const int g = d.image_size / d.patch;
    h->P = g * g;
    h->Nv = 1 + 2 * h->P;
    h->patch_dim = 3 * d.patch * d.patch;
    h->patch_ld = (int)align_up(h->patch_dim, 8);
    { const char* e = getenv("B200MART_STREAMS"); if (e && atoi(e) == 0) h->split_streams = false; }
    build_layout(h);
    // AdamW segments (one per parameter, covering its padded extent)
    std::vector<AdamSeg> segs[2];
    for (int mode = 0; mode < 2; ++mode)
        for (size_t i = 0; i < h->params.size(); ++i) {
            const PInfo& p = h->params[i];
            AdamSeg sg;
            sg.begin = p.offset;
            sg.end = (i + 1 < h->params.size()) ? h->params[i + 1].offset : h->lay.total;
            sg.decay = p.decay;
            sg.skip = p.unused_ft || (mode == 1 && p.name.find("adaptive_weight") != std::string::npos);
            segs[mode].push_back(sg);
        }
    h->nseg = (int)segs[0].size();
    h->segs_host[0] = segs[0];
    h->segs_host[1] = segs[1];
    *out = h;
    return 0;
}

extern "C" void b200mart_destroy(b200mart_handle* h) { delete h; }

extern "C" int32_t b200mart_param_count(const b200mart_handle* h) { return h ? (int32_t)h->params.size() : 0; }
extern "C" int64_t b200mart_param_total(const b200mart_handle* h) { return h ? h->lay.total : 0; }
extern "C" int32_t b200mart_param_info(const b200mart_handle* h, int32_t i, const char** name, int64_t* offset, int32_t* ndim,
                                       int64_t shape[4], int32_t* decay, int32_t* unused_ft) {
    B2_CHECK(h && i >= 0 && i < (int)h->params.size(), "param_info: index %d out of range", i);
    const PInfo& p = h->params[i];
    if (name) *name = p.name.c_str();
    if (offset) *offset = p.offset;
    if (ndim) *ndim = p.ndim;
    if (shape) for (int k = 0; k < 4; ++k) shape[k] = p.shape[k];
    if (decay) *decay = p.decay;
    if (unused_ft) *unused_ft = p.unused_ft;
    return 0;
}

extern "C" int32_t b200mart_layer_extent(const b200mart_handle* h, int32_t layer, int64_t* begin, int64_t* end) {
    B2_CHECK(h && begin && end && layer >= 0 && layer < (int)h->layer_begin.size(), "layer_extent: layer %d out of range", layer);
    *begin = h->layer_begin[layer];
    *end = h->layer_end[layer];
    return 0;
}

extern "C" int32_t b200mart_set_option(b200mart_handle* h, const char* name, int32_t value) {
    B2_CHECK(h && name, "set_option: null argument");
    if (strcmp(name, "streams") == 0) { h->split_streams = value != 0; return 0; }
    if (strcmp(name, "training") == 0) { h->training = value != 0; return 0; }
    return b200mart_set_err(1, "set_option: unknown option '%s'", name);
}

extern "C" int32_t b200mart_set_dropout(b200mart_handle* h, float p_hidden, float p_attn, uint32_t seed) {
    B2_CHECK(h && p_hidden >= 0.f && p_hidden < 1.f && p_attn >= 0.f && p_attn < 1.f, "set_dropout: probabilities must be in [0, 1)");
    h->p_hidden = p_hidden; h->p_attn = p_attn; h->seed0 = seed;
    return 0;
}

extern "C" uint32_t b200mart_last_dropout_seed(const b200mart_handle* h) { return h ? h->cur_seed : 0u; }

extern "C" int32_t b200mart_cast_params(const b200mart_handle* h, const float* p32, void* p16, void* stream) {
    B2_CHECK(h && p32 && p16, "cast_params: null argument");
    return launch_cast_bf16(p32, reinterpret_cast<bf16*>(p16), h->lay.total, (cudaStream_t)stream);
}

extern "C" int32_t b200mart_cast_params_range(const b200mart_handle* h, const float* p32, void* p16, int64_t begin, int64_t end,
                                              void* stream) {
    B2_CHECK(h && p32 && p16 && begin >= 0 && begin <= end && end <= h->lay.total, "cast_params_range: bad arguments [%ld, %ld)",
             (long)begin, (long)end);
    if (end == begin) return 0;
    return launch_cast_bf16(p32 + begin, reinterpret_cast<bf16*>(p16) + begin, end - begin, (cudaStream_t)stream);
}

extern "C" int32_t b200mart_multimem_add(const float* src, float* mc_dst, int64_t n, float scale, void* stream) {
    B2_CHECK(src && mc_dst && n >= 0, "multimem_add: bad arguments");
    return launch_multimem_add(src, mc_dst, (long)n, scale, (cudaStream_t)stream);
}

extern "C" int64_t b200mart_workspace_bytes(const b200mart_handle* h, int32_t B, int32_t L) {
    if (!h || B <= 0 || L <= 0) return 0;
    Workspace w;
    carve(h, B, L, nullptr, w);
    return (int64_t)w.bytes;
}

extern "C" int32_t b200mart_status(b200mart_handle* h, const b200mart_buffers* bufs, void* stream, int32_t* flags) {
    B2_CHECK(h && bufs && flags && h->ws_base == bufs->workspace, "status: workspace not bound (run forward first)");
    int v = 0;
    CUDA_TRY(cudaMemcpyAsync(&v, h->ws.status, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    *flags = v & 15;
    return 0;
}

extern "C" int32_t b200mart_read_hidden(b200mart_handle* h, const b200mart_buffers* bufs, int32_t stream_kind, int32_t layer, float* out,
                                        void* stream) {
    B2_CHECK(h && bufs && out && h->ws_base == bufs->workspace && h->ws.B > 0, "read_hidden: run forward on this workspace first");
    B2_CHECK((stream_kind == 0 || stream_kind == 1) && layer >= 0 && layer <= h->d.layers, "read_hidden: bad stream kind %d / layer %d",
             stream_kind, layer);
    const Workspace& w = h->ws;
    const size_t rows = (size_t)w.B * (stream_kind == 0 ? w.L : h->Nv);
    const float* src = stream_kind == 0 ? w.T32[layer] : w.V32[layer];
    CUDA_TRY(cudaMemcpyAsync(out, src, sizeof(float) * rows * h->d.hidden, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

// --------------------------------------------------------------------------------------- forward
extern "C" int32_t b200mart_forward(b200mart_handle* h, const b200mart_buffers* bufs, const b200mart_batch* batch, float* z_out,
                                    float* logits_out, void* stream) {
    B2_CHECK(h && bufs && bufs->params32 && bufs->params16, "forward: null handle / parameter buffers");
    B2_TRY(check_batch(h, batch));
    const b200mart_model_desc& d = h->d;
    const int B = batch->B, L = batch->L, H = d.hidden, F = d.intermediate, nh = d.heads, Nv = h->Nv, P = h->P;
    const int Mt = B * L, Mv = B * Nv, D = d.layers;
    B2_TRY(bind_workspace(h, bufs, B, L));
    Workspace& w = h->ws;
    const ParamLayout& l = h->lay;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(w.status, 0, 8 * sizeof(int), s));
    Chains ch;
    B2_TRY(chains_begin(h, s, ch));
    const cudaStream_t sv = ch.v, stx = ch.t;          // vision chain / text chain
    Run rv{h, bufs, sv}, rt{h, bufs, stx};
    if (ch.split) { rv.cap = chain_cap(0); rt.cap = chain_cap(1); }
    std::vector<cudaEvent_t> ev_tqkv(D, nullptr), ev_vout(D, nullptr);
    // dropout state of THIS forward (backward regenerates the same masks)
    h->cur_ph = h->training ? h->p_hidden : 0.f;
    h->cur_pa = h->training ? h->p_attn : 0.f;
    ++h->fwd_count;
    h->cur_seed = hash32(h->seed0 ^ hash32((uint32_t)h->fwd_count) ^ (uint32_t)(h->fwd_count >> 32));
    const float ph = h->cur_ph, pa = h->cur_pa;
    const uint32_t seed = h->cur_seed;
    const uint32_t* salt = h->step_block;

    // ---- vision embeddings: im2col -> patch GEMM -> (+cls, +pos) -> pre-LN     (modeling_unimo.py:119-132, :711)
    B2_TRY(launch_im2col(batch->pixel_values, w.patches16, B, d.image_size, d.patch, h->patch_ld, sv));
    {
        const bf16* wp = rv.p16(l.patch_w);
        if (h->patch_ld != h->patch_dim) {   // e.g. ViT-L/14: 588-element rows are not 16-byte multiples -> padded copy
            B2_TRY(launch_cast_pad_bf16(rv.p32(l.patch_w), h->patch_dim, w.wp16, h->patch_ld, H, h->patch_dim, sv));
            wp = w.wp16;
        }
        GemmOp op;
        op.M = B * 2 * P; op.N = H; op.K1 = h->patch_dim;
        op.A = w.patches16; op.lda = h->patch_ld; op.B = wp; op.ldb = h->patch_ld;
        op.out_f32 = w.patch_out; op.ldo32 = H;
        B2_TRY(run_gemm(h, op, sv));
    }
    B2_TRY(launch_vision_embed_ln(w.patch_out, rv.p32(l.cls), rv.p32(l.vpos), rv.p32(l.preln_w), rv.p32(l.preln_b), d.vision_eps,
                                  w.vx0, w.V32[0], w.vmean0, w.vrstd0, B, P, H, sv));
    // ---- text embeddings (:152-186)
    B2_TRY(launch_text_embed_ln(batch->input_ids, batch->token_type_ids, rt.p32(l.word), rt.p32(l.tpos), rt.p32(l.ttype),
                                rt.p32(l.eln_w), rt.p32(l.eln_b), d.text_eps, w.tpre, w.T32[0], w.T16[0], w.emean, w.erstd, B, L,
                                H, d.vocab, d.type_vocab, w.status, make_drop(ph, seed, 0, salt), stx));

    for (int i = 0; i < D; ++i) {
        const bool fused = i >= d.fuse_from;
        // ================= vision layer i (CLIPEncoderLayer :490-527)
        {
            const VisLayerP& p = l.vis[i];
            VisLayerA& a = w.vl[i];
            B2_TRY(launch_layernorm_fwd(w.V32[i], rv.p32(p.ln1_w), rv.p32(p.ln1_b), d.vision_eps, nullptr, a.u16, a.mean1, a.rstd1, Mv, H, sv));
            B2_TRY(rv.linear(Mv, 3 * H, H, a.u16, p.qkv_w, p.qkv_b, 0, nullptr, nullptr, nullptr, a.qkv16));
            AttnFwd at{};
            at.q = a.qkv16; at.ldq = 3 * H;
            if (fused) {   // K/V of text layer i-1 prepended (:227-229, :616)
                at.k1 = w.tl[i - 1].qkv16 + H; at.v1 = w.tl[i - 1].qkv16 + 2 * H; at.ld1 = 3 * H; at.n1 = L;
            }
            at.k2 = a.qkv16 + H; at.v2 = a.qkv16 + 2 * H; at.ld2 = 3 * H; at.n2 = Nv;
            at.B = B; at.Lq = Nv; at.nh = nh; at.scale = 0.125f;
            at.out = a.ctx16; at.ldo = H; at.lse = a.lse;
            if (fused) ch.wait(sv, ev_tqkv[i - 1]);     // needs K/V of text layer i-1
            { ProfScope ps(h, sv, CAT_ATTN); B2_TRY(launch_attention_fwd(at, nullptr, sv)); }
            B2_TRY(rv.linear(Mv, H, H, a.ctx16, p.o_w, p.o_b, 0, nullptr, w.V32[i], a.x1, nullptr));
            B2_TRY(launch_layernorm_fwd(a.x1, rv.p32(p.ln2_w), rv.p32(p.ln2_b), d.vision_eps, nullptr, a.u2_16, a.mean2, a.rstd2, Mv, H, sv));
            B2_TRY(rv.linear(Mv, F, H, a.u2_16, p.fc1_w, p.fc1_b, 2, a.pre16, nullptr, nullptr, a.h16));
            B2_TRY(rv.linear(Mv, H, F, a.h16, p.fc2_w, p.fc2_b, 0, nullptr, a.x1, w.V32[i + 1], nullptr));
            if (i + 1 >= d.fuse_from || i + 1 == D) ev_vout[i] = ch.mark(sv);
        }
        // ================= text layer i (BertLayer :540-577)
        {
            const TxtLayerP& p = l.txt[i];
            TxtLayerA& a = w.tl[i];
            B2_TRY(rt.linear(Mt, 3 * H, H, w.T16[i], p.qkv_w, p.qkv_b, 0, nullptr, nullptr, nullptr, a.qkv16));
            if (i + 1 >= d.fuse_from && i + 1 < D) ev_tqkv[i] = ch.mark(stx);
            AttnFwd at{};
            at.q = a.qkv16; at.ldq = 3 * H;
            at.k2 = a.qkv16 + H; at.v2 = a.qkv16 + 2 * H; at.ld2 = 3 * H; at.n2 = L;
            at.B = B; at.Lq = L; at.nh = nh; at.scale = 0.125f;
            at.key_mask = batch->attention_mask;
            if (batch->sep_idx) { at.sep_idx = batch->sep_idx + 2; at.sep_stride = 6; at.w0 = rt.p32(p.aw0); at.w1 = rt.p32(p.aw1); }
            at.out = a.ctx16; at.ldo = H; at.lse = a.lse;
            at.drop = make_drop(pa, seed, 1 + 3 * i, salt);
            { ProfScope ps(h, stx, CAT_ATTN); B2_TRY(launch_attention_fwd(at, fused ? a.ctx32 : nullptr, stx)); }
            if (fused) ch.wait(stx, ev_vout[i]);       // needs the vision output of this layer index
            if (fused)   // BertFusion on the context with the vision output of THIS layer index (:373, :627)
                { ProfScope ps(h, stx, CAT_FUSION); B2_TRY(launch_fusion_fwd(a.ctx32, w.V32[i + 1], a.g16, a.flse, B, L, Nv, H, stx)); }
            rt.drop_next = make_drop(ph, seed, 2 + 3 * i, salt);
            B2_TRY(rt.linear(Mt, H, H, a.ctx16, p.ao_w, p.ao_b, 0, nullptr, w.T32[i], a.apre, nullptr));
            B2_TRY(launch_layernorm_fwd(a.apre, rt.p32(p.aln_w), rt.p32(p.aln_b), d.text_eps, a.a32, a.a16, a.amean, a.arstd, Mt, H, stx));
            if (fused)
                B2_TRY(rt.linear(Mt, F, H, a.a16, p.i_w, p.i_b, 1, a.ipre16, nullptr, nullptr, a.i16, a.g16, p.f_w, p.f_b));
            else
                B2_TRY(rt.linear(Mt, F, H, a.a16, p.i_w, p.i_b, 1, a.ipre16, nullptr, nullptr, a.i16));
            rt.drop_next = make_drop(ph, seed, 3 + 3 * i, salt);
            B2_TRY(rt.linear(Mt, H, F, a.i16, p.o_w, p.o_b, 0, nullptr, a.a32, a.opre, nullptr));
            B2_TRY(launch_layernorm_fwd(a.opre, rt.p32(p.oln_w), rt.p32(p.oln_b), d.text_eps, w.T32[i + 1], w.T16[i + 1], a.omean, a.orstd, Mt, H, stx));
        }
    }
    // ---- MLM head transform (:972-976): z = LN(gelu(T W^T + b)).  (text_pooler :748 is never consumed: skipped.)
    B2_TRY(rt.linear(Mt, H, H, w.T16[D], l.tr_w, l.tr_b, 1, w.hpre16, nullptr, w.hact, nullptr));
    B2_TRY(launch_layernorm_fwd(w.hact, rt.p32(l.trln_w), rt.p32(l.trln_b), d.text_eps, w.z32, w.z16, w.hmean, w.hrstd, Mt, H, stx));
    if (z_out) CUDA_TRY(cudaMemcpyAsync(z_out, w.z32, sizeof(float) * Mt * H, cudaMemcpyDeviceToDevice, stx));
    if (logits_out) {   // dense compat head: logits = z E^T + b over the whole vocabulary (:958)
        GemmOp op;
        op.M = Mt; op.N = d.vocab; op.K1 = H;
        op.A = w.z16; op.lda = H; op.B = rt.p16(l.word); op.ldb = H;
        op.bias = rt.p32(l.dec_b);
        op.out_f32 = logits_out; op.ldo32 = d.vocab;
        B2_TRY(run_gemm(h, op, stx));
    }
    B2_TRY(chains_end(ch));
    return 0;
}

// --------------------------------------------------------------------------------------- score head + loss
extern "C" int32_t b200mart_head_loss(b200mart_handle* h, const b200mart_buffers* bufs, const b200mart_batch* batch,
                                      const b200mart_head* hd, float* loss_out, float* mask_logits_out, int64_t* ranks_out,
                                      int32_t compute_grad, void* stream) {
    B2_CHECK(h && bufs && hd, "head_loss: null argument");
    B2_TRY(check_batch(h, batch));
    B2_CHECK(h->ws_base == bufs->workspace && h->ws.B == batch->B && h->ws.L == batch->L, "head_loss: run forward on this batch first");
    B2_CHECK(!compute_grad || bufs->grads32, "head_loss: gradients requested without a gradient buffer");
    B2_CHECK(hd->label, "head_loss: labels missing");
    const b200mart_model_desc& d = h->d;
    const int B = batch->B, L = batch->L, H = d.hidden;
    Workspace& w = h->ws;
    const ParamLayout& l = h->lay;
    cudaStream_t s = (cudaStream_t)stream;
    Run r{h, bufs, s};
    const long Vp = align_up(d.vocab, 8);

    CUDA_TRY(cudaMemsetAsync(w.loss, 0, 8 * sizeof(float), s));
    B2_TRY(launch_find_mask(batch->input_ids, w.mask_pos, B, L, hd->mask_token_id, w.status, s));
    B2_TRY(launch_gather_mask_rows(w.z16, w.mask_pos, w.zm16, B, L, H, s));
    if (compute_grad) CUDA_TRY(cudaMemsetAsync(w.dz, 0, sizeof(float) * B * L * H, s));

    struct Group { int C; const bf16* E; const float* bias; int select; long col0; bool gathered; long out_off; };
    Group groups[2];
    int ng = 0;
    if (hd->mode == 0) {
        B2_CHECK(hd->col_ids && hd->n_cols > 0 && hd->n_cols <= d.vocab, "head_loss: fine-tune mode needs col_ids");
        B2_CHECK(hd->rel_idx && hd->q_head_idx && hd->a_head_idx, "head_loss: fine-tune mode needs rel_idx, q_head_idx, a_head_idx");
        B2_TRY(launch_gather_rows_bf16(r.p16(l.word), hd->col_ids, w.esel16, hd->n_cols, H, d.vocab, w.status, s));
        B2_TRY(launch_gather_f32(r.p32(l.dec_b), hd->col_ids, w.bias_sel, hd->n_cols, d.vocab, s));
        groups[ng++] = Group{hd->n_cols, w.esel16, w.bias_sel, 0, 0, true, 0};
    } else {
        B2_CHECK(hd->pre_type, "head_loss: pre-train mode needs pre_type");
        B2_CHECK(0 <= hd->ent_st && hd->ent_st < hd->ent_ed && hd->ent_ed <= d.vocab && 0 <= hd->rel_st && hd->rel_st < hd->rel_ed &&
                     hd->rel_ed <= d.vocab, "head_loss: bad entity / relation id ranges");
        groups[ng++] = Group{hd->ent_ed - hd->ent_st, r.p16(l.word) + (long)hd->ent_st * H, r.p32(l.dec_b) + hd->ent_st, 1, hd->ent_st, false, 0};
        groups[ng++] = Group{hd->rel_ed - hd->rel_st, r.p16(l.word) + (long)hd->rel_st * H, r.p32(l.dec_b) + hd->rel_st, 2, hd->rel_st, false,
                             (long)(hd->ent_ed - hd->ent_st)};
    }
    const long ld_out = (hd->mode == 0) ? hd->n_cols : (long)(hd->ent_ed - hd->ent_st) + (hd->rel_ed - hd->rel_st);
    if (compute_grad) CUDA_TRY(cudaMemsetAsync(w.dzm, 0, sizeof(float) * B * H, s));
    for (int gi = 0; gi < ng; ++gi) {
        const Group& g = groups[gi];
        const long Cp = align_up(g.C, 8);
        {   // logits of the [MASK] rows over this column group (K11 of SURVEY.md §2)
            GemmOp op;
            op.M = B; op.N = g.C; op.K1 = H;
            op.A = w.zm16; op.lda = H; op.B = g.E; op.ldb = H;
            op.bias = g.bias;
            op.out_f32 = w.logits; op.ldo32 = Cp;
            B2_TRY(run_gemm(h, op, s));
        }
        if (mask_logits_out)
            CUDA_TRY(cudaMemcpy2DAsync(mask_logits_out + g.out_off, ld_out * sizeof(float), w.logits, Cp * sizeof(float),
                                       g.C * sizeof(float), B, cudaMemcpyDeviceToDevice, s));
        if (ranks_out)
            B2_TRY(launch_ranks(w.logits, Cp, g.C, hd->label, hd->pre_type, g.select, ranks_out, B, s));
        LossGroup lg{};
        lg.logits = w.logits; lg.ld = Cp; lg.C = g.C;
        lg.dlogits16 = compute_grad ? w.dlogits16 : nullptr; lg.ldd = Cp;
        lg.select = g.select;
        B2_TRY(launch_ls_ce(lg, hd->label, hd->pre_type, hd->label_smoothing, w.loss, B, w.status, s, compute_grad ? hd->grad_scale : nullptr));
        if (compute_grad) {
            // d zm += dlogits E_sel        (M = B, N = H, K = C; B operand stored [K = C][N = H])
            GemmOp op;
            op.M = B; op.N = H; op.K1 = g.C;
            op.A = w.dlogits16; op.lda = Cp; op.B = g.E; op.ldb = H; op.b_mn = 1;
            op.out_f32 = w.dzm; op.ldo32 = H; op.accumulate = 1; op.split_k = pick_split(B, H, g.C);
            B2_TRY(run_gemm(h, op, s));
            // d E_sel = dlogits^T zm       (M = C, N = H, K = B)
            GemmOp ow;
            ow.M = g.C; ow.N = H; ow.K1 = B;
            ow.A = w.dlogits16; ow.lda = Cp; ow.a_mn = 1; ow.B = w.zm16; ow.ldb = H; ow.b_mn = 1;
            if (g.gathered) {
                ow.out_f32 = w.desel; ow.ldo32 = H;
                B2_TRY(run_gemm(h, ow, s));
                B2_TRY(launch_scatter_add_rows_f32(w.desel, hd->col_ids, r.g32(l.word), g.C, H, d.vocab, s));
                CUDA_TRY(cudaMemsetAsync(w.dbias_sel, 0, sizeof(float) * g.C, s));
                B2_TRY(launch_colsum_bf16(w.dlogits16, Cp, w.dbias_sel, B, g.C, s));
                B2_TRY(launch_scatter_add_f32(w.dbias_sel, hd->col_ids, r.g32(l.dec_b), g.C, d.vocab, s));
            } else {
                ow.out_f32 = r.g32(l.word) + g.col0 * H; ow.ldo32 = H; ow.accumulate = 1;
                B2_TRY(run_gemm(h, ow, s));
                B2_TRY(launch_colsum_bf16(w.dlogits16, Cp, r.g32(l.dec_b) + g.col0, B, g.C, s));
            }
        }
    }
    if (hd->mode == 0)
        B2_TRY(launch_relaxation(w.z32, hd->rel_idx, hd->q_head_idx, hd->a_head_idx, hd->alpha, w.loss, compute_grad ? w.dz : nullptr,
                                 B, L, H, w.status, s, compute_grad ? hd->grad_scale : nullptr));
    if (compute_grad) B2_TRY(launch_scatter_mask_rows(w.dzm, w.mask_pos, w.dz, B, L, H, s));
    if (loss_out) CUDA_TRY(cudaMemcpyAsync(loss_out, w.loss, sizeof(float), cudaMemcpyDeviceToDevice, s));
    return 0;
}

// --------------------------------------------------------------------------------------- backward
extern "C" int32_t b200mart_backward(b200mart_handle* h, const b200mart_buffers* bufs, const b200mart_batch* batch,
                                     int32_t use_head_grad, const float* dz_extra, const float* dlogits, void* dlogits_scratch,
                                     void* const* layer_events, void* stream) {
    B2_CHECK(h && bufs && bufs->params32 && bufs->params16 && bufs->grads32, "backward: null handle / buffers");
    B2_TRY(check_batch(h, batch));
    B2_CHECK(h->ws_base == bufs->workspace && h->ws.B == batch->B && h->ws.L == batch->L, "backward: run forward on this batch first");
    B2_CHECK(use_head_grad || dz_extra || dlogits, "backward: no incoming gradient");
    B2_CHECK(!dlogits || dlogits_scratch, "backward: dense dlogits need a bf16 scratch buffer");
    const b200mart_model_desc& d = h->d;
    const int B = batch->B, L = batch->L, H = d.hidden, F = d.intermediate, nh = d.heads, Nv = h->Nv, P = h->P;
    const int Mt = B * L, Mv = B * Nv, D = d.layers;
    Workspace& w = h->ws;
    const ParamLayout& l = h->lay;
    cudaStream_t s = (cudaStream_t)stream;
    Run r{h, bufs, s};

    // ---- gradient wrt z (trans_hidden_states)
    float* dz = w.dz;
    if (!use_head_grad) {
        if (dz_extra) CUDA_TRY(cudaMemcpyAsync(dz, dz_extra, sizeof(float) * Mt * H, cudaMemcpyDeviceToDevice, s));
        else CUDA_TRY(cudaMemsetAsync(dz, 0, sizeof(float) * Mt * H, s));
    } else if (dz_extra) {
        B2_TRY(launch_axpy_f32(dz_extra, dz, (long)Mt * H, s));
    }
    if (dlogits) {   // dense compat head: dz += dlogits E ; dE += dlogits^T z ; db += colsum(dlogits)
        const long Vp = align_up(d.vocab, 8);
        bf16* dl16 = reinterpret_cast<bf16*>(dlogits_scratch);
        B2_TRY(launch_cast_pad_bf16(dlogits, d.vocab, dl16, Vp, Mt, d.vocab, s));
        GemmOp op;
        op.M = Mt; op.N = H; op.K1 = d.vocab;
        op.A = dl16; op.lda = Vp; op.B = r.p16(l.word); op.ldb = H; op.b_mn = 1;
        op.residual = dz; op.ldr = H; op.out_f32 = dz; op.ldo32 = H;
        B2_TRY(run_gemm(h, op, s));
        GemmOp ow;
        ow.M = d.vocab; ow.N = H; ow.K1 = Mt;
        ow.A = dl16; ow.lda = Vp; ow.a_mn = 1; ow.B = w.z16; ow.ldb = H; ow.b_mn = 1;
        ow.out_f32 = r.g32(l.word); ow.ldo32 = H; ow.accumulate = 1;
        B2_TRY(run_gemm(h, ow, s));
        B2_TRY(launch_colsum_bf16(dl16, Vp, r.g32(l.dec_b), Mt, d.vocab, s));
    }
    // ---- head transform backward: z = LN(gelu(T W^T + b))
    B2_TRY(launch_layernorm_bwd(dz, w.hact, w.hmean, w.hrstd, r.p32(l.trln_w), nullptr, w.hpre16, 1, nullptr, w.t16a[0],
                                r.g32(l.trln_w), r.g32(l.trln_b), r.g32(l.tr_b), Mt, H, s));
    B2_TRY(r.grad_weight(Mt, H, H, w.t16a[0], H, w.T16[D], H, l.tr_w, -1));
    B2_TRY(r.grad_input(Mt, H, H, w.t16a[0], l.tr_w, nullptr, 0, nullptr, w.dT, nullptr));

    const bool any_fusion = D > d.fuse_from;
    // fork: text chain / vision chain (see b200mart_forward) plus one weight-gradient side stream per chain: dW = dY^T X
    // (and the remaining bias column sums) only feed the optimizer, so they leave the dX critical path.  dY scratch
    // buffers alternate by layer parity; a chain re-uses a buffer only after the side stream has consumed it.
    Chains ch;
    B2_TRY(chains_begin(h, s, ch));
    const cudaStream_t sv = ch.v, stx = ch.t, svw = ch.vw, stw = ch.tw;
    Run rv{h, bufs, sv}, rt{h, bufs, stx}, rvw{h, bufs, svw}, rtw{h, bufs, stw};
    if (ch.split) { rv.cap = chain_cap(0); rt.cap = chain_cap(1); }
    std::vector<cudaEvent_t> ev_vdone(D + 1, nullptr), ev_tfus(D, nullptr), ev_tattn(D, nullptr);
    std::vector<cudaEvent_t> ev_tw_done(D + 2, nullptr), ev_vw_done(D + 2, nullptr);
    const float ph = h->cur_ph, pa = h->cur_pa;      // dropout of the forward whose activations are in the workspace
    const uint32_t seed = h->cur_seed;
    const uint32_t* salt = h->step_block;
    const DropCfg nodrop{0u, 0u, 1.f};
    for (int i = D - 1; i >= 0; --i) {
        const bool fused = i >= d.fuse_from;
        const int pb = i & 1;                         // parity of this layer's dY scratch set
        // ================= text layer i backward
        {
            const TxtLayerP& p = l.txt[i];
            TxtLayerA& a = w.tl[i];
            bf16 *t16a = w.t16a[pb], *t16b = w.t16b[pb], *tF16 = w.tF16[pb], *tdqkv16 = w.tdqkv16[pb];
            ch.wait(stx, ev_tw_done[i + 2]);          // the side stream is done with this parity's buffers
            // output LayerNorm
            B2_TRY(launch_layernorm_bwd(w.dT, a.opre, a.omean, a.orstd, rt.p32(p.oln_w), nullptr, nullptr, 0, w.t32a, t16a,
                                        rt.g32(p.oln_w), rt.g32(p.oln_b), rt.g32(p.o_b), Mt, H, stx, nodrop, make_drop(ph, seed, 3 + 3 * i, salt)));
            ch.wait(stw, ch.mark(stx));
            B2_TRY(rtw.grad_weight(Mt, H, F, t16a, H, a.i16, F, p.o_w, -1));
            // d(pre-GELU) = (dY Wout) * gelu'(pre); its column sums are the intermediate (and fusion_dense) bias gradients
            B2_TRY(rt.grad_input(Mt, H, F, t16a, p.o_w, a.ipre16, 3, nullptr, nullptr, tF16, p.i_b, fused ? p.f_b : -1));
            ch.wait(stw, ch.mark(stx));
            B2_TRY(rtw.grad_weight(Mt, F, H, tF16, F, a.a16, H, p.i_w, -1));
            if (fused) B2_TRY(rtw.grad_weight(Mt, F, H, tF16, F, a.g16, H, p.f_w, -1));
            B2_TRY(rt.grad_input(Mt, F, H, tF16, p.i_w, nullptr, 0, w.t32a, w.t32b, nullptr));        // d a32 (incl. residual)
            if (fused) B2_TRY(rt.grad_input(Mt, F, H, tF16, p.f_w, nullptr, 0, nullptr, w.t32c, nullptr));  // d fusion output
            // attention-output LayerNorm
            B2_TRY(launch_layernorm_bwd(w.t32b, a.apre, a.amean, a.arstd, rt.p32(p.aln_w), nullptr, nullptr, 0, w.t32a, t16b,
                                        rt.g32(p.aln_w), rt.g32(p.aln_b), rt.g32(p.ao_b), Mt, H, stx, nodrop, make_drop(ph, seed, 2 + 3 * i, salt)));
            ch.wait(stw, ch.mark(stx));
            B2_TRY(rtw.grad_weight(Mt, H, H, t16b, H, a.ctx16, H, p.ao_w, -1));
            B2_TRY(rt.grad_input(Mt, H, H, t16b, p.ao_w, nullptr, 0, nullptr, w.t32b, nullptr));      // d ctx
            // the rows pass of the fusion backward only needs this chain's own tensors; dV (gradient wrt the vision output of
            // layer i, which the keys pass adds into) and dk1/dv1 come from vision layer i+1's backward: wait as late as possible
            if (fused)
                { ProfScope ps(h, stx, CAT_FUSION); B2_TRY(launch_fusion_bwd_rows(w.t32c, a.ctx32, w.V32[i + 1], a.flse, w.pbuf, w.dsbuf, w.t32b, B, L, Nv, H, stx)); }
            if (i + 1 < D && i + 1 >= d.fuse_from) ch.wait(stx, ev_vdone[i + 1]);
            if (fused) {
                { ProfScope ps(h, stx, CAT_FUSION); B2_TRY(launch_fusion_bwd_keys(w.t32c, a.ctx32, w.pbuf, w.dsbuf, (i == D - 1) ? nullptr : w.dV[pb], w.dV[pb], w.dV16[pb], B, L, Nv, H, stx)); }
            }
            if (fused) ev_tfus[i] = ch.mark(stx);      // dV / dV16 now include this layer's fusion contribution
            AttnBwd ab{};
            AttnFwd& at = ab.f;
            at.q = a.qkv16; at.ldq = 3 * H;
            at.k2 = a.qkv16 + H; at.v2 = a.qkv16 + 2 * H; at.ld2 = 3 * H; at.n2 = L;
            at.B = B; at.Lq = L; at.nh = nh; at.scale = 0.125f;
            at.key_mask = batch->attention_mask;
            if (batch->sep_idx) { at.sep_idx = batch->sep_idx + 2; at.sep_stride = 6; at.w0 = rt.p32(p.aw0); at.w1 = rt.p32(p.aw1); ab.dw0 = rt.g32(p.aw0); ab.dw1 = rt.g32(p.aw1); }
            at.lse = a.lse;
            at.drop = make_drop(pa, seed, 1 + 3 * i, salt);
            ab.dout = w.t32b; ab.lddo = H;
            ab.dq = tdqkv16; ab.dk2 = tdqkv16 + H; ab.dv2 = tdqkv16 + 2 * H; ab.ldd = 3 * H;
            if (i + 1 < D && i + 1 >= d.fuse_from) { ab.dk2_add = w.dk1; ab.dv2_add = w.dv1; ab.ldadd = H; }   // from vision layer i+1
            { ProfScope ps(h, stx, CAT_ATTN); B2_TRY(launch_attention_bwd(ab, a.ctx16, H, w.delta, stx)); }
            if (ab.dk2_add) ev_tattn[i] = ch.mark(stx);   // dk1/dv1 consumed: vision layer i may overwrite them
            ch.wait(stw, ch.mark(stx));
            B2_TRY(rtw.grad_weight(Mt, 3 * H, H, tdqkv16, 3 * H, w.T16[i], H, p.qkv_w, p.qkv_b));
            B2_TRY(rt.grad_input(Mt, 3 * H, H, tdqkv16, p.qkv_w, nullptr, 0, w.t32a, w.dT, nullptr));  // d T^i
            ev_tw_done[i] = ch.mark(stw);
        }
        // ================= vision layer i backward (only reached by gradient through the fusion / KV-prefix paths)
        if (any_fusion) {
            const VisLayerP& p = l.vis[i];
            VisLayerA& a = w.vl[i];
            const int qb = (i + 1) & 1;               // parity of the gradient buffer this layer hands to layer i-1
            float* dVin = w.dV[pb]; bf16* dVin16 = w.dV16[pb];
            bf16 *v16a = w.v16a[pb], *vF16 = w.vF16[pb], *vdqkv16 = w.vdqkv16[pb];
            if (fused) ch.wait(sv, ev_tfus[i]);
            ch.wait(sv, ev_vw_done[i + 2]);
            // fc2 bias gradient = column sums of dV: fused into the LayerNorm-1 backward of layer i+1 when no text-layer
            // fusion adds into dV afterwards (i < fuse_from), a column-sum pass otherwise
            ch.wait(svw, ch.mark(sv));
            B2_TRY(rvw.grad_weight(Mv, H, F, dVin16, H, a.h16, F, p.fc2_w, fused ? p.fc2_b : -1));
            B2_TRY(rv.grad_input(Mv, H, F, dVin16, p.fc2_w, a.pre16, 3, nullptr, nullptr, vF16, p.fc1_b));
            ch.wait(svw, ch.mark(sv));
            B2_TRY(rvw.grad_weight(Mv, F, H, vF16, F, a.u2_16, H, p.fc1_w, -1));
            B2_TRY(rv.grad_input(Mv, F, H, vF16, p.fc1_w, nullptr, 0, nullptr, w.v32a, nullptr));
            B2_TRY(launch_layernorm_bwd(w.v32a, a.x1, a.mean2, a.rstd2, rv.p32(p.ln2_w), dVin, nullptr, 0, w.v32b, v16a,
                                        rv.g32(p.ln2_w), rv.g32(p.ln2_b), rv.g32(p.o_b), Mv, H, sv));
            ch.wait(svw, ch.mark(sv));
            B2_TRY(rvw.grad_weight(Mv, H, H, v16a, H, a.ctx16, H, p.o_w, -1));
            B2_TRY(rv.grad_input(Mv, H, H, v16a, p.o_w, nullptr, 0, nullptr, w.v32a, nullptr));
            AttnBwd ab{};
            AttnFwd& at = ab.f;
            at.q = a.qkv16; at.ldq = 3 * H;
            if (fused) { at.k1 = w.tl[i - 1].qkv16 + H; at.v1 = w.tl[i - 1].qkv16 + 2 * H; at.ld1 = 3 * H; at.n1 = L; }
            at.k2 = a.qkv16 + H; at.v2 = a.qkv16 + 2 * H; at.ld2 = 3 * H; at.n2 = Nv;
            at.B = B; at.Lq = Nv; at.nh = nh; at.scale = 0.125f;
            at.lse = a.lse;
            ab.dout = w.v32a; ab.lddo = H;
            ab.dq = vdqkv16; ab.dk2 = vdqkv16 + H; ab.dv2 = vdqkv16 + 2 * H; ab.ldd = 3 * H;
            if (fused) { ab.dk1 = w.dk1; ab.dv1 = w.dv1; ab.ldd1 = H; ch.wait(sv, ev_tattn[i]); }
            { ProfScope ps(h, sv, CAT_ATTN); B2_TRY(launch_attention_bwd(ab, a.ctx16, H, w.delta_v, sv)); }
            ch.wait(svw, ch.mark(sv));
            B2_TRY(rvw.grad_weight(Mv, 3 * H, H, vdqkv16, 3 * H, a.u16, H, p.qkv_w, p.qkv_b));
            B2_TRY(rv.grad_input(Mv, 3 * H, H, vdqkv16, p.qkv_w, nullptr, 0, nullptr, w.v32a, nullptr));
            ev_vw_done[i] = ch.mark(svw);
            ch.wait(sv, ev_vw_done[i + 1]);           // dV16[qb] was layer i+1's input: its dW(fc2) must have read it
            B2_TRY(launch_layernorm_bwd(w.v32a, w.V32[i], a.mean1, a.rstd1, rv.p32(p.ln1_w), w.v32b, nullptr, 0, w.dV[qb], w.dV16[qb],
                                        rv.g32(p.ln1_w), rv.g32(p.ln1_b),
                                        (i >= 1 && i - 1 < d.fuse_from) ? rv.g32(l.vis[i - 1].fc2_b) : nullptr, Mv, H, sv));
            ev_vdone[i] = ch.mark(sv);
        }
        if (layer_events) {   // layer i's gradients are final once both chains AND both side streams pass this point
            ch.wait(stw, ch.mark(stx));
            CUDA_TRY(cudaEventRecord((cudaEvent_t)layer_events[2 * i], stw));
            if (any_fusion) ch.wait(svw, ch.mark(sv));
            CUDA_TRY(cudaEventRecord((cudaEvent_t)layer_events[2 * i + 1], any_fusion ? svw : stw));
        }
    }
    // ---- embeddings
    B2_TRY(launch_layernorm_bwd(w.dT, w.tpre, w.emean, w.erstd, rt.p32(l.eln_w), nullptr, nullptr, 0, w.t32a, nullptr,
                                rt.g32(l.eln_w), rt.g32(l.eln_b), nullptr, Mt, H, stx, make_drop(ph, seed, 0, salt), nodrop));
    B2_TRY(launch_text_embed_bwd(w.t32a, batch->input_ids, batch->token_type_ids, rt.g32(l.word), rt.g32(l.tpos), rt.g32(l.ttype), B, L, H, d.vocab, d.type_vocab, stx));
    if (any_fusion) {
        B2_TRY(launch_layernorm_bwd(w.dV[1], w.vx0, w.vmean0, w.vrstd0, rv.p32(l.preln_w), nullptr, nullptr, 0, w.v32a, nullptr,
                                    rv.g32(l.preln_w), rv.g32(l.preln_b), nullptr, Mv, H, sv));
        B2_TRY(launch_vision_embed_bwd(w.v32a, rv.g32(l.cls), rv.g32(l.vpos), w.dpatch16, B, P, H, sv));
        B2_TRY(rv.grad_weight(B * 2 * P, H, h->patch_dim, w.dpatch16, H, w.patches16, h->patch_ld, l.patch_w, -1));
    }
    B2_TRY(chains_end(ch));
    return 0;
}

// --------------------------------------------------------------------------------------- optimizer
extern "C" int32_t b200mart_adamw(b200mart_handle* h, float* params32, void* params16, float* grads32, float* exp_avg,
                                  float* exp_avg_sq, int32_t step, float lr, float beta1, float beta2, float eps,
                                  float weight_decay, int32_t pretrain, int64_t begin, int64_t end, int32_t zero_grads,
                                  void* stream) {
    B2_CHECK(h && params32 && grads32 && exp_avg && exp_avg_sq && step >= 1, "adamw: bad arguments");
    if (end <= 0) end = h->lay.total;
    B2_CHECK(begin >= 0 && begin < end && end <= h->lay.total, "adamw: bad range [%ld, %ld)", (long)begin, (long)end);
    const int mode = pretrain ? 1 : 0;
    if (!h->segs_dev[mode]) {   // tiny one-off table: parameter extents + decay / skip flags
        CUDA_TRY(cudaMalloc(&h->segs_dev[mode], sizeof(AdamSeg) * h->nseg));
        CUDA_TRY(cudaMemcpy(h->segs_dev[mode], h->segs_host[mode].data(), sizeof(AdamSeg) * h->nseg, cudaMemcpyHostToDevice));
    }
    const double bc1d = 1.0 - pow((double)beta1, (double)step), bc2d = 1.0 - pow((double)beta2, (double)step);
    const float bc1 = (float)bc1d, bc2 = (float)bc2d;
    return launch_adamw(params32, grads32, exp_avg, exp_avg_sq, reinterpret_cast<bf16*>(params16), h->segs_dev[mode],
                        h->nseg, begin, end, zero_grads, lr, beta1, beta2, eps, weight_decay, bc1, bc2, (cudaStream_t)stream,
                        reinterpret_cast<const float*>(h->step_block));
}

extern "C" int32_t b200mart_set_step_block(b200mart_handle* h, void* device_block) {
    B2_CHECK(h, "set_step_block: null handle");
    B2_CHECK(((uintptr_t)device_block & 15) == 0, "set_step_block: the block must be 16-byte aligned");
    h->step_block = reinterpret_cast<uint32_t*>(device_block);
    return 0;
}

extern "C" int32_t b200mart_adamw_multicast(b200mart_handle* h, const float* params32, float* exp_avg, float* exp_avg_sq,
                                            float* grads_mc, float* params32_mc, void* params16_mc, int32_t step, float lr,
                                            float beta1, float beta2, float eps, float weight_decay, int32_t pretrain, int64_t begin,
                                            int64_t end, int32_t world, int32_t max_ctas, void* stream) {
    B2_CHECK(h && params32 && exp_avg && exp_avg_sq && grads_mc && params32_mc && params16_mc && step >= 1 && world >= 1,
             "adamw_multicast: bad arguments");
    B2_CHECK(begin >= 0 && begin < end && end <= h->lay.total, "adamw_multicast: bad range [%ld, %ld)", (long)begin, (long)end);
    const int mode = pretrain ? 1 : 0;
    if (!h->segs_dev[mode]) {
        CUDA_TRY(cudaMalloc(&h->segs_dev[mode], sizeof(AdamSeg) * h->nseg));
        CUDA_TRY(cudaMemcpy(h->segs_dev[mode], h->segs_host[mode].data(), sizeof(AdamSeg) * h->nseg, cudaMemcpyHostToDevice));
    }
    const double bc1d = 1.0 - pow((double)beta1, (double)step), bc2d = 1.0 - pow((double)beta2, (double)step);
    return launch_adamw_mc(params32, exp_avg, exp_avg_sq, grads_mc, params32_mc, reinterpret_cast<bf16*>(params16_mc), h->segs_dev[mode],
                           h->nseg, begin, end, 1.0f / (float)world, lr, beta1, beta2, eps, weight_decay, (float)bc1d, (float)bc2d,
                           (cudaStream_t)stream, max_ctas);
}

// --------------------------------------------------------------------------------------- profiling
extern "C" double b200mart_profile_gemm_bytes(const b200mart_handle* h) { return h ? h->prof_bytes : 0.0; }

extern "C" int32_t b200mart_profile_begin(b200mart_handle* h) {
    B2_CHECK(h, "profile_begin: null handle");
    h->profiling = true;
    h->prof_used = 0;
    h->prof_flops = 0.0;
    h->prof_bytes = 0.0;
    return 0;
}

extern "C" int32_t b200mart_profile_end(b200mart_handle* h, void* stream, double ms_by_cat[3], int64_t launches_by_cat[3],
                                        double* gemm_flops) {
    B2_CHECK(h && ms_by_cat && launches_by_cat && gemm_flops, "profile_end: null argument");
    h->profiling = false;
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    for (int c = 0; c < CAT_COUNT; ++c) { ms_by_cat[c] = 0.0; launches_by_cat[c] = 0; }
    for (size_t i = 0; i < h->prof_used; ++i) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->prof[i].e0, h->prof[i].e1));
        ms_by_cat[h->prof[i].cat] += ms;
        launches_by_cat[h->prof[i].cat] += 1;
    }
    *gemm_flops = h->prof_flops;
    return 0;
}


// ======================================================================================= pre-LN encoder stack
// Generic pre-LayerNorm transformer stack = the block FLAVA and ViLT are made of (and the CLIP layers above):
//   FlavaLayer   MarT/models/modeling_flava.py:620-665  (FlavaSelfAttention :441-521 with adaptive scaling on rows 1:s)
//   ViltLayer    MarT/models/modeling_vilt.py:491-528   (ViltSelfAttention :329-400, rows img_length+1 : s')
//   h = x + drop(Wo attn(LN1 x) + bo);  y = h + drop(W2 act(W1 LN2 h + b1) + b2)
// attn: S = QK^T / 8, adaptive block scaling (rows [row_lo, s) x cols >= s by clamp(w0,0,.5); rows >= s by clamp(w1,.5,1))
// BEFORE the additive -1e4 key mask, softmax, dropout, PV.  Same kernels as the MKGformer path.
struct b200mart_stack {
    b200mart_handle core;      // plan cache / profiling / stream plumbing of the GEMM runner
    b200mart_stack_desc d;
    struct LayerP { long ln1_w, ln1_b, qkv_w, qkv_b, aw0, aw1, o_w, o_b, ln2_w, ln2_b, fc1_w, fc1_b, fc2_w, fc2_b; };
    struct LayerA { bf16 *u16, *qkv16, *ctx16, *u2_16, *pre16, *h16; float *mean1, *rstd1, *lse, *x1, *mean2, *rstd2;
                    float *h1 = nullptr, *x2 = nullptr; };     // post-LN only: LayerNorm(x1) in fp32 (residual of the MLP), pre-LN sum of the MLP
    std::vector<LayerP> lay;
    // workspace carving
    int B = 0, T = 0; void* ws_base = nullptr; size_t ws_bytes = 0;
    std::vector<float*> X;     // layers + 1 residual streams
    std::vector<LayerA> act;
    float *s32a, *s32b, *dX; bf16 *dX16, *s16a, *sF16, *sdqkv16; float* delta;
};

namespace {

void stack_carve(b200mart_stack* st, int B, int T, void* base) {
    const size_t H = st->d.hidden, F = st->d.intermediate, nh = st->d.heads, M = (size_t)B * T;
    Carver c{reinterpret_cast<uint8_t*>(base)};
    st->X.assign(st->d.layers + 1, nullptr);
    for (auto& p : st->X) p = c.take<float>(M * H);
    st->act.assign(st->d.layers, b200mart_stack::LayerA());
    for (auto& a : st->act) {
        a.u16 = c.take<bf16>(M * H); a.mean1 = c.take<float>(M); a.rstd1 = c.take<float>(M);
        a.qkv16 = c.take<bf16>(M * 3 * H); a.lse = c.take<float>((size_t)B * nh * T);
        a.ctx16 = c.take<bf16>(M * H); a.x1 = c.take<float>(M * H);
        a.u2_16 = c.take<bf16>(M * H); a.mean2 = c.take<float>(M); a.rstd2 = c.take<float>(M);
        a.pre16 = c.take<bf16>(M * F); a.h16 = c.take<bf16>(M * F);
        if (st->d.post_ln) { a.h1 = c.take<float>(M * H); a.x2 = c.take<float>(M * H); }
    }
    st->s32a = c.take<float>(M * H); st->s32b = c.take<float>(M * H); st->dX = c.take<float>(M * H);
    st->dX16 = c.take<bf16>(M * H); st->s16a = c.take<bf16>(M * H); st->sF16 = c.take<bf16>(M * F);
    st->sdqkv16 = c.take<bf16>(M * 3 * H); st->delta = c.take<float>((size_t)B * nh * T);
    st->B = B; st->T = T; st->ws_base = base; st->ws_bytes = c.off + 256;
}

int stack_bind(b200mart_stack* st, const b200mart_buffers* bufs, int B, int T) {
    B2_CHECK(bufs && bufs->workspace && bufs->params32 && bufs->params16, "stack: null buffers");
    if (st->ws_base != bufs->workspace || st->B != B || st->T != T) stack_carve(st, B, T, bufs->workspace);
    B2_CHECK((long)st->ws_bytes <= bufs->workspace_bytes, "stack: workspace too small: need %zu bytes, have %ld", st->ws_bytes,
             (long)bufs->workspace_bytes);
    return 0;
}

void stack_attn(const b200mart_stack* st, const b200mart_stack_io* io, const Run& r, int i, AttnFwd& at) {
    const int H = st->d.hidden;
    const b200mart_stack::LayerA& a = st->act[i];
    at.q = a.qkv16; at.ldq = 3 * H;
    at.k2 = a.qkv16 + H; at.v2 = a.qkv16 + 2 * H; at.ld2 = 3 * H; at.n2 = io->T;
    at.B = io->B; at.Lq = io->T; at.nh = st->d.heads; at.scale = 0.125f;
    at.key_mask = io->key_mask;
    if (io->split) {
        at.sep_idx = io->split; at.sep_stride = 1; at.row_lo = io->row_lo;
        at.w0 = r.p32(st->lay[i].aw0); at.w1 = r.p32(st->lay[i].aw1);
    }
    at.lse = a.lse;
    at.drop = make_drop(io->p_attn, io->seed, 1 + 3 * i);
}

}  // namespace

extern "C" int32_t b200mart_stack_create(const b200mart_stack_desc* desc, b200mart_stack** out) {
    B2_CHECK(desc && out, "stack_create: null argument");
    *out = nullptr;
    const b200mart_stack_desc& d = *desc;
    B2_CHECK(d.hidden % 128 == 0 && d.hidden <= 1024 && d.heads > 0 && d.hidden == d.heads * 64,
             "stack: hidden %d must be heads * 64, a multiple of 128 and <= 1024", d.hidden);
    B2_CHECK(d.intermediate % 64 == 0 && d.layers > 0 && (d.act == 1 || d.act == 2), "stack: bad intermediate / layers / act");
    b200mart_stack* st = new b200mart_stack();
    st->core.split_streams = false;     // one stream: nothing to share the SMs with, weight-gradient GEMMs use the whole machine
    st->d = d;
    b200mart_handle* h = &st->core;
    const long H = d.hidden, F = d.intermediate;
    char buf[160];
    for (int i = 0; i < d.layers; ++i) {
        b200mart_stack::LayerP p;
        auto nm = [&](const char* s) { snprintf(buf, sizeof(buf), "layer.%d.%s", i, s); return std::string(buf); };
        // pre-LN: ViT-style names (FLAVA / ViLT); post-LN: BERT-style names (VisualBERT)
        const std::string sa = d.post_ln ? "attention.self." : "attention.attention.";
        p.qkv_w = add_param(h, nm((sa + "query.weight").c_str()), {H, H});
        add_param(h, nm((sa + "key.weight").c_str()), {H, H});
        add_param(h, nm((sa + "value.weight").c_str()), {H, H});
        p.qkv_b = add_param(h, nm((sa + "query.bias").c_str()), {H});
        add_param(h, nm((sa + "key.bias").c_str()), {H});
        add_param(h, nm((sa + "value.bias").c_str()), {H});
        p.aw0 = add_param(h, nm((sa + "adaptive_weight.0").c_str()), {1});
        p.aw1 = add_param(h, nm((sa + "adaptive_weight.1").c_str()), {1});
        p.o_w = add_param(h, nm("attention.output.dense.weight"), {H, H});
        p.o_b = add_param(h, nm("attention.output.dense.bias"), {H});
        p.fc1_w = add_param(h, nm("intermediate.dense.weight"), {F, H});
        p.fc1_b = add_param(h, nm("intermediate.dense.bias"), {F});
        p.fc2_w = add_param(h, nm("output.dense.weight"), {H, F});
        p.fc2_b = add_param(h, nm("output.dense.bias"), {H});
        p.ln1_w = add_param(h, nm(d.post_ln ? "attention.output.LayerNorm.weight" : "layernorm_before.weight"), {H});
        p.ln1_b = add_param(h, nm(d.post_ln ? "attention.output.LayerNorm.bias" : "layernorm_before.bias"), {H});
        p.ln2_w = add_param(h, nm(d.post_ln ? "output.LayerNorm.weight" : "layernorm_after.weight"), {H});
        p.ln2_b = add_param(h, nm(d.post_ln ? "output.LayerNorm.bias" : "layernorm_after.bias"), {H});
        st->lay.push_back(p);
    }
    *out = st;
    return 0;
}

extern "C" void b200mart_stack_destroy(b200mart_stack* st) { delete st; }
extern "C" int32_t b200mart_stack_param_count(const b200mart_stack* st) { return st ? (int32_t)st->core.params.size() : 0; }
extern "C" int64_t b200mart_stack_param_total(const b200mart_stack* st) { return st ? st->core.lay.total : 0; }
extern "C" int32_t b200mart_stack_param_info(const b200mart_stack* st, int32_t i, const char** name, int64_t* offset, int32_t* ndim,
                                             int64_t shape[4]) {
    B2_CHECK(st, "stack_param_info: null handle");
    return b200mart_param_info(&st->core, i, name, offset, ndim, shape, nullptr, nullptr);
}
extern "C" int64_t b200mart_stack_workspace_bytes(const b200mart_stack* st, int32_t B, int32_t T) {
    if (!st || B <= 0 || T <= 0) return 0;
    b200mart_stack tmp;
    tmp.d = st->d;
    stack_carve(&tmp, B, T, nullptr);
    return (int64_t)tmp.ws_bytes;
}

extern "C" int32_t b200mart_stack_forward(b200mart_stack* st, const b200mart_buffers* bufs, const b200mart_stack_io* io, void* stream) {
    B2_CHECK(st && io && io->x && io->y && io->B > 0 && io->T > 0, "stack_forward: bad arguments");
    B2_TRY(stack_bind(st, bufs, io->B, io->T));
    const b200mart_stack_desc& d = st->d;
    const int H = d.hidden, F = d.intermediate, M = io->B * io->T;
    cudaStream_t s = (cudaStream_t)stream;
    Run r{&st->core, bufs, s};
    CUDA_TRY(cudaMemcpyAsync(st->X[0], io->x, sizeof(float) * M * H, cudaMemcpyDeviceToDevice, s));
    if (d.post_ln) {
        // BertLayer order (VisualBertLayer, MarT/models/modeling_visual_bert.py:381-421): x1 = x + drop(Wo attn(x) + bo); h1 = LN(x1);
        // x2 = h1 + drop(W2 act(W1 h1 + b1) + b2); y = LN(x2)
        B2_TRY(launch_cast_bf16(io->x, st->act[0].u16, (long)M * H, s));
        for (int i = 0; i < d.layers; ++i) {
            const b200mart_stack::LayerP& p = st->lay[i];
            b200mart_stack::LayerA& a = st->act[i];
            B2_TRY(r.linear(M, 3 * H, H, a.u16, p.qkv_w, p.qkv_b, 0, nullptr, nullptr, nullptr, a.qkv16));
            AttnFwd at{};
            stack_attn(st, io, r, i, at);
            at.out = a.ctx16; at.ldo = H;
            B2_TRY(launch_attention_fwd(at, nullptr, s));
            r.drop_next = make_drop(io->p_hidden, io->seed, 2 + 3 * i);
            B2_TRY(r.linear(M, H, H, a.ctx16, p.o_w, p.o_b, 0, nullptr, st->X[i], a.x1, nullptr));
            B2_TRY(launch_layernorm_fwd(a.x1, r.p32(p.ln1_w), r.p32(p.ln1_b), d.eps, a.h1, a.u2_16, a.mean1, a.rstd1, M, H, s));
            B2_TRY(r.linear(M, F, H, a.u2_16, p.fc1_w, p.fc1_b, d.act, a.pre16, nullptr, nullptr, a.h16));
            r.drop_next = make_drop(io->p_hidden, io->seed, 3 + 3 * i);
            B2_TRY(r.linear(M, H, F, a.h16, p.fc2_w, p.fc2_b, 0, nullptr, a.h1, a.x2, nullptr));
            B2_TRY(launch_layernorm_fwd(a.x2, r.p32(p.ln2_w), r.p32(p.ln2_b), d.eps, st->X[i + 1], i + 1 < d.layers ? st->act[i + 1].u16 : st->s16a,
                                        a.mean2, a.rstd2, M, H, s));
        }
        CUDA_TRY(cudaMemcpyAsync(io->y, st->X[d.layers], sizeof(float) * M * H, cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    for (int i = 0; i < d.layers; ++i) {
        const b200mart_stack::LayerP& p = st->lay[i];
        b200mart_stack::LayerA& a = st->act[i];
        B2_TRY(launch_layernorm_fwd(st->X[i], r.p32(p.ln1_w), r.p32(p.ln1_b), d.eps, nullptr, a.u16, a.mean1, a.rstd1, M, H, s));
        B2_TRY(r.linear(M, 3 * H, H, a.u16, p.qkv_w, p.qkv_b, 0, nullptr, nullptr, nullptr, a.qkv16));
        AttnFwd at{};
        stack_attn(st, io, r, i, at);
        at.out = a.ctx16; at.ldo = H;
        B2_TRY(launch_attention_fwd(at, nullptr, s));
        r.drop_next = make_drop(io->p_hidden, io->seed, 2 + 3 * i);
        B2_TRY(r.linear(M, H, H, a.ctx16, p.o_w, p.o_b, 0, nullptr, st->X[i], a.x1, nullptr));
        B2_TRY(launch_layernorm_fwd(a.x1, r.p32(p.ln2_w), r.p32(p.ln2_b), d.eps, nullptr, a.u2_16, a.mean2, a.rstd2, M, H, s));
        B2_TRY(r.linear(M, F, H, a.u2_16, p.fc1_w, p.fc1_b, d.act, a.pre16, nullptr, nullptr, a.h16));
        r.drop_next = make_drop(io->p_hidden, io->seed, 3 + 3 * i);
        B2_TRY(r.linear(M, H, F, a.h16, p.fc2_w, p.fc2_b, 0, nullptr, a.x1, st->X[i + 1], nullptr));
    }
    CUDA_TRY(cudaMemcpyAsync(io->y, st->X[d.layers], sizeof(float) * M * H, cudaMemcpyDeviceToDevice, s));
    return 0;
}

extern "C" int32_t b200mart_stack_backward(b200mart_stack* st, const b200mart_buffers* bufs, const b200mart_stack_io* io, const float* dy,
                                           float* dx, void* stream) {
    B2_CHECK(st && io && dy && bufs && bufs->grads32, "stack_backward: bad arguments");
    B2_CHECK(st->ws_base == bufs->workspace && st->B == io->B && st->T == io->T, "stack_backward: run stack_forward on this input first");
    const b200mart_stack_desc& d = st->d;
    const int H = d.hidden, F = d.intermediate, M = io->B * io->T, D = d.layers;
    cudaStream_t s = (cudaStream_t)stream;
    Run r{&st->core, bufs, s};
    const DropCfg nodrop{0u, 0u, 1.f};
    if (d.post_ln) {
        CUDA_TRY(cudaMemcpyAsync(st->dX, dy, sizeof(float) * M * H, cudaMemcpyDeviceToDevice, s));
        for (int i = D - 1; i >= 0; --i) {
            const b200mart_stack::LayerP& p = st->lay[i];
            b200mart_stack::LayerA& a = st->act[i];
            // output LayerNorm: d x2 (fp32: the residual branch into h1) + its masked bf16 copy (the fc2 branch) + fc2 bias gradient
            B2_TRY(launch_layernorm_bwd(st->dX, a.x2, a.mean2, a.rstd2, r.p32(p.ln2_w), nullptr, nullptr, 0, st->s32a, st->dX16,
                                        r.g32(p.ln2_w), r.g32(p.ln2_b), r.g32(p.fc2_b), M, H, s, nodrop, make_drop(io->p_hidden, io->seed, 3 + 3 * i)));
            B2_TRY(r.grad_weight(M, H, F, st->dX16, H, a.h16, F, p.fc2_w, -1));
            B2_TRY(r.grad_input(M, H, F, st->dX16, p.fc2_w, a.pre16, 3, nullptr, nullptr, st->sF16, p.fc1_b));
            B2_TRY(r.grad_weight(M, F, H, st->sF16, F, a.u2_16, H, p.fc1_w, -1));
            B2_TRY(r.grad_input(M, F, H, st->sF16, p.fc1_w, nullptr, 0, st->s32a, st->s32b, nullptr));        // d h1 (incl. residual)
            // attention-output LayerNorm: d x1 (fp32: the residual branch into x) + masked bf16 copy (the Wo branch) + bo gradient
            B2_TRY(launch_layernorm_bwd(st->s32b, a.x1, a.mean1, a.rstd1, r.p32(p.ln1_w), nullptr, nullptr, 0, st->s32a, st->s16a,
                                        r.g32(p.ln1_w), r.g32(p.ln1_b), r.g32(p.o_b), M, H, s, nodrop, make_drop(io->p_hidden, io->seed, 2 + 3 * i)));
            B2_TRY(r.grad_weight(M, H, H, st->s16a, H, a.ctx16, H, p.o_w, -1));
            B2_TRY(r.grad_input(M, H, H, st->s16a, p.o_w, nullptr, 0, nullptr, st->s32b, st->dX16));          // d ctx (+ bf16 copy; dX16 is free here)
            AttnBwd ab{};
            stack_attn(st, io, r, i, ab.f);
            if (io->split) { ab.dw0 = r.g32(p.aw0); ab.dw1 = r.g32(p.aw1); }
            ab.dout = st->s32b; ab.lddo = H; ab.dout16 = st->dX16; ab.lddo16 = H;
            ab.dq = st->sdqkv16; ab.dk2 = st->sdqkv16 + H; ab.dv2 = st->sdqkv16 + 2 * H; ab.ldd = 3 * H;
            B2_TRY(launch_attention_bwd(ab, a.ctx16, H, st->delta, s));
            B2_TRY(r.grad_weight(M, 3 * H, H, st->sdqkv16, 3 * H, a.u16, H, p.qkv_w, p.qkv_b));
            B2_TRY(r.grad_input(M, 3 * H, H, st->sdqkv16, p.qkv_w, nullptr, 0, st->s32a, st->dX, nullptr));     // d x (incl. residual)
        }
        if (dx) CUDA_TRY(cudaMemcpyAsync(dx, st->dX, sizeof(float) * M * H, cudaMemcpyDeviceToDevice, s));
        return 0;
    }
    // gradient wrt the stack output: fp32 copy (residual path) + masked bf16 copy / bias gradient of the top fc2
    CUDA_TRY(cudaMemcpyAsync(st->dX, dy, sizeof(float) * M * H, cudaMemcpyDeviceToDevice, s));
    B2_TRY(launch_mask_cast_colsum(dy, st->dX16, r.g32(st->lay[D - 1].fc2_b), M, H, make_drop(io->p_hidden, io->seed, 3 + 3 * (D - 1)), s));
    for (int i = D - 1; i >= 0; --i) {
        const b200mart_stack::LayerP& p = st->lay[i];
        b200mart_stack::LayerA& a = st->act[i];
        B2_TRY(r.grad_weight(M, H, F, st->dX16, H, a.h16, F, p.fc2_w, -1));
        B2_TRY(r.grad_input(M, H, F, st->dX16, p.fc2_w, a.pre16, 3, nullptr, nullptr, st->sF16, p.fc1_b));
        B2_TRY(r.grad_weight(M, F, H, st->sF16, F, a.u2_16, H, p.fc1_w, -1));
        B2_TRY(r.grad_input(M, F, H, st->sF16, p.fc1_w, nullptr, 0, nullptr, st->s32a, nullptr));
        B2_TRY(launch_layernorm_bwd(st->s32a, a.x1, a.mean2, a.rstd2, r.p32(p.ln2_w), st->dX, nullptr, 0, st->s32b, st->s16a,
                                    r.g32(p.ln2_w), r.g32(p.ln2_b), r.g32(p.o_b), M, H, s, nodrop, make_drop(io->p_hidden, io->seed, 2 + 3 * i)));
        B2_TRY(r.grad_weight(M, H, H, st->s16a, H, a.ctx16, H, p.o_w, -1));
        B2_TRY(r.grad_input(M, H, H, st->s16a, p.o_w, nullptr, 0, nullptr, st->s32a, st->dX16));    // d ctx (+ bf16 copy; dX16 is free here)
        AttnBwd ab{};
        stack_attn(st, io, r, i, ab.f);
        if (io->split) { ab.dw0 = r.g32(p.aw0); ab.dw1 = r.g32(p.aw1); }
        ab.dout = st->s32a; ab.lddo = H; ab.dout16 = st->dX16; ab.lddo16 = H;
        ab.dq = st->sdqkv16; ab.dk2 = st->sdqkv16 + H; ab.dv2 = st->sdqkv16 + 2 * H; ab.ldd = 3 * H;
        B2_TRY(launch_attention_bwd(ab, a.ctx16, H, st->delta, s));
        B2_TRY(r.grad_weight(M, 3 * H, H, st->sdqkv16, 3 * H, a.u16, H, p.qkv_w, p.qkv_b));
        B2_TRY(r.grad_input(M, 3 * H, H, st->sdqkv16, p.qkv_w, nullptr, 0, nullptr, st->s32a, nullptr));
        // gradient wrt this layer's input; its bf16 copy carries the dropout mask of the fc2 output of the layer below
        B2_TRY(launch_layernorm_bwd(st->s32a, st->X[i], a.mean1, a.rstd1, r.p32(p.ln1_w), st->s32b, nullptr, 0, st->dX, st->dX16,
                                    r.g32(p.ln1_w), r.g32(p.ln1_b), i > 0 ? r.g32(st->lay[i - 1].fc2_b) : nullptr, M, H, s, nodrop,
                                    i > 0 ? make_drop(io->p_hidden, io->seed, 3 + 3 * (i - 1)) : nodrop));
    }
    if (dx) CUDA_TRY(cudaMemcpyAsync(dx, st->dX, sizeof(float) * M * H, cudaMemcpyDeviceToDevice, s));
    return 0;
}

// --------------------------------------------------------------------------------------- stand-alone operators
extern "C" int32_t b200mart_attention_fwd(const b200mart_attn_desc* d, void* stream) {
    B2_CHECK(d, "attention_fwd: null descriptor");
    AttnFwd a{};
    a.q = (const bf16*)d->q; a.ldq = d->ldq;
    a.k1 = (const bf16*)d->prefix_k; a.v1 = (const bf16*)d->prefix_v; a.ld1 = d->ld_prefix; a.n1 = d->prefix_k ? d->n_prefix : 0;
    a.k2 = (const bf16*)d->k; a.v2 = (const bf16*)d->v; a.ld2 = d->ldkv; a.n2 = d->n_keys;
    a.B = d->B; a.Lq = d->Lq; a.nh = d->heads; a.scale = d->scale;
    a.key_mask = d->key_mask; a.sep_idx = d->sep_idx ? d->sep_idx + 2 : nullptr; a.sep_stride = 6; a.w0 = d->w0; a.w1 = d->w1;
    a.out = (bf16*)d->out; a.ldo = d->ldo; a.lse = d->lse;
    B2_CHECK(a.out && a.lse, "attention_fwd: out / lse missing");
    return launch_attention_fwd(a, d->out_f32, (cudaStream_t)stream);
}

extern "C" int32_t b200mart_attention_bwd(const b200mart_attn_desc* d, void* stream) {
    B2_CHECK(d, "attention_bwd: null descriptor");
    AttnBwd g{};
    AttnFwd& a = g.f;
    a.q = (const bf16*)d->q; a.ldq = d->ldq;
    a.k1 = (const bf16*)d->prefix_k; a.v1 = (const bf16*)d->prefix_v; a.ld1 = d->ld_prefix; a.n1 = d->prefix_k ? d->n_prefix : 0;
    a.k2 = (const bf16*)d->k; a.v2 = (const bf16*)d->v; a.ld2 = d->ldkv; a.n2 = d->n_keys;
    a.B = d->B; a.Lq = d->Lq; a.nh = d->heads; a.scale = d->scale;
    a.key_mask = d->key_mask; a.sep_idx = d->sep_idx ? d->sep_idx + 2 : nullptr; a.sep_stride = 6; a.w0 = d->w0; a.w1 = d->w1;
    a.lse = d->lse;
    g.dout = d->dout; g.lddo = d->lddo;
    g.dq = (bf16*)d->dq; g.dk2 = (bf16*)d->dk; g.dv2 = (bf16*)d->dv; g.ldd = d->ldd;
    g.dk2_add = d->dk_add; g.dv2_add = d->dv_add; g.ldadd = d->ldadd;
    g.dk1 = d->dprefix_k; g.dv1 = d->dprefix_v; g.ldd1 = d->ldd_prefix;
    g.dw0 = d->dw0; g.dw1 = d->dw1;
    B2_CHECK(!a.sep_idx || (g.dw0 && g.dw1), "attention_bwd: adaptive scaling needs dw0 / dw1");
    B2_CHECK(d->out, "attention_bwd: the forward output (out) is needed");
    return launch_attention_bwd(g, (const bf16*)d->out, d->ldo, d->delta, (cudaStream_t)stream);
}

extern "C" int32_t b200mart_fusion_fwd(const float* ctx, const float* vh, void* out_bf16, float* lse, int32_t B, int32_t L, int32_t Nv,
                                       int32_t H, void* stream) {
    B2_CHECK(ctx && vh && out_bf16 && lse, "fusion_fwd: null argument");
    return launch_fusion_fwd(ctx, vh, (bf16*)out_bf16, lse, B, L, Nv, H, (cudaStream_t)stream);
}

extern "C" int32_t b200mart_fusion_bwd(const float* dg, const float* ctx, const float* vh, const float* lse, float* pbuf, float* dsbuf,
                                       float* dctx_accum, const float* dvh_in, float* dvh_out, int32_t B, int32_t L, int32_t Nv, int32_t H,
                                       void* stream) {
    B2_CHECK(dg && ctx && vh && lse && pbuf && dsbuf && dctx_accum && dvh_out, "fusion_bwd: null argument");
    B2_TRY(launch_fusion_bwd_rows(dg, ctx, vh, lse, pbuf, dsbuf, dctx_accum, B, L, Nv, H, (cudaStream_t)stream));
    return launch_fusion_bwd_keys(dg, ctx, pbuf, dsbuf, dvh_in, dvh_out, nullptr, B, L, Nv, H, (cudaStream_t)stream);
}

extern "C" int32_t b200mart_layernorm_fwd(const float* x, const float* gamma, const float* beta, float eps, float* y32, void* y16,
                                          float* mean, float* rstd, int32_t M, int32_t H, void* stream) {
    B2_CHECK(x && gamma && beta && mean && rstd, "layernorm_fwd: null argument");
    return launch_layernorm_fwd(x, gamma, beta, eps, y32, (bf16*)y16, mean, rstd, M, H, (cudaStream_t)stream);
}

extern "C" int32_t b200mart_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                                          const float* add, float* dx32, void* dx16, float* dgamma, float* dbeta, int32_t M, int32_t H,
                                          void* stream) {
    B2_CHECK(dy && x && mean && rstd && gamma && dgamma && dbeta, "layernorm_bwd: null argument");
    return launch_layernorm_bwd(dy, x, mean, rstd, gamma, add, nullptr, 0, dx32, (bf16*)dx16, dgamma, dbeta, nullptr, M, H, (cudaStream_t)stream);
}

extern "C" int32_t b200mart_ls_ce(const float* logits, int32_t B, int32_t C, const int64_t* label, float eps, float* loss, float* dlogits,
                                  void* stream) {
    B2_CHECK(logits && label && loss, "ls_ce: null argument");
    LossGroup g{};
    g.logits = logits; g.ld = C; g.C = C; g.dlogits32 = dlogits; g.ldd = C; g.select = 0;
    return launch_ls_ce(g, label, nullptr, eps, loss, B, nullptr, (cudaStream_t)stream);
}

extern "C" int32_t b200mart_gather_images(const void* table, int32_t table_is_bf16, const int64_t* idx, float* out, int32_t n_images,
                                          int64_t image_elems, int64_t n_entities, int32_t* status, void* stream) {
    return launch_gather_images(table, table_is_bf16, idx, out, n_images, (long)image_elems, (long)n_entities, status, (cudaStream_t)stream);
}
