"""CPU ORACLE (test infrastructure, NOT a product path) — fp32 restatement of the reference MKGformer step.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import
this file.  The product (mkg_analogy_b200/) never does: it fails loudly when its CUDA library is missing.

What is restated (reference file:line at each function): the forward of `UnimoForMaskedLM`
(/root/reference/MarT/models/modeling_unimo.py:848-893 and everything it calls), the fine-tune and pre-train
losses of `TransformerLitModel.training_step` (MarT/lit_models/transformer.py:59-113) with
`LabelSmoothSoftmaxCEV1` (MarT/lit_models/utils.py:42-66), the rank computation of `_eval` (:115-166) and the
optimizer of `configure_optimizers` (:224-241).  Plain functional PyTorch, fp32, no in-place score edits;
gradients come from torch autograd over this restatement.

Pinning: the reference ships no tests or golden vectors for this path, so the oracle is
pinned against OUTPUTS OF THE REFERENCE ITSELF executed in the build container:
  * tests/test_oracle_vs_reference.py runs the reference modules from /root/reference side by side (skipped
    where the reference tree is absent, i.e. on the GPU box);
  * oracle/make_golden.py dumps reference outputs for seeded inputs into tests/golden/*.pt, and
    tests/test_oracle_golden.py checks this restatement against them everywhere.
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------- pieces
def layer_norm(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def quick_gelu(x):
    """transformers.activations.ACT2FN["quick_gelu"], used by CLIPMLP (modeling_unimo.py:275-287)."""
    return x * torch.sigmoid(1.702 * x)


def vision_embed(P, cfg, pixel_values):
    """CLIPVisionEmbeddings.forward, modeling_unimo.py:119-132 (+ vision_pre_layrnorm, :711).
    Conv with kernel = stride = patch and no bias == per-patch matmul; image 2 reuses patch positions 1..P."""
    B = pixel_values.shape[0]
    Wp = P["unimo.vision_embeddings.patch_embedding.weight"]
    pe = []
    for img in (0, 1):
        x = F.conv2d(pixel_values[:, img], Wp, stride=cfg.patch)           # (B, H, g, g)
        pe.append(x.flatten(2).transpose(1, 2))                            # (B, P, H)
    cls = P["unimo.vision_embeddings.class_embedding"].expand(B, 1, -1)
    pos = P["unimo.vision_embeddings.position_embedding.weight"]
    x = torch.cat([cls, pe[0], pe[1]], dim=1) + torch.cat([pos, pos[1:]], dim=0)[None]
    return layer_norm(x, P["unimo.vision_pre_layrnorm.weight"], P["unimo.vision_pre_layrnorm.bias"], cfg.vision_eps)


def text_embed(P, cfg, input_ids, token_type_ids, drop=None):
    """BertEmbeddings.forward, modeling_unimo.py:152-186.  `drop(site, tensor)` (optional) applies the dropout of :185
    with an explicit mask; None = eval mode."""
    L = input_ids.shape[1]
    e = (P["unimo.text_embeddings.word_embeddings.weight"][input_ids]
         + P["unimo.text_embeddings.token_type_embeddings.weight"][token_type_ids]
         + P["unimo.text_embeddings.position_embeddings.weight"][:L][None])
    out = layer_norm(e, P["unimo.text_embeddings.LayerNorm.weight"], P["unimo.text_embeddings.LayerNorm.bias"],
                     cfg.text_eps)
    return drop(0, out) if drop is not None else out


def _heads(x, nh):
    B, T, H = x.shape
    return x.view(B, T, nh, H // nh).permute(0, 2, 1, 3)


def vision_layer(P, cfg, i, x, text_kv):
    """CLIPEncoderLayer.forward :490-527 with CLIPAttention.forward :212-272 and CLIPMLP :275-287.
    text_kv = (K, V) of text layer i-1, each (B, nh, L, d), prepended on the key axis when i >= fuse_from
    (:227-229); no mask exists on this path, so padded text keys are attended."""
    p = f"unimo.encoder.vision_layers.{i}."
    nh, d = cfg.heads, cfg.head_dim
    u = layer_norm(x, P[p + "layer_norm1.weight"], P[p + "layer_norm1.bias"], cfg.vision_eps)
    q = F.linear(u, P[p + "self_attn.q_proj.weight"], P[p + "self_attn.q_proj.bias"]) * d ** -0.5
    k = F.linear(u, P[p + "self_attn.k_proj.weight"], P[p + "self_attn.k_proj.bias"])
    v = F.linear(u, P[p + "self_attn.v_proj.weight"], P[p + "self_attn.v_proj.bias"])
    q, k, v = _heads(q, nh), _heads(k, nh), _heads(v, nh)
    if text_kv is not None:
        k = torch.cat([text_kv[0], k], dim=2)
        v = torch.cat([text_kv[1], v], dim=2)
    pr = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    ctx = (pr @ v).permute(0, 2, 1, 3).reshape(x.shape)
    x = x + F.linear(ctx, P[p + "self_attn.out_proj.weight"], P[p + "self_attn.out_proj.bias"])
    u = layer_norm(x, P[p + "layer_norm2.weight"], P[p + "layer_norm2.bias"], cfg.vision_eps)
    h = quick_gelu(F.linear(u, P[p + "mlp.fc1.weight"], P[p + "mlp.fc1.bias"]))
    return x + F.linear(h, P[p + "mlp.fc2.weight"], P[p + "mlp.fc2.bias"])


# When set to a list, text_layer appends (layer, raw attention scores) with retain_grad(): the gradient of an adaptive weight
# is sum(dS * S_raw) over its block, a heavily cancelling sum whose forward-error bound is eps * sum(|dS * S_raw|).
SCORE_TAP = None


def adaptive_weight_conditioning(tap, sep_split, w0s, w1s):
    """{layer: ((dw0, sum|terms|), (dw1, sum|terms|))} from the tapped raw scores and their gradients (call after backward):
    terms = dS_scaled * S_raw = S_raw.grad * S_raw / factor over rows < s x cols >= s (w0) and rows >= s x cols >= s (w1)."""
    out = {}
    for (i, sc), w0, w1 in zip(tap, w0s, w1s):
        B, nh, L, _ = sc.shape
        ar = torch.arange(L, device=sc.device)
        s = sep_split.view(B, 1, 1, 1)
        col = ar.view(1, 1, 1, L) >= s
        row_lo = ar.view(1, 1, L, 1) < s
        res = []
        for blk, w, lo, hi in ((col & row_lo, w0, 0.0, 0.5), (col & ~row_lo, w1, 0.5, 1.0)):
            f = torch.clamp(w.detach(), lo, hi)
            t = (sc.grad * sc.detach() / f) * blk
            res.append((t.sum().item(), t.abs().sum().item()))
        out[i] = tuple(res)
    return out


def adaptive_scale(scores, sep_split, w0, w1):
    """modeling_unimo.py:342-349 without the per-sample Python loop: rows < s, cols >= s are multiplied by
    clamp(w0, 0, 0.5); rows >= s, cols >= s by clamp(w1, 0.5, 1).  Applied BEFORE the additive mask."""
    B, nh, L, _ = scores.shape
    ar = torch.arange(L, device=scores.device)
    s = sep_split.view(B, 1, 1, 1)
    col = (ar.view(1, 1, 1, L) >= s)
    row_lo = (ar.view(1, 1, L, 1) < s)
    w0c = torch.clamp(w0, 0, 0.5)
    w1c = torch.clamp(w1, 0.5, 1)
    one = torch.ones((), device=scores.device, dtype=scores.dtype)
    factor = torch.where(col & row_lo, w0c, torch.where(col & ~row_lo, w1c, one))
    return scores * factor


def text_layer(P, cfg, i, x, ext_mask, vis, sep_split, drop=None):
    """BertLayer.forward :540-577 = BertSelfAttention :317-377 (+ BertFusion :400-414 on the context when a
    visual state is given) -> BertSelfOutput :380-391 -> BertIntermediate :448-464 -> BertOutput :467-478.
    Returns (hidden, (K, V)) — the layer's K/V are what the next vision layer prepends."""
    p = f"unimo.encoder.text_layer.{i}."
    nh, d = cfg.heads, cfg.head_dim
    q = _heads(F.linear(x, P[p + "attention.self.query.weight"], P[p + "attention.self.query.bias"]), nh)
    k = _heads(F.linear(x, P[p + "attention.self.key.weight"], P[p + "attention.self.key.bias"]), nh)
    v = _heads(F.linear(x, P[p + "attention.self.value.weight"], P[p + "attention.self.value.bias"]), nh)
    sc = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    if SCORE_TAP is not None and sc.requires_grad:      # tests: raw scores + their gradient -> conditioning of d adaptive_weight
        sc.retain_grad()
        SCORE_TAP.append((i, sc))
    if sep_split is not None:
        sc = adaptive_scale(sc, sep_split, P[p + "attention.self.adaptive_weight.0"],
                            P[p + "attention.self.adaptive_weight.1"])
    sc = sc + ext_mask
    probs = torch.softmax(sc, dim=-1)
    if drop is not None:
        probs = drop(1 + 3 * i, probs)                      # :362
    ctx = (probs @ v).permute(0, 2, 1, 3).reshape(x.shape)
    fusion = None
    if vis is not None:  # BertFusion: no 1/sqrt(d), no mask
        fusion = torch.softmax(ctx @ vis.transpose(-1, -2), dim=-1) @ vis
    ao = F.linear(ctx, P[p + "attention.output.dense.weight"], P[p + "attention.output.dense.bias"])
    if drop is not None:
        ao = drop(2 + 3 * i, ao)                            # :389
    a = layer_norm(x + ao, P[p + "attention.output.LayerNorm.weight"], P[p + "attention.output.LayerNorm.bias"], cfg.text_eps)
    h = F.linear(a, P[p + "intermediate.dense.weight"], P[p + "intermediate.dense.bias"])
    if fusion is not None:
        h = h + F.linear(fusion, P[p + "intermediate.fusion_dense.weight"], P[p + "intermediate.fusion_dense.bias"])
    h = F.gelu(h)
    fo = F.linear(h, P[p + "output.dense.weight"], P[p + "output.dense.bias"])
    if drop is not None:
        fo = drop(3 + 3 * i, fo)                            # :476
    out = layer_norm(a + fo, P[p + "output.LayerNorm.weight"], P[p + "output.LayerNorm.bias"], cfg.text_eps)
    return out, (k, v)


def encoder(P, cfg, vis, txt, ext_mask, sep_split, drop=None, hidden=None):
    """UnimoEncoder.forward :589-658: interleaved loop; layer idx >= fuse_from: the vision layer gets the
    (K, V) of text layer idx-1 and the text layer fuses with the vision OUTPUT of layer idx.
    hidden: optional dict; receives the lists "text" / "vision" of the hidden states entering each layer plus the final ones
    (`output_hidden_states`, :603-647)."""
    kv = None
    if hidden is not None:
        hidden["text"], hidden["vision"] = [txt], [vis]
    for i in range(cfg.layers):
        vis = vision_layer(P, cfg, i, vis, kv if i >= cfg.fuse_from else None)
        txt, kv = text_layer(P, cfg, i, txt, ext_mask, vis if i >= cfg.fuse_from else None, sep_split, drop)
        if hidden is not None:
            hidden["text"].append(txt)
            hidden["vision"].append(vis)
    return vis, txt


class bf16_operands:
    """Context manager: every dense product of the oracle (F.linear, the patch convolution) rounds BOTH operands to bf16 and
    accumulates in fp32 - the arithmetic contract of a bf16 tensor-core implementation, everything else staying fp32.  Used by
    the parity tests to separate 'what bf16 operands cost on this model' from implementation error."""

    def __enter__(self):
        self.lin, self.conv = F.linear, F.conv2d
        r = lambda t: t.to(torch.bfloat16).to(torch.float32)
        lin, conv = self.lin, self.conv
        F.linear = lambda x, w, b=None: lin(r(x), r(w), b)
        F.conv2d = lambda x, w, *a, **k: conv(r(x), r(w), *a, **k)
        return self

    def __exit__(self, *exc):
        F.linear, F.conv2d = self.lin, self.conv
        return False


def head_transform(P, cfg, x):
    """BertPredictionHeadTransform :972-976: LN(gelu(x Wt + b)) — returned as trans_hidden_states."""
    return layer_norm(F.gelu(F.linear(x, P["cls.predictions.transform.dense.weight"],
                                      P["cls.predictions.transform.dense.bias"])),
                      P["cls.predictions.transform.LayerNorm.weight"], P["cls.predictions.transform.LayerNorm.bias"],
                      cfg.text_eps)


def forward(P, cfg, batch, full_logits=True, drop=None, hidden=None):
    """UnimoForMaskedLM.forward :848-893.  Returns (logits (B, L, V) or None, trans_hidden (B, L, H))."""
    ids, att = batch["input_ids"], batch["attention_mask"]
    tt = batch.get("token_type_ids")
    if tt is None:
        tt = torch.zeros_like(ids)
    sep = batch.get("sep_idx")
    vis = vision_embed(P, cfg, batch["pixel_values"])
    txt = text_embed(P, cfg, ids, tt, drop)
    ext_mask = (1.0 - att[:, None, None, :].to(txt.dtype)) * -10000.0        # get_extended_attention_mask :21-57
    _, txt = encoder(P, cfg, vis, txt, ext_mask, None if sep is None else sep[:, 2], drop, hidden)
    z = head_transform(P, cfg, txt)
    logits = None
    if full_logits:
        logits = F.linear(z, P["unimo.text_embeddings.word_embeddings.weight"], P["cls.predictions.bias"])
    return logits, z


def mask_positions(input_ids, mask_token_id=103):
    """(input_ids == mask).nonzero, one mask per row asserted (transformer.py:71-74, :94)."""
    rows, pos = (input_ids == mask_token_id).nonzero(as_tuple=True)
    assert rows.shape[0] == input_ids.shape[0], "exactly one [MASK] per row expected"
    return pos


def mask_row_logits(P, z, mpos, col_ids):
    """logits[arange, mask_idx][:, col_ids] (transformer.py:95) computed on the mask rows only."""
    B = z.shape[0]
    zm = z[torch.arange(B, device=z.device), mpos]                                            # (B, H)
    E = P["unimo.text_embeddings.word_embeddings.weight"][col_ids]
    return zm @ E.t() + P["cls.predictions.bias"][col_ids]


def label_smooth_ce(logits, label, eps=0.1, ignore_index=-100):
    """LabelSmoothSoftmaxCEV1.forward, lit_models/utils.py:42-66.  The target is NOT a distribution: the
    scatter overwrites, so q[label] = 1-eps and q[other] = eps/C (sum = 1 - eps/C). Mean over valid rows."""
    C = logits.shape[1]
    ignore = label == ignore_index
    n_valid = (~ignore).sum()
    lab = label.clone()
    lab[ignore] = 0
    q = torch.full_like(logits, eps / C).scatter_(1, lab[:, None], 1.0 - eps)
    logs = torch.log_softmax(logits, dim=1)
    loss = -(logs * q).sum(dim=1)
    loss = loss.masked_fill(ignore, 0.0)
    return loss.sum() / n_valid


def relaxation_loss(z, rel_idx, q_head_idx, a_head_idx):
    """transformer.py:103-108: relu(cos(z[q_head], z[a_head])) + 1 - cos(z[rel0], z[rel1]), mean over batch."""
    ar = torch.arange(z.shape[0], device=z.device)
    r0, r1 = z[ar, rel_idx[:, 0]], z[ar, rel_idx[:, 1]]
    hq, ha = z[ar, q_head_idx], z[ar, a_head_idx]
    sim = torch.relu(F.cosine_similarity(hq, ha, dim=-1)) + 1.0 - F.cosine_similarity(r0, r1, dim=-1)
    return sim.mean()


def finetune_loss(P, cfg, batch, analogy_entity_ids, alpha=0.43, label_smoothing=0.1, drop=None):
    """training_step fine-tune branch, transformer.py:92-109.  Returns (loss, mask_logits (B, C), z)."""
    _, z = forward(P, cfg, batch, full_logits=False, drop=drop)
    mpos = mask_positions(batch["input_ids"])
    cols = torch.as_tensor(analogy_entity_ids, device=z.device)
    x = mask_row_logits(P, z, mpos, cols)
    if label_smoothing != 0.0:
        ce = label_smooth_ce(x, batch["label"], label_smoothing)
    else:
        ce = F.cross_entropy(x, batch["label"])
    sim = relaxation_loss(z, batch["rel_idx"], batch["q_head_idx"], batch["a_head_idx"])
    return ce + alpha * sim, x, z


def pretrain_loss(P, cfg, batch, entity_range, relation_range, label_smoothing=0.1):
    """training_step pre-train branch, transformer.py:72-90: rows split by pre_type (2 = relation), CE over the
    contiguous entity / relation id ranges, summed."""
    _, z = forward(P, cfg, batch, full_logits=False)
    mpos = mask_positions(batch["input_ids"])
    is_rel = batch["pre_type"] == 2
    loss = z.new_zeros(())
    lf = (lambda a, b: label_smooth_ce(a, b, label_smoothing)) if label_smoothing != 0.0 else F.cross_entropy
    for sel, (st, ed) in ((~is_rel, entity_range), (is_rel, relation_range)):
        if sel.any():
            x = mask_row_logits(P, z[sel], mpos[sel], torch.arange(st, ed, device=z.device))
            loss = loss + lf(x, batch["label"][sel])
    return loss


# ----------------------------------------------------------------------------------------------- FLAVA / ViLT encoder
def adaptive_scale_rows(scores, sep_split, w0, w1, row_lo):
    """FlavaSelfAttention modeling_flava.py:491-496 (row_lo = 1) / ViltSelfAttention modeling_vilt.py:370-375
    (row_lo = img_length + 1): rows [row_lo, s) x cols >= s scaled by clamp(w0, 0, .5); rows >= s x cols >= s by
    clamp(w1, .5, 1); rows < row_lo untouched.  (modeling_unimo.py:342-349 is the row_lo = 0 case.)"""
    B, nh, T, _ = scores.shape
    ar = torch.arange(T, device=scores.device)
    s = sep_split.view(B, 1, 1, 1)
    col = ar.view(1, 1, 1, T) >= s
    row = ar.view(1, 1, T, 1)
    one = torch.ones((), device=scores.device, dtype=scores.dtype)
    factor = torch.where(col & (row >= s), torch.clamp(w1, 0.5, 1),
                         torch.where(col & (row >= row_lo) & (row < s), torch.clamp(w0, 0, 0.5), one))
    return scores * factor


def preln_stack(P, x, heads, eps=1e-12, act="gelu", attention_mask=None, sep_split=None, row_lo=1, drop=None, prefix=""):
    """FlavaEncoder (modeling_flava.py:668-722 over FlavaLayer :620-665) / ViltEncoder (modeling_vilt.py:531-583 over
    ViltLayer :491-528): pre-LayerNorm blocks.  P: {"layer.i....": tensor} with the HF parameter names; attention_mask
    (B, T) 1 = keep -> additive (1 - m) * -10000 as the callers build it; drop: optional explicit-mask dropout callback
    (sites 1+3i attention probabilities, 2+3i attention output, 3+3i MLP output)."""
    B, T, H = x.shape
    d = H // heads
    n_layers = 1 + max(int(k[len(prefix):].split(".")[1]) for k in P if k.startswith(prefix + "layer."))
    ext = None if attention_mask is None else (1.0 - attention_mask[:, None, None, :].to(x.dtype)) * -10000.0
    fact = F.gelu if act == "gelu" else quick_gelu
    for i in range(n_layers):
        p = f"{prefix}layer.{i}."
        u = layer_norm(x, P[p + "layernorm_before.weight"], P[p + "layernorm_before.bias"], eps)
        q = _heads(F.linear(u, P[p + "attention.attention.query.weight"], P[p + "attention.attention.query.bias"]), heads)
        k = _heads(F.linear(u, P[p + "attention.attention.key.weight"], P[p + "attention.attention.key.bias"]), heads)
        v = _heads(F.linear(u, P[p + "attention.attention.value.weight"], P[p + "attention.attention.value.bias"]), heads)
        sc = (q @ k.transpose(-1, -2)) / math.sqrt(d)
        if sep_split is not None:
            sc = adaptive_scale_rows(sc, sep_split, P[p + "attention.attention.adaptive_weight.0"],
                                     P[p + "attention.attention.adaptive_weight.1"], row_lo)
        if ext is not None:
            sc = sc + ext
        pr = torch.softmax(sc, dim=-1)
        if drop is not None:
            pr = drop(1 + 3 * i, pr)
        ctx = (pr @ v).permute(0, 2, 1, 3).reshape(B, T, H)
        ao = F.linear(ctx, P[p + "attention.output.dense.weight"], P[p + "attention.output.dense.bias"])
        if drop is not None:
            ao = drop(2 + 3 * i, ao)
        h = x + ao
        u2 = layer_norm(h, P[p + "layernorm_after.weight"], P[p + "layernorm_after.bias"], eps)
        f = fact(F.linear(u2, P[p + "intermediate.dense.weight"], P[p + "intermediate.dense.bias"]))
        fo = F.linear(f, P[p + "output.dense.weight"], P[p + "output.dense.bias"])
        if drop is not None:
            fo = drop(3 + 3 * i, fo)
        x = h + fo
    return x


def postln_stack(P, x, heads, eps=1e-12, attention_mask=None, sep_split=None, row_lo=1, drop=None, prefix=""):
    """VisualBertEncoder (modeling_visual_bert.py:424-488 over VisualBertLayer :381-421, VisualBertSelfAttention :203-286,
    VisualBertSelfOutput :289-300, VisualBertOutput :367-378): post-LayerNorm BERT blocks with the adaptive scaling of :255-260
    (rows img_length+1 : s' -> row_lo = img_length + 1).  P: {"layer.i....": tensor} with the HF names; attention_mask (B, T) 1 = keep;
    drop: explicit-mask dropout callback, sites as in preln_stack."""
    B, T, H = x.shape
    d = H // heads
    n_layers = 1 + max(int(k[len(prefix):].split(".")[1]) for k in P if k.startswith(prefix + "layer."))
    ext = None if attention_mask is None else (1.0 - attention_mask[:, None, None, :].to(x.dtype)) * -10000.0
    for i in range(n_layers):
        p = f"{prefix}layer.{i}."
        q = _heads(F.linear(x, P[p + "attention.self.query.weight"], P[p + "attention.self.query.bias"]), heads)
        k = _heads(F.linear(x, P[p + "attention.self.key.weight"], P[p + "attention.self.key.bias"]), heads)
        v = _heads(F.linear(x, P[p + "attention.self.value.weight"], P[p + "attention.self.value.bias"]), heads)
        sc = (q @ k.transpose(-1, -2)) / math.sqrt(d)
        if sep_split is not None:
            sc = adaptive_scale_rows(sc, sep_split, P[p + "attention.self.adaptive_weight.0"], P[p + "attention.self.adaptive_weight.1"], row_lo)
        if ext is not None:
            sc = sc + ext
        pr = torch.softmax(sc, dim=-1)
        if drop is not None:
            pr = drop(1 + 3 * i, pr)
        ctx = (pr @ v).permute(0, 2, 1, 3).reshape(B, T, H)
        ao = F.linear(ctx, P[p + "attention.output.dense.weight"], P[p + "attention.output.dense.bias"])
        if drop is not None:
            ao = drop(2 + 3 * i, ao)
        h = layer_norm(x + ao, P[p + "attention.output.LayerNorm.weight"], P[p + "attention.output.LayerNorm.bias"], eps)
        f = F.gelu(F.linear(h, P[p + "intermediate.dense.weight"], P[p + "intermediate.dense.bias"]))
        fo = F.linear(f, P[p + "output.dense.weight"], P[p + "output.dense.bias"])
        if drop is not None:
            fo = drop(3 + 3 * i, fo)
        x = layer_norm(h + fo, P[p + "output.LayerNorm.weight"], P[p + "output.LayerNorm.bias"], eps)
    return x


def hash_dropout(seed, p_hidden, p_attn, salt=0):
    """drop(site, tensor) callback reproducing the CUDA path's stateless dropout masks (csrc/common.cuh: keep element e of
    site s iff hash32(e ^ hash32(seed + 0x9E3779B9 (s + 1))) >= p 2^32; kept values scaled by 1/(1-p) in fp32), so the
    training-mode graph can be checked against the oracle mask for mask.  Sites: 0 embeddings, 1+3i attention probabilities
    (p_attn), 2+3i / 3+3i sub-layer outputs.  salt: the per-step word of b200mart_set_step_block (XORed into every site seed)."""
    import numpy as np
    M = 0xFFFFFFFF

    def h32(x):
        x = x ^ (x >> 16)
        x = (x * 0x7feb352d) & M
        x = x ^ (x >> 15)
        x = (x * 0x846ca68b) & M
        return x ^ (x >> 16)

    def drop(site, t):
        p = np.float32(p_attn if (site >= 1 and (site - 1) % 3 == 0) else p_hidden)
        if p <= 0:
            return t
        thresh = int(np.float64(p) * 4294967296.0)
        scale = float(np.float32(1.0) / (np.float32(1.0) - p))
        site_seed = int(h32(torch.tensor((seed + 0x9E3779B9 * (site + 1)) & M, dtype=torch.int64))) ^ (salt & M)
        idx = torch.arange(t.numel(), dtype=torch.int64, device=t.device)
        keep = h32(idx ^ site_seed) >= thresh
        return t * (keep.to(t.dtype) * scale).view(t.shape)
    return drop


def ranks_of_gold(x, label):
    """_eval, transformer.py:162-164: argsort(argsort(-x))[gold] + 1 == 1 + #{c: x_c > x_gold} when there are no
    ties."""
    gold = x.gather(1, label[:, None])
    return (x > gold).sum(dim=1) + 1


def adamw_step(params, grads, state, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
    """torch.optim.AdamW as configured at transformer.py:224-241: decay 0.01 except names containing "bias" or
    "LayerNorm.weight" (so CLIP layer_norm1/2 and all embeddings ARE decayed). In-place on `params`."""
    b1, b2 = betas
    for name, p in params.items():
        g = grads.get(name)
        if g is None:
            continue
        wd = 0.0 if ("bias" in name or "LayerNorm.weight" in name) else weight_decay
        m, v = state.setdefault(name, (torch.zeros_like(p), torch.zeros_like(p)))
        p.mul_(1.0 - lr * wd)
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m, denom, value=-lr / bc1)


def linear_warmup_lr(step, total_steps, base_lr, warmup_ratio=0.1):
    """get_linear_schedule_with_warmup as used at transformer.py:233-240 (stepped every batch)."""
    warm = total_steps * warmup_ratio        # the reference passes the float product (transformer.py:234)
    if step < warm:
        return base_lr * float(step) / float(max(1, warm))
    return base_lr * max(0.0, float(total_steps - step) / float(max(1, total_steps - warm)))


def collate_images(entity2visual_features, entities, head_ent, tail_ent):
    """Image half of the reference collator, MarT/data/data_module.py:127-160 (tensor branch = MKGformer / FLAVA / ViLT):
    per sample two (C, S, S) tensors — `entity2visual_features[entities.index(name)]`, or zeros for a missing entity —
    stacked to (B, 2, C, S, S)."""
    pixel_images = []
    for head, tail in zip(head_ent, tail_ent):
        pixel_image = []
        if head and tail:                                            # :129-132
            pixel_image.append(entity2visual_features[entities.index(head)])
            pixel_image.append(entity2visual_features[entities.index(tail)])
        else:                                                        # :138-145
            entity = head if head is not None else tail
            if entity:
                pixel_image.append(entity2visual_features[entities.index(entity)])
            else:
                pixel_image.append(torch.zeros(entity2visual_features.shape[1:]))
            pixel_image.append(torch.zeros(entity2visual_features.shape[1:]))
        pixel_images.append(torch.stack(pixel_image))                # :156
    return torch.stack(pixel_images)                                 # :161
