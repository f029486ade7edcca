"""VisualBERT on the B200 library: the post-LayerNorm mode of the encoder stack through the C ABI against the reference
VisualBertEncoder (tests/golden/postln_stack.pt) and the oracle restatement (incl. training-mode dropout, mask for mask), the
VisualBertKGC drop-in against the reference VisualBertForMaskedLM on a tiny config (visual_bert_tiny.pt) and at its REAL size
(visual_bert_base.pt: 12 blocks, hidden 768, 64 text tokens + 72 region features of 2048, vocabulary 42,007; loss of the reference
`TransformerLitModel.training_step`, every gradient norm).  Fixtures: oracle/make_golden_visual_bert.py.
Tolerance (bf16 operands, fp32 accumulation): hidden states / logits 2e-2 relative on the random-weight tiny cases (as for the MKGformer
path), 1e-2 at base size; gradients 6e-2; the cancelling adaptive-weight sums 25-50 %."""
import os

import pytest
import torch
import torch.nn.functional as F

from mkg_analogy_b200.data import MARS_VOCAB, make_finetune_batch, make_region_features

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def checksum(ts):
    return float(sum(t.double().abs().sum().item() for t in ts))


def _stack_from(c, **kw):
    from mkg_analogy_b200.stack import EncoderStack
    st = EncoderStack(128, 3, 2, 256, eps=1e-12, act="gelu", device="cuda", post_ln=True, **kw)
    st.load_state_dict(c["params"], strict=True)
    return st


def test_postln_stack_matches_reference_encoder():
    c = torch.load(os.path.join(GOLD, "postln_stack.pt"), weights_only=False)
    st = _stack_from(c).eval()
    x = c["x"].cuda().requires_grad_(True)
    y = st(x, attention_mask=c["mask"].cuda(), sep_idx=c["sep"].cuda(), row_lo=c["row_lo"])
    y.backward(c["gy"].cuda())
    torch.cuda.synchronize()
    print(f"post-LN stack: y rel err {rel(y.detach().cpu(), c['y']):.3e}, dx rel err {rel(x.grad.cpu(), c['dx']):.3e}")
    assert rel(y.detach().cpu(), c["y"]) <= 2e-2
    assert rel(x.grad.cpu(), c["dx"]) <= 6e-2
    named = dict(st.named_parameters())
    for n, gr in c["grads"].items():
        e = rel(named[n].grad.cpu(), gr)
        assert e <= (0.25 if "adaptive_weight" in n else 6e-2), (n, e)
    for n, ref_norm in c["grad_norm"].items():
        if ref_norm < 1e-7:
            continue
        got = named[n].grad.norm().item()
        assert abs(got - ref_norm) <= (0.25 if "adaptive_weight" in n else 0.06) * ref_norm + 1e-6, (n, got, ref_norm)


def test_postln_stack_training_dropout_matches_oracle_masks():
    from oracle import mkgformer_oracle as O
    c = torch.load(os.path.join(GOLD, "postln_stack.pt"), weights_only=False)
    st = _stack_from(c, hidden_dropout=0.1, attn_dropout=0.1).train()
    x = c["x"].cuda().requires_grad_(True)
    y = st(x, attention_mask=c["mask"].cuda(), sep_idx=c["sep"].cuda(), row_lo=c["row_lo"])
    y.backward(c["gy"].cuda())
    drop = O.hash_dropout(st._last_seed, 0.1, 0.1)
    P = {k: v.cuda().clone().requires_grad_(True) for k, v in c["params"].items()}
    xr = c["x"].cuda().requires_grad_(True)
    yr = O.postln_stack(P, xr, 2, attention_mask=c["mask"].cuda(), sep_split=c["sep"][:, 2].cuda(), row_lo=c["row_lo"], drop=drop)
    yr.backward(c["gy"].cuda())
    assert rel(yr.detach().cpu(), c["y"]) > 1e-2                 # the masks matter
    assert rel(y.detach(), yr.detach()) <= 2e-2 and rel(x.grad, xr.grad) <= 6e-2
    named = dict(st.named_parameters())
    for n in ("layer.0.attention.self.query.weight", "layer.2.output.dense.bias", "layer.1.attention.output.dense.bias",
              "layer.0.intermediate.dense.bias", "layer.2.output.LayerNorm.weight", "layer.1.attention.output.LayerNorm.bias"):
        assert rel(named[n].grad, P[n].grad) <= 6e-2, n


def test_encoder_adapter_keeps_the_visual_bert_encoder_call_signature():
    """EncoderAdapter(post_ln=True) called the way VisualBertModel calls its encoder (additive 4-D mask, sep_idx already shifted,
    img_length; modeling_visual_bert.py:895-905) reproduces the reference encoder's golden output."""
    from mkg_analogy_b200.accelerate import EncoderAdapter
    c = torch.load(os.path.join(GOLD, "postln_stack.pt"), weights_only=False)
    ad = EncoderAdapter(128, 3, 2, 256, eps=1e-12, act="gelu", device="cuda", post_ln=True)
    ad.load_state_dict(c["params"], strict=True)
    ad.eval()
    ext = (1.0 - c["mask"][:, None, None, :].float()) * -10000.0
    with torch.no_grad():
        out = ad(c["x"].cuda(), attention_mask=ext.cuda(), head_mask=[None] * 3, sep_idx=c["sep"].cuda(), img_length=c["row_lo"] - 1,
                 return_dict=True)
    assert rel(out.last_hidden_state.cpu(), c["y"]) <= 2e-2 and rel(out[0].cpu(), c["y"]) <= 2e-2


def test_visual_bert_dropin_vs_reference_golden():
    from mkg_analogy_b200.modeling_visual_bert import VisualBertKGC
    c = torch.load(os.path.join(GOLD, "visual_bert_tiny.pt"), weights_only=False)
    m = VisualBertKGC(device="cpu", **c["cfg"])
    missing, unexpected = m.load_state_dict(c["params"], strict=False)
    assert not missing and not unexpected
    m.tie_weights()
    m = m.cuda().eval()
    kw = {k: c[k].cuda() for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "visual_attention_mask", "sep_idx")}
    out = m(**kw)
    logits, trans = out[0].logits, out[1]
    ((logits * c["gl"].cuda()).sum() + (trans * c["gt"].cuda()).sum()).backward()
    print(f"VisualBERT tiny: logits rel err {rel(logits.detach().cpu(), c['logits']):.3e}, trans {rel(trans.detach().cpu(), c['trans']):.3e}")
    assert rel(logits.detach().cpu(), c["logits"]) <= 2e-2 and rel(trans.detach().cpu(), c["trans"]) <= 2e-2
    prm = dict(m.named_parameters())
    for n, gr in c["grads"].items():
        e = rel(prm[n].grad.cpu(), gr)
        assert e <= (0.25 if "adaptive_weight" in n else 6e-2), (n, e)


def test_visual_bert_base_vs_reference():
    from mkg_analogy_b200 import nn as BN
    from mkg_analogy_b200.modeling_visual_bert import VisualBertKGC
    c = torch.load(os.path.join(GOLD, "visual_bert_base.pt"), weights_only=False)
    torch.manual_seed(c["seed"])
    m = VisualBertKGC(device="cpu", vocab_size=MARS_VOCAB.vocab, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    sd = m.state_dict()
    chk = checksum([sd["visual_bert.embeddings.word_embeddings.weight"], sd["visual_bert.encoder.layer.11.output.dense.weight"]])
    assert abs(chk - c["param_checksum"]) < 1e-6 * c["param_checksum"], "seeded weights drifted from the fixture"
    m = m.cuda().eval()
    hb = make_finetune_batch(c["B"], c["L"], 32, MARS_VOCAB, seed=c["seed_b"], min_len=c["min_len"])
    hb["pixel_values"], hb["visual_attention_mask"] = make_region_features(c["B"], seed=c["seed_b"] + 1)
    assert abs(checksum([hb["input_ids"], hb["pixel_values"], hb["sep_idx"], hb["visual_attention_mask"]]) - c["batch_checksum"]) \
        < 1e-6 * c["batch_checksum"]
    b = {k: v.cuda() for k, v in hb.items()}
    _, trans = m(**{k: b[k] for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "visual_attention_mask", "sep_idx")},
                 output_logits=False)
    B, img_len = trans.shape[0], b["pixel_values"].shape[1]
    assert trans.shape[1] == c["L"] + img_len
    _, mpos = (b["input_ids"] == 103).nonzero(as_tuple=True)
    cols = torch.as_tensor(MARS_VOCAB.analogy_entity_ids(), device="cuda")
    E, bias = m.get_input_embeddings().weight, m.get_output_embeddings().bias
    ce, ml = BN.mask_row_loss(trans, mpos, E, bias, cols, b["label"])                       # fused route (transformer.py:92-109)
    ar = torch.arange(B, device="cuda")
    rel_idx, qh, ah = b["rel_idx"] + img_len, b["q_head_idx"] + img_len, b["a_head_idx"] + img_len     # transformer.py:98-102
    sim = torch.relu(F.cosine_similarity(trans[ar, qh], trans[ar, ah], dim=-1)) + 1 - \
        F.cosine_similarity(trans[ar, rel_idx[:, 0]], trans[ar, rel_idx[:, 1]], dim=-1)
    loss = ce + 0.43 * sim.mean()
    loss.backward()
    e_t, e_l = rel(trans.detach().cpu(), c["trans"]), rel(ml.detach().cpu(), c["mask_logits_entities"])
    worst = 0.0
    for n, p in m.named_parameters():
        ref = c["grad_norm"].get(n)
        if ref is None:
            assert p.grad is None or p.grad.abs().max().item() == 0.0, n
            continue
        if ref < 1e-7:
            continue
        got = p.grad.norm().item()
        # adaptive weights: cancelling sums of ~10^6 terms dS * S_raw whose total is ~1e-7 .. 1e-5 here (near-uniform attention of a randomly
        # initialised model) while the terms are bf16-rounded: bounded absolutely (other gradient norms are 1e-3 .. 1)
        if "adaptive_weight" in n:
            assert abs(got - ref) <= 0.50 * ref + 2e-5, (n, got, ref)
            continue
        assert abs(got - ref) <= 0.05 * ref + 1e-7, (n, got, ref)
        worst = max(worst, abs(got - ref) / ref)
    print(f"VisualBERT-base: trans rel err {e_t:.3e}, [MASK]-row entity logits {e_l:.3e}, loss {loss.item():.5f} vs reference "
          f"{c['loss'].item():.5f}, worst gradient-norm deviation {worst:.3e}")
    assert e_t <= 1e-2 and e_l <= 1e-2 and abs(loss.item() - c["loss"].item()) <= 1e-2
