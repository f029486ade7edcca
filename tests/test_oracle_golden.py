"""The oracle restatement (oracle/mkgformer_oracle.py) against golden outputs of the REFERENCE itself
(tests/golden/*.pt, produced by oracle/make_golden.py in the build container). fp32 on CPU; tolerance:
abs 2e-5 on logits/hidden (different but equivalent op order), rel 2e-4 on gradients."""
import os

import pytest
import torch

from mkg_analogy_b200.config import MKGformerConfig, init_params
from mkg_analogy_b200.data import TINY_VOCAB, make_finetune_batch, make_pretrain_batch
from oracle import mkgformer_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def tiny():
    g = torch.load(os.path.join(GOLD, "mkgformer_tiny.pt"), weights_only=False)
    cfg = MKGformerConfig(**g["cfg"])
    params = init_params(cfg, g["seed_w"])
    batch = make_finetune_batch(g["B"], g["L"], cfg.image_size, TINY_VOCAB, seed=g["seed_b"], min_len=16)
    chk = float(sum(t.double().abs().sum().item() for t in params.values()))
    assert abs(chk - g["param_checksum"]) < 1e-6 * g["param_checksum"], "seeded weights drifted from the fixture"
    bchk = float(sum(t.double().abs().sum().item() for t in (batch["input_ids"], batch["pixel_values"], batch["sep_idx"])))
    assert abs(bchk - g["batch_checksum"]) < 1e-6 * g["batch_checksum"], "seeded batch drifted from the fixture"
    return g, cfg, params, batch


def test_forward_logits_and_hidden(tiny):
    g, cfg, P, batch = tiny
    with torch.no_grad():
        logits, z = O.forward(P, cfg, batch, full_logits=True)
    mpos = O.mask_positions(batch["input_ids"])
    B = z.shape[0]
    assert torch.allclose(z, g["ft_trans"], atol=2e-5, rtol=1e-4)
    assert torch.allclose(logits[torch.arange(B), mpos], g["ft_mask_logits_full"], atol=2e-5, rtol=1e-4)
    assert torch.allclose(logits[0], g["ft_logits_row0"], atol=2e-5, rtol=1e-4)


def test_finetune_loss_grads_ranks(tiny):
    g, cfg, P, batch = tiny
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    loss, x, _ = O.finetune_loss(Pg, cfg, batch, TINY_VOCAB.analogy_entity_ids(), alpha=0.43, label_smoothing=0.1)
    assert abs(loss.item() - g["ft_loss"].item()) < 1e-5
    loss.backward()
    for name, ref_norm in g["ft_grad_norm"].items():
        if name.startswith("cls.predictions.decoder"):
            continue
        gr = Pg[name].grad
        if ref_norm is None:
            assert gr is None or gr.abs().max().item() == 0.0, f"{name} must have no gradient"
            continue
        assert gr is not None, name
        assert abs(gr.norm().item() - ref_norm) <= 2e-4 * ref_norm + 1e-7, (name, gr.norm().item(), ref_norm)
    for name, ref in g["ft_grad_full"].items():
        gr = Pg[name].grad
        if name.endswith("patch_embedding.weight"):
            gr = gr[:8]
        assert torch.allclose(gr, ref, atol=1e-6, rtol=2e-4), name
    ranks = O.ranks_of_gold(x.detach(), batch["label"])
    assert torch.equal(ranks, g["ft_ranks"])


def test_pretrain_loss_and_grad_norms(tiny):
    g, cfg, P, _ = tiny
    pb = make_pretrain_batch(g["B"], g["L"], cfg.image_size, TINY_VOCAB, seed=g["seed_b"] + 1)
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    loss = O.pretrain_loss(Pg, cfg, pb, (TINY_VOCAB.entity_id_st, TINY_VOCAB.entity_id_ed),
                           (TINY_VOCAB.relation_id_st, TINY_VOCAB.relation_id_ed), label_smoothing=0.1)
    assert abs(loss.item() - g["pt_loss"].item()) < 1e-5
    loss.backward()
    for name, ref_norm in g["pt_grad_norm"].items():
        if name.startswith("cls.predictions.decoder"):
            continue
        gr = Pg[name].grad
        if ref_norm is None:
            assert gr is None or gr.abs().max().item() == 0.0, name
        else:
            assert abs(gr.norm().item() - ref_norm) <= 2e-4 * ref_norm + 1e-7, name


def test_label_smooth_ce_known_answers():
    g = torch.load(os.path.join(GOLD, "ls_ce.pt"), weights_only=False)
    for tag, c in g.items():
        x = c["x"].clone().requires_grad_(True)
        loss = O.label_smooth_ce(x, c["y"], 0.1)
        (gx,) = torch.autograd.grad(loss, x)
        assert torch.allclose(loss, c["loss"], atol=1e-6, rtol=1e-6), tag
        assert torch.allclose(gx, c["grad"], atol=1e-7, rtol=1e-5), tag
        # closed form of SURVEY.md App. A: dx = ((sum q) softmax(x) - q) / n_valid
        C = x.shape[1]
        valid = c["y"] != -100
        q = torch.full_like(x, 0.1 / C).scatter_(1, c["y"].clamp(min=0)[:, None], 0.9)
        closed = ((1 - 0.1 / C) * torch.softmax(x, 1) - q) * valid[:, None] / valid.sum()
        assert torch.allclose(closed.detach(), c["grad"], atol=1e-7, rtol=1e-5), tag


def test_preln_stack_oracle_vs_reference_flava_and_vilt_encoders():
    """oracle.preln_stack against outputs / gradients of the reference FlavaEncoder (modeling_flava.py:668-722) and
    ViltEncoder (modeling_vilt.py:531-583) stored by oracle/make_golden.py::stack_case."""
    g = torch.load(os.path.join(GOLD, "preln_stack.pt"), weights_only=False)
    for tag, c in g.items():
        P = {k: v.clone().requires_grad_(True) for k, v in c["params"].items()}
        x = c["x"].clone().requires_grad_(True)
        y = O.preln_stack(P, x, 2, attention_mask=c["mask"], sep_split=c["sep"][:, 2], row_lo=c["row_lo"])
        y.backward(c["gy"])
        assert torch.allclose(y, c["y"], atol=1e-5, rtol=1e-5), tag
        assert torch.allclose(x.grad, c["dx"], atol=1e-5, rtol=1e-4), tag
        for n, gr in c["grads"].items():
            assert torch.allclose(P[n].grad, gr, atol=1e-5, rtol=1e-4), (tag, n)
        for n, ref_norm in c["grad_norm"].items():
            assert abs(P[n].grad.norm().item() - ref_norm) <= 2e-4 * ref_norm + 1e-7, (tag, n)


def _vilt_with_oracle_encoder(c):
    """ViltKGC's PyTorch plumbing (embeddings, final LN, head) with the ORACLE pre-LN stack injected in place of the CUDA
    encoder — test infrastructure only: lets the CPU suite pin the plumbing against the reference ViltForMaskedLM."""
    from mkg_analogy_b200.modeling_vilt import ViltKGC
    m = ViltKGC(device="cpu", **c["cfg"])
    m.load_state_dict(c["params"])
    m.tie_weights()

    class OracleEncoder(torch.nn.Module):
        def __init__(self, st):
            super().__init__()
            self.st = st

        def forward(self, x, attention_mask=None, sep_idx=None, row_lo=1):
            P = dict(self.st.named_parameters())
            return O.preln_stack(P, x, c["cfg"]["num_attention_heads"], attention_mask=attention_mask,
                                 sep_split=None if sep_idx is None else sep_idx[:, 2], row_lo=row_lo)
    m.vilt.encoder = OracleEncoder(m.vilt.encoder)
    return m.eval()


def test_vilt_dropin_plumbing_vs_reference_golden():
    """ViltKGC (embeddings incl. the 40-position quirk, [text, img1, img2] layout, modality embeddings, sep_idx + img_length,
    final LN, first-L rows, tied MLM head) == reference ViltForMaskedLM on the text rows, although the reference permutes the
    image patches randomly (self-attention is permutation-equivariant)."""
    c = torch.load(os.path.join(GOLD, "vilt_tiny.pt"), weights_only=False)
    m = _vilt_with_oracle_encoder(c)
    sep = c["sep"].clone()
    out = m(input_ids=c["ids"], attention_mask=c["mask"], token_type_ids=c["tt"], pixel_values=c["px"], sep_idx=sep)
    assert torch.equal(sep, c["sep"])                     # unlike the reference, the caller's sep_idx is not mutated
    assert torch.allclose(out[0].logits, c["logits"], atol=2e-5, rtol=1e-4)
    assert torch.allclose(out[1], c["trans"], atol=2e-5, rtol=1e-4)
    ((out[0].logits * c["gl"]).sum() + (out[1] * c["gt"]).sum()).backward()
    prm = dict(m.named_parameters())
    for n, gr in c["grads"].items():
        key = n.replace("vilt.encoder.", "vilt.encoder.st.")
        assert torch.allclose(prm[key].grad, gr, atol=1e-5, rtol=2e-4), n
    assert m.get_output_embeddings().weight is m.get_input_embeddings().weight
    emb = m.resize_token_embeddings(305)
    assert emb.weight.shape[0] == 305 and m.get_output_embeddings().weight is emb.weight and m.mlm_score.bias.shape[0] == 305


def test_flava_dropin_plumbing_vs_reference_golden():
    """FlavaKGC's PyTorch plumbing (image embeddings with the image-2 position quirk, text embeddings, mm projections on the
    pre-final-LN states, [cls ; image ; text] multimodal input without a mask, final LN, last-L rows, tied MLM head) with the
    ORACLE pre-LN stack injected as the three encoders (test infrastructure only) == reference FlavaForMaskedLM."""
    from mkg_analogy_b200.modeling_flava import FlavaKGC
    c = torch.load(os.path.join(GOLD, "flava_tiny.pt"), weights_only=False)
    m = FlavaKGC(device="cpu", **c["cfg"])
    m.load_state_dict(c["params"])
    m.tie_weights()

    class OracleEncoder(torch.nn.Module):
        def __init__(self, st):
            super().__init__()
            self.st = st

        def forward(self, x, attention_mask=None, sep_idx=None, row_lo=1):
            return O.preln_stack(dict(self.st.named_parameters()), x, 2, attention_mask=attention_mask,
                                 sep_split=None if sep_idx is None else sep_idx[:, 2], row_lo=row_lo)
    for tower in (m.flava.image_model, m.flava.text_model, m.flava.multimodal_model):
        tower.encoder = OracleEncoder(tower.encoder)
    m.eval()
    out = m(input_ids=c["ids"], attention_mask=c["mask"], token_type_ids=c["tt"], pixel_values=c["px"], sep_idx=c["sep"].clone())
    assert torch.allclose(out[0].logits, c["logits"], atol=2e-5, rtol=1e-4)
    assert torch.allclose(out[1], c["trans"], atol=2e-5, rtol=1e-4)
    ((out[0].logits * c["gl"]).sum() + (out[1] * c["gt"]).sum()).backward()
    prm = dict(m.named_parameters())
    for n, gr in c["grads"].items():
        key = n.replace(".encoder.", ".encoder.st.")
        assert torch.allclose(prm[key].grad, gr, atol=2e-5, rtol=2e-4), n
    # parameters the reference never reaches stay without gradient here too (pooler, contrastive heads, final LNs of the towers)
    for n in c["unused"]:
        if "adaptive_weight" in n:
            continue                      # oracle stack returns zeros for scaling weights of unscaled towers
        assert prm[n.replace(".encoder.", ".encoder.st.")].grad is None, n
    assert m.get_output_embeddings().weight is m.get_input_embeddings().weight
    emb = m.resize_token_embeddings(305)
    assert emb.weight.shape[0] == 305 and m.get_output_embeddings().weight is emb.weight and m.cls.bias.shape[0] == 305


# ------------------------------------------------------------------------------------------ base size (the benchmarked shape)
@pytest.fixture(scope="module")
def base_fixture():
    from mkg_analogy_b200.config import BASE
    g = torch.load(os.path.join(GOLD, "mkgformer_base.pt"), weights_only=False)
    params = init_params(BASE, g["seed_w"])
    chk = float(sum(params[k].double().abs().sum().item() for k in
                    ("unimo.text_embeddings.word_embeddings.weight", "unimo.encoder.text_layer.11.output.dense.weight")))
    assert abs(chk - g["param_checksum"]) < 1e-6 * g["param_checksum"], "seeded weights drifted from the fixture"
    return g, BASE, params


def test_base_config1_oracle_vs_reference_fixture(base_fixture):
    """BASELINE.json config 1 (MKGformer-base, 4 synthetic MARS examples): the oracle reproduces the reference's [MASK]-row
    logits over the whole vocabulary, trans_hidden_states, training_step loss, ranks and gradient norms."""
    from mkg_analogy_b200.data import MARS_VOCAB
    g, cfg, P = base_fixture
    d = g["c1"]
    batch = make_finetune_batch(d["B"], d["L"], cfg.image_size, MARS_VOCAB, seed=d["seed_b"], min_len=d["min_len"])
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    logits, z = O.forward(Pg, cfg, batch, full_logits=True)
    mpos = O.mask_positions(batch["input_ids"])
    assert torch.allclose(z.detach(), d["trans"], atol=5e-5, rtol=1e-4)
    assert torch.allclose(logits.detach()[torch.arange(d["B"]), mpos], d["mask_logits_full"], atol=5e-5, rtol=1e-4)
    del logits
    loss, x, _ = O.finetune_loss(Pg, cfg, batch, MARS_VOCAB.analogy_entity_ids(), alpha=0.43, label_smoothing=0.1)
    assert abs(loss.item() - d["loss"].item()) < 2e-5
    assert torch.equal(O.ranks_of_gold(x.detach(), batch["label"]), d["ranks"].to(torch.int64))
    loss.backward()
    for name, ref_norm in d["grad_norm"].items():
        if name.startswith("cls.predictions.decoder"):
            continue
        gr = Pg[name].grad
        if ref_norm is None:
            assert gr is None or gr.abs().max().item() == 0.0, name
        elif ref_norm >= 1e-7:
            assert abs(gr.norm().item() - ref_norm) <= 5e-4 * ref_norm + 1e-8, (name, gr.norm().item(), ref_norm)
    for name, ref in d["grad_full"].items():
        if "adaptive_weight" in name:
            assert torch.allclose(Pg[name].grad, ref, atol=1e-7, rtol=2e-3), (name, Pg[name].grad, ref)


def test_base_eval_split_oracle_ranks_first_batch(base_fixture):
    """The held-out decidable split: oracle ranks == reference `_eval` ranks (fixture) on the first batch, Hits@1 = 1."""
    from mkg_analogy_b200.data import MARS_VOCAB, make_decidable_batch
    g, cfg, P = base_fixture
    s = torch.load(os.path.join(GOLD, "base_eval_split.pt"), weights_only=False)
    P = dict(P)
    E = P["unimo.text_embeddings.word_embeddings.weight"].clone()
    bias = P["cls.predictions.bias"].clone()
    E[s["gold_token_ids"]] = s["rows"]
    bias[s["gold_token_ids"]] = s["bias"]
    P["unimo.text_embeddings.word_embeddings.weight"], P["cls.predictions.bias"] = E, bias
    b = make_decidable_batch(s["bs"], s["L"], cfg.image_size, MARS_VOCAB, seed=s["eval_seed0"], n_classes=s["n_classes"])
    chk = float(b["input_ids"].double().abs().sum().item())
    assert abs(chk - s["batch_checksum"]) < 1e-6 * s["batch_checksum"]
    with torch.no_grad():
        _, x, _ = O.finetune_loss(P, cfg, b, MARS_VOCAB.analogy_entity_ids(), alpha=0.43, label_smoothing=0.1)
    r = O.ranks_of_gold(x, b["label"])
    assert torch.equal(r, s["ranks"][: s["bs"]].to(torch.int64))
    assert s["gold_margin"].min().item() > 0.25 and bool((s["ranks"] == 1).all())     # the split is decidable by construction


def test_postln_stack_oracle_vs_reference_visual_bert_encoder():
    """oracle.postln_stack against outputs / gradients of the reference VisualBertEncoder (modeling_visual_bert.py:424-488, adaptive
    scaling :255-260 with img_length = 16) stored by oracle/make_golden_visual_bert.py::stack_case."""
    c = torch.load(os.path.join(GOLD, "postln_stack.pt"), weights_only=False)
    P = {k: v.clone().requires_grad_(True) for k, v in c["params"].items()}
    x = c["x"].clone().requires_grad_(True)
    y = O.postln_stack(P, x, 2, attention_mask=c["mask"], sep_split=c["sep"][:, 2], row_lo=c["row_lo"])
    y.backward(c["gy"])
    assert torch.allclose(y, c["y"], atol=1e-5, rtol=1e-5)
    assert torch.allclose(x.grad, c["dx"], atol=1e-5, rtol=1e-4)
    for n, gr in c["grads"].items():
        assert torch.allclose(P[n].grad, gr, atol=1e-5, rtol=1e-4), n
    for n, ref_norm in c["grad_norm"].items():
        assert abs(P[n].grad.norm().item() - ref_norm) <= 2e-4 * ref_norm + 1e-7, n


def test_visual_bert_dropin_plumbing_vs_reference_golden():
    """VisualBertKGC's plumbing (text + visual embeddings with ONE LayerNorm over [text ; regions], visual token type 1 / visual position
    0, [attention_mask ; visual_attention_mask], sep_idx + img_length on a copy, head over all L + n_regions rows, tied decoder) with the
    ORACLE post-LN stack injected as the encoder (test infrastructure only) == reference VisualBertForMaskedLM."""
    from mkg_analogy_b200.modeling_visual_bert import VisualBertKGC
    c = torch.load(os.path.join(GOLD, "visual_bert_tiny.pt"), weights_only=False)
    m = VisualBertKGC(device="cpu", **c["cfg"])
    missing, unexpected = m.load_state_dict(c["params"], strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    m.tie_weights()

    class OracleEncoder(torch.nn.Module):
        def __init__(self, st):
            super().__init__()
            self.st = st

        def forward(self, x, attention_mask=None, sep_idx=None, row_lo=1):
            return O.postln_stack(dict(self.st.named_parameters()), x, c["cfg"]["num_attention_heads"], attention_mask=attention_mask,
                                  sep_split=None if sep_idx is None else sep_idx[:, 2], row_lo=row_lo)
    m.visual_bert.encoder = OracleEncoder(m.visual_bert.encoder)
    m.eval()
    sep = c["sep_idx"].clone()
    out = m(input_ids=c["input_ids"], attention_mask=c["attention_mask"], token_type_ids=c["token_type_ids"], pixel_values=c["pixel_values"],
            visual_attention_mask=c["visual_attention_mask"], sep_idx=sep)
    assert torch.equal(sep, c["sep_idx"])                 # unlike the reference (:876), the caller's sep_idx is not mutated
    assert out[0].logits.shape == c["logits"].shape       # (B, L + n_regions, V)
    assert torch.allclose(out[0].logits, c["logits"], atol=2e-5, rtol=1e-4)
    assert torch.allclose(out[1], c["trans"], atol=2e-5, rtol=1e-4)
    ((out[0].logits * c["gl"]).sum() + (out[1] * c["gt"]).sum()).backward()
    prm = dict(m.named_parameters())
    for n, gr in c["grads"].items():
        key = n.replace("visual_bert.encoder.", "visual_bert.encoder.st.")
        assert torch.allclose(prm[key].grad, gr, atol=2e-5, rtol=2e-4), n
    assert m.get_output_embeddings().weight is m.get_input_embeddings().weight
    emb = m.resize_token_embeddings(420)
    assert emb.weight.shape[0] == 420 and m.get_output_embeddings().weight is emb.weight and m.cls.predictions.bias.shape[0] == 420
