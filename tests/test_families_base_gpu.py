"""FLAVA-base and ViLT-base drop-ins at their REAL sizes (hidden 768; FLAVA 393 image + 64 text tokens over 12 + 12 + 6 blocks,
ViLT 290 image + 128 text tokens over 12 blocks; vocabulary 42,007) against fixtures generated from the reference models
(`oracle/make_golden_families.py`: `FlavaForMaskedLM` / `ViltForMaskedLM` forward, the reference `TransformerLitModel.training_step`
loss, every gradient norm).  Both sides build the drop-in on the CPU under the same seed, so no weights are stored.
Tolerance (bf16 operands, fp32 accumulation, well-conditioned pre-LN stacks): trans / logits 1e-2 relative, loss |d| 1e-2,
gradient norms within 5 % (50 % for the cancelling adaptive-weight sums)."""
import os

import pytest
import torch
import torch.nn.functional as F

from mkg_analogy_b200.data import MARS_VOCAB, make_finetune_batch, make_pretrain_batch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def checksum(ts):
    return float(sum(t.double().abs().sum().item() for t in ts))


def _build(kind, seed):
    torch.manual_seed(seed)
    if kind == "flava":
        from mkg_analogy_b200.modeling_flava import FlavaKGC
        m = FlavaKGC(device="cpu", text_config=dict(vocab_size=MARS_VOCAB.vocab))
    else:
        from mkg_analogy_b200.modeling_vilt import ViltKGC
        m = ViltKGC(device="cpu", vocab_size=MARS_VOCAB.vocab)
    return m


def _check_grad_norms(model, ref_norms):
    worst = 0.0
    for n, p in model.named_parameters():
        ref = ref_norms.get(n)
        if ref is None:
            assert p.grad is None or p.grad.abs().max().item() == 0.0, n
            continue
        if ref < 1e-7:
            continue
        got = p.grad.norm().item()
        # the two adaptive weights of a block are cancelling sums of ~10^6 terms dS * S_raw (forward-error bound 2^-9 * sum|terms|,
        # asserted with the oracle's own terms in tests/test_base_gpu.py::test_base_shape_well_conditioned): a loose relative bound here
        if "adaptive_weight" in n:
            # ... and an absolute floor: these sums come out at 1e-5 .. 1e-4 here while their terms are bf16-rounded and the partial sums
            # arrive through atomics in a different order every run (observed run to run: 24 % .. 56 % of a 4.8e-5 gradient)
            assert abs(got - ref) <= 0.50 * ref + 3e-5, (n, got, ref)
            continue
        assert abs(got - ref) <= 0.05 * ref + 1e-7, (n, got, ref)
        worst = max(worst, abs(got - ref) / ref)
    return worst


def test_flava_base_vs_reference():
    from mkg_analogy_b200 import nn as BN
    c = torch.load(os.path.join(GOLD, "flava_base.pt"), weights_only=False)
    m = _build("flava", c["seed"])
    sd = m.state_dict()
    chk = checksum([sd["flava.text_model.embeddings.word_embeddings.weight"], sd["flava.multimodal_model.encoder.layer.5.output.dense.weight"]])
    assert abs(chk - c["param_checksum"]) < 1e-6 * c["param_checksum"], "seeded weights drifted from the fixture"
    m = m.cuda().eval()
    hb = make_finetune_batch(c["B"], c["L"], 224, MARS_VOCAB, seed=c["seed_b"], min_len=c["min_len"])
    assert abs(checksum([hb["input_ids"], hb["pixel_values"], hb["sep_idx"]]) - c["batch_checksum"]) < 1e-6 * c["batch_checksum"]
    b = {k: v.cuda() for k, v in hb.items()}
    _, trans = m(**{k: b[k] for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "sep_idx")}, output_logits=False)
    B = trans.shape[0]
    _, mpos = (b["input_ids"] == 103).nonzero(as_tuple=True)
    cols = torch.as_tensor(MARS_VOCAB.analogy_entity_ids(), device="cuda")
    E, bias = m.get_input_embeddings().weight, m.get_output_embeddings().bias
    ce, ml = BN.mask_row_loss(trans, mpos, E, bias, cols, b["label"])                       # fused route (transformer.py:92-109)
    ar = torch.arange(B, device="cuda")
    sim = torch.relu(F.cosine_similarity(trans[ar, b["q_head_idx"]], trans[ar, b["a_head_idx"]], dim=-1)) + 1 - \
        F.cosine_similarity(trans[ar, b["rel_idx"][:, 0]], trans[ar, b["rel_idx"][:, 1]], dim=-1)
    loss = ce + 0.43 * sim.mean()
    loss.backward()
    e_t, e_l = rel(trans.detach().cpu(), c["trans"]), rel(ml.detach().cpu(), c["mask_logits_entities"])
    worst = _check_grad_norms(m, c["grad_norm"])
    print(f"FLAVA-base: trans rel err {e_t:.3e}, [MASK]-row entity logits {e_l:.3e}, loss {loss.item():.5f} vs reference {c['loss'].item():.5f}, "
          f"worst gradient-norm deviation {worst:.3e}")
    assert e_t <= 1e-2 and e_l <= 1e-2 and abs(loss.item() - c["loss"].item()) <= 1e-2


def test_vilt_base_vs_reference():
    from mkg_analogy_b200 import nn as BN
    c = torch.load(os.path.join(GOLD, "vilt_base.pt"), weights_only=False)
    m = _build("vilt", c["seed"])
    sd = m.state_dict()
    chk = checksum([sd["vilt.embeddings.text_embeddings.word_embeddings.weight"], sd["vilt.encoder.layer.11.output.dense.weight"]])
    assert abs(chk - c["param_checksum"]) < 1e-6 * c["param_checksum"], "seeded weights drifted from the fixture"
    m = m.cuda().eval()
    hb = make_pretrain_batch(c["B"], c["L"], 384, MARS_VOCAB, seed=c["seed_b"], min_len=c["min_len"])
    assert abs(checksum([hb["input_ids"], hb["pixel_values"]]) - c["batch_checksum"]) < 1e-6 * c["batch_checksum"]
    hb["sep_idx"] = torch.full((c["B"], 6), c["L"], dtype=torch.long)       # the reference ViLT cannot run without one; no-op split
    b = {k: v.cuda() for k, v in hb.items()}
    _, trans = m(**{k: b[k] for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "sep_idx")}, output_logits=False)
    B = trans.shape[0]
    _, mpos = (b["input_ids"] == 103).nonzero(as_tuple=True)
    E, bias = m.get_input_embeddings().weight, m.get_output_embeddings().bias
    voc = MARS_VOCAB
    ent = b["pre_type"] != 2                                                                # pre-train branch (transformer.py:72-90)
    loss = trans.new_zeros(())
    if ent.any():
        loss = loss + BN.mask_row_loss(trans[ent], mpos[ent], E, bias, (voc.entity_id_st, voc.entity_id_ed), b["label"][ent])[0]
    if (~ent).any():
        loss = loss + BN.mask_row_loss(trans[~ent], mpos[~ent], E, bias, (voc.relation_id_st, voc.relation_id_ed), b["label"][~ent])[0]
    loss.backward()
    with torch.no_grad():
        zm = trans[torch.arange(B, device="cuda"), mpos]
        ml = zm @ E[voc.entity_id_st:voc.relation_id_ed].t() + bias[voc.entity_id_st:voc.relation_id_ed]
    e_t, e_l = rel(trans.detach().cpu(), c["trans"]), rel(ml.cpu(), c["mask_logits_ent_rel"])
    worst = _check_grad_norms(m, c["grad_norm"])
    print(f"ViLT-base: trans rel err {e_t:.3e}, [MASK]-row entity+relation logits {e_l:.3e}, loss {loss.item():.5f} vs reference "
          f"{c['loss'].item():.5f}, worst gradient-norm deviation {worst:.3e}")
    assert e_t <= 1e-2 and e_l <= 1e-2 and abs(loss.item() - c["loss"].item()) <= 1e-2
