"""Pins the oracle restatement against the reference code executed live from /root/reference (build container
only; skipped on the GPU box). Covers a second seed / shape than the committed golden fixture, and the AdamW
restatement against torch.optim.AdamW with the reference's parameter grouping (lit_models/transformer.py:224-241)."""
import pytest
import torch

from mkg_analogy_b200.config import MKGformerConfig, init_params
from mkg_analogy_b200.data import TINY_VOCAB, make_finetune_batch
from oracle import mkgformer_oracle as O
from oracle import reference_loader as RL

pytestmark = pytest.mark.reference


def test_forward_matches_reference_other_seed():
    cfg = MKGformerConfig(hidden=128, layers=9, heads=2, intermediate=256, vocab=TINY_VOCAB.vocab, max_position=64,
                          image_size=64, patch=32)
    P = init_params(cfg, 3)
    model = RL.build_reference_model(cfg, P)
    batch = make_finetune_batch(3, 32, cfg.image_size, TINY_VOCAB, seed=99, min_len=20)
    with torch.no_grad():
        mo = model(**{k: batch[k] for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "sep_idx")},
                   return_dict=True)
        logits, z = O.forward(P, cfg, batch)
    assert torch.allclose(logits, mo[0].logits, atol=2e-5, rtol=1e-4)
    assert torch.allclose(z, mo[1], atol=2e-5, rtol=1e-4)


def test_state_dict_contract_matches_reference():
    from mkg_analogy_b200.config import TINY, param_shapes
    model = RL.build_reference_model(TINY)
    ref = {k: tuple(v.shape) for k, v in model.named_parameters()}
    mine = param_shapes(TINY)
    assert set(ref) == set(mine), (set(ref) ^ set(mine))
    for k in ref:
        assert tuple(mine[k]) == ref[k], k


def test_adamw_restatement_matches_torch():
    g = torch.Generator().manual_seed(0)
    names = ["a.weight", "a.bias", "x.LayerNorm.weight", "v.layer_norm1.weight"]
    params = {n: torch.randn(5, 3, generator=g) for n in names}
    grads = {n: torch.randn(5, 3, generator=g) for n in names}
    tp = {n: torch.nn.Parameter(p.clone()) for n, p in params.items()}
    no_decay = ["bias", "LayerNorm.weight"]
    opt = torch.optim.AdamW([
        {"params": [p for n, p in tp.items() if not any(nd in n for nd in no_decay)], "weight_decay": 0.01},
        {"params": [p for n, p in tp.items() if any(nd in n for nd in no_decay)], "weight_decay": 0.0}],
        lr=5e-5, eps=1e-8)
    state = {}
    for step in (1, 2, 3):
        for n in names:
            tp[n].grad = grads[n].clone() * step
        opt.step()
        O.adamw_step(params, {n: grads[n] * step for n in names}, state, step, 5e-5)
    for n in names:
        assert torch.allclose(params[n], tp[n].detach(), atol=1e-7, rtol=1e-6), n


def test_reference_litmodel_accepts_dropin_module():
    """The UNMODIFIED reference TransformerLitModel (lit_models/transformer.py:18-57) constructs around the drop-in
    module and runs its vocabulary surgery (`resize_token_embeddings`, `_init_relation_word` incl. its tie asserts);
    the no-decay parameter grouping of configure_optimizers (:225-230) sees the reference's names.  CPU: storage only."""
    from mkg_analogy_b200.config import TINY
    from mkg_analogy_b200.modeling import MKGformerKGC
    cfg = MKGformerConfig(**{**TINY.to_dict(), "vocab": TINY_VOCAB.vocab - 1})
    model = MKGformerKGC(cfg, device="cpu")
    lit = RL.build_reference_lit(model, TINY_VOCAB)
    lit.tokenizer.n = TINY_VOCAB.vocab - 1           # before "[R]" is added
    lit.analogy_relation_ids = TINY_VOCAB.analogy_relation_ids()
    lit._init_relation_word()                        # resize to vocab, set E["[R]"] = mean of relation rows, assert tie
    assert model.get_input_embeddings().weight.shape[0] == TINY_VOCAB.vocab
    E = model.get_input_embeddings().weight
    assert torch.allclose(E[TINY_VOCAB.vocab - 1], E[TINY_VOCAB.analogy_relation_ids()].mean(0))
    ref_model = RL.build_reference_model(TINY)
    assert set(dict(model.named_parameters())) == set(dict(ref_model.named_parameters()))
    assert set(model.state_dict()) == set(ref_model.state_dict())
    nd = ["bias", "LayerNorm.weight"]
    mine = sorted(n for n, p in model.named_parameters() if p.requires_grad and any(x in n for x in nd))
    ref = sorted(n for n, p in ref_model.named_parameters() if p.requires_grad and any(x in n for x in nd))
    assert mine == ref
    # reference checkpoints load (main.py:133-134): state_dict of the reference module -> drop-in module
    model.load_state_dict(ref_model.state_dict())
    assert torch.equal(model.state_dict()["unimo.encoder.text_layer.3.attention.self.query.weight"],
                       ref_model.state_dict()["unimo.encoder.text_layer.3.attention.self.query.weight"])


def test_swap_encoders_into_reference_flava_and_vilt_models():
    """mkg_analogy_b200.accelerate.swap_encoders on the REAL reference models (FlavaForMaskedLM: image / text / multimodal
    encoders; ViltForMaskedLM): every pre-LN encoder is replaced, state_dict keys and values are preserved, and the oracle
    restatement of the stack reproduces the reference encoder the adapter replaces.  (CPU: structure only — the adapter's
    numerics are GPU-tested against the same reference encoders' golden outputs, tests/test_stack_gpu.py.)"""
    from transformers import FlavaConfig, ViltConfig
    from mkg_analogy_b200.accelerate import EncoderAdapter, swap_encoders
    Fm, Vm = RL.load_flava(), RL.load_vilt()
    small = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256)
    fc = FlavaConfig(image_config=dict(small, image_size=32, patch_size=16), text_config=dict(small, vocab_size=300),
                     multimodal_config=dict(small), hidden_size=128, projection_dim=128)
    for cfg, cls in ((fc, Fm.FlavaModel), (ViltConfig(**small, vocab_size=300, image_size=64, patch_size=32), Vm.ViltModel)):
        model = cls(cfg)
        before = {k: v.clone() for k, v in model.state_dict().items()}
        n_enc = sum(1 for m in model.modules() if type(m).__name__ in ("FlavaEncoder", "ViltEncoder"))
        swap_encoders(model, device="cpu")
        adapters = [m for m in model.modules() if isinstance(m, EncoderAdapter)]
        assert len(adapters) == n_enc and n_enc >= 1
        assert not any(type(m).__name__ in ("FlavaEncoder", "ViltEncoder") for m in model.modules())
        after = model.state_dict()
        assert set(after) == set(before)
        assert all(torch.equal(after[k], before[k]) for k in before)


def test_preln_stack_oracle_matches_live_reference_encoder_other_seed():
    from transformers import FlavaTextConfig
    Fm = RL.load_flava()
    cfg = FlavaTextConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                          hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    torch.manual_seed(5)
    enc = Fm.FlavaEncoder(cfg).eval()
    x = torch.randn(2, 21, 128)
    mask = torch.ones(2, 21, dtype=torch.long)
    mask[1, 15:] = 0
    sep = torch.zeros(2, 6, dtype=torch.long)
    sep[:, 2] = torch.tensor([9, 5])
    with torch.no_grad():
        ref = enc(x, attention_mask=(1.0 - mask[:, None, None, :].float()) * -10000.0, sep_idx=sep).last_hidden_state
        mine = O.preln_stack(dict(enc.state_dict()), x, 2, attention_mask=mask, sep_split=sep[:, 2], row_lo=1)
    assert torch.allclose(mine, ref, atol=1e-5, rtol=1e-5)


def test_vilt_dropin_resizes_position_table_like_the_reference_for_other_image_sizes():
    """Images of another size than config.image_size: the reference resizes the patch position table bilinearly to the actual patch
    grid (modeling_vilt.py:121-136).  ViltKGC's plumbing with the ORACLE encoder injected (test infrastructure) == the live reference
    on the text rows (which do not depend on the reference's random patch order), for a 3 x 3 grid under a 2 x 2 config."""
    from transformers import ViltConfig
    from mkg_analogy_b200.modeling_vilt import ViltKGC
    Vm = RL.load_vilt()
    kw = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, vocab_size=300, image_size=64,
              patch_size=32)
    torch.manual_seed(7)
    ref = Vm.ViltForMaskedLM(ViltConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **kw)).eval()
    with torch.no_grad():
        for n, prm in ref.named_parameters():
            prm.normal_(1.0 if ("norm" in n.lower() and n.endswith("weight")) else 0.0, 0.05)
    ref.mlm_score.decoder.weight = ref.vilt.embeddings.text_embeddings.word_embeddings.weight
    m = ViltKGC(device="cpu", **kw)
    missing, unexpected = m.load_state_dict(ref.state_dict(), strict=False)
    assert not unexpected, unexpected
    m.tie_weights()

    class OracleEncoder(torch.nn.Module):
        def __init__(self, st):
            super().__init__()
            self.st = st

        def forward(self, x, attention_mask=None, sep_idx=None, row_lo=1):
            return O.preln_stack(dict(self.st.named_parameters()), x, 2, attention_mask=attention_mask,
                                 sep_split=None if sep_idx is None else sep_idx[:, 2], row_lo=row_lo)
    m.vilt.encoder = OracleEncoder(m.vilt.encoder)
    m.eval()
    g = torch.Generator().manual_seed(8)
    Bn, L = 2, 24
    ids = torch.randint(5, 300, (Bn, L), generator=g)
    mask = torch.ones(Bn, L, dtype=torch.long)
    px = torch.randn(Bn, 2, 3, 96, 96, generator=g)             # 3 x 3 patches per image; the table holds 2 x 2
    sep = torch.full((Bn, 6), L, dtype=torch.long)              # no-op split (the reference cannot run without one)
    with torch.no_grad():
        torch.manual_seed(1)
        ro = ref(input_ids=ids, attention_mask=mask, token_type_ids=torch.zeros_like(ids), pixel_values=px, sep_idx=sep.clone(), return_dict=True)
        mo = m(input_ids=ids, attention_mask=mask, token_type_ids=torch.zeros_like(ids), pixel_values=px, sep_idx=sep)
    assert torch.allclose(mo[1], ro[1], atol=2e-5, rtol=1e-4)
    assert torch.allclose(mo[0].logits, ro[0].logits, atol=2e-5, rtol=1e-4)


def test_postln_stack_oracle_matches_live_reference_visual_bert_encoder_other_seed():
    Vb = RL.load_visual_bert()
    cfg = RL.visual_bert_config(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                                hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    torch.manual_seed(6)
    enc = Vb.VisualBertEncoder(cfg).eval()
    x = torch.randn(2, 27, 128)
    mask = torch.ones(2, 27, dtype=torch.long)
    mask[1, 20:] = 0
    img_length = 6
    sep = torch.zeros(2, 6, dtype=torch.long)
    sep[:, 2] = torch.tensor([15, 11])                       # already shifted by img_length, as VisualBertModel passes it
    with torch.no_grad():
        ref = enc(x, attention_mask=(1.0 - mask[:, None, None, :].float()) * -10000.0, sep_idx=sep.clone(), img_length=img_length,
                  return_dict=True).last_hidden_state
        mine = O.postln_stack(dict(enc.state_dict()), x, 2, attention_mask=mask, sep_split=sep[:, 2], row_lo=img_length + 1)
    assert torch.allclose(mine, ref, atol=1e-5, rtol=1e-5)


def test_visual_bert_dropin_loads_under_the_reference_litmodel_name_check():
    """TransformerLitModel keys the rel_idx / q_head_idx / a_head_idx shift on the CLASS NAME (lit_models/transformer.py:98): the drop-in
    must be called VisualBertKGC, as models/model.py:13 names the reference class."""
    from mkg_analogy_b200.modeling_visual_bert import VisualBertKGC
    m = VisualBertKGC(device="cpu", vocab_size=120, hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=256,
                      visual_embedding_dim=32, max_position_embeddings=40)
    assert "VisualBertKGC" in str(type(m))
    ref_names = set(RL.load_visual_bert().VisualBertForMaskedLM(RL.visual_bert_config(
        vocab_size=120, hidden_size=128, num_hidden_layers=1, num_attention_heads=2, intermediate_size=256, visual_embedding_dim=32,
        max_position_embeddings=40)).state_dict())
    mine = set(m.state_dict())
    assert mine == ref_names, (sorted(mine ^ ref_names)[:8])


def test_swap_encoders_into_reference_visual_bert_model():
    """swap_encoders on the REAL reference VisualBertModel: the post-LN encoder (with its adaptive weights) is replaced by an
    EncoderAdapter(post_ln=True), state_dict keys and values are preserved.  (CPU: structure only; numerics: tests/test_visual_bert_gpu.py.)"""
    from mkg_analogy_b200.accelerate import EncoderAdapter, swap_encoders
    Vb = RL.load_visual_bert()
    cfg = RL.visual_bert_config(vocab_size=200, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
                                visual_embedding_dim=32, max_position_embeddings=40)
    model = Vb.VisualBertModel(cfg, add_pooling_layer=False)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    swap_encoders(model, device="cpu")
    adapters = [m for m in model.modules() if isinstance(m, EncoderAdapter)]
    assert len(adapters) == 1 and not any(type(m).__name__ == "VisualBertEncoder" for m in model.modules())
    after = model.state_dict()
    assert set(after) == set(before) and all(torch.equal(after[k], before[k]) for k in before)
