"""ORACLE TOOLING (test infrastructure) — runs the REAL reference code from /root/reference on CPU.

Only usable in the build container (the reference tree does not travel to the GPU box).  Used by
oracle/make_golden.py to generate tests/golden/*.pt and by tests/test_oracle_vs_reference.py to pin the
restatement in oracle/mkgformer_oracle.py.  Shims follow SURVEY.md §8c2/c3 and Appendix C: the reference pins
transformers 4.19 / pytorch_lightning 1.3.1; this image has transformers 5.5 and no lightning.
Nothing is copied from the reference: its files are imported in place, by path.
"""
import importlib.util
import os
import sys
import types

import torch

REF_ROOT = "/root/reference/MarT"


def available():
    return os.path.isdir(REF_ROOT)


def _install_shims():
    import transformers.modeling_utils as mu
    from transformers import pytorch_utils as pu
    if not hasattr(mu, "apply_chunking_to_forward"):
        mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    # ViLT / FLAVA files import helpers that moved or vanished between transformers 4.19 and 5.x
    for mod in (mu, pu):
        if not hasattr(mod, "find_pruneable_heads_and_indices"):
            def _no_prune(*a, **k):
                raise NotImplementedError("head pruning is not part of the hot path")
            mod.find_pruneable_heads_and_indices = _no_prune
        if not hasattr(mod, "prune_linear_layer"):
            mod.prune_linear_layer = pu.prune_linear_layer
    # 4.19 signatures the vendored ViLT / FLAVA models rely on
    if not getattr(mu.PreTrainedModel, "_b200_shimmed", False):
        mu.PreTrainedModel.get_head_mask = lambda self, hm, n, is_attention_chunked=False: [None] * n
        mu.PreTrainedModel.get_extended_attention_mask = lambda self, m, shape=None, device=None, dtype=None: \
            (1.0 - (m[:, None, None, :] if m.dim() == 2 else m[:, None, :, :]).float()) * -10000.0
        mu.PreTrainedModel._b200_shimmed = True
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")

        class _LM(torch.nn.Module):
            def save_hyperparameters(self, *a, **k):
                pass

            def log(self, *a, **k):
                pass

        pl.LightningModule = _LM
        sys.modules["pytorch_lightning"] = pl


def load_module(name, relpath):
    _install_shims()
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load_flava():
    return load_module("ref_flava", "models/modeling_flava.py")


def load_vilt():
    return load_module("ref_vilt", "models/modeling_vilt.py")


def load_visual_bert():
    """modeling_visual_bert.py under transformers 5.x: its classes call the 4.x `init_weights()` from `__init__` (5.x expects
    `post_init()` to have run first), so the module's base class gets the 4.x behaviour back (apply `_init_weights`; the callers tie the
    decoder themselves); configs additionally need `is_decoder` / `torchscript` attributes (see `visual_bert_config`)."""
    m = load_module("ref_visual_bert", "models/modeling_visual_bert.py")
    m.VisualBertPreTrainedModel.init_weights = lambda self: self.apply(self._init_weights)
    return m


def visual_bert_config(**kw):
    from transformers import VisualBertConfig
    cfg = VisualBertConfig(**kw)
    cfg.is_decoder = False            # read at modeling_visual_bert.py:1640; no longer a config attribute in 5.x
    cfg.torchscript = False
    return cfg


def build_reference_model(cfg, params=None):
    """Instantiate the reference `UnimoForMaskedLM` for `cfg` (MarT/main.py:76-86 construction, then
    resize_token_embeddings as lit_models/transformer.py:33,43 does) and optionally load `params`."""
    from transformers import BertConfig, CLIPVisionConfig
    unimo = load_module("ref_unimo", "models/modeling_unimo.py")
    # The real flow ALWAYS resizes (lit_models/transformer.py:33,43), which rebuilds the word table as a plain
    # nn.Embedding without padding_idx (modeling_unimo.py:823): start one row short so that the
    # resize path is taken for every config.
    base_vocab = min(cfg.vocab - 1, 30522)
    tc = BertConfig(vocab_size=base_vocab, hidden_size=cfg.hidden, num_hidden_layers=cfg.layers,
                    num_attention_heads=cfg.heads, intermediate_size=cfg.intermediate,
                    max_position_embeddings=cfg.max_position, type_vocab_size=cfg.type_vocab,
                    layer_norm_eps=cfg.text_eps)
    tc.torchscript = False
    vc = CLIPVisionConfig(hidden_size=cfg.hidden, intermediate_size=cfg.intermediate, num_hidden_layers=cfg.layers,
                          num_attention_heads=cfg.heads, image_size=cfg.image_size, patch_size=cfg.patch,
                          layer_norm_eps=cfg.vision_eps)
    vc.device = "cpu"
    model = unimo.UnimoForMaskedLM(vc, tc)
    if cfg.vocab != base_vocab:
        model.resize_token_embeddings(cfg.vocab)
    if params is not None:
        sd = model.state_dict()
        new = {}
        for k, v in sd.items():
            if k in params:
                new[k] = params[k].detach().clone()
            elif k == "cls.predictions.decoder.weight":
                new[k] = params["unimo.text_embeddings.word_embeddings.weight"].detach().clone()
            elif k == "cls.predictions.decoder.bias":
                new[k] = params["cls.predictions.bias"].detach().clone()
            else:
                new[k] = v  # position_ids buffers
        model.load_state_dict(new)
        model.tie_weights()
    model.eval()
    return model


class FakeTokenizer:
    """Just enough tokenizer for TransformerLitModel."""

    def __init__(self, n_tokens, mask_token_id=103):
        self.n = n_tokens
        self.mask_token_id = mask_token_id
        self.pad_token_id = 0

    def __len__(self):
        return self.n

    def add_special_tokens(self, d):
        self.n += len(d.get("additional_special_tokens", []))
        return 1

    def __call__(self, *a, **k):
        return {"input_ids": [[self.n - 1]]}

    def batch_decode(self, *a, **k):
        return [""]


def build_reference_lit(model, vocab_layout, alpha=0.43, label_smoothing=0.1, pretrain=0):
    """The reference LightningModule itself (lit_models/transformer.py:18-57) around `model`.
    The model must ALREADY have the final vocabulary (incl. "[R]"): _init_relation_word is NOT re-run here so
    that seeded weights stay as given."""
    from argparse import Namespace
    _install_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import lit_models.transformer as T  # noqa
    args = Namespace(label_smoothing=label_smoothing, alpha=alpha, pretrain=pretrain, lr=5e-5, optimizer="AdamW",
                     weight_decay=0.01, warm_up_radio=0.1)
    tok = FakeTokenizer(vocab_layout.vocab)
    data_config = dict(analogy_entity_ids=vocab_layout.analogy_entity_ids(),
                       analogy_relation_ids=vocab_layout.analogy_relation_ids(),
                       entity_id_st=vocab_layout.entity_id_st, entity_id_ed=vocab_layout.entity_id_ed,
                       relation_id_st=vocab_layout.relation_id_st, relation_id_ed=vocab_layout.relation_id_ed)
    lit = T.TransformerLitModel(model=model, args=args, tokenizer=tok, data_config=data_config)
    return lit


def load_collator():
    """The reference's data package imported in place (as `ref_data`), for its collator `DataCollatorForSeq2Seq`
    (MarT/data/data_module.py:46-182).  The collator reads the module globals `entities` / `entity2visual_features`
    (set by KGC.__init__, :195-209): callers assign them on the returned module."""
    _install_shims()
    pl = sys.modules["pytorch_lightning"]
    if not hasattr(pl, "LightningDataModule"):
        class _DM:
            def __init__(self, *a, **k):
                pass
        pl.LightningDataModule = _DM
    if "ref_data" not in sys.modules:
        root = os.path.join(REF_ROOT, "data")
        spec = importlib.util.spec_from_file_location("ref_data", os.path.join(root, "__init__.py"), submodule_search_locations=[root])
        pkg = importlib.util.module_from_spec(spec)
        sys.modules["ref_data"] = pkg
        spec.loader.exec_module(pkg)
    return sys.modules["ref_data.data_module"]


def make_bert_tokenizer(tmpdir, n_words=60):
    """A real HF BertTokenizer over a tiny vocabulary ([PAD] = 0, [MASK] = 4): supplies the genuine `tokenizer.pad`
    the reference collator calls (data_module.py:117-123)."""
    from transformers import BertTokenizer
    path = os.path.join(tmpdir, "vocab.txt")
    with open(path, "w") as f:
        f.write("\n".join(["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + [f"w{i}" for i in range(n_words)]))
    return BertTokenizer(path)
