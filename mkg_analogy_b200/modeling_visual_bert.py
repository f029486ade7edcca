"""VisualBERT drop-in (`models.VisualBertKGC`, MarT/models/model.py:13-17 over `VisualBertForMaskedLM`,
MarT/models/modeling_visual_bert.py:1632-1714) with the transformer encoder on the B200 library.

The 12 post-LayerNorm BERT blocks over [text ; 72 region features] run in `EncoderStack(post_ln=True)` (tcgen05 GEMMs, fused
attention with the reference's adaptive scaling, modeling_visual_bert.py:255-260).  Embeddings and the MLM head are thin plumbing on the
library's Linear / LayerNorm kernels (`mkg_analogy_b200.nn`) and keep the reference's `state_dict()` names, so reference checkpoints
load.  Restated from the reference (not copied):
  * embeddings (:72-199): text = word + token-type + position; visual = `visual_projection(pixel_values)` (Faster-RCNN features,
    (B, 72, 2048) from the collator, MarT/data/data_module.py:129-157) + visual token-type (ids default to ONE) + visual position
    (ids all ZERO; MarT never passes `image_text_alignment`); sequence = [text ; visual], ONE LayerNorm + dropout over the whole;
  * attention mask = [attention_mask ; visual_attention_mask] -> additive -1e4 (:828-838), applied inside the fused attention kernel;
  * `sep_idx += img_length` (:875-876) is applied to a copy — the reference mutates the caller's tensor in place.  The scaled blocks are
    rows `img_length+1 : s'` x columns `>= s'` (clamp(w0, 0, .5)) and rows/columns `>= s'` (clamp(w1, .5, 1)) with s' = sep_idx[:, 2] +
    img_length, exactly as the reference indexes them (its indices assume the image tokens come first although the embeddings put them
    last; the drop-in reproduces the reference's arithmetic, not an intention);
  * head `VisualBertOnlyMLMHead` (:507-565) over ALL L + 72 rows: returns (MaskedLMOutput(logits (B, L+72, V)), trans_hidden (B, L+72, H)),
    which is what `TransformerLitModel.training_step` indexes after shifting rel_idx / q_head_idx / a_head_idx by img_len
    (MarT/lit_models/transformer.py:98-102).
`bypass_transformer` (unused by MarT) is not supported and raises.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import nn as B
from .modeling import MaskedLMOutput
from .pretrained import PretrainedMixin, init_hf_style
from .stack import EncoderStack


class _Embeddings(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        H = cfg["hidden_size"]
        self.word_embeddings = nn.Embedding(cfg["vocab_size"], H, padding_idx=cfg["pad_token_id"])
        self.position_embeddings = nn.Embedding(cfg["max_position_embeddings"], H)
        self.token_type_embeddings = nn.Embedding(cfg["type_vocab_size"], H)
        self.LayerNorm = nn.LayerNorm(H, eps=cfg["layer_norm_eps"])
        self.dropout = nn.Dropout(cfg["hidden_dropout_prob"])
        self.register_buffer("position_ids", torch.arange(cfg["max_position_embeddings"]).expand((1, -1)))
        self.visual_token_type_embeddings = nn.Embedding(cfg["type_vocab_size"], H)
        self.visual_position_embeddings = nn.Embedding(cfg["max_position_embeddings"], H)
        self.visual_projection = nn.Linear(cfg["visual_embedding_dim"], H)

    def forward(self, input_ids, token_type_ids, visual_embeds, visual_token_type_ids):
        L = input_ids.shape[1]
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        text = self.word_embeddings(input_ids) + self.token_type_embeddings(token_type_ids) + self.position_embeddings(self.position_ids[:, :L])
        if visual_token_type_ids is None:
            vtype = self.visual_token_type_embeddings.weight[1]                       # ids default to one (:141-144)
        else:
            vtype = self.visual_token_type_embeddings(visual_token_type_ids)
        vis = B.linear(visual_embeds, self.visual_projection.weight, self.visual_projection.bias) \
            + self.visual_position_embeddings.weight[0] + vtype                        # position ids all zero (:187-190)
        x = torch.cat([text, vis], dim=1)
        return self.dropout(B.layer_norm(x, self.LayerNorm))


class _VisualBert(nn.Module):
    def __init__(self, cfg, device):
        super().__init__()
        self.embeddings = _Embeddings(cfg)
        self.encoder = EncoderStack(cfg["hidden_size"], cfg["num_hidden_layers"], cfg["num_attention_heads"], cfg["intermediate_size"],
                                    eps=cfg["layer_norm_eps"], act="gelu", hidden_dropout=cfg["hidden_dropout_prob"],
                                    attn_dropout=cfg["attention_probs_dropout_prob"], device=device,
                                    initializer_range=cfg["initializer_range"], post_ln=True)


class _Transform(nn.Module):
    def __init__(self, hidden, eps):
        super().__init__()
        self.dense = nn.Linear(hidden, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=eps)

    def forward(self, x):
        return B.layer_norm(F.gelu(B.linear(x, self.dense.weight, self.dense.bias)), self.LayerNorm)


class _Predictions(nn.Module):
    def __init__(self, hidden, vocab, eps):
        super().__init__()
        self.transform = _Transform(hidden, eps)
        self.decoder = nn.Linear(hidden, vocab, bias=False)
        self.bias = nn.Parameter(torch.zeros(vocab))
        self.decoder.bias = self.bias


class _Cls(nn.Module):
    def __init__(self, hidden, vocab, eps):
        super().__init__()
        self.predictions = _Predictions(hidden, vocab, eps)


DEFAULTS = dict(vocab_size=30522, hidden_size=768, visual_embedding_dim=2048, num_hidden_layers=12, num_attention_heads=12,
                intermediate_size=3072, max_position_embeddings=512, type_vocab_size=2, layer_norm_eps=1e-12, pad_token_id=1,
                hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, initializer_range=0.02, special_visual_initialize=True,
                bypass_transformer=False)


class VisualBertKGC(PretrainedMixin, nn.Module):
    def __init__(self, config=None, device=None, **overrides):
        super().__init__()
        cfg = dict(DEFAULTS)
        if config is not None:
            cfg.update({k: getattr(config, k) for k in DEFAULTS if hasattr(config, k)})
        cfg.update(overrides)
        if cfg.get("hidden_act", "gelu") != "gelu":
            raise ValueError("VisualBERT uses the erf GELU")
        if cfg["bypass_transformer"]:
            raise NotImplementedError("config.bypass_transformer is not part of the MarT path")
        self.cfg = cfg
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"
        self.visual_bert = _VisualBert(cfg, device)
        self.cls = _Cls(cfg["hidden_size"], cfg["vocab_size"], cfg["layer_norm_eps"])
        init_hf_style(self, cfg["initializer_range"])
        emb = self.visual_bert.embeddings
        if cfg["special_visual_initialize"]:                     # :98-104
            with torch.no_grad():
                emb.visual_token_type_embeddings.weight.copy_(emb.token_type_embeddings.weight)
                emb.visual_position_embeddings.weight.copy_(emb.position_embeddings.weight)
        self.tie_weights()
        for m in (emb, self.cls):
            m.to(device)

    @staticmethod
    def add_to_argparse(parser):
        parser.add_argument("--pretrain", type=int, default=0, help="")
        return parser

    # ---- HF-style embedding API used by TransformerLitModel (lit_models/transformer.py:33,43-54)
    def get_input_embeddings(self):
        return self.visual_bert.embeddings.word_embeddings

    def get_output_embeddings(self):
        return self.cls.predictions.decoder

    def tie_weights(self):
        self.cls.predictions.decoder.weight = self.get_input_embeddings().weight

    def resize_token_embeddings(self, n):
        old = self.get_input_embeddings()
        if n is None or n == old.num_embeddings:
            return old
        new = nn.Embedding(n, old.embedding_dim, padding_idx=old.padding_idx).to(old.weight.device)
        pr = self.cls.predictions
        with torch.no_grad():
            new.weight.normal_(0.0, self.cfg["initializer_range"])
            k = min(n, old.num_embeddings)
            new.weight[:k] = old.weight[:k]
            bias = torch.zeros(n, device=old.weight.device)
            bias[:k] = pr.bias[:k]
        self.visual_bert.embeddings.word_embeddings = new
        pr.decoder = nn.Linear(old.embedding_dim, n, bias=False).to(old.weight.device)
        pr.bias = nn.Parameter(bias)
        pr.decoder.bias = pr.bias
        self.tie_weights()
        self.cfg["vocab_size"] = n
        return new

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, pixel_values=None, visual_attention_mask=None,
                visual_token_type_ids=None, sep_idx=None, return_dict=None, output_logits=True, **unused):
        """pixel_values: the region features (B, n_regions, visual_embedding_dim).  output_logits=False skips the dense
        (B, L + n_regions, V) decoder product (the fused route scores the [MASK] rows only)."""
        emb = self.visual_bert.embeddings
        dev = emb.word_embeddings.weight.device
        input_ids = input_ids.to(dev)
        Bn, L = input_ids.shape
        vis = pixel_values.to(dev, torch.float32)
        img_length = vis.shape[1]
        if attention_mask is None:
            attention_mask = torch.ones(Bn, L, dtype=torch.long, device=dev)
        if visual_attention_mask is None:
            visual_attention_mask = torch.ones(Bn, img_length, dtype=torch.long, device=dev)
        x = emb(input_ids, None if token_type_ids is None else token_type_ids.to(dev), vis,
                None if visual_token_type_ids is None else visual_token_type_ids.to(dev))
        mask = torch.cat([attention_mask.to(dev).long(), visual_attention_mask.to(dev).long()], dim=1)
        sep = None
        if sep_idx is not None:
            sep = sep_idx.to(dev).clone()
            sep += img_length
        y = self.visual_bert.encoder(x, attention_mask=mask, sep_idx=sep, row_lo=img_length + 1)
        pr = self.cls.predictions
        trans = pr.transform(y)
        logits = B.linear(trans, pr.decoder.weight, pr.bias) if output_logits else None
        return MaskedLMOutput(logits=logits), trans
