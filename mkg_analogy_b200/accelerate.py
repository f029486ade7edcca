"""Plug the B200 encoder stack into the reference's FLAVA / ViLT / VisualBERT models.

The reference's `FlavaKGC` / `ViltKGC` (MarT/models/model.py:27-37) are HF-style modules whose encoders are
`FlavaEncoder` (modeling_flava.py:668-722; three instances: image, text, multimodal) and `ViltEncoder`
(modeling_vilt.py:531-583) — over 90 % of their FLOPs.  `swap_encoders(model)` replaces every such encoder inside an
already constructed reference model by `EncoderAdapter`, an `EncoderStack` that keeps the encoder's call signature and
return type (`encoder(hidden_states, attention_mask=<additive 4-D mask>, ..., sep_idx=..., img_length=...)` ->
`BaseModelOutput`-like), its `state_dict()` names and its weights.  Embeddings, poolers, projections and the MLM head stay the
reference's own PyTorch modules (they run on the GPU as they do today); only the encoders change.  `VisualBertEncoder`
(modeling_visual_bert.py:424-488, post-LayerNorm BERT blocks carrying the reference's `adaptive_weight`s) is swapped the same way.

    model = FlavaKGC.from_pretrained(...)          # reference code, unchanged
    model = swap_encoders(model).cuda()
"""
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn

from .stack import EncoderStack


@dataclass
class EncoderOutput:
    """BaseModelOutput-like: indexable ([0] = last_hidden_state) and attribute access, as the reference parents use it."""
    last_hidden_state: torch.Tensor
    hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    attentions: Optional[Tuple[torch.Tensor, ...]] = None

    def __getitem__(self, i):
        return tuple(v for v in (self.last_hidden_state, self.hidden_states, self.attentions) if v is not None)[i]


class EncoderAdapter(EncoderStack):
    """EncoderStack with the forward signature of FlavaEncoder.forward (modeling_flava.py:680-722) and
    ViltEncoder.forward (modeling_vilt.py:541-583)."""

    def __init__(self, *a, scale_rows_from=1, **k):
        super().__init__(*a, **k)
        self.scale_rows_from = scale_rows_from      # FLAVA text encoder: rows 1:s (modeling_flava.py:494)

    def forward(self, hidden_states, attention_mask=None, head_mask=None, output_attentions=False, output_hidden_states=False,
                return_dict=True, sep_idx=None, img_length=None):
        if output_attentions:
            raise NotImplementedError("attention probabilities never leave the SM in the fused kernel")
        if head_mask is not None and any(h is not None for h in head_mask):
            raise NotImplementedError("head masks are not part of the MarT path (get_head_mask -> [None] * n)")
        key_mask = None
        if attention_mask is not None:
            # parents pass the additive extension (B, 1, 1, T) = (1 - mask) * -10000 (get_extended_attention_mask);
            # an all-zero extension (FLAVA multimodal encoder: all-ones mask) needs no masking at all
            m = attention_mask
            if m.dim() == 4:
                m = m[:, 0, 0, :]
            key_mask = (m > -1.0).to(torch.int64)
            if bool(key_mask.all()):
                key_mask = None
        row_lo = self.scale_rows_from if img_length is None else int(img_length) + 1    # modeling_vilt.py:373
        y = super().forward(hidden_states, attention_mask=key_mask, sep_idx=sep_idx, row_lo=row_lo)
        out = EncoderOutput(last_hidden_state=y, hidden_states=(y,) if output_hidden_states else None)
        return out if return_dict else tuple(v for v in (out.last_hidden_state, out.hidden_states) if v is not None)


def _is_preln_encoder(m):
    """A FlavaEncoder / ViltEncoder: `.layer` ModuleList of blocks with layernorm_before / layernorm_after."""
    layers = getattr(m, "layer", None)
    return isinstance(layers, nn.ModuleList) and len(layers) > 0 and hasattr(layers[0], "layernorm_before") and \
        hasattr(layers[0], "layernorm_after") and hasattr(layers[0], "intermediate")


def _is_postln_encoder(m):
    """A VisualBertEncoder: `.layer` ModuleList of BERT blocks whose self-attention carries the reference's adaptive weights
    (a plain BertEncoder has none and is left alone)."""
    layers = getattr(m, "layer", None)
    if not (isinstance(layers, nn.ModuleList) and len(layers) > 0):
        return False
    att = getattr(getattr(layers[0], "attention", None), "self", None)
    return att is not None and hasattr(att, "adaptive_weight") and hasattr(layers[0].attention, "output") and \
        hasattr(layers[0].attention.output, "LayerNorm") and hasattr(layers[0], "intermediate") and hasattr(layers[0].output, "LayerNorm")


def adapter_from(encoder, device=None):
    """EncoderAdapter with the shapes and weights of a reference FlavaEncoder / ViltEncoder / VisualBertEncoder instance."""
    l0 = encoder.layer[0]
    if _is_postln_encoder(encoder):
        att = l0.attention.self
        ad = EncoderAdapter(att.query.in_features, len(encoder.layer), att.num_attention_heads, l0.intermediate.dense.out_features,
                            eps=l0.output.LayerNorm.eps, act="gelu", hidden_dropout=l0.output.dropout.p, attn_dropout=att.dropout.p,
                            device=device if device is not None else att.query.weight.device, post_ln=True)
        ad.load_state_dict(encoder.state_dict(), strict=True)
        ad.train(encoder.training)
        return ad
    att = l0.attention.attention
    hidden = att.query.in_features
    act = l0.intermediate.intermediate_act_fn
    act_name = "quick_gelu" if "quick" in type(act).__name__.lower() else "gelu"
    ad = EncoderAdapter(hidden, len(encoder.layer), att.num_attention_heads, l0.intermediate.dense.out_features,
                        eps=l0.layernorm_before.eps, act=act_name, hidden_dropout=l0.output.dropout.p,
                        attn_dropout=att.dropout.p, device=device if device is not None else att.query.weight.device)
    ad.load_state_dict(encoder.state_dict(), strict=True)
    ad.train(encoder.training)
    return ad


def swap_encoders(model, device=None):
    """Replace every FlavaEncoder / ViltEncoder / VisualBertEncoder inside `model` (in place) by an EncoderAdapter carrying its weights.
    Returns the model; `model.state_dict()` keys are unchanged."""
    for name, child in list(model.named_children()):
        if _is_preln_encoder(child) or _is_postln_encoder(child):
            setattr(model, name, adapter_from(child, device))
        else:
            swap_encoders(child, device)
    return model
