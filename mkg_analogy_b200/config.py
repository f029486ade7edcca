"""Shapes of the MKGformer ("Unimo") dual-stream encoder, SURVEY.md §8.

Field meanings follow the reference's two HF configs: `BertConfig` for the text stream and
`CLIPVisionConfig` for the vision stream (MarT/main.py:76-86).  `fuse_from` is the hard-coded layer
index 8 of MarT/models/modeling_unimo.py:616,627 (K/V export starts at fuse_from-1, :628).
"""
from dataclasses import dataclass, asdict


@dataclass(frozen=True)
class MKGformerConfig:
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    intermediate: int = 3072
    vocab: int = 42007            # 30522 + 11292 entities + 192 relations + "[R]"
    max_position: int = 512
    type_vocab: int = 2
    image_size: int = 224
    patch: int = 32
    fuse_from: int = 8
    text_eps: float = 1e-12
    vision_eps: float = 1e-5
    hidden_dropout: float = 0.1   # BertConfig.hidden_dropout_prob
    attn_dropout: float = 0.1     # BertConfig.attention_probs_dropout_prob

    @property
    def head_dim(self):
        return self.hidden // self.heads

    @property
    def patches(self):
        return (self.image_size // self.patch) ** 2

    @property
    def vision_tokens(self):
        return 1 + 2 * self.patches

    @property
    def patch_dim(self):
        return 3 * self.patch * self.patch

    def to_dict(self):
        return asdict(self)


BASE = MKGformerConfig()
LARGE = MKGformerConfig(hidden=1024, layers=24, heads=16, intermediate=4096, patch=14)
# small shape used by the golden fixtures (same code paths: 10 layers so that layers 8, 9 fuse)
TINY = MKGformerConfig(hidden=128, layers=10, heads=2, intermediate=256, vocab=691, max_position=64,
                       image_size=64, patch=32)


def param_shapes(cfg: MKGformerConfig):
    """Ordered {state_dict name: shape} of the reference `UnimoForMaskedLM` (SURVEY.md §8a state-dict contract;
    MarT/models/modeling_unimo.py:100-186, :189-287, :290-478, :661-673, :942-976)."""
    H, F, V = cfg.hidden, cfg.intermediate, cfg.vocab
    s = {}
    s["unimo.vision_embeddings.class_embedding"] = (H,)
    s["unimo.vision_embeddings.patch_embedding.weight"] = (H, 3, cfg.patch, cfg.patch)
    s["unimo.vision_embeddings.position_embedding.weight"] = (cfg.patches + 1, H)
    s["unimo.vision_pre_layrnorm.weight"] = (H,)
    s["unimo.vision_pre_layrnorm.bias"] = (H,)
    s["unimo.vision_post_layernorm.weight"] = (H,)
    s["unimo.vision_post_layernorm.bias"] = (H,)
    s["unimo.text_embeddings.word_embeddings.weight"] = (V, H)
    s["unimo.text_embeddings.position_embeddings.weight"] = (cfg.max_position, H)
    s["unimo.text_embeddings.token_type_embeddings.weight"] = (cfg.type_vocab, H)
    s["unimo.text_embeddings.LayerNorm.weight"] = (H,)
    s["unimo.text_embeddings.LayerNorm.bias"] = (H,)
    for i in range(cfg.layers):
        p = f"unimo.encoder.vision_layers.{i}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            s[p + f"self_attn.{n}.weight"] = (H, H)
            s[p + f"self_attn.{n}.bias"] = (H,)
        s[p + "layer_norm1.weight"] = (H,)
        s[p + "layer_norm1.bias"] = (H,)
        s[p + "mlp.fc1.weight"] = (F, H)
        s[p + "mlp.fc1.bias"] = (F,)
        s[p + "mlp.fc2.weight"] = (H, F)
        s[p + "mlp.fc2.bias"] = (H,)
        s[p + "layer_norm2.weight"] = (H,)
        s[p + "layer_norm2.bias"] = (H,)
    for i in range(cfg.layers):
        p = f"unimo.encoder.text_layer.{i}."
        for n in ("query", "key", "value"):
            s[p + f"attention.self.{n}.weight"] = (H, H)
            s[p + f"attention.self.{n}.bias"] = (H,)
        s[p + "attention.self.adaptive_weight.0"] = (1,)
        s[p + "attention.self.adaptive_weight.1"] = (1,)
        s[p + "attention.output.dense.weight"] = (H, H)
        s[p + "attention.output.dense.bias"] = (H,)
        s[p + "attention.output.LayerNorm.weight"] = (H,)
        s[p + "attention.output.LayerNorm.bias"] = (H,)
        s[p + "intermediate.dense.weight"] = (F, H)
        s[p + "intermediate.dense.bias"] = (F,)
        s[p + "intermediate.fusion_dense.weight"] = (F, H)
        s[p + "intermediate.fusion_dense.bias"] = (F,)
        s[p + "output.dense.weight"] = (H, F)
        s[p + "output.dense.bias"] = (H,)
        s[p + "output.LayerNorm.weight"] = (H,)
        s[p + "output.LayerNorm.bias"] = (H,)
    s["unimo.text_pooler.dense.weight"] = (H, H)
    s["unimo.text_pooler.dense.bias"] = (H,)
    s["cls.predictions.bias"] = (V,)
    s["cls.predictions.transform.dense.weight"] = (H, H)
    s["cls.predictions.transform.dense.bias"] = (H,)
    s["cls.predictions.transform.LayerNorm.weight"] = (H,)
    s["cls.predictions.transform.LayerNorm.bias"] = (H,)
    return s


def init_params(cfg: MKGformerConfig, seed: int = 0, device="cpu"):
    """Seeded synthetic weights: normal(0, 0.02) for Linear/Embedding/conv, LayerNorm (1, 0),
    class_embedding ~ N(0, 1), adaptive weights as the reference initialises them
    (modeling_unimo.py:305-310: U(0, 0.5) and exactly 0.5). Returns {name: fp32 tensor}. Deterministic on CPU."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    out = {}
    for name, shape in param_shapes(cfg).items():
        if "LayerNorm.weight" in name or "layer_norm1.weight" in name or "layer_norm2.weight" in name or \
                name.endswith("layrnorm.weight") or name.endswith("layernorm.weight"):
            t = torch.ones(shape)
        elif name.endswith("adaptive_weight.0"):
            t = torch.rand(shape, generator=g) * 0.5
        elif name.endswith("adaptive_weight.1"):
            t = torch.full(shape, 0.5)
        elif name.endswith("class_embedding"):
            t = torch.randn(shape, generator=g)
        elif name.endswith(".bias") and "predictions.bias" not in name:
            # non-zero biases so that bias handling is actually exercised by parity tests
            t = torch.randn(shape, generator=g) * 0.02
        elif name == "cls.predictions.bias":
            t = torch.randn(shape, generator=g) * 0.02
        else:
            t = torch.randn(shape, generator=g) * 0.02
        out[name] = t.to(device)
    return out
