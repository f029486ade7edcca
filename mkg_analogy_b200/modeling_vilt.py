"""ViLT drop-in (`models.ViltKGC`, MarT/models/model.py:27-31 over `ViltForMaskedLM`, MarT/models/modeling_vilt.py:896-968)
with the transformer encoder on the B200 library.

The 12 pre-LN blocks over [text ; image 1 ; image 2] — all of the model's attention and MLP work — run in
`EncoderStack` (tcgen05 GEMMs, fused attention with the reference's adaptive scaling on rows `img_length+1 : s'`,
modeling_vilt.py:370-375).  Embeddings and the MLM head are thin PyTorch plumbing around it and keep the reference's
`state_dict()` names, so reference checkpoints load.  Restated from the reference (not copied):
  * text embeddings (:246-302): word + type + position embeddings for the first `max_position_embeddings` (40) tokens only;
  * image embeddings (:112-196, :214-243): conv patches + cls, position embeddings, the SAME modules for both images;
    sequence = [text, image 1, image 2]; modality embeddings 0 (text) / 1 (image).
    With the all-ones `pixel_mask` the reference builds when none is given (:834-835), its bilinear position-embedding resize
    is the identity for images of `config.image_size` (other sizes: the same `interpolate(..., align_corners=True)` to the
    actual patch grid) and its `torch.multinomial` "selection" is a random PERMUTATION of the patches of each image; self-attention
    is permutation-equivariant and only the text rows are read (:950), so WITHOUT `sep_idx` (pre-training) the permutation can be
    omitted with equal outputs.  WITH `sep_idx` (fine-tuning) the adaptive scaling indexes rows `img_length+1 : s'` and columns
    `>= s'` of the [text ; image] sequence (:373-375), which reaches image-token columns: the reference then scales a RANDOM
    subset of the image-2 patches (whatever the permutation put last), this drop-in always the last raster-order patches, so
    parity there holds up to the reference's own random patch order, not bit for bit.  A partial `pixel_mask` and a positive
    `config.max_image_length` smaller than the patch count (which sub-samples patches in the reference, :166-178) are not
    supported and raise;
  * `sep_idx + img_length` (:843-844) is applied to a copy — the reference mutates the caller's tensor in place;
  * final LayerNorm, first L rows (:950), head `ViltMLMHead` (:988-1004): returns (MaskedLMOutput(logits), trans_hidden).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import nn as B
from .modeling import MaskedLMOutput
from .pretrained import PretrainedMixin, init_hf_style
from .stack import EncoderStack


class _TextEmbeddings(nn.Module):
    def __init__(self, vocab, hidden, max_pos, type_vocab, eps, p):
        super().__init__()
        self.word_embeddings = nn.Embedding(vocab, hidden)
        self.position_embeddings = nn.Embedding(max_pos, hidden)
        self.token_type_embeddings = nn.Embedding(type_vocab, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=eps)
        self.dropout = nn.Dropout(p)
        self.register_buffer("position_ids", torch.arange(max_pos).expand((1, -1)))

    def forward(self, input_ids, token_type_ids):
        L = input_ids.shape[1]
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        e = self.word_embeddings(input_ids) + self.token_type_embeddings(token_type_ids)
        n = min(L, self.position_embeddings.num_embeddings)      # only the first 40 positions get one (:296-299)
        pos = self.position_embeddings(self.position_ids[:, :n])
        e = torch.cat([e[:, :n] + pos, e[:, n:]], dim=1)
        return self.dropout(B.layer_norm(e, self.LayerNorm))


class _PatchEmbeddings(nn.Module):
    def __init__(self, hidden, patch, channels):
        super().__init__()
        self.projection = nn.Conv2d(channels, hidden, kernel_size=patch, stride=patch)


class _Embeddings(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        H = cfg["hidden_size"]
        self.text_embeddings = _TextEmbeddings(cfg["vocab_size"], H, cfg["max_position_embeddings"], cfg["type_vocab_size"],
                                               cfg["layer_norm_eps"], cfg["hidden_dropout_prob"])
        self.cls_token = nn.Parameter(torch.zeros(1, 1, H))
        self.patch_embeddings = _PatchEmbeddings(H, cfg["patch_size"], cfg["num_channels"])
        g = cfg["image_size"] // cfg["patch_size"]
        self.position_embeddings = nn.Parameter(torch.zeros(1, g * g + 1, H))
        self.token_type_embeddings = nn.Embedding(cfg["modality_type_vocab_size"], H)
        self.dropout = nn.Dropout(cfg["hidden_dropout_prob"])

    def image_tokens(self, pixels):
        x = B.patch_embed(pixels, self.patch_embeddings.projection)                      # (B, g*g, H), row-major patches; im2col + tcgen05 GEMM
        pos = self.position_embeddings
        g_cfg = int(round((pos.shape[1] - 1) ** 0.5))
        gh, gw = pixels.shape[2] // self.patch_embeddings.projection.kernel_size[0], pixels.shape[3] // self.patch_embeddings.projection.kernel_size[1]
        if (gh, gw) != (g_cfg, g_cfg):
            # images of another size than config.image_size: the patch position table is resized bilinearly to the actual patch
            # grid (modeling_vilt.py:121-136; with the all-ones pixel_mask the valid extent is the whole grid)
            sp = pos[:, 1:].transpose(1, 2).reshape(1, pos.shape[2], g_cfg, g_cfg)
            sp = F.interpolate(sp, size=(gh, gw), mode="bilinear", align_corners=True)
            pos = torch.cat([pos[:, :1], sp.flatten(2).transpose(1, 2)], dim=1)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1) + pos
        return self.dropout(x)


class _Pooler(nn.Module):
    def __init__(self, hidden):
        super().__init__()
        self.dense = nn.Linear(hidden, hidden)


class _Transform(nn.Module):
    def __init__(self, hidden, eps):
        super().__init__()
        self.dense = nn.Linear(hidden, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=eps)

    def forward(self, x):
        return B.layer_norm(F.gelu(B.linear(x, self.dense.weight, self.dense.bias)), self.LayerNorm)


class _MLMHead(nn.Module):
    def __init__(self, hidden, vocab, eps):
        super().__init__()
        self.transform = _Transform(hidden, eps)
        self.decoder = nn.Linear(hidden, vocab, bias=False)
        self.bias = nn.Parameter(torch.zeros(vocab))
        self.decoder.bias = self.bias


class _Vilt(nn.Module):
    def __init__(self, cfg, device):
        super().__init__()
        H = cfg["hidden_size"]
        self.embeddings = _Embeddings(cfg)
        self.encoder = EncoderStack(H, cfg["num_hidden_layers"], cfg["num_attention_heads"], cfg["intermediate_size"],
                                    eps=cfg["layer_norm_eps"], act="gelu", hidden_dropout=cfg["hidden_dropout_prob"],
                                    attn_dropout=cfg["attention_probs_dropout_prob"], device=device)
        self.layernorm = nn.LayerNorm(H, eps=cfg["layer_norm_eps"])
        self.pooler = _Pooler(H)        # parameters kept for checkpoint compatibility; its output is never consumed by MarT


DEFAULTS = dict(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                max_position_embeddings=40, type_vocab_size=2, modality_type_vocab_size=2, image_size=384, patch_size=32,
                num_channels=3, layer_norm_eps=1e-12, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                initializer_range=0.02)


class ViltKGC(PretrainedMixin, nn.Module):
    def __init__(self, config=None, device=None, **overrides):
        super().__init__()
        cfg = dict(DEFAULTS)
        if config is not None:
            cfg.update({k: getattr(config, k) for k in DEFAULTS if hasattr(config, k)})
        cfg.update(overrides)
        if cfg.get("hidden_act", "gelu") != "gelu":
            raise ValueError("ViLT uses the erf GELU")
        mil = getattr(config, "max_image_length", -1) if config is not None else overrides.get("max_image_length", -1)
        n_patches = (cfg["image_size"] // cfg["patch_size"]) ** 2
        if mil is not None and 0 < mil < n_patches:
            raise NotImplementedError(f"config.max_image_length = {mil} < {n_patches} patches: the reference then sub-samples patches "
                                      "(modeling_vilt.py:166-178); this drop-in keeps every patch")
        self.cfg = cfg
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"
        self.vilt = _Vilt(cfg, device)
        self.mlm_score = _MLMHead(cfg["hidden_size"], cfg["vocab_size"], cfg["layer_norm_eps"])
        init_hf_style(self, cfg["initializer_range"])
        self.tie_weights()
        for m in (self.vilt.embeddings, self.vilt.layernorm, self.vilt.pooler, self.mlm_score):
            m.to(device)

    @staticmethod
    def add_to_argparse(parser):
        parser.add_argument("--pretrain", type=int, default=0, help="")
        return parser

    # ---- HF-style embedding API used by TransformerLitModel (lit_models/transformer.py:33,43-54)
    def get_input_embeddings(self):
        return self.vilt.embeddings.text_embeddings.word_embeddings

    def get_output_embeddings(self):
        return self.mlm_score.decoder

    def tie_weights(self):
        self.mlm_score.decoder.weight = self.get_input_embeddings().weight      # config.tie_word_embeddings (main.py:121)

    def resize_token_embeddings(self, n):
        old = self.get_input_embeddings()
        if n is None or n == old.num_embeddings:
            return old
        new = nn.Embedding(n, old.embedding_dim).to(old.weight.device)
        with torch.no_grad():
            new.weight.normal_(0.0, self.cfg["initializer_range"])
            k = min(n, old.num_embeddings)
            new.weight[:k] = old.weight[:k]
            bias = torch.zeros(n, device=old.weight.device)
            bias[:k] = self.mlm_score.bias[:k]
        self.vilt.embeddings.text_embeddings.word_embeddings = new
        self.mlm_score.decoder = nn.Linear(old.embedding_dim, n, bias=False).to(old.weight.device)
        self.mlm_score.bias = nn.Parameter(bias)
        self.mlm_score.decoder.bias = self.mlm_score.bias
        self.tie_weights()
        self.cfg["vocab_size"] = n
        return new

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, pixel_values=None, pixel_mask=None, sep_idx=None,
                return_dict=None, output_logits=True, **unused):
        """output_logits=False skips the dense (B, L, V) decoder product (the fused route scores the [MASK] rows only)."""
        if pixel_mask is not None and not bool(pixel_mask.all()):
            raise NotImplementedError("partial pixel masks are not part of the MarT path")
        emb = self.vilt.embeddings
        dev = emb.cls_token.device
        input_ids = input_ids.to(dev)
        Bn, L = input_ids.shape
        if attention_mask is None:
            attention_mask = torch.ones(Bn, L, dtype=torch.long, device=dev)
        text = emb.text_embeddings(input_ids, None if token_type_ids is None else token_type_ids.to(dev))
        px = pixel_values.to(dev, torch.float32)
        image = torch.cat([emb.image_tokens(px[:, 0]), emb.image_tokens(px[:, 1])], dim=1)
        text = text + emb.token_type_embeddings.weight[0]
        image = image + emb.token_type_embeddings.weight[1]
        x = torch.cat([text, image], dim=1)
        mask = torch.cat([attention_mask.to(dev, torch.long), torch.ones(Bn, image.shape[1], dtype=torch.long, device=dev)], dim=1)
        img_length = image.shape[1]
        sep = None
        if sep_idx is not None:
            sep = sep_idx.to(dev).clone()
            sep[:, 2] += img_length
        y = self.vilt.encoder(x, attention_mask=mask, sep_idx=sep, row_lo=img_length + 1)
        seq = B.layer_norm(y[:, :L].contiguous(), self.vilt.layernorm)          # LayerNorm is per row: the text rows only
        trans = self.mlm_score.transform(seq)
        logits = B.linear(trans, self.mlm_score.decoder.weight, self.mlm_score.bias) if output_logits else None
        return MaskedLMOutput(logits=logits), trans
