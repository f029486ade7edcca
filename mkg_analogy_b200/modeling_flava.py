"""FLAVA drop-in (`models.FlavaKGC`, MarT/models/model.py:33-37 over `FlavaForMaskedLM`, MarT/models/modeling_flava.py:2127-2204)
with its three transformer encoders on the B200 library.

image encoder (12 pre-LN blocks over 1 + 2 x patches tokens), text encoder (12 blocks, adaptive analogy scaling on rows 1:s,
:491-496) and multimodal encoder (6 blocks over [cls ; image ; text], all-ones mask) are `EncoderStack`s — all of the model's
attention and MLP work.  Embeddings, the two projections and the MLM head are thin PyTorch plumbing with the reference's
`state_dict()` names.  Restated from the reference (not copied), quirks kept:
  * image embeddings (:308-343): both images through the same conv; [cls ; image 1] + position_embeddings, image 2 +
    position_embeddings[:, :patches] (i.e. the cls slot and the first patches-1 patch slots), no cls for image 2;
  * text embeddings (:380-440): word (padding_idx = pad_token_id until the table is resized) + position + type, LayerNorm;
  * `image_to_mm_projection` / `text_to_mm_projection` read the encoders' LAST HIDDEN STATES BEFORE the final LayerNorm
    (:1430, :1450); the final LayerNorms of the image / text models only feed outputs MarT never reads;
  * multimodal encoder input = [cls ; image_mm ; text_mm], no padding mask (:1187-1188, :1456); final LayerNorm; the last L
    rows go to the head (:2187); returns (MaskedLMOutput(logits), trans_hidden_states) (:2199-2204).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import nn as B
from .modeling import MaskedLMOutput
from .pretrained import PretrainedMixin, init_hf_style
from .stack import EncoderStack

DEFAULTS = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072, layer_norm_eps=1e-12,
                hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)


def _sub(cfg, **extra):
    d = dict(DEFAULTS)
    d.update(extra)
    if cfg is not None:
        src = cfg if isinstance(cfg, dict) else {k: getattr(cfg, k) for k in list(d) if hasattr(cfg, k)}
        d.update({k: v for k, v in src.items() if k in d})
    return d


def _stack(c, device):
    return EncoderStack(c["hidden_size"], c["num_hidden_layers"], c["num_attention_heads"], c["intermediate_size"],
                        eps=c["layer_norm_eps"], act="gelu", hidden_dropout=c["hidden_dropout_prob"],
                        attn_dropout=c["attention_probs_dropout_prob"], device=device)


class _Pooler(nn.Module):
    def __init__(self, hidden):
        super().__init__()
        self.dense = nn.Linear(hidden, hidden)      # kept for checkpoint compatibility; never consumed by MarT


class _Patch(nn.Module):
    def __init__(self, hidden, patch, channels):
        super().__init__()
        self.projection = nn.Conv2d(channels, hidden, kernel_size=patch, stride=patch)


class _ImageEmbeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        H = c["hidden_size"]
        self.cls_token = nn.Parameter(torch.zeros(1, 1, H))
        self.mask_token = nn.Parameter(torch.zeros(1, 1, H))
        n = (c["image_size"] // c["patch_size"]) ** 2
        self.position_embeddings = nn.Parameter(torch.zeros(1, n + 1, H))
        self.patch_embeddings = _Patch(H, c["patch_size"], c["num_channels"])
        self.dropout = nn.Dropout(c["hidden_dropout_prob"])

    def forward(self, px):
        head = B.patch_embed(px[:, 0], self.patch_embeddings.projection)       # im2col + tcgen05 GEMM (no cuDNN)
        tail = B.patch_embed(px[:, 1], self.patch_embeddings.projection)
        x = torch.cat([self.cls_token.expand(head.shape[0], -1, -1), head], dim=1) + self.position_embeddings
        tail = tail + self.position_embeddings[:, :tail.shape[1]]
        return self.dropout(torch.cat([x, tail], dim=1))


class _TextEmbeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        H = c["hidden_size"]
        self.word_embeddings = nn.Embedding(c["vocab_size"], H, padding_idx=c["pad_token_id"])   # :388; resize drops it (4.19)
        self.position_embeddings = nn.Embedding(c["max_position_embeddings"], H)
        self.token_type_embeddings = nn.Embedding(c["type_vocab_size"], H)
        self.LayerNorm = nn.LayerNorm(H, eps=c["layer_norm_eps"])
        self.dropout = nn.Dropout(c["hidden_dropout_prob"])
        self.register_buffer("position_ids", torch.arange(c["max_position_embeddings"]).expand((1, -1)))

    def forward(self, ids, tt):
        L = ids.shape[1]
        if tt is None:
            tt = torch.zeros_like(ids)
        e = self.word_embeddings(ids) + self.token_type_embeddings(tt) + self.position_embeddings(self.position_ids[:, :L])
        return self.dropout(B.layer_norm(e, self.LayerNorm))


class _Tower(nn.Module):
    def __init__(self, c, device, embeddings=None, cls_token=False):
        super().__init__()
        if embeddings is not None:
            self.embeddings = embeddings
        if cls_token:
            self.cls_token = nn.Parameter(torch.zeros(1, 1, c["hidden_size"]))
        self.encoder = _stack(c, device)
        self.layernorm = nn.LayerNorm(c["hidden_size"], eps=c["layer_norm_eps"])
        self.pooler = _Pooler(c["hidden_size"])


class _Flava(nn.Module):
    def __init__(self, ic, tc, mc, hidden, proj, device):
        super().__init__()
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.6592)
        self.text_model = _Tower(tc, device, _TextEmbeddings(tc))
        self.image_model = _Tower(ic, device, _ImageEmbeddings(ic))
        self.multimodal_model = _Tower(mc, device, cls_token=True)
        self.image_projection = nn.Linear(ic["hidden_size"], proj)      # contrastive heads: unused by FlavaForMaskedLM
        self.text_projection = nn.Linear(tc["hidden_size"], proj)
        self.image_to_mm_projection = nn.Linear(ic["hidden_size"], hidden)
        self.text_to_mm_projection = nn.Linear(tc["hidden_size"], hidden)


class _Transform(nn.Module):
    def __init__(self, hidden, eps):
        super().__init__()
        self.dense = nn.Linear(hidden, hidden)
        self.LayerNorm = nn.LayerNorm(hidden, eps=eps)

    def forward(self, x):
        return B.layer_norm(F.gelu(B.linear(x, self.dense.weight, self.dense.bias)), self.LayerNorm)


class _Head(nn.Module):
    def __init__(self, hidden, vocab, eps):
        super().__init__()
        self.transform = _Transform(hidden, eps)
        self.decoder = nn.Linear(hidden, vocab, bias=False)
        self.bias = nn.Parameter(torch.zeros(vocab))
        self.decoder.bias = self.bias


class FlavaKGC(PretrainedMixin, nn.Module):
    def __init__(self, config=None, device=None, image_config=None, text_config=None, multimodal_config=None, hidden_size=None,
                 projection_dim=None):
        super().__init__()
        if config is not None:
            image_config, text_config, multimodal_config = config.image_config, config.text_config, config.multimodal_config
            hidden_size, projection_dim = config.hidden_size, config.projection_dim
        ic = _sub(image_config, image_size=224, patch_size=16, num_channels=3)
        tc = _sub(text_config, vocab_size=30522, max_position_embeddings=512, type_vocab_size=2, initializer_range=0.02,
                  pad_token_id=0)
        mc = _sub(multimodal_config, num_hidden_layers=6)
        hidden_size = hidden_size or mc["hidden_size"]
        if mc["hidden_size"] != hidden_size:
            raise ValueError("multimodal hidden size must equal config.hidden_size")
        self.ic, self.tc, self.mc = ic, tc, mc
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"
        self.flava = _Flava(ic, tc, mc, hidden_size, projection_dim or 768, device)
        self.cls = _Head(tc["hidden_size"], tc["vocab_size"], tc["layer_norm_eps"])
        init_hf_style(self, tc["initializer_range"])
        self.tie_weights()
        self.to(device)          # EncoderStack._apply is a no-op for its own device

    @staticmethod
    def add_to_argparse(parser):
        parser.add_argument("--pretrain", type=int, default=0, help="")
        return parser

    def get_input_embeddings(self):
        return self.flava.text_model.embeddings.word_embeddings

    def get_output_embeddings(self):
        return self.cls.decoder

    def tie_weights(self):
        self.cls.decoder.weight = self.get_input_embeddings().weight

    def resize_token_embeddings(self, n):
        """4.19 semantics: copy old rows, new rows ~ N(0, initializer_range), zero-padded bias, tied."""
        old = self.get_input_embeddings()
        if n is None or n == old.num_embeddings:
            return old
        dev = old.weight.device
        new = nn.Embedding(n, old.embedding_dim).to(dev)
        with torch.no_grad():
            new.weight.normal_(0.0, self.tc["initializer_range"])
            k = min(n, old.num_embeddings)
            new.weight[:k] = old.weight[:k]
            bias = torch.zeros(n, device=dev)
            bias[:k] = self.cls.bias[:k]
        self.flava.text_model.embeddings.word_embeddings = new
        self.cls.decoder = nn.Linear(old.embedding_dim, n, bias=False).to(dev)
        self.cls.bias = nn.Parameter(bias)
        self.cls.decoder.bias = self.cls.bias
        self.tie_weights()
        self.tc["vocab_size"] = n
        return new

    def forward(self, input_ids=None, pixel_values=None, attention_mask=None, token_type_ids=None, sep_idx=None, return_dict=None,
                output_logits=True, **unused):
        """output_logits=False skips the dense (B, L, V) decoder product (the fused route scores the [MASK] rows only)."""
        fl = self.flava
        dev = fl.multimodal_model.cls_token.device
        ids = input_ids.to(dev)
        Bn, L = ids.shape
        img = fl.image_model.encoder(fl.image_model.embeddings(pixel_values.to(dev, torch.float32)))            # pre-final-LN
        txt = fl.text_model.encoder(fl.text_model.embeddings(ids, None if token_type_ids is None else token_type_ids.to(dev)),
                                    attention_mask=None if attention_mask is None else attention_mask.to(dev),
                                    sep_idx=None if sep_idx is None else sep_idx.to(dev), row_lo=1)
        ip, tp = fl.image_to_mm_projection, fl.text_to_mm_projection
        mm_in = torch.cat([fl.multimodal_model.cls_token.expand(Bn, -1, -1), B.linear(img, ip.weight, ip.bias),
                           B.linear(txt, tp.weight, tp.bias)], dim=1)
        mm = B.layer_norm(fl.multimodal_model.encoder(mm_in), fl.multimodal_model.layernorm)
        trans = self.cls.transform(mm[:, -L:])
        logits = B.linear(trans, self.cls.decoder.weight, self.cls.bias) if output_logits else None
        return MaskedLMOutput(logits=logits), trans
