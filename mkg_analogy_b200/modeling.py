"""Drop-in replacement of the reference model class at its plugin seam.

The reference loads its model by name (`models.<--model_class>`, MarT/main.py:12-17,:73) and drives it from
`TransformerLitModel` (MarT/lit_models/transformer.py).  `MKGformerKGC` below mirrors the interface that code relies on
 — same constructor `(vision_config, text_config)` (main.py:86), same `state_dict()` names and shapes
as `UnimoForMaskedLM` (MarT/models/modeling_unimo.py:839-930), `resize_token_embeddings`, tied
`get_input_embeddings()/get_output_embeddings()`, `add_to_argparse`, and `forward(**batch, return_dict=True)` returning
the 2-tuple `(MaskedLMOutput-like with .logits (B, L, V), trans_hidden_states (B, L, H))` (modeling_unimo.py:888-893) —
while every FLOP runs in libb200mart.so on the GPU.  Parameters are ordinary `nn.Parameter`s that are *views* of the
engine's flat fp32 buffer; their `.grad`s are views of the flat gradient buffer.

Arithmetic: bf16 tensor-core operands, fp32 accumulation; `module.train()` enables the BERT stream's dropout (reference
p = 0.1; stateless hash masks, so the random stream differs from torch's), `module.eval()` is the deterministic graph.
"""
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn as nn

from .config import MKGformerConfig
from .engine import Engine


@dataclass
class MaskedLMOutput:
    loss: Optional[torch.Tensor] = None
    logits: Optional[torch.Tensor] = None
    hidden_states: Optional[tuple] = None
    attentions: Optional[tuple] = None


class _Holder(nn.Module):
    """Plain container so that dotted reference names resolve (unimo.encoder.text_layer.3.attention.self.query.weight)."""


def _holder_path(root, parts):
    m = root
    for p in parts:
        if p not in m._modules:
            m.add_module(p, _Holder())
        m = m._modules[p]
    return m


def config_from_hf(vision_config, text_config):
    """MKGformerConfig from the two HF configs the reference passes (main.py:78-86)."""
    g = lambda c, k, d=None: getattr(c, k, d)
    if g(vision_config, "hidden_size") != g(text_config, "hidden_size") or \
            g(vision_config, "num_hidden_layers") != g(text_config, "num_hidden_layers"):
        raise ValueError("MKGformer needs equal hidden size and depth in both streams (modeling_unimo.py:600)")
    return MKGformerConfig(
        hidden=text_config.hidden_size, layers=text_config.num_hidden_layers, heads=text_config.num_attention_heads,
        intermediate=text_config.intermediate_size, vocab=text_config.vocab_size,
        max_position=text_config.max_position_embeddings, type_vocab=text_config.type_vocab_size,
        image_size=vision_config.image_size, patch=vision_config.patch_size,
        text_eps=float(text_config.layer_norm_eps), vision_eps=float(vision_config.layer_norm_eps),
        hidden_dropout=float(g(text_config, "hidden_dropout_prob", 0.1)),
        attn_dropout=float(g(text_config, "attention_probs_dropout_prob", 0.1)))


class _ForwardFn(torch.autograd.Function):
    """Autograd node of the compat route: dense logits + hidden states out, their gradients in."""

    @staticmethod
    def forward(ctx, anchor, module, batch, want_logits):
        z, logits = module.engine.forward(batch, want_z=True, want_logits=want_logits)
        ctx.module, ctx.batch = module, batch
        if logits is None:
            logits = z.new_zeros(())
        ctx.want_logits = want_logits
        return logits, z

    @staticmethod
    def backward(ctx, dlogits, dz):
        m = ctx.module
        m._prepare_grads()
        m.engine.backward(ctx.batch, use_head_grad=False, dz=dz, dlogits=dlogits if ctx.want_logits else None)
        m._attach_grads()
        return None, None, None, None


class MKGformerKGC(nn.Module):
    """`models.MKGformerKGC` of the reference (MarT/models/model.py:9-13), B200-native."""

    def __init__(self, vision_config, text_config=None, device=None):
        super().__init__()
        cfg = vision_config if isinstance(vision_config, MKGformerConfig) else config_from_hf(vision_config, text_config)
        self.vision_config, self.text_config = vision_config, text_config
        self.initializer_range = float(getattr(text_config, "initializer_range", 0.02)) if text_config is not None else 0.02
        if device is None:
            device = "cuda" if torch.cuda.is_available() else "cpu"
        self._anchor = None
        self._build(cfg, torch.device(device), None)

    # ------------------------------------------------------------------ construction
    @staticmethod
    def add_to_argparse(parser):
        parser.add_argument("--pretrain", type=int, default=0, help="")
        return parser

    def _build(self, cfg, device, old_state):
        self.cfg = cfg
        self.engine = Engine(cfg, device)
        for name in list(self._modules):
            del self._modules[name]
        self._param_names = []
        with torch.no_grad():
            for info in self.engine.infos:
                v = self.engine.view(info.name)
                if old_state is not None and info.name in old_state and old_state[info.name].shape == v.shape:
                    v.copy_(old_state[info.name])
                else:
                    self._init_tensor(info.name, v)
        for info in self.engine.infos:   # reference registration order is irrelevant for lookups by name
            parts = info.name.split(".")
            holder = _holder_path(self, parts[:-1])
            holder.register_parameter(parts[-1], nn.Parameter(self.engine.view(info.name)))
            self._param_names.append(info.name)
        # tied decoder (modeling_unimo.py:904-926): aliases of the embedding table and of cls.predictions.bias
        dec = _holder_path(self, ["cls", "predictions", "decoder"])
        dec.weight = self.unimo.text_embeddings.word_embeddings.weight
        dec.bias = self.cls.predictions.bias
        self.unimo.text_embeddings.register_buffer("position_ids", torch.arange(cfg.max_position, device=device).expand((1, -1)))
        self.unimo.vision_embeddings.register_buffer("position_ids", torch.arange(cfg.patches + 1, device=device).expand((1, -1)))
        self._anchor = torch.zeros((), device=device, requires_grad=True)
        self.engine.mirror_stale = True
        self.engine.set_dropout(cfg.hidden_dropout, cfg.attn_dropout, int(torch.initial_seed()) & 0x7FFFFFFF)

    def _init_tensor(self, name, v):
        if "LayerNorm.weight" in name or "layer_norm1.weight" in name or "layer_norm2.weight" in name or \
                name.endswith("layrnorm.weight") or name.endswith("layernorm.weight"):
            v.fill_(1.0)
        elif name.endswith("adaptive_weight.0"):
            v.uniform_(0.0, 0.5)       # modeling_unimo.py:305-310
        elif name.endswith("adaptive_weight.1"):
            v.fill_(0.5)
        elif name.endswith("class_embedding"):
            v.normal_(0.0, 1.0)
        elif name.endswith(".bias"):
            v.zero_()
        else:
            v.normal_(0.0, self.initializer_range)

    def _params(self):
        for name in self._param_names:
            parts = name.split(".")
            m = self
            for p in parts[:-1]:
                m = m._modules[p]
            yield name, m, parts[-1]

    def _rebind(self):
        """Point every nn.Parameter at the engine's (possibly moved) flat buffer."""
        for name, holder, leaf in self._params():
            holder._parameters[leaf].data = self.engine.view(name)
            holder._parameters[leaf].grad = None
        self._anchor = torch.zeros((), device=self.engine.device, requires_grad=True)

    def _apply(self, fn, recurse=True):
        probe = fn(torch.empty(0, device=self.engine.device, dtype=torch.float32))
        if probe.dtype != torch.float32:
            raise TypeError("MKGformerKGC keeps fp32 master parameters (bf16 operands are mirrored internally); "
                            "dtype casts of the module are not supported")
        if probe.device != self.engine.device:
            self.engine.move(probe.device)
            self._rebind()
        for m in self.modules():
            for k, b in m._buffers.items():
                if b is not None:
                    m._buffers[k] = fn(b)
        return self

    # ------------------------------------------------------------------ HF-style embedding API
    def get_input_embeddings(self):
        return self.unimo.text_embeddings.word_embeddings

    def get_output_embeddings(self):
        return self.cls.predictions.decoder

    def resize_token_embeddings(self, new_num_tokens=None):
        """modeling_unimo.py:782-836, :928-930: grow/shrink the tied table (new rows ~ N(0, initializer_range)),
        zero-pad the decoder bias, keep the tie."""
        if new_num_tokens is None or new_num_tokens == self.cfg.vocab:
            return self.get_input_embeddings()
        old = self.engine.state()
        old_v = self.cfg.vocab
        cfg = MKGformerConfig(**{**self.cfg.to_dict(), "vocab": int(new_num_tokens)})
        device = self.engine.device
        self.engine = None
        self._build(cfg, device, old)
        n = min(old_v, cfg.vocab)
        with torch.no_grad():
            self.engine.view("unimo.text_embeddings.word_embeddings.weight")[:n].copy_(
                old["unimo.text_embeddings.word_embeddings.weight"][:n])
            b = self.engine.view("cls.predictions.bias")
            b.zero_()
            b[:n].copy_(old["cls.predictions.bias"][:n])
        if self.text_config is not None:
            self.text_config.vocab_size = cfg.vocab
        return self.get_input_embeddings()

    def tie_weights(self):
        pass   # tied by construction

    # ------------------------------------------------------------------ gradients <-> flat buffer
    def _prepare_grads(self):
        """Called right before a backward accumulates into the flat gradient buffer."""
        fresh = True
        for name, holder, leaf in self._params():
            g = holder._parameters[leaf].grad
            if g is not None:
                fresh = False
                if g.data_ptr() != self.engine.grad_view(name).data_ptr():
                    self.engine.grad_view(name).copy_(g)      # someone replaced .grad: keep its value
        if fresh:
            self.engine.zero_grad()                            # zero_grad(set_to_none=True) semantics
        else:
            for name, holder, leaf in self._params():
                if holder._parameters[leaf].grad is None:
                    self.engine.grad_view(name).zero_()

    def _attach_grads(self, pretrain=False):
        for info, (name, holder, leaf) in zip(self.engine.infos, self._params()):
            p = holder._parameters[leaf]
            # parameters the reference never reaches keep .grad = None, so torch optimizers skip them as they do there
            if info.unused_ft or not p.requires_grad:
                continue
            p.grad = self.engine.grad_view(name)

    # ------------------------------------------------------------------ forward
    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, pixel_values=None, sep_idx=None,
                return_dict=None, output_logits=True, **unused):
        dev = self.engine.device
        batch = {"input_ids": input_ids.to(dev).contiguous(), "attention_mask": attention_mask.to(dev, torch.int64).contiguous(),
                 "pixel_values": pixel_values.to(dev, torch.float32).contiguous()}
        if token_type_ids is not None:
            batch["token_type_ids"] = token_type_ids.to(dev).contiguous()
        if sep_idx is not None:
            batch["sep_idx"] = sep_idx.to(dev).contiguous()
        self.engine.mirror_stale = True     # parameters may have been updated by a torch optimizer since the last call
        if self.engine.device.type == "cuda":
            self.engine.train(self.training)    # module.train() / .eval() -> dropout of the BERT stream on / off
        if torch.is_grad_enabled():
            logits, z = _ForwardFn.apply(self._anchor, self, batch, output_logits)
        else:
            z, logits = self.engine.forward(batch, want_z=True, want_logits=output_logits)
        return MaskedLMOutput(logits=logits if output_logits else None), z
