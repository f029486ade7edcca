"""`from_pretrained` / HF-style initialisation for the FLAVA and ViLT drop-ins.

The reference builds these two families with `model_class.from_pretrained(args.visual_model_path[, config=config])`
(MarT/main.py:119-125) from a LOCAL checkpoint directory (config.json + pytorch_model.bin or model.safetensors).  This
mixin restates that contract for the drop-in classes: read the config, build the module, initialise what the checkpoint
does not hold the way `_init_weights` does (modeling_flava.py:895-912, modeling_vilt.py:602-618: N(0, initializer_range)
weights, zero biases, unit LayerNorms, zeroed padding row), load matching tensors by name, tie the decoder."""
import json
import os

import torch
import torch.nn as nn


def init_hf_style(module, std=0.02):
    """Initialise the torch plumbing (everything outside the encoder stacks, which initialise themselves)."""
    from .stack import EncoderStack
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, EncoderStack):
                continue
            if isinstance(m, (nn.Linear, nn.Conv2d)):
                m.weight.normal_(0.0, std)
                if m.bias is not None:
                    m.bias.zero_()
            elif isinstance(m, nn.Embedding):
                m.weight.normal_(0.0, std)
                if m.padding_idx is not None:
                    m.weight[m.padding_idx].zero_()
            elif isinstance(m, nn.LayerNorm):
                m.weight.fill_(1.0)
                m.bias.zero_()


def read_checkpoint(path):
    """State dict of a local HF checkpoint directory (or a single weights file)."""
    if os.path.isdir(path):
        idx = [f for f in ("model.safetensors.index.json", "pytorch_model.bin.index.json") if os.path.exists(os.path.join(path, f))]
        if idx:
            with open(os.path.join(path, idx[0])) as f:
                files = sorted(set(json.load(f)["weight_map"].values()))
        else:
            files = [f for f in ("model.safetensors", "pytorch_model.bin") if os.path.exists(os.path.join(path, f))][:1]
        if not files:
            raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin under {path}")
        files = [os.path.join(path, f) for f in files]
    else:
        files = [path]
    sd = {}
    for f in files:
        if f.endswith(".safetensors"):
            from safetensors.torch import load_file
            sd.update(load_file(f))
        else:
            sd.update(torch.load(f, map_location="cpu", weights_only=True))
    return sd


class PretrainedMixin:
    @classmethod
    def from_pretrained(cls, path, config=None, device=None, **kwargs):
        if config is None:
            from transformers import AutoConfig
            config = AutoConfig.from_pretrained(path)
        model = cls(config, device=device, **kwargs)
        sd = read_checkpoint(path)
        own = model.state_dict()
        take = {k: v for k, v in sd.items() if k in own and tuple(own[k].shape) == tuple(v.shape)}
        missing = sorted(k for k in own if k not in take)
        unexpected = sorted(k for k in sd if k not in take)
        model.load_state_dict(take, strict=False)
        model.tie_weights()
        model.loading_info = {"missing_keys": missing, "unexpected_keys": unexpected}
        if missing:
            print(f"{cls.__name__}.from_pretrained: {len(missing)} tensors not in the checkpoint were newly initialised "
                  f"(first: {missing[:3]})")
        return model
