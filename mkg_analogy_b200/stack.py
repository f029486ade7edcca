"""Encoder stack on the B200 library: the pre-LayerNorm block FLAVA and ViLT are built from, and (post_ln=True) the BERT block of VisualBERT.

`EncoderStack` mirrors the HF encoders the reference vendors — `FlavaEncoder` (MarT/models/modeling_flava.py:668-722: image,
text and multimodal encoders are three instances) and `ViltEncoder` (MarT/models/modeling_vilt.py:531-583) — same
`state_dict()` names (`layer.<i>.attention.attention.query.weight`, ...), same call shape (hidden states in, hidden states
out), with the reference's adaptive analogy scaling (`sep_idx`, FLAVA rows 1:s / ViLT rows img_length+1:s) and additive
padding mask inside the fused attention kernel.  `post_ln=True` gives `VisualBertEncoder` (MarT/models/modeling_visual_bert.py:424-488:
`layer.<i>.attention.self.query.weight`, `...attention.output.LayerNorm.weight`, ...; scaling on rows img_length+1:s', :255-260).  Parameters are `nn.Parameter` views of a flat fp32 buffer; the backward is
`b200mart_stack_backward` (hand-derived), hooked into torch autograd so that embeddings / heads written in PyTorch around it
train normally.  No CPU fallback: computing needs a CUDA device.
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .modeling import _Holder, _holder_path


class _StackFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, stack, key_mask, split, row_lo):
        y, io = stack._forward(x, key_mask, split, row_lo)
        ctx.stack, ctx.io_args = stack, (x.shape, key_mask, split, row_lo, io.seed, io.p_hidden, io.p_attn)
        return y

    @staticmethod
    def backward(ctx, dy):
        st = ctx.stack
        dx = st._backward(dy.contiguous(), *ctx.io_args)
        return dx, None, None, None, None


class EncoderStack(nn.Module):
    def __init__(self, hidden, layers, heads, intermediate, eps=1e-12, act="gelu", hidden_dropout=0.0, attn_dropout=0.0,
                 device=None, initializer_range=0.02, post_ln=False):
        super().__init__()
        self.l = _lib.lib()
        d = _lib.StackDesc()
        d.hidden, d.layers, d.heads, d.intermediate = hidden, layers, heads, intermediate
        d.post_ln = 1 if post_ln else 0
        d.act = {"gelu": 1, "quick_gelu": 2}[act]
        d.eps = eps
        self.h = _lib.vp()
        _lib.check(self.l.b200mart_stack_create(C.byref(d), C.byref(self.h)))
        self.hidden, self.n_layers = hidden, layers
        self.p_hidden, self.p_attn = float(hidden_dropout), float(attn_dropout)
        self._calls = 0
        self._seed0 = int(torch.initial_seed()) & 0x7FFFFFFF
        device = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
        self.total = self.l.b200mart_stack_param_total(self.h)
        self.infos = []
        for i in range(self.l.b200mart_stack_param_count(self.h)):
            name, off, nd, shape = C.c_char_p(), _lib.i64(), _lib.i32(), (_lib.i64 * 4)()
            _lib.check(self.l.b200mart_stack_param_info(self.h, i, C.byref(name), C.byref(off), C.byref(nd), C.byref(shape)))
            self.infos.append((name.value.decode(), off.value, tuple(shape[k] for k in range(nd.value))))
        self._alloc(device)
        with torch.no_grad():
            for name, off, shape in self.infos:
                v = self._view(self.params32, off, shape)
                if name.lower().endswith(("layernorm_before.weight", "layernorm_after.weight", "layernorm.weight")):
                    v.fill_(1.0)
                elif name.endswith("adaptive_weight.0"):
                    v.uniform_(0.0, 0.5)
                elif name.endswith("adaptive_weight.1"):
                    v.fill_(0.5)
                elif name.endswith(".bias"):
                    v.zero_()
                else:
                    v.normal_(0.0, initializer_range)
        for name, off, shape in self.infos:
            parts = name.split(".")
            _holder_path(self, parts[:-1]).register_parameter(parts[-1], nn.Parameter(self._view(self.params32, off, shape)))
        self.ws = None

    def __del__(self):
        try:
            if self.h:
                self.l.b200mart_stack_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ storage
    @staticmethod
    def _view(flat, off, shape):
        n = 1
        for s in shape:
            n *= s
        return flat[off:off + n].view(shape)

    def _alloc(self, device):
        self.dev = device
        old = getattr(self, "params32", None)
        self.params32 = torch.zeros(self.total, device=device) if old is None else old.to(device)
        self.params16 = self.grads32 = None
        if device.type == "cuda":
            self.params16 = torch.zeros(self.total, device=device, dtype=torch.bfloat16)
            self.grads32 = torch.zeros(self.total, device=device)
        self.ws = None

    def _named(self):
        for name, off, shape in self.infos:
            parts = name.split(".")
            m = self
            for p in parts[:-1]:
                m = m._modules[p]
            yield name, off, shape, m._parameters[parts[-1]]

    def _apply(self, fn, recurse=True):
        probe = fn(torch.empty(0, device=self.dev, dtype=torch.float32))
        if probe.dtype != torch.float32:
            raise TypeError("EncoderStack keeps fp32 master parameters; dtype casts are not supported")
        if probe.device != self.dev:
            self._alloc(probe.device)
            for name, off, shape, p in self._named():
                p.data = self._view(self.params32, off, shape)
                p.grad = None
        return self

    # ------------------------------------------------------------------ compute
    def _bufs(self, B, T):
        if self.dev.type != "cuda":
            raise _lib.B200MartError("b200mart computes on a CUDA device (sm_100a) only; there is no CPU fallback")
        n = self.l.b200mart_stack_workspace_bytes(self.h, B, T)
        if self.ws is None or self.ws.numel() < n:
            self.ws = torch.empty(n, device=self.dev, dtype=torch.uint8)
        b = _lib.Buffers()
        b.params32, b.params16, b.grads32 = self.params32.data_ptr(), self.params16.data_ptr(), self.grads32.data_ptr()
        b.workspace, b.workspace_bytes = self.ws.data_ptr(), self.ws.numel()
        return b

    def _io(self, shape, key_mask, split, row_lo, seed, ph, pa):
        io = _lib.StackIO()
        io.B, io.T = shape[0], shape[1]
        io.key_mask, io.split, io.row_lo = _lib.ptr(key_mask), _lib.ptr(split), int(row_lo)
        io.p_hidden, io.p_attn, io.seed = ph, pa, seed
        return io

    def _forward(self, x, key_mask, split, row_lo):
        B, T, H = x.shape
        assert H == self.hidden and x.is_cuda and x.dtype == torch.float32
        x = x.contiguous()
        bufs = self._bufs(B, T)
        self.params16.copy_(self.params32)            # bf16 operand mirror (parameters may have been stepped by torch)
        self._calls += 1
        train = self.training
        seed = (self._seed0 * 2654435761 + self._calls * 40503) & 0xFFFFFFFF
        io = self._io(x.shape, key_mask, split, row_lo, seed, self.p_hidden if train else 0.0, self.p_attn if train else 0.0)
        y = torch.empty_like(x)
        io.x, io.y = x.data_ptr(), y.data_ptr()
        self._last_seed = seed
        _lib.check(self.l.b200mart_stack_forward(self.h, C.byref(bufs), C.byref(io), _lib.cur_stream()))
        return y, io

    def _backward(self, dy, shape, key_mask, split, row_lo, seed, ph, pa):
        bufs = self._bufs(shape[0], shape[1])
        io = self._io(shape, key_mask, split, row_lo, seed, ph, pa)
        fresh = all(p.grad is None for _, _, _, p in self._named())
        if fresh:
            self.grads32.zero_()
        dx = torch.empty(shape, device=self.dev)
        _lib.check(self.l.b200mart_stack_backward(self.h, C.byref(bufs), C.byref(io), dy.data_ptr(), dx.data_ptr(), _lib.cur_stream()))
        for name, off, shp, p in self._named():
            if "adaptive_weight" in name and split is None:
                continue                               # no scaling -> the reference leaves .grad = None
            p.grad = self._view(self.grads32, off, shp)
        return dx

    def forward(self, hidden_states, attention_mask=None, sep_idx=None, row_lo=1):
        """hidden_states (B, T, H) fp32; attention_mask (B, T) 1 = keep (the additive -1e4 extension is applied in-kernel);
        sep_idx (B, 6) as the reference passes it (split = sep_idx[:, 2]); row_lo: first scaled row (FLAVA 1, ViLT
        img_length + 1).  Returns the last hidden states (B, T, H)."""
        km = None if attention_mask is None else attention_mask.to(self.dev, torch.int64).contiguous()
        split = None if sep_idx is None else sep_idx[:, 2].to(self.dev, torch.int64).contiguous()
        x = hidden_states.to(self.dev, torch.float32)
        if torch.is_grad_enabled():
            if not x.requires_grad:
                x = x.detach().requires_grad_(True)       # keeps the autograd node alive for the parameter gradients
            return _StackFn.apply(x, self, km, split, row_lo)
        return self._forward(x, km, split, row_lo)[0]
