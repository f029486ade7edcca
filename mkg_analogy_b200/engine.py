"""Host-side owner of one b200mart handle: flat parameter / gradient / optimizer buffers in HBM, the
workspace, and thin wrappers of the C-ABI calls (include/b200mart.h).  PyTorch is used for device memory
and streams only; every computation happens inside libb200mart.so on the current CUDA stream.

Data layout in HBM (per GPU):
  params32  flat fp32 [total]   master parameters; nn.Parameters of the drop-in module are views of it
  params16  flat bf16 [total]   tensor-core operand mirror (same offsets)
  grads32   flat fp32 [total]   gradients (same offsets) — one contiguous buffer, so the data-parallel
                                exchange is a handful of large NCCL all-reduces, not one per tensor
  exp_avg / exp_avg_sq          AdamW state (allocated on first optimizer step)
  workspace                     activations saved for backward + scratch, carved by the library
"""
import ctypes as C

import torch

from . import _lib
from .config import MKGformerConfig


class ParamInfo:
    __slots__ = ("name", "offset", "shape", "decay", "unused_ft")

    def __init__(self, name, offset, shape, decay, unused_ft):
        self.name, self.offset, self.shape, self.decay, self.unused_ft = name, offset, shape, decay, unused_ft

    @property
    def numel(self):
        n = 1
        for s in self.shape:
            n *= s
        return n


def query_layout(cfg: MKGformerConfig):
    """Parameter layout of the library for `cfg` (host-only: works without a GPU)."""
    l = _lib.lib()
    d = _model_desc(cfg)
    h = _lib.vp()
    _lib.check(l.b200mart_create(C.byref(d), C.byref(h)))
    try:
        return _read_layout(l, h)
    finally:
        l.b200mart_destroy(h)


def _model_desc(cfg):
    d = _lib.ModelDesc()
    d.hidden, d.layers, d.heads, d.intermediate = cfg.hidden, cfg.layers, cfg.heads, cfg.intermediate
    d.vocab, d.max_position, d.type_vocab = cfg.vocab, cfg.max_position, cfg.type_vocab
    d.image_size, d.patch, d.fuse_from = cfg.image_size, cfg.patch, cfg.fuse_from
    d.text_eps, d.vision_eps = cfg.text_eps, cfg.vision_eps
    return d


def _read_layout(l, h):
    infos = []
    n = l.b200mart_param_count(h)
    for i in range(n):
        name, off, nd = C.c_char_p(), _lib.i64(), _lib.i32()
        shape = (_lib.i64 * 4)()
        decay, unused = _lib.i32(), _lib.i32()
        _lib.check(l.b200mart_param_info(h, i, C.byref(name), C.byref(off), C.byref(nd), C.byref(shape), C.byref(decay),
                                         C.byref(unused)))
        infos.append(ParamInfo(name.value.decode(), off.value, tuple(shape[k] for k in range(nd.value)), bool(decay.value),
                               bool(unused.value)))
    return infos, l.b200mart_param_total(h)


class Engine:
    def __init__(self, cfg: MKGformerConfig, device="cuda"):
        """device 'cpu' gives parameter STORAGE only (state_dict surgery before .cuda(), as MarT/main.py does);
        every compute call then raises: there is no CPU fallback."""
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.l = _lib.lib()
        self.h = _lib.vp()
        d = _model_desc(cfg)
        _lib.check(self.l.b200mart_create(C.byref(d), C.byref(self.h)))
        self.infos, self.total = _read_layout(self.l, self.h)
        self.by_name = {p.name: p for p in self.infos}
        self.params32 = torch.zeros(self.total, device=self.device, dtype=torch.float32)
        self.params16 = self.grads32 = None
        if self.device.type == "cuda":
            self.params16 = torch.zeros(self.total, device=self.device, dtype=torch.bfloat16)
            self.grads32 = torch.zeros(self.total, device=self.device, dtype=torch.float32)
        self.exp_avg = None
        self.exp_avg_sq = None
        self.ws = None
        self.ws_key = None
        self.mirror_stale = True
        self.launches = 0   # bookkeeping for bench.py: C-ABI calls issued
        self.forward_calls = 0   # the workspace holds the activations of forward number `forward_calls`

    def __del__(self):
        try:
            if self.h:
                self.l.b200mart_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _need_cuda(self):
        if self.device.type != "cuda":
            raise _lib.B200MartError("b200mart computes on a CUDA device (sm_100a) only: move the module with .cuda(); "
                                     "there is no CPU fallback")

    def move(self, device):
        """Re-home the flat buffers (parameters and optimizer state keep their values; gradients are reset)."""
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.params32 = self.params32.to(device)
        self.device = device
        self.params16 = self.grads32 = None
        if device.type == "cuda":
            self.params16 = torch.zeros(self.total, device=device, dtype=torch.bfloat16)
            self.grads32 = torch.zeros(self.total, device=device, dtype=torch.float32)
        if self.exp_avg is not None:
            self.exp_avg, self.exp_avg_sq = self.exp_avg.to(device), self.exp_avg_sq.to(device)
        self.ws = self.ws_key = None
        self.mirror_stale = True

    # ------------------------------------------------------------------ parameter access
    def view(self, name, flat=None):
        p = self.by_name[name]
        flat = self.params32 if flat is None else flat
        return flat[p.offset:p.offset + p.numel].view(p.shape)

    def grad_view(self, name):
        self._need_cuda()
        return self.view(name, self.grads32)

    def load(self, params):
        """Copy {reference state_dict name: tensor} into the flat buffer and refresh the bf16 mirror."""
        missing = [n for n in self.by_name if n not in params]
        if missing:
            raise KeyError(f"missing parameters: {missing[:5]} ...")
        with torch.no_grad():
            for n in self.by_name:
                self.view(n).copy_(params[n].to(self.device, torch.float32))
        if self.device.type == "cuda":
            self.sync_mirror()

    def state(self):
        return {n: self.view(n).clone() for n in self.by_name}

    def sync_mirror(self):
        self._need_cuda()
        _lib.check(self.l.b200mart_cast_params(self.h, self.params32.data_ptr(), self.params16.data_ptr(), _lib.cur_stream()))
        self.mirror_stale = False

    def adamw_multicast(self, step, lr, grads_mc, params32_mc, params16_mc, world, begin, end, betas=(0.9, 0.999), eps=1e-8,
                        weight_decay=0.01, pretrain=False, max_ctas=0):
        """This rank's shard [begin, end) of the fused reduce-scatter + AdamW + all-gather over NVSwitch multicast
        (b200mart_adamw_multicast); *_mc are the multicast base addresses of the symmetric gradient / parameter buffers."""
        self._need_cuda()
        if self.exp_avg is None:
            self.exp_avg = torch.zeros_like(self.params32)
            self.exp_avg_sq = torch.zeros_like(self.params32)
        _lib.check(self.l.b200mart_adamw_multicast(self.h, self.params32.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                                   int(grads_mc), int(params32_mc), int(params16_mc), int(step), float(lr),
                                                   float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(pretrain),
                                                   int(begin), int(end), int(world), int(max_ctas), _lib.cur_stream()))
        self.mirror_stale = False

    def cast_range(self, begin, end):
        """bf16 mirror of params32[begin:end) (parameters delivered by another rank's optimizer shard)."""
        if end > begin:
            _lib.check(self.l.b200mart_cast_params_range(self.h, self.params32.data_ptr(), self.params16.data_ptr(), int(begin),
                                                         int(end), _lib.cur_stream()))

    def zero_grad(self):
        self.grads32.zero_()

    # ------------------------------------------------------------------ C-ABI plumbing
    def _bufs(self, B, L, need_grad=True):
        self._need_cuda()
        key = (B, L)
        if self.ws_key != key:
            nbytes = self.l.b200mart_workspace_bytes(self.h, B, L)
            if self.ws is None or self.ws.numel() < nbytes:
                self.ws = None
                self.ws = torch.empty(nbytes, device=self.device, dtype=torch.uint8)
            self.ws_key = key
        b = _lib.Buffers()
        b.params32, b.params16 = self.params32.data_ptr(), self.params16.data_ptr()
        b.grads32 = self.grads32.data_ptr() if need_grad else None
        b.workspace, b.workspace_bytes = self.ws.data_ptr(), self.ws.numel()
        return b

    @staticmethod
    def _batch(batch):
        ids = batch["input_ids"]
        b = _lib.Batch()
        b.B, b.L = ids.shape
        for k in ("input_ids", "attention_mask", "pixel_values"):
            t = batch[k]
            assert t.is_cuda and t.is_contiguous(), f"{k} must be a contiguous CUDA tensor"
        assert ids.dtype == torch.int64 and batch["attention_mask"].dtype == torch.int64
        assert batch["pixel_values"].dtype == torch.float32
        b.input_ids, b.attention_mask = ids.data_ptr(), batch["attention_mask"].data_ptr()
        tt = batch.get("token_type_ids")
        b.token_type_ids = _lib.ptr(tt)
        b.pixel_values = batch["pixel_values"].data_ptr()
        sep = batch.get("sep_idx")
        if sep is not None:
            assert sep.is_cuda and sep.is_contiguous() and sep.dtype == torch.int64 and sep.shape[1] == 6
        b.sep_idx = _lib.ptr(sep)
        return b

    # ------------------------------------------------------------------ the path
    def forward(self, batch, want_z=True, want_logits=False):
        """UnimoForMaskedLM.forward. Returns (z (B,L,H) fp32 or None, logits (B,L,V) fp32 or None)."""
        self._need_cuda()
        if self.mirror_stale:
            self.sync_mirror()
        cb = self._batch(batch)
        bufs = self._bufs(cb.B, cb.L)
        z = torch.empty(cb.B, cb.L, self.cfg.hidden, device=self.device) if want_z else None
        logits = torch.empty(cb.B, cb.L, self.cfg.vocab, device=self.device) if want_logits else None
        self.forward_calls += 1
        _lib.check(self.l.b200mart_forward(self.h, C.byref(bufs), C.byref(cb), _lib.ptr(z), _lib.ptr(logits), _lib.cur_stream()))
        return z, logits

    def head_loss(self, batch, head, compute_grad=True, want_logits=False, want_ranks=False):
        """Score head + loss on the [MASK] rows (after forward on the same batch).
        `head` is a dict: mode 'finetune' {col_ids, label, rel_idx, q_head_idx, a_head_idx, alpha, label_smoothing}
        or 'pretrain' {entity_range, relation_range, label, pre_type, label_smoothing}; mask_token_id optional."""
        cb = self._batch(batch)
        bufs = self._bufs(cb.B, cb.L)
        hd = _lib.Head()
        hd.label = head["label"].data_ptr()
        hd.label_smoothing = float(head.get("label_smoothing", 0.1))
        hd.mask_token_id = int(head.get("mask_token_id", 103))
        gs = head.get("grad_scale")                 # device scalar (fp32) multiplying every gradient seeded here; None = 1
        if gs is not None:
            assert gs.is_cuda and gs.dtype == torch.float32 and gs.numel() == 1
        hd.grad_scale = _lib.ptr(gs)
        if head["mode"] == "finetune":
            hd.mode = 0
            hd.col_ids, hd.n_cols = head["col_ids"].data_ptr(), head["col_ids"].numel()
            hd.rel_idx, hd.q_head_idx, hd.a_head_idx = (head["rel_idx"].data_ptr(), head["q_head_idx"].data_ptr(),
                                                       head["a_head_idx"].data_ptr())
            hd.alpha = float(head.get("alpha", 0.43))
            width = hd.n_cols
        else:
            hd.mode = 1
            (hd.ent_st, hd.ent_ed), (hd.rel_st, hd.rel_ed) = head["entity_range"], head["relation_range"]
            hd.pre_type = head["pre_type"].data_ptr()
            width = (hd.ent_ed - hd.ent_st) + (hd.rel_ed - hd.rel_st)
        loss = torch.empty(1, device=self.device)
        ml = torch.empty(cb.B, width, device=self.device) if want_logits else None
        ranks = torch.zeros(cb.B, device=self.device, dtype=torch.int64) if want_ranks else None
        _lib.check(self.l.b200mart_head_loss(self.h, C.byref(bufs), C.byref(cb), C.byref(hd), loss.data_ptr(), _lib.ptr(ml),
                                             _lib.ptr(ranks), int(compute_grad), _lib.cur_stream()))
        return loss, ml, ranks

    def read_hidden(self, kind, layer):
        """fp32 residual stream saved by the last forward: kind "text" (B, L, H) / "vision" (B, Nv, H) entering layer `layer`
        (layer == cfg.layers: the stream's output) — the reference's `output_hidden_states`."""
        B, L = self.ws_key
        rows = L if kind == "text" else self.cfg.vision_tokens
        out = torch.empty(B, rows, self.cfg.hidden, device=self.device)
        bufs = self._bufs(B, L)
        _lib.check(self.l.b200mart_read_hidden(self.h, C.byref(bufs), 0 if kind == "text" else 1, int(layer), out.data_ptr(),
                                               _lib.cur_stream()))
        return out

    def layer_extents(self):
        """[(begin, end)] flat extents of each layer index's parameters (vision + text layer)."""
        out = []
        for i in range(self.cfg.layers):
            b, e = _lib.i64(), _lib.i64()
            _lib.check(self.l.b200mart_layer_extent(self.h, i, C.byref(b), C.byref(e)))
            out.append((b.value, e.value))
        return out

    def set_dropout(self, p_hidden, p_attn, seed=0):
        _lib.check(self.l.b200mart_set_dropout(self.h, float(p_hidden), float(p_attn), int(seed) & 0xFFFFFFFF))

    def set_step_block(self, block):
        """Per-step scalars in device memory (b200mart_set_step_block): a 16-byte CUDA tensor {uint32 salt, float lr, float 1 - beta1^t,
        float 1 - beta2^t} the kernels read dropout salt / AdamW scalars from, or None to restore by-value arguments."""
        if block is not None:
            assert block.is_cuda and block.numel() * block.element_size() == 16 and block.is_contiguous()
        self._step_block = block                      # keeps the storage alive
        _lib.check(self.l.b200mart_set_step_block(self.h, _lib.ptr(block)))

    def train(self, mode=True):
        """nn.Module.train()/eval(): dropout on/off for the following forward calls."""
        self.set_option("training", 1 if mode else 0)

    def last_dropout_seed(self):
        """Per-forward seed of the most recent forward (tests rebuild the masks from it)."""
        return int(self.l.b200mart_last_dropout_seed(self.h))

    def set_option(self, name, value):
        _lib.check(self.l.b200mart_set_option(self.h, name.encode(), int(value)))

    def backward(self, batch, use_head_grad=True, dz=None, dlogits=None, layer_events=None):
        """layer_events: optional list of 2*layers torch.cuda.Event (already created by a record()); see b200mart.h."""
        cb = self._batch(batch)
        bufs = self._bufs(cb.B, cb.L)
        scratch = None
        if dlogits is not None:
            vp8 = (self.cfg.vocab + 7) // 8 * 8
            scratch = torch.empty(cb.B * cb.L * vp8, device=self.device, dtype=torch.bfloat16)
            dlogits = dlogits.contiguous()
        if dz is not None:
            dz = dz.contiguous()
        evs = None
        if layer_events is not None:
            assert len(layer_events) == 2 * self.cfg.layers
            evs = (_lib.vp * len(layer_events))(*[e.cuda_event for e in layer_events])
        _lib.check(self.l.b200mart_backward(self.h, C.byref(bufs), C.byref(cb), int(use_head_grad), _lib.ptr(dz),
                                            _lib.ptr(dlogits), _lib.ptr(scratch), evs, _lib.cur_stream()))

    def adamw(self, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, pretrain=False, begin=0, end=0,
              zero_grads=False):
        """AdamW over the flat range [begin, end) (default: everything) on the current stream."""
        self._need_cuda()
        if self.exp_avg is None:
            self.exp_avg = torch.zeros_like(self.params32)
            self.exp_avg_sq = torch.zeros_like(self.params32)
        _lib.check(self.l.b200mart_adamw(self.h, self.params32.data_ptr(), self.params16.data_ptr(), self.grads32.data_ptr(),
                                         self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), int(step), float(lr),
                                         float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(pretrain),
                                         int(begin), int(end), int(zero_grads), _lib.cur_stream()))
        self.mirror_stale = False

    # ------------------------------------------------------------------ instrumentation (bench / profiles only)
    def launch_count(self):
        return self.l.b200mart_launch_count()

    def profile_begin(self):
        _lib.check(self.l.b200mart_profile_begin(self.h))

    def profile_end(self):
        """{'gemm'|'attention'|'fusion': (device ms, launches)}, gemm flops."""
        ms, n, fl = (C.c_double * 3)(), (_lib.i64 * 3)(), C.c_double()
        _lib.check(self.l.b200mart_profile_end(self.h, _lib.cur_stream(), C.byref(ms), C.byref(n), C.byref(fl)))
        self.last_profile_gemm_bytes = float(self.l.b200mart_profile_gemm_bytes(self.h))
        return {k: (ms[i], n[i]) for i, k in enumerate(("gemm", "attention", "fusion"))}, fl.value

    def status(self):
        if self.ws is None:
            return 0
        bufs = self._bufs(*self.ws_key)
        flags = _lib.i32()
        _lib.check(self.l.b200mart_status(self.h, C.byref(bufs), _lib.cur_stream(), C.byref(flags)))
        return flags.value
