"""Fused training / evaluation step of the MarT hot path on one GPU (one process per GPU when data-parallel).

This is the B200-native counterpart of what `TransformerLitModel.training_step` + `loss.backward()` + the AdamW /
linear-warmup scheduler of `configure_optimizers` do in the reference (MarT/lit_models/transformer.py:59-113, :224-241;
MarT/lit_models/base.py:30-32,:86-91), with the score head evaluated on the [MASK] rows only (SURVEY.md §8 b4
"fused" mode).  Data parallelism: every rank holds a replica, processes its own shard of the global
batch, and the flat fp32 gradient buffer is averaged across ranks with NCCL all-reduce before the optimizer step.
"""
import torch

from .config import MKGformerConfig
from .data import VocabLayout
from .engine import Engine

BATCH_KEYS = ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "pixel_idx", "sep_idx", "label", "rel_idx",
              "q_head_idx", "a_head_idx", "pre_type")


def linear_warmup_lr(step, total_steps, base_lr, warmup_ratio):
    """get_linear_schedule_with_warmup as used at transformer.py:233-240: `step` optimizer steps already taken."""
    warm = total_steps * warmup_ratio
    if step < warm:
        return base_lr * step / max(1.0, warm)
    return base_lr * max(0.0, (total_steps - step) / max(1.0, total_steps - warm))


def step_salt(salt0, t):
    """32-bit word XORed into every dropout site seed at optimizer step t (b200mart_set_step_block): hash32 of
    (salt0 + 0x9E3779B9 t), the same finaliser the kernels use for the masks (csrc/common.cuh)."""
    x = (salt0 + 0x9E3779B9 * t) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def pack_step_block(block, salt, lr, t, betas=(0.9, 0.999)):
    """Fill a 4 x int32 host tensor with {uint32 salt, float lr, float 1 - beta1^t, float 1 - beta2^t} (include/b200mart.h)."""
    block[0] = salt - (1 << 32) if salt >= (1 << 31) else salt
    f = block.view(torch.float32)
    f[1], f[2], f[3] = lr, 1.0 - betas[0] ** t, 1.0 - betas[1] ** t
    return block


def pick_bucket(L, buckets):
    """Smallest sequence-length bucket that holds L (graph replay pads the text rows up to it)."""
    for b in buckets:
        if b >= L:
            return b
    raise ValueError(f"sequence length {L} exceeds the largest graph bucket {buckets[-1]}")


def bucket_bounds(n, buckets, align=1024):
    """Contiguous [start, end) ranges covering a flat buffer of n elements in about `buckets` aligned pieces."""
    per = (n + buckets - 1) // max(1, buckets)
    per = max(align, (per + align - 1) // align * align)
    return [(st, min(n, st + per)) for st in range(0, n, per)]


def allreduce_mean_(flat, group=None, buckets=8):
    """In-place mean over the data-parallel ranks of a flat gradient buffer, a few large buckets at a time
    (equal-sized shards, so the mean of per-rank batch-mean gradients is the global-batch gradient).
    NCCL averages in the collective itself; other backends (gloo in the CPU tests) sum, then scale."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    avg = dist.get_backend(group) == "nccl"
    for st, ed in bucket_bounds(flat.numel(), buckets):
        piece = flat[st:ed]
        if avg:
            dist.all_reduce(piece, op=dist.ReduceOp.AVG, group=group)
        else:
            dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=group)
            piece.mul_(1.0 / world)
    return flat


def shard_seed(base_seed, rank, step=0):
    """Seed of the synthetic shard a rank draws at a step: distinct across ranks and steps, reproducible."""
    return base_seed + 100003 * rank + step


def init_weights_(engine: Engine, seed=0):
    """Seeded synthetic weights generated on the device: N(0, 0.02) for matrices / embeddings / biases,
    LayerNorm (1, 0), class_embedding ~ N(0, 1), adaptive weights U(0, 0.5) and exactly 0.5 (modeling_unimo.py:305-310)."""
    g = torch.Generator(device=engine.device).manual_seed(seed)
    with torch.no_grad():
        engine.params32.normal_(0.0, 0.02, generator=g)
        for info in engine.infos:
            n, v = info.name, engine.view(info.name)
            if "LayerNorm.weight" in n or "layer_norm1.weight" in n or "layer_norm2.weight" in n or \
                    n.endswith("layrnorm.weight") or n.endswith("layernorm.weight"):
                v.fill_(1.0)
            elif "LayerNorm.bias" in n or "layer_norm1.bias" in n or "layer_norm2.bias" in n or \
                    n.endswith("layrnorm.bias") or n.endswith("layernorm.bias"):
                v.zero_()
            elif n.endswith("adaptive_weight.0"):
                v.uniform_(0.0, 0.5, generator=g)
            elif n.endswith("adaptive_weight.1"):
                v.fill_(0.5)
            elif n.endswith("class_embedding"):
                v.normal_(0.0, 1.0, generator=g)
    engine.sync_mirror()


class MartTrainer:
    def __init__(self, cfg: MKGformerConfig, vocab: VocabLayout, *, pretrain=False, lr=5e-5, weight_decay=0.01,
                 warmup_ratio=0.1, total_steps=1000, alpha=0.43, label_smoothing=0.1, device=None, process_group=None,
                 grad_buckets=8, overlap_comm=True, dropout=True, seed=0, status_every=200, shard_optimizer=False, multicast=None):
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.cfg, self.vocab, self.pretrain = cfg, vocab, bool(pretrain)
        self.engine = Engine(cfg, device)
        self.device = self.engine.device
        self.lr, self.weight_decay, self.warmup_ratio, self.total_steps = lr, weight_decay, warmup_ratio, total_steps
        self.alpha, self.label_smoothing = alpha, label_smoothing
        self.col_ids = torch.as_tensor(vocab.analogy_entity_ids(), device=self.device, dtype=torch.int64)
        self.steps_done = 0
        # device-side index guards (token id / type, label, relaxation position, [MASK] count) only raise status bits; the
        # reference raises IndexError / asserts on the spot, so the flags are polled every `status_every` training steps,
        # at every evaluation step and before a checkpoint is written (one stream synchronisation each)
        self.status_every = int(status_every)
        self.pg = process_group
        self.world = 1
        if process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            self.world = torch.distributed.get_world_size(process_group)
        self.grad_buckets = grad_buckets
        # N > 1, NCCL route, optional: reduce-scatter -> AdamW on this rank's 1/N of each extent -> all-gather of the fp32 masters
        # (SURVEY.md §8 f1).  Measured on B200 (profiles/scaling_r2.txt): slower than all-reduce + replicated AdamW at N = 2 and
        # N = 8 (three dependent stages per extent on the exchange stream), so off by default; the fused multicast kernel below is
        # the sharded optimizer that pays.
        self.shard_optimizer = bool(shard_optimizer) and self.world > 1
        self.rank = torch.distributed.get_rank(process_group) if self.world > 1 else 0
        # NVSwitch multicast exchange (default for N >= 2 when the box supports it; B200MART_MULTICAST=1 / 0 or multicast=True /
        # False force it on / off): gradients (two buffers, alternating per step), fp32 masters and the bf16 mirror live in
        # symmetric memory; ONE kernel per layer extent pulls the mean gradient of the rank's shard with multimem.ld_reduce, applies
        # AdamW and stores the result in every replica with multimem.st.  Measured (profiles/scaling_r2.txt): 8 GPUs 8.17 ms / step
        # against 9.54 ms with NCCL all-reduce; 2 GPUs 8.17 against 8.33 ms.
        self._mc = None
        self._mc_prev = None
        if multicast is None:
            import os
            env = os.environ.get("B200MART_MULTICAST")
            multicast = (self.world >= 2) if env is None else (env != "0")
        if multicast and self.world > 1 and overlap_comm:
            self._setup_multicast()
        self._dev_batch = None
        self._graphs = None          # CUDA-graph replay of the step (enable_graphs)
        self._step_dev = None
        # training-mode dropout of the BERT stream (reference: hidden 0.1, attention 0.1); per-rank seed
        self.dropout = bool(dropout)
        rank = torch.distributed.get_rank(process_group) if self.world > 1 else 0
        self.engine.set_dropout(cfg.hidden_dropout if dropout else 0.0, cfg.attn_dropout if dropout else 0.0,
                                shard_seed(seed, rank))
        # pipelined tail of the step: per-layer "gradients final" events recorded by the library; on a side stream each
        # layer's gradient exchange (N > 1) and AdamW update run while the layers below are still in backward
        self.overlap_comm = bool(overlap_comm)
        self._comm_stream = self._layer_events = None
        self._grads_clean = False
        if self.overlap_comm:
            import os
            self._tail_chunks = max(1, int(os.environ.get("B200MART_TAIL_CHUNKS", "1")))
            self._comm_stream = torch.cuda.Stream(device=self.device)
            self._update_stream = torch.cuda.Stream(device=self.device)
            self._layer_events = [torch.cuda.Event() for _ in range(2 * cfg.layers)]
            for e in self._layer_events:
                e.record()                       # materialise the CUDA event handles
            self._extents = self.engine.layer_extents()
            self.engine.exp_avg = torch.zeros_like(self.engine.params32)
            self.engine.exp_avg_sq = torch.zeros_like(self.engine.params32)

    def _setup_multicast(self):
        """Re-home grads32 / params32 / params16 in torch symmetric memory and fetch their multicast addresses.  Any failure
        (no NVSwitch multicast on this box, old driver) leaves the NCCL path in place."""
        import os
        import torch.distributed as dist
        try:
            import torch.distributed._symmetric_memory as symm_mem
            eng, group = self.engine, (self.pg if self.pg is not None else dist.group.WORLD)
            out = {}
            for name in ("grads32", "params32", "params16"):
                old = getattr(eng, name)
                t = symm_mem.empty(old.numel(), dtype=old.dtype, device=self.device)
                t.copy_(old)
                hdl = symm_mem.rendezvous(t, group=group)
                if not hdl.multicast_ptr:
                    raise RuntimeError("no multicast address (NVSwitch multicast unavailable)")
                out[name] = (t, hdl)
            # second gradient buffer: step t accumulates into one while the other (consumed by step t - 1) is cleared off the
            # critical path
            g2 = symm_mem.empty(eng.grads32.numel(), dtype=eng.grads32.dtype, device=self.device)
            g2.zero_()
            h2 = symm_mem.rendezvous(g2, group=group)
            if not h2.multicast_ptr:
                raise RuntimeError("no multicast address (NVSwitch multicast unavailable)")
            for name, (t, _) in out.items():
                setattr(eng, name, t)
            eng.ws_key = None
            self._mc = {k: h for k, (_, h) in out.items()}
            self._mc_grads = [(out["grads32"][0], out["grads32"][1]), (g2, h2)]
            self._mc_cur = 0
            self._mc_tail_ctas = int(os.environ.get("B200MART_MC_TAIL_CTAS", "128"))
            torch.cuda.synchronize(self.device)
        except Exception as e:      # noqa: BLE001
            import warnings
            warnings.warn(f"NVSwitch multicast exchange unavailable ({type(e).__name__}: {e}); using NCCL")
            self._mc = None

    # ------------------------------------------------------------------ pieces
    def _head(self, batch):
        if self.pretrain:
            return dict(mode="pretrain", entity_range=(self.vocab.entity_id_st, self.vocab.entity_id_ed),
                        relation_range=(self.vocab.relation_id_st, self.vocab.relation_id_ed), label=batch["label"],
                        pre_type=batch["pre_type"], label_smoothing=self.label_smoothing)
        return dict(mode="finetune", col_ids=self.col_ids, label=batch["label"], rel_idx=batch["rel_idx"],
                    q_head_idx=batch["q_head_idx"], a_head_idx=batch["a_head_idx"], alpha=self.alpha,
                    label_smoothing=self.label_smoothing)

    def _model_inputs(self, batch):
        b = {k: batch[k] for k in ("input_ids", "attention_mask", "pixel_values")}
        if batch.get("token_type_ids") is not None:
            b["token_type_ids"] = batch["token_type_ids"]
        if not self.pretrain and batch.get("sep_idx") is not None:
            b["sep_idx"] = batch["sep_idx"]      # the pre-train collator supplies no sep_idx
        return b

    def allreduce_grads(self):
        """Average the flat gradient buffer over the data-parallel ranks (NCCL over NVLink). A few large buckets:
        launch latency, not link count, is what matters on NVSwitch."""
        if self.world == 1:
            return
        allreduce_mean_(self.engine.grads32, self.pg, self.grad_buckets)

    def _backward_overlapped(self, mb, update=None):
        """Backward with the tail of the step pipelined per layer: as soon as layer i's backward has run on both chains
        (layers are walked top-down) a side stream averages that layer's gradient extent over the ranks (one NCCL
        all-reduce of about 66 MB at base size, N > 1 only) and, when `update` = (step, lr) is given, applies AdamW to
        it and clears its gradients, while the layers below are still in backward.  Embeddings / head, whose gradients
        are final only at the end, go last.  A layer's parameters are dead once both chains have left that layer, so
        updating them (and their bf16 mirror) early is safe."""
        import torch.distributed as dist
        eng, g = self.engine, self.engine.grads32
        main, comm = torch.cuda.current_stream(), self._comm_stream
        eng.backward(mb, layer_events=self._layer_events)

        # N > 1: the exchange (NVLink-bound) and the update (HBM-bound) of consecutive layers overlap on two side streams
        upd = self._update_stream if (self.world > 1 and update is not None) else comm

        def finish(b, e, tail=False):
            if self._mc is not None and update is not None and (e - b) % (8 * self.world) == 0:
                # fused exchange + update of the extent [b, e) over NVSwitch multicast; this rank owns [sb, se).
                # The barrier (every rank's local gradients of this extent are complete) runs on the exchange stream, AHEAD of the
                # update stream: waiting for the slowest rank's layer never delays the optimizer kernel of the extent before.
                S = (e - b) // self.world
                sb, se = b + self.rank * S, b + (self.rank + 1) * S
                hg = self._mc_grads[self._mc_cur][1]
                hg.barrier(channel=0)
                ev = torch.cuda.Event()
                ev.record(comm)
                upd.wait_event(ev)
                with torch.cuda.stream(upd):
                    eng.adamw_multicast(update[0], update[1], hg.multicast_ptr, self._mc["params32"].multicast_ptr,
                                        self._mc["params16"].multicast_ptr, self.world, sb, se, weight_decay=self.weight_decay,
                                        pretrain=self.pretrain, max_ctas=self._mc_tail_ctas if tail else 0)
                self._mc_prev = (b, e)
                return
            if self.shard_optimizer and update is not None and (e - b) % (8 * self.world) == 0:
                # sharded update of the extent [b, e): this rank owns [sb, se)
                S = (e - b) // self.world
                sb, se = b + self.rank * S, b + (self.rank + 1) * S
                dist.reduce_scatter_tensor(g[sb:se], g[b:e], op=dist.ReduceOp.AVG, group=self.pg)
                ev = torch.cuda.Event()
                ev.record(comm)
                upd.wait_event(ev)
                with torch.cuda.stream(upd):
                    eng.adamw(update[0], update[1], weight_decay=self.weight_decay, pretrain=self.pretrain, begin=sb, end=se,
                              zero_grads=True)
                    if sb > b:
                        g[b:sb].zero_()              # the other ranks' shards of the (consumed) local gradient
                    if se < e:
                        g[se:e].zero_()
                    ev2 = torch.cuda.Event()
                    ev2.record(upd)
                comm.wait_event(ev2)
                dist.all_gather_into_tensor(eng.params32[b:e], eng.params32[sb:se], group=self.pg)
                ev3 = torch.cuda.Event()
                ev3.record(comm)
                upd.wait_event(ev3)
                with torch.cuda.stream(upd):         # bf16 mirror of what the other ranks delivered (AdamW wrote this rank's part)
                    eng.cast_range(b, sb)
                    eng.cast_range(se, e)
                return
            if self.world > 1:
                dist.all_reduce(g[b:e], op=dist.ReduceOp.AVG, group=self.pg)
            if update is not None:
                if upd is not comm:
                    ev = torch.cuda.Event()
                    ev.record(comm)
                    upd.wait_event(ev)
                with torch.cuda.stream(upd):
                    eng.adamw(update[0], update[1], weight_decay=self.weight_decay, pretrain=self.pretrain, begin=b, end=e,
                              zero_grads=True)

        with torch.cuda.stream(comm):
            for i in reversed(range(self.cfg.layers)):
                comm.wait_event(self._layer_events[2 * i])
                comm.wait_event(self._layer_events[2 * i + 1])
                finish(*self._extents[i])
            comm.wait_stream(main)               # the whole backward (embeddings included) is ordered before this point
            lo, hi = self._extents[0][0], self._extents[-1][1]
            # B200MART_TAIL_CHUNKS=n (default 1): the embedding / head extents, which only become final at the end of backward,
            # go in n pieces so that (N > 1) the exchange of piece k+1 overlaps the update of piece k
            for b0, e0 in ((0, lo), (hi, g.numel())):
                if e0 > b0:
                    for b, e in bucket_bounds(e0 - b0, self._tail_chunks):
                        finish(b0 + b, b0 + e, tail=True)
        if self._mc is not None and self._mc_prev is not None:
            # closing barrier of the step, behind this rank's last optimizer kernel: every rank has read every shard of this step's
            # gradient buffer and delivered every new parameter.  The next step may start; the consumed buffer is cleared on the
            # update stream while the next step's forward runs (that step accumulates into the other buffer).
            with torch.cuda.stream(upd):
                self._mc_grads[self._mc_cur][1].barrier(channel=1)
                done = torch.cuda.Event()
                done.record(upd)
                g.zero_()
            main.wait_event(done)
            main.wait_stream(comm)
            self._mc_prev = None
            self._mc_cur ^= 1
            eng.grads32 = self._mc_grads[self._mc_cur][0]
            return
        if upd is not comm:
            main.wait_stream(upd)
        main.wait_stream(comm)

    def check_status(self):
        """Raise IndexError (as the reference's indexing would) when a device-side guard fired since the last forward."""
        flags = self.engine.status()
        if self._graphs is not None:          # graphs of other buckets run on the workspace they were captured on (status word first)
            for ent in self._graphs["entries"].values():
                ws = ent.get("ws")
                if ws is not None and ws is not self.engine.ws:
                    flags |= int(ws[:4].view(torch.int32).item()) & 15
        if flags:
            names = [n for bit, n in ((1, "token id / score-head column out of range"), (2, "token type out of range"),
                                      (4, "a row without exactly one [MASK]"), (8, "label or relaxation position out of range"))
                     if flags & bit]
            raise IndexError(f"b200mart status flags {flags:#x}: " + "; ".join(names))

    # ------------------------------------------------------------------ public API
    def _step_body(self, batch, step, lr):
        """forward -> [MASK]-row head + loss -> backward with the per-layer exchange / AdamW tail.  Everything here is stream work with
        device-resident arguments (capturable into a CUDA graph once the shapes are warm)."""
        eng = self.engine
        mb = self._model_inputs(batch)
        eng.train(True)
        if not self._grads_clean:
            eng.zero_grad()
        eng.forward(mb, want_z=False)
        loss, _, _ = eng.head_loss(mb, self._head(batch), compute_grad=True)
        if self.overlap_comm:
            self._backward_overlapped(mb, update=(step, lr))
            self._grads_clean = True             # AdamW cleared every gradient range it consumed
        else:
            eng.backward(mb)
            self.allreduce_grads()
            eng.adamw(step, lr, weight_decay=self.weight_decay, pretrain=self.pretrain)
        return loss

    def _next_step(self):
        """Advance the optimizer step counter; returns (step, lr) and, in step-block mode, publishes {salt, lr, 1 - beta1^t,
        1 - beta2^t} to the device block the kernels read (stream-ordered before the step)."""
        lr = linear_warmup_lr(self.steps_done, self.total_steps, self.lr, self.warmup_ratio)
        self.steps_done += 1
        t = self.steps_done
        if self._step_dev is not None:
            k = t % len(self._step_host)
            if self._step_ev[k] is not None:
                self._step_ev[k].synchronize()            # the copy that last read this pinned slot has run (never blocks in practice)
            self.last_salt = step_salt(self._salt0, t)            # this step's dropout masks
            h = pack_step_block(self._step_host[k], self.last_salt, lr, t)
            self._step_dev.copy_(h, non_blocking=True)
            self._step_ev[k] = torch.cuda.Event()
            self._step_ev[k].record()
        return t, lr

    def train_step(self, batch):
        """One optimizer step on a device-resident batch. Returns the loss as a device tensor (1,)."""
        if self._graphs is not None:
            return self._train_step_graphed(batch)
        step, lr = self._next_step()
        loss = self._step_body(batch, step, lr)
        if self.status_every and self.steps_done % self.status_every == 0:
            self.check_status()
        return loss

    # ------------------------------------------------------------------ CUDA-graph replay of the step
    def enable_graphs(self, buckets=(32, 64, 96, 128), warmup=2):
        """Replay the training step as ONE CUDA graph per (batch size, sequence-length bucket): the ~520 launches of a step, the forks /
        joins of the library's five streams and the pipelined AdamW tail are captured once (after `warmup` eager steps of that shape
        have created the GEMM plans and the workspace) and re-issued with a single cudaGraphLaunch.  What changes per step lives in
        memory: the batch is copied into the graph's static input buffers (text rows zero-padded up to the bucket: padded keys are
        masked by the additive -1e4, padded rows feed nothing, so the real rows' results are those of the unpadded call), and the
        dropout salt / learning rate / AdamW bias corrections are read from the device step block (b200mart_set_step_block).
        One process per GPU without a gradient exchange only (world size 1); returns self."""
        if self.world != 1:
            raise NotImplementedError("graph replay covers the single-GPU step (the exchange paths run eagerly)")
        self._graphs = {"buckets": tuple(sorted(int(b) for b in buckets)), "warmup": int(warmup), "entries": {}}
        self._enable_step_block()
        return self

    def _enable_step_block(self):
        if self._step_dev is None:
            self._step_host = [torch.zeros(4, dtype=torch.int32).pin_memory() for _ in range(8)]
            self._step_ev = [None] * 8
            self._step_dev = torch.zeros(4, dtype=torch.int32, device=self.device)
            self._salt0 = (0x51ED27 * (self.rank + 1) + 977 * int(torch.initial_seed())) & 0xFFFFFFFF
            self.last_salt = 0
            self.engine.set_step_block(self._step_dev)

    def _train_step_graphed(self, batch):
        g = self._graphs
        B, L = batch["input_ids"].shape
        Lb = pick_bucket(L, g["buckets"])
        keys = tuple(sorted(k for k in batch if k in BATCH_KEYS and k != "pixel_idx"))
        ent = g["entries"].get((B, Lb, keys))
        if ent is None:
            static = {}
            for k in keys:
                v = batch[k]
                shape = (B, Lb) if k in ("input_ids", "attention_mask", "token_type_ids") else tuple(v.shape)
                static[k] = torch.zeros(shape, dtype=v.dtype, device=self.device)
            ent = g["entries"][(B, Lb, keys)] = {"static": static, "n": 0, "graph": None, "loss": None}
        for k, dst in ent["static"].items():
            src = batch[k]
            if dst.shape != src.shape:                  # text rows: pad up to the bucket ([PAD] = 0, attention_mask 0)
                dst[:, :L].copy_(src)
                dst[:, L:].zero_()
            elif dst.data_ptr() != src.data_ptr():
                dst.copy_(src)
        step, lr = self._next_step()
        if ent["n"] < g["warmup"]:
            ent["loss"] = self._step_body(ent["static"], step, lr)
        else:
            if ent["graph"] is None:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    ent["loss"] = self._step_body(ent["static"], step, lr)
                ent["graph"] = graph
                # the captured kernels carry raw pointers into the engine's workspace: keep that allocation alive for as long as
                # the graph exists (a later, larger bucket makes the engine allocate a new workspace and drop its own reference)
                ent["ws"] = self.engine.ws
            ent["graph"].replay()
        ent["n"] += 1
        if self.status_every and self.steps_done % self.status_every == 0:
            self.check_status()
        return ent["loss"]

    def attach_image_table(self, table):
        """Serve `pixel_values` from a device-resident `EntityImageTable`: host batches then carry `pixel_idx`
        (B, 2) int64 table rows (-1 = the all-zero "[PAD] image", as `MarsCollator(images="index")` produces) instead of 38.5 MB
        of pixels; the gather runs on the training stream right before the step."""
        self.image_table = table
        self._px = {}

    def _resolve_images(self, batch, slot=0):
        if "pixel_idx" not in batch or "pixel_values" in batch:
            return batch
        if getattr(self, "image_table", None) is None:
            raise ValueError("the batch carries pixel_idx: attach an EntityImageTable first (attach_image_table)")
        idx = batch["pixel_idx"]
        shape = (idx.shape[0], 2) + self.image_table.image_shape
        if slot not in self._px or self._px[slot].shape != shape:
            self._px[slot] = torch.empty(shape, device=self.device, dtype=torch.float32)
        out = dict(batch)
        out["pixel_values"] = self.image_table.gather(idx, out=self._px[slot])
        return out

    def train_step_host(self, host_batch):
        """End-to-end step from (pinned) host tensors: H2D copies, step, D2H read of the loss. Returns a float."""
        return float(self.train_step(self._resolve_images(self.to_device(host_batch))).item())

    def train_step_host_async(self, host_batch):
        """Pipelined end-to-end step (what a training loop with asynchronous logging does): this step's inputs are copied
        host -> device on a copy stream into one of two device staging batches while the previous step is still
        computing, the step is launched behind that copy, and its loss is copied device -> host into pinned memory.
        Returns the PREVIOUS step's loss as a float (None on the first call): every step still pays one H2D copy of its
        inputs and one D2H read of a loss, but neither sits on the critical path.  `drain()` returns the last loss."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._stage = [None, None]
            self._stage_free = [None, None]
            self._loss_host = [torch.zeros(1).pin_memory(), torch.zeros(1).pin_memory()]
            self._loss_ev = [None, None]
            self._n_async = 0
        k = self._n_async & 1
        main = torch.cuda.current_stream()
        keys = {n for n in host_batch if n in BATCH_KEYS}
        if self._stage[k] is None or set(self._stage[k]) != keys or any(self._stage[k][n].shape != host_batch[n].shape for n in keys):
            self._stage[k] = {n: torch.empty(v.shape, dtype=v.dtype, device=self.device) for n, v in host_batch.items() if n in BATCH_KEYS}
            # The caching allocator may hand out blocks that tensors freed on the main stream used moments ago while work
            # reading / writing them is still queued there; the copy stream is not ordered behind that work.  One full
            # synchronisation when a staging batch is (re)allocated closes the window (steady state never allocates).
            torch.cuda.synchronize(self.device)
        with torch.cuda.stream(self._copy_stream):
            if self._stage_free[k] is not None:
                self._copy_stream.wait_event(self._stage_free[k])      # the step that last used this staging batch is done
            for n, d in self._stage[k].items():
                d.copy_(host_batch[n], non_blocking=True)
        main.wait_stream(self._copy_stream)
        loss = self.train_step(self._resolve_images(self._stage[k], slot=k))
        self._stage_free[k] = torch.cuda.Event()
        self._stage_free[k].record(main)
        self._loss_host[k].copy_(loss, non_blocking=True)
        self._loss_ev[k] = torch.cuda.Event()
        self._loss_ev[k].record(main)
        self._n_async += 1
        prev = k ^ 1
        if self._loss_ev[prev] is None:
            return None
        self._loss_ev[prev].synchronize()
        return float(self._loss_host[prev].item())

    def drain(self):
        """Loss of the most recent train_step_host_async call."""
        k = (self._n_async - 1) & 1
        self._loss_ev[k].synchronize()
        return float(self._loss_host[k].item())

    def to_device(self, host_batch):
        keys = {k for k in host_batch if k in BATCH_KEYS}
        if self._dev_batch is None or set(self._dev_batch) != keys or any(self._dev_batch[k].shape != host_batch[k].shape for k in keys):
            self._dev_batch = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host_batch.items()
                               if k in BATCH_KEYS}
        for k, d in self._dev_batch.items():
            d.copy_(host_batch[k], non_blocking=True)
        return self._dev_batch

    @torch.no_grad()
    def eval_step(self, batch):
        """_eval (transformer.py:115-166): ranks of the gold entity (1-based), and the [MASK]-row scores."""
        eng = self.engine
        mb = self._model_inputs(batch)
        eng.train(False)
        eng.forward(mb, want_z=False)
        loss, logits, ranks = eng.head_loss(mb, self._head(batch), compute_grad=False, want_logits=True, want_ranks=True)
        self.check_status()
        return ranks, logits, loss

    # ------------------------------------------------------------------ checkpoints (reference layout, main.py:133-159)
    def save_checkpoint(self, path):
        """`{"state_dict": {"model." + name: tensor}}` as the reference's ModelCheckpoint(save_weights_only=True) writes it."""
        from .checkpoint import save_checkpoint
        self.check_status()
        torch.cuda.synchronize(self.device)
        save_checkpoint(self.engine, path, global_step=self.steps_done)

    def load_checkpoint(self, path, strict=False):
        """main.py:134: load_state_dict(torch.load(path)["state_dict"], strict=False). Returns (missing, unexpected)."""
        from .checkpoint import load_checkpoint
        torch.cuda.synchronize(self.device)
        return load_checkpoint(self.engine, path, strict=strict)

    @staticmethod
    def h2d_bytes(host_batch):
        return int(sum(v.numel() * v.element_size() for k, v in host_batch.items() if k in BATCH_KEYS))
