"""Epoch-level evaluation metrics of the MarT path and their data-parallel aggregation.

Reference: `TransformerLitModel.validation_epoch_end` / `test_epoch_end`, MarT/lit_models/transformer.py:173-222 — the
per-batch rank arrays returned by `_eval` (:115-166) are concatenated and reduced to Hits@{1,3,5,10,20}, mean rank and
MRR.  The reference evaluates on one device; under data parallelism every rank holds the ranks of its own shard, so they
are all-gathered first (shards may differ in length: the last batch of an epoch is ragged)."""
import numpy as np
import torch

HITS_AT = (1, 3, 5, 10, 20)


def epoch_metrics(ranks, prefix="Eval_entity"):
    """ranks: 1-based ranks of the gold entity (any int array-like, possibly empty).  Returns the dictionary the reference
    logs, with the same keys (`Eval_entity/hits1` ... `Eval_entity/mean_rank`, `Eval_entity/mrr`); {} when empty, as the
    reference logs nothing for an empty list (:175)."""
    r = np.asarray(torch.as_tensor(ranks).cpu() if torch.is_tensor(ranks) else ranks, dtype=np.int64).reshape(-1)
    if r.size == 0:
        return {}
    if (r < 1).any():
        raise ValueError("ranks are 1-based")
    out = {f"{prefix}/hits{k}": float((r <= k).mean()) for k in HITS_AT}
    out[f"{prefix}/mean_rank"] = float(r.mean())
    out[f"{prefix}/mrr"] = float((1.0 / r).mean())
    return out


def gather_ranks(ranks, group=None):
    """All ranks' rank vectors, concatenated in rank order, on every rank.  `ranks`: 1-D integer tensor on the device the
    process group communicates on (CUDA for NCCL, CPU for gloo); lengths may differ between ranks."""
    import torch.distributed as dist
    ranks = ranks.reshape(-1).to(torch.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return ranks
    world = dist.get_world_size(group)
    n = torch.tensor([ranks.numel()], dtype=torch.int64, device=ranks.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    padded = torch.zeros(cap, dtype=torch.int64, device=ranks.device)
    padded[: ranks.numel()] = ranks
    parts = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])


def evaluate(trainer, batches, group=None, prefix="Eval_entity"):
    """Evaluation epoch of a `MartTrainer` over an iterable of device batches: `_eval` per batch (fused [MASK]-row head +
    rank-of-gold kernel), ranks gathered over the data-parallel ranks, then `epoch_metrics`.  Returns (metrics, ranks)."""
    chunks = [trainer.eval_step(b)[0].reshape(-1) for b in batches]
    local = torch.cat(chunks) if chunks else torch.zeros(0, dtype=torch.int64, device=trainer.device)
    allr = gather_ranks(local, group if group is not None else trainer.pg)
    return epoch_metrics(allr, prefix), allr
