"""Epoch metrics (Hits@k / mean rank / MRR) and their data-parallel aggregation — CPU only.
Reference: TransformerLitModel.validation_epoch_end, MarT/lit_models/transformer.py:173-193."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mkg_analogy_b200.metrics import epoch_metrics, gather_ranks


def _reference_epoch_end(outputs):
    """transformer.py:173-193 restated with the reference's own numpy expressions (logged keys -> values)."""
    entity_ranks = [_["entity_ranks"] for _ in outputs if "entity_ranks" in _]
    log = {}
    if len(entity_ranks) > 0:
        entity_ranks = np.concatenate(entity_ranks)
        log["Eval_entity/hits20"] = (entity_ranks <= 20).mean()
        log["Eval_entity/hits10"] = (entity_ranks <= 10).mean()
        log["Eval_entity/hits5"] = (entity_ranks <= 5).mean()
        log["Eval_entity/hits3"] = (entity_ranks <= 3).mean()
        log["Eval_entity/hits1"] = (entity_ranks <= 1).mean()
        log["Eval_entity/mean_rank"] = entity_ranks.mean()
        log["Eval_entity/mrr"] = (1. / entity_ranks).mean()
    return log


def test_epoch_metrics_match_reference_expressions():
    rng = np.random.default_rng(0)
    outputs = [dict(entity_ranks=rng.integers(1, 2064, size=n)) for n in (32, 32, 7)]   # ragged last batch
    ref = _reference_epoch_end(outputs)
    got = epoch_metrics(np.concatenate([o["entity_ranks"] for o in outputs]))
    assert set(got) == set(ref)
    for k in ref:
        assert got[k] == pytest.approx(float(ref[k]), rel=0, abs=1e-12), k
    # known answers
    m = epoch_metrics(torch.tensor([1, 2, 4, 11, 21]))
    assert m["Eval_entity/hits1"] == 0.2 and m["Eval_entity/hits3"] == 0.4 and m["Eval_entity/hits5"] == 0.6
    assert m["Eval_entity/hits10"] == 0.6 and m["Eval_entity/hits20"] == 0.8
    assert m["Eval_entity/mean_rank"] == pytest.approx(39 / 5) and m["Eval_entity/mrr"] == pytest.approx((1 + .5 + .25 + 1 / 11 + 1 / 21) / 5)


def test_epoch_metrics_edge_cases():
    assert epoch_metrics([]) == {}                       # the reference logs nothing for an empty epoch
    assert epoch_metrics(np.array([1]))["Eval_entity/mrr"] == 1.0
    with pytest.raises(ValueError):
        epoch_metrics([0, 3])                            # ranks are 1-based (argsort(argsort) + 1)
    assert "Test_entity/hits1" in epoch_metrics([2], prefix="Test_entity")


def test_gather_ranks_single_process_is_identity():
    r = torch.tensor([[3, 1], [9, 2]])
    assert torch.equal(gather_ranks(r), torch.tensor([3, 1, 9, 2]))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    local = torch.arange(1, 4 + 3 * rank) + 100 * rank           # lengths 3 and 6: ragged shards
    allr = gather_ranks(local)
    empty = gather_ranks(torch.zeros(0, dtype=torch.int64) if rank == 0 else torch.tensor([5]))
    if rank == 0:
        torch.save(dict(allr=allr, empty=empty), out)
    dist.destroy_process_group()


def test_gather_ranks_world2_ragged(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    g = torch.load(out)
    assert torch.equal(g["allr"], torch.cat([torch.arange(1, 4), torch.arange(1, 7) + 100]))
    assert torch.equal(g["empty"], torch.tensor([5]))            # an empty shard contributes nothing
    m = epoch_metrics(g["allr"])
    assert m["Eval_entity/hits3"] == pytest.approx(3 / 9)
