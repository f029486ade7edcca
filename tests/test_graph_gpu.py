"""CUDA-graph replay of the training step and sequence-length bucketing: `MartTrainer.enable_graphs()` captures
forward + [MASK]-row head + backward + the pipelined per-layer AdamW once per (batch size, L bucket) and replays it, with the per-step
scalars (dropout salt, learning rate, AdamW bias corrections) read from the device step block (b200mart_set_step_block).  Checked:
replayed steps == eager steps on the same batches (summation-order noise of the atomics only), batches of different lengths share a
bucket's graph and match their unpadded eager results, the salted dropout masks match the oracle mask for mask, and replays draw a fresh
mask every step."""
import pytest
import torch

from mkg_analogy_b200.config import TINY, init_params
from mkg_analogy_b200.data import TINY_VOCAB, make_finetune_batch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def to_cuda(batch):
    return {k: v.cuda().contiguous() for k, v in batch.items()}


def _trainer(params, **kw):
    from mkg_analogy_b200.trainer import MartTrainer
    tr = MartTrainer(TINY, TINY_VOCAB, total_steps=40, lr=1e-3, **kw)
    tr.engine.load(params)
    return tr


def test_graph_replay_matches_eager_steps_and_buckets_share_a_graph():
    params = init_params(TINY, 11)
    lens = [24, 19, 22, 24, 17, 30, 21, 24]                         # 30 falls in the next bucket
    batches = [to_cuda(make_finetune_batch(4, L, TINY.image_size, TINY_VOCAB, seed=50 + i, min_len=min(16, L))) for i, L in enumerate(lens)]
    eager = _trainer(params, dropout=False)
    l_eager = [eager.train_step(b).item() for b in batches]
    graphed = _trainer(params, dropout=False).enable_graphs(buckets=(24, 32), warmup=2)
    l_graph = [graphed.train_step(b).item() for b in batches]
    torch.cuda.synchronize()
    ents = graphed._graphs["entries"]
    assert sorted(k[1] for k in ents) == [24, 32]
    e24 = next(v for k, v in ents.items() if k[1] == 24)
    assert e24["graph"] is not None and e24["n"] == 7                  # two eager warm-up steps, then capture + replays
    print("eager  ", [f"{x:.4f}" for x in l_eager])
    print("graphed", [f"{x:.4f}" for x in l_graph])
    assert max(abs(a - b) for a, b in zip(l_eager, l_graph)) <= 3e-3   # atomics / padded-tile summation order only
    d = (eager.engine.params32 - graphed.engine.params32).abs()
    # Adam's update is sign-like: where a gradient is ~0, summation-order noise can move a parameter by up to 2 lr per step
    assert d.max().item() <= len(lens) * 2e-3 + 1e-6 and d.mean().item() <= 5e-5, (d.max().item(), d.mean().item())
    assert graphed.engine.grads32.abs().max().item() == 0.0
    assert graphed.steps_done == len(lens)


def test_step_block_dropout_masks_match_oracle_and_change_per_replay():
    from oracle import mkgformer_oracle as O
    params = init_params(TINY, 11)
    batch = to_cuda(make_finetune_batch(4, 24, TINY.image_size, TINY_VOCAB, seed=30, min_len=16))
    tr = _trainer(params, dropout=True, seed=5)
    tr._enable_step_block()
    eng = tr.engine
    # eager forward / backward with a salt in the step block: masks = hash(idx ^ site_seed ^ salt)
    tr._next_step()
    salt = tr.last_salt
    eng.train(True)
    eng.zero_grad()
    z, _ = eng.forward(batch, want_z=True)
    seed = eng.last_dropout_seed()
    head = dict(mode="finetune", col_ids=tr.col_ids, label=batch["label"], rel_idx=batch["rel_idx"], q_head_idx=batch["q_head_idx"],
                a_head_idx=batch["a_head_idx"], alpha=0.43, label_smoothing=0.1)
    loss, _, _ = eng.head_loss(batch, head, compute_grad=True)
    eng.backward(batch)
    torch.cuda.synchronize()
    Pg = {k: v.cuda().clone().requires_grad_(True) for k, v in params.items()}
    ref_loss, _, zr = O.finetune_loss(Pg, TINY, batch, TINY_VOCAB.analogy_entity_ids(), alpha=0.43, label_smoothing=0.1,
                                      drop=O.hash_dropout(seed, TINY.hidden_dropout, TINY.attn_dropout, salt=salt))
    ref_loss.backward()
    _, _, z0 = O.finetune_loss({k: v.cuda() for k, v in params.items()}, TINY, batch, TINY_VOCAB.analogy_entity_ids(), alpha=0.43,
                               label_smoothing=0.1, drop=O.hash_dropout(seed, TINY.hidden_dropout, TINY.attn_dropout))
    print(f"salted dropout: loss cuda {loss.item():.5f} oracle {ref_loss.item():.5f}, z rel err {rel(z, zr.detach()):.3e} "
          f"(unsalted masks: {rel(z, z0.detach()):.3e})")
    assert salt != 0 and rel(z, zr.detach()) <= 2e-2 and abs(loss.item() - ref_loss.item()) <= 2e-2
    assert rel(z, z0.detach()) > 5e-2                                   # the salt does change the masks
    for name in ("unimo.encoder.text_layer.1.attention.self.query.weight", "unimo.encoder.text_layer.3.output.dense.weight",
                 "unimo.encoder.text_layer.0.attention.output.dense.bias", "unimo.text_embeddings.LayerNorm.weight"):
        assert rel(eng.grad_view(name), Pg[name].grad) <= 6e-2, name
    # replays of one captured graph on one batch with lr = 0: the loss moves only because each replay reads a new salt
    tr2 = _trainer(params, dropout=True, seed=5)
    tr2.lr = 0.0
    tr2.weight_decay = 0.0
    tr2.enable_graphs(buckets=(24,), warmup=2)
    losses = [tr2.train_step(batch).item() for _ in range(6)]
    torch.cuda.synchronize()
    assert len({round(x, 4) for x in losses[2:]}) >= 3, losses         # replayed steps 3..6 see different masks
    tr3 = _trainer(params, dropout=False)
    tr3.lr = 0.0
    tr3.weight_decay = 0.0
    tr3.enable_graphs(buckets=(24,), warmup=2)
    same = [tr3.train_step(batch).item() for _ in range(5)]
    assert max(same) - min(same) <= 2e-3, same                         # without dropout nothing changes (lr = 0)
