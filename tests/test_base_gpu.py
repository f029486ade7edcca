"""Whole-model parity at the BENCHMARKED shape: MKGformer-base (12 x 768, 99 vision tokens, 42,007-token vocabulary).

Fixtures come from the reference itself (oracle/make_golden_base.py, run in the build container):
  tests/golden/mkgformer_base.pt   BASELINE.json config 1 (4 synthetic MARS examples, L = 64): [MASK]-row logits over the whole
                                   vocabulary, trans_hidden_states, the reference training_step loss, every gradient norm, the
                                   small gradients in full, `_eval` ranks; an odd length (L = 37); the L = 128 pre-train shape.
  tests/golden/base_eval_split.pt  the held-out decidable evaluation split (256 samples): reference `_eval` ranks.
Tolerances: bf16 tensor-core operands with fp32 accumulation against an fp32 reference — relative Frobenius error <= 1e-2 on
hidden states and logits (measured 2-4e-3; printed), loss |d| <= 1e-2, gradients per tensor <= 6e-2 (cosine >= 0.995), the
adaptive-weight scalars <= 5e-2 without any absolute floor.  Hits@1 / ranks on the evaluation split: EQUAL ON EVERY ROW.
"""
import os

import pytest
import torch

from mkg_analogy_b200.config import BASE, init_params
from mkg_analogy_b200.data import MARS_VOCAB, analogy_classes, make_decidable_batch, make_finetune_batch, make_pretrain_batch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
MODEL_KEYS = ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "sep_idx")


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def to_cuda(batch):
    return {k: v.cuda().contiguous() for k, v in batch.items()}


def checksum(ts):
    return float(sum(t.double().abs().sum().item() for t in ts))


@pytest.fixture(scope="module")
def base():
    from mkg_analogy_b200.engine import Engine
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.load(os.path.join(GOLD, "mkgformer_base.pt"), weights_only=False)
    params = init_params(BASE, g["seed_w"])
    chk = checksum([params["unimo.text_embeddings.word_embeddings.weight"], params["unimo.encoder.text_layer.11.output.dense.weight"]])
    assert abs(chk - g["param_checksum"]) < 1e-6 * g["param_checksum"], "seeded weights drifted from the fixture"
    eng = Engine(BASE)
    eng.load(params)
    return g, params, eng


def ft_head(batch, **kw):
    d = dict(mode="finetune", col_ids=torch.as_tensor(MARS_VOCAB.analogy_entity_ids(), device="cuda"), label=batch["label"],
             rel_idx=batch["rel_idx"], q_head_idx=batch["q_head_idx"], a_head_idx=batch["a_head_idx"], alpha=0.43, label_smoothing=0.1)
    d.update(kw)
    return d


def _case_batch(d):
    batch = make_finetune_batch(d["B"], d["L"], BASE.image_size, MARS_VOCAB, seed=d["seed_b"], min_len=d["min_len"])
    bchk = checksum([batch["input_ids"], batch["pixel_values"], batch["sep_idx"]])
    assert abs(bchk - d["batch_checksum"]) < 1e-6 * d["batch_checksum"], "seeded batch drifted from the fixture"
    return batch


def test_config1_forward_logits_vs_reference(base):
    """BASELINE.json config 1: logits (full 42,007 columns on the [MASK] rows, through the dense compat head) and trans."""
    g, params, eng = base
    d = g["c1"]
    batch = to_cuda(_case_batch(d))
    z, logits = eng.forward({k: batch[k] for k in MODEL_KEYS}, want_z=True, want_logits=True)
    torch.cuda.synchronize()
    assert eng.status() == 0
    B = d["B"]
    _, mpos = (batch["input_ids"] == 103).nonzero(as_tuple=True)
    ml = logits[torch.arange(B, device="cuda"), mpos].cpu()
    e_z, e_l = rel(z.cpu(), d["trans"]), rel(ml, d["mask_logits_full"])
    ent = torch.as_tensor(MARS_VOCAB.analogy_entity_ids())
    e_abs = (ml[:, ent] - d["mask_logits_full"][:, ent]).abs().max().item()
    print(f"base config 1: rel err trans {e_z:.3e}  [MASK]-row logits (42007 cols) {e_l:.3e}  max abs err on the 2063 entity columns "
          f"{e_abs:.3e} (logit std {d['mask_logits_full'][:, ent].std().item():.3f})")
    assert e_z <= 1e-2 and e_l <= 1e-2
    del logits


@pytest.mark.parametrize("tag", ["c1", "odd"])
def test_finetune_loss_and_every_gradient(base, tag):
    """Loss against the reference training_step; gradient norms against the reference's; every gradient against the oracle run
    in fp32 on the GPU (the oracle itself is pinned to the same fixture on the CPU, tests/test_oracle_golden.py)."""
    from oracle import mkgformer_oracle as O
    from test_model_gpu import _compare_grads
    g, params, eng = base
    d = g[tag]
    batch = to_cuda(_case_batch(d))
    mb = {k: batch[k] for k in MODEL_KEYS}
    eng.zero_grad()
    z, _ = eng.forward(mb, want_z=True)
    loss, ml, ranks = eng.head_loss(mb, ft_head(batch), compute_grad=True, want_logits=True, want_ranks=True)
    eng.backward(mb)
    torch.cuda.synchronize()
    assert eng.status() == 0
    print(f"{tag}: loss cuda {loss.item():.6f} reference {d['loss'].item():.6f}; trans rel err {rel(z.cpu(), d['trans']):.3e}")
    assert abs(loss.item() - d["loss"].item()) <= 1e-2
    assert rel(z.cpu(), d["trans"]) <= 1e-2
    if tag == "odd":
        assert rel(ml.cpu(), d["mask_logits_entities"]) <= 1e-2
    # gradient norms against the reference's own
    worst = 0.0
    for name, ref_norm in d["grad_norm"].items():
        if name.startswith("cls.predictions.decoder"):
            continue
        n = eng.grad_view(name).norm().item()
        if ref_norm is None:
            assert n == 0.0, (name, n)
            continue
        if ref_norm < 1e-7:
            continue            # mathematically zero (softmax shift invariance): rounding noise in the reference
        tol = 0.05
        assert abs(n - ref_norm) <= tol * ref_norm + 1e-7, (name, n, ref_norm)
        worst = max(worst, abs(n - ref_norm) / ref_norm)
    # the adaptive-interaction scalars: full value against the reference, NO absolute floor
    aw = []
    for name, ref in d["grad_full"].items():
        if "adaptive_weight" in name:
            mine = eng.grad_view(name).cpu()
            aw.append((abs(mine.item() - ref.item()) / abs(ref.item()), name, mine.item(), ref.item()))
    aw.sort(reverse=True)
    print(f"{tag}: worst gradient-norm deviation {worst:.3e}; adaptive_weight gradients worst rel err {aw[0][0]:.3e} ({aw[0][1]}: "
          f"{aw[0][2]:.4e} vs {aw[0][3]:.4e})")
    assert aw[0][0] <= 5e-2, aw[:4]
    # every gradient against the fp32 oracle on the same device
    Pg = {k: v.cuda().clone().requires_grad_(True) for k, v in params.items()}
    rl, x, _ = O.finetune_loss(Pg, BASE, batch, MARS_VOCAB.analogy_entity_ids(), alpha=0.43, label_smoothing=0.1)
    rl.backward()
    assert abs(rl.item() - d["loss"].item()) <= 2e-4            # the oracle reproduces the reference loss on this device
    bad = _compare_grads(eng, {k: (None if v.grad is None else v.grad.detach()) for k, v in Pg.items()}, [], tol_main=6e-2,
                         tol_adaptive=5e-2)
    assert not bad, bad[:10]


def test_pretrain_shape_L128(base):
    """The L = 128 pre-train shape (no sep_idx; entity + relation column groups)."""
    g, params, eng = base
    d = g["pretrain"]
    pb = make_pretrain_batch(d["B"], d["L"], BASE.image_size, MARS_VOCAB, seed=d["seed_b"], min_len=d["min_len"])
    assert abs(checksum([pb["input_ids"], pb["pixel_values"]]) - d["batch_checksum"]) < 1e-6 * d["batch_checksum"]
    pb = to_cuda(pb)
    mb = {k: pb[k] for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values")}
    eng.zero_grad()
    z, _ = eng.forward(mb, want_z=True)
    head = dict(mode="pretrain", entity_range=(MARS_VOCAB.entity_id_st, MARS_VOCAB.entity_id_ed),
                relation_range=(MARS_VOCAB.relation_id_st, MARS_VOCAB.relation_id_ed), label=pb["label"], pre_type=pb["pre_type"],
                label_smoothing=0.1)
    loss, ml, _ = eng.head_loss(mb, head, compute_grad=True, want_logits=True)
    eng.backward(mb)
    torch.cuda.synchronize()
    assert eng.status() == 0
    B = d["B"]
    _, mpos = (pb["input_ids"] == 103).nonzero(as_tuple=True)
    e_z = rel(z[torch.arange(B, device="cuda"), mpos].cpu(), d["trans_mask_rows"])
    e_l = rel(ml.cpu(), d["mask_logits_ent_rel"])
    print(f"pre-train L=128: loss cuda {loss.item():.6f} reference {d['loss'].item():.6f}; [MASK]-row trans {e_z:.3e} logits {e_l:.3e}")
    assert abs(loss.item() - d["loss"].item()) <= 1e-2 and e_z <= 1e-2 and e_l <= 1e-2
    for name, ref_norm in d["grad_norm"].items():
        if name.startswith("cls.predictions.decoder"):
            continue
        n = eng.grad_view(name).norm().item()
        if ref_norm is None:
            assert n == 0.0, (name, n)
        elif ref_norm >= 1e-7:
            assert abs(n - ref_norm) <= 0.05 * ref_norm + 1e-7, (name, n, ref_norm)


def test_eval_split_hits1_and_ranks_equal_on_every_row(base):
    """The held-out synthetic evaluation split: 256 samples of the decidable task at base size, the 16
    gold-entity rows / biases fitted on the REFERENCE's features (fixture), ranks from the reference `TransformerLitModel._eval`.
    The CUDA evaluation path (fused [MASK]-row head + rank kernel, 8 batches of 32) must give the same rank ON EVERY ROW —
    no margin filter.  (The fixture records how decidable the split is: min |gold - nearest competitor| = 0.67 logits.)"""
    from mkg_analogy_b200.trainer import MartTrainer
    g, params, _ = base
    s = torch.load(os.path.join(GOLD, "base_eval_split.pt"), weights_only=False)
    assert s["seed_w"] == g["seed_w"]
    P = dict(params)
    P["unimo.text_embeddings.word_embeddings.weight"] = P["unimo.text_embeddings.word_embeddings.weight"].clone()
    P["cls.predictions.bias"] = P["cls.predictions.bias"].clone()
    P["unimo.text_embeddings.word_embeddings.weight"][s["gold_token_ids"]] = s["rows"]
    P["cls.predictions.bias"][s["gold_token_ids"]] = s["bias"]
    tr = MartTrainer(BASE, MARS_VOCAB, total_steps=10, dropout=False)
    tr.engine.load(P)
    label_idx = analogy_classes(MARS_VOCAB, s["n_classes"])[0]
    ranks, ranks_s, err = [], [], 0.0
    for i in range(s["n_eval"] // s["bs"]):
        hb = make_decidable_batch(s["bs"], s["L"], BASE.image_size, MARS_VOCAB, seed=s["eval_seed0"] + i, n_classes=s["n_classes"])
        if i == 0:
            assert abs(checksum([hb["input_ids"]]) - s["batch_checksum"]) < 1e-6 * s["batch_checksum"]
        b = to_cuda(hb)
        r, ml, _ = tr.eval_step(b)
        ranks.append(r.reshape(-1).cpu())
        b2 = dict(b)
        b2["label"] = label_idx[(hb["class"] + s["shift"]) % s["n_classes"]].cuda()
        r2, _, _ = tr.eval_step(b2)
        ranks_s.append(r2.reshape(-1).cpu())
    torch.cuda.synchronize()
    assert tr.engine.status() == 0
    ranks, ranks_s = torch.cat(ranks), torch.cat(ranks_s)
    hits1, hits1_ref = (ranks == 1).float().mean().item(), (s["ranks"] == 1).float().mean().item()
    print(f"eval split: Hits@1 cuda {hits1:.4f} reference {hits1_ref:.4f}; rows with equal rank {(ranks == s['ranks']).sum().item()}/"
          f"{ranks.numel()}; fixture min top-1 margin {s['gold_margin'].min().item():.3f}")
    assert torch.equal(ranks, s["ranks"].to(ranks.dtype))              # every row, no filter
    assert hits1 == hits1_ref
    # rotated labels put the scored entity anywhere in the ranking (2 .. 2063): in the bulk of 2,000 untrained candidates logits
    # are ~0.01 apart, so equality is not decidable in bf16 there; the rank may move by a few places, never by more
    dev = (ranks_s - s["ranks_shifted"].to(ranks_s.dtype)).abs()
    top = s["ranks_shifted"] <= s["n_classes"]
    print(f"rotated labels: reference ranks {s['ranks_shifted'].min().item()}..{s['ranks_shifted'].max().item()}, equal on "
          f"{(dev == 0).float().mean().item():.3f} of rows, max deviation {dev.max().item()} places; among the trained entities "
          f"(rank <= {s['n_classes']}): {int(top.sum())} rows, max deviation {dev[top].max().item() if top.any() else 0}")
    assert dev.max().item() <= 8 and (dev == 0).float().mean().item() >= 0.5
