"""Whole-model parity at the BENCHMARKED shape: MKGformer-base (12 x 768, 99 vision tokens, 42,007-token vocabulary).

Fixtures come from the reference itself (oracle/make_golden_base.py, run in the build container):
  tests/golden/mkgformer_base.pt   BASELINE.json config 1 (4 synthetic MARS examples, L = 64): [MASK]-row logits over the whole
                                   vocabulary, trans_hidden_states, the reference training_step loss, every gradient norm, the
                                   small gradients in full, `_eval` ranks; an odd length (L = 37); the L = 128 pre-train shape.
  tests/golden/base_eval_split.pt  the held-out decidable evaluation split (256 samples): reference `_eval` ranks.
Tolerances.  The CUDA path multiplies bf16 operands with fp32 accumulation.  Below the fusion depth (layers 0-7) that costs
<= 4e-3 relative on the hidden states (asserted layer by layer, 8e-3).  The four fused layers then pass the text stream through
BertFusion's UNSCALED softmax over 768-dimensional dot products (modeling_unimo.py:400-414), which with N(0, 0.02) random
weights is close to one-hot and amplifies ANY operand rounding: the oracle itself, run with bf16-rounded GEMM operands and
everything else in fp32 (`O.bf16_operands`), moves by 1.5e-2 on trans_hidden_states at this shape.  So at base size with random
weights: trans / logits relative error <= 8e-2 and within 4x of what the bf16-operand oracle shows (both printed), loss |d| <=
3e-2, gradients per tensor <= 0.15 with cosine >= 0.99 (tiny / 9-layer configs keep 6e-2).  Hits@1 / ranks on the evaluation
split: EQUAL ON EVERY ROW.
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
    from oracle import mkgformer_oracle as O
    z, logits = eng.forward({k: batch[k] for k in MODEL_KEYS}, want_z=True, want_logits=True)
    torch.cuda.synchronize()
    assert eng.status() == 0
    B = d["B"]
    _, mpos = (batch["input_ids"] == 103).nonzero(as_tuple=True)
    ml = logits[torch.arange(B, device="cuda"), mpos].cpu()
    del logits
    e_z, e_l = rel(z.cpu(), d["trans"]), rel(ml, d["mask_logits_full"])
    ent = torch.as_tensor(MARS_VOCAB.analogy_entity_ids())
    e_abs = (ml[:, ent] - d["mask_logits_full"][:, ent]).abs().max().item()
    # the oracle on this device: fp32 (must reproduce the reference fixture) and with bf16-rounded GEMM operands
    Pc = {k: v.cuda() for k, v in params.items()}
    hid = {}
    with torch.no_grad():
        _, z32 = O.forward(Pc, BASE, batch, full_logits=False, hidden=hid)
        with O.bf16_operands():
            _, z16 = O.forward(Pc, BASE, batch, full_logits=False)
        with torch.autocast("cuda", dtype=torch.bfloat16):       # what stock PyTorch bf16 training does to the same graph
            _, zac = O.forward(Pc, BASE, batch, full_logits=False)
    assert rel(z32.cpu(), d["trans"]) <= 1e-4
    e_emul, e_ac = rel(z16.cpu(), d["trans"]), rel(zac.float().cpu(), d["trans"])
    print(f"base config 1: rel err trans {e_z:.3e}  [MASK]-row logits (42007 cols) {e_l:.3e}  max abs err on the 2063 entity columns "
          f"{e_abs:.3e} (logit std {d['mask_logits_full'][:, ent].std().item():.3f}); the oracle with bf16-rounded GEMM operands: "
          f"trans {e_emul:.3e}; the oracle under torch.autocast(bf16): trans {e_ac:.3e}")
    # layer by layer (the reference's output_hidden_states): where does the error come from?
    D = BASE.layers
    et = [rel(eng.read_hidden("text", i), hid["text"][i]) for i in range(D + 1)]
    ev = [rel(eng.read_hidden("vision", i), hid["vision"][i]) for i in range(D + 1)]
    print("  hidden-state rel err by layer   text: " + " ".join(f"{e:.1e}" for e in et))
    print("                                vision: " + " ".join(f"{e:.1e}" for e in ev))
    assert max(ev) <= 8e-3, ev                                   # the vision stream never meets the unscaled softmax
    assert max(et[: BASE.fuse_from + 1]) <= 8e-3, et             # neither does the text stream below the fusion depth
    assert e_z <= 8e-2 and e_l <= 8e-2 and e_z <= 4 * e_emul + 5e-3 and e_z <= 2 * e_ac + 5e-3


@pytest.mark.parametrize("tag", ["c1", "odd"])
def test_finetune_loss_and_every_gradient(base, tag):
    """Loss against the reference training_step; every gradient against the oracle run in fp32 on the GPU (the oracle itself is
    pinned to the same fixture on the CPU, tests/test_oracle_golden.py) and gradient norms against the reference's own.
    Error budget: with N(0, 0.02) random weights the near-one-hot BertFusion softmax of the fused layers makes this model
    chaotic in ANY reduced precision - the same oracle under `torch.autocast(bfloat16)` (what stock PyTorch bf16 training does to
    the reference graph) and with bf16-rounded GEMM operands only (`O.bf16_operands`) are run next to the CUDA path; every CUDA
    gradient must be within 6e-2 or 1.5x the autocast oracle's own error (measured: the CUDA path is MORE accurate than
    autocast - median 0.26 against 0.32 at L = 64, 0.37 against 1.28 at L = 37).  Kernel-level gradient parity at these SHAPES with
    the chaos switched off is `test_base_shape_well_conditioned` below (6e-2 per tensor)."""
    from oracle import mkgformer_oracle as O
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

    def oracle_grads(emulate):
        Pg = {k: v.cuda().clone().requires_grad_(True) for k, v in params.items()}
        if emulate == "autocast":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                rl, _, _ = O.finetune_loss(Pg, BASE, batch, MARS_VOCAB.analogy_entity_ids(), alpha=0.43, label_smoothing=0.1)
            rl.float().backward()
        elif emulate:
            with O.bf16_operands():
                rl, _, _ = O.finetune_loss(Pg, BASE, batch, MARS_VOCAB.analogy_entity_ids(), alpha=0.43, label_smoothing=0.1)
                rl.backward()
        else:
            rl, _, _ = O.finetune_loss(Pg, BASE, batch, MARS_VOCAB.analogy_entity_ids(), alpha=0.43, label_smoothing=0.1)
            rl.backward()
        return rl.item(), {k: (None if v.grad is None else v.grad.detach()) for k, v in Pg.items()}

    l32, G32 = oracle_grads(False)
    l16, G16 = oracle_grads(True)
    lac, Gac = oracle_grads("autocast")
    assert abs(l32 - d["loss"].item()) <= 2e-4                  # the fp32 oracle reproduces the reference loss on this device
    print(f"{tag}: loss cuda {loss.item():.6f} reference {d['loss'].item():.6f} bf16-operand oracle {l16:.6f}; trans rel err "
          f"{rel(z.cpu(), d['trans']):.3e}")
    rms = sorted(gr.double().pow(2).mean().sqrt().item() for gr in G32.values() if gr is not None)
    typical = rms[len(rms) // 2]
    rows, bad = [], []
    for name, gr in G32.items():
        mine = eng.grad_view(name)
        if gr is None:
            assert mine.abs().max().item() == 0.0, name
            continue
        if gr.double().pow(2).mean().sqrt().item() < 1e-5 * typical:      # mathematically zero (softmax shift invariance)
            assert mine.double().pow(2).mean().sqrt().item() <= 1e-2 * typical, name
            continue
        e_c, e_16, e_a = rel(mine, gr), rel(G16[name], gr), rel(Gac[name].float(), gr)
        cos = torch.nn.functional.cosine_similarity(mine.flatten().double(), gr.flatten().double(), dim=0).item()
        rows.append((e_c, e_16, e_a, cos, name))
        if "adaptive_weight" not in name and e_c > max(6e-2, 1.5 * e_a):
            bad.append((name, e_c, e_16, e_a, cos))
    rows.sort(reverse=True)
    print(f"{tag}: loss bf16-operand oracle {l16:.6f}, autocast oracle {lac.item() if hasattr(lac, 'item') else lac:.6f}")
    print(f"{tag}: gradient rel err vs the fp32 oracle, worst 16 of {len(rows)} tensors (cuda | bf16-operand oracle | autocast oracle | cosine)")
    for e_c, e_16, e_a, cos, name in rows[:16]:
        print(f"    {e_c:.3e} | {e_16:.3e} | {e_a:.3e} | {cos:.5f}  {name}")
    med = [sorted(r[k] for r in rows)[len(rows) // 2] for k in range(3)]
    print(f"{tag}: median gradient rel err cuda {med[0]:.3e}, bf16-operand oracle {med[1]:.3e}, autocast oracle {med[2]:.3e}")
    assert abs(loss.item() - d["loss"].item()) <= max(3e-2, 2.0 * abs(float(lac) - d["loss"].item()))
    assert rel(z.cpu(), d["trans"]) <= 8e-2
    if tag == "odd":
        assert rel(ml.cpu(), d["mask_logits_entities"]) <= 8e-2
    assert not bad, bad[:10]
    assert med[0] <= max(2e-2, 1.5 * med[2])


def test_base_shape_well_conditioned(base):
    """The benchmarked SHAPES (12 x 768, 99 vision tokens, L = 64, 42,007-token vocabulary, B = 4) with the one source of chaos
    removed: the value projections of the four fused text layers are scaled by 0.1, which takes BertFusion's unscaled softmax
    out of its near-one-hot regime (score std 3.9 -> 0.4).  Forward, loss and EVERY gradient against the fp32 oracle at the
    tolerances of the small configurations: hidden states / logits 1e-2, gradients 6e-2 per tensor (cosine >= 0.995).  The
    adaptive-interaction scalars are cancelling sums of ~10^6 terms dS * S_raw: 6e-2 relative, or the forward-error bound of the
    sum under bf16 operand rounding, 2^-9 * sum|dS * S_raw|, computed from the oracle's own tapped scores (no absolute floor)."""
    from mkg_analogy_b200.engine import Engine
    from oracle import mkgformer_oracle as O
    g, params, _ = base
    P = dict(params)
    for i in range(BASE.fuse_from, BASE.layers):
        for n in ("weight", "bias"):
            k = f"unimo.encoder.text_layer.{i}.attention.self.value.{n}"
            P[k] = P[k] * 0.1
    eng = Engine(BASE)
    eng.load(P)
    batch = to_cuda(make_finetune_batch(4, 64, BASE.image_size, MARS_VOCAB, seed=77, min_len=40))
    mb = {k: batch[k] for k in MODEL_KEYS}
    eng.zero_grad()
    z, _ = eng.forward(mb, want_z=True)
    loss, ml, _ = eng.head_loss(mb, ft_head(batch), compute_grad=True, want_logits=True)
    eng.backward(mb)
    torch.cuda.synchronize()
    assert eng.status() == 0
    Pg = {k: v.cuda().clone().requires_grad_(True) for k, v in P.items()}
    O.SCORE_TAP = []
    try:
        rl, x, zr = O.finetune_loss(Pg, BASE, batch, MARS_VOCAB.analogy_entity_ids(), alpha=0.43, label_smoothing=0.1)
        rl.backward()
        tap = O.SCORE_TAP
    finally:
        O.SCORE_TAP = None
    aw_name = "unimo.encoder.text_layer.{}.attention.self.adaptive_weight.{}"
    cond = O.adaptive_weight_conditioning(tap, batch["sep_idx"][:, 2], [Pg[aw_name.format(i, 0)] for i in range(BASE.layers)],
                                          [Pg[aw_name.format(i, 1)] for i in range(BASE.layers)])
    e_z, e_l = rel(z, zr.detach()), rel(ml, x.detach())
    rms = sorted(v.grad.double().pow(2).mean().sqrt().item() for v in Pg.values() if v.grad is not None)
    typical = rms[len(rms) // 2]
    rows, bad, aw_report = [], [], []
    for name, v in Pg.items():
        mine = eng.grad_view(name)
        if v.grad is None:
            assert mine.abs().max().item() == 0.0, name
            continue
        gr = v.grad
        if gr.double().pow(2).mean().sqrt().item() < 1e-5 * typical:
            assert mine.double().pow(2).mean().sqrt().item() <= 1e-2 * typical, name
            continue
        e = rel(mine, gr)
        cos = torch.nn.functional.cosine_similarity(mine.flatten().double(), gr.flatten().double(), dim=0).item()
        rows.append((e, cos, name))
        if "adaptive_weight" in name:
            # d w = sum over (sample, head, row, col) of dS * S_raw, a cancelling sum: its error bound under operand rounding eps is
            # eps * sum |dS * S_raw| (eps = 2^-9 for bf16); the oracle's own tapped terms give both the sum and the bound
            layer, which = int(name.split(".")[3]), int(name[-1])
            total, abs_sum = cond[layer][which]
            assert abs(total - gr.item()) <= 1e-3 * abs(gr.item()) + 1e-9, (name, total, gr.item())    # the tap reproduces the gradient
            err = abs(mine.item() - gr.item())
            aw_report.append((err / (2.0 ** -9 * abs_sum), abs_sum / abs(gr.item()), e, name))
            if err > max(6e-2 * abs(gr.item()), 2.0 ** -9 * abs_sum):
                bad.append((name, e, err / (2.0 ** -9 * abs_sum)))
        elif e > 6e-2 or cos < 0.995:
            bad.append((name, e, cos))
    rows.sort(reverse=True)
    aw = [(e, n) for e, c, n in rows if "adaptive_weight" in n]
    print(f"well-conditioned base shape: loss cuda {loss.item():.6f} oracle {rl.item():.6f}; rel err trans {e_z:.3e} entity logits {e_l:.3e}; "
          f"gradients: worst {rows[0][0]:.3e} ({rows[0][2]}), median {rows[len(rows) // 2][0]:.3e}; adaptive_weight worst {aw[0][0]:.3e}")
    for e, cos, name in rows[:8]:
        print(f"    {e:.3e} cos {cos:.6f}  {name}")
    aw_report.sort(reverse=True)
    print("  adaptive_weight gradients: |error| / (2^-9 sum|dS S_raw|)   condition sum|terms|/|sum|   rel err")
    for r, c, e, name in aw_report[:6]:
        print(f"    {r:.3f}   {c:9.1f}   {e:.3e}  {name}")
    assert abs(loss.item() - rl.item()) <= 1e-2 and e_z <= 1e-2 and e_l <= 1e-2
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
    assert abs(loss.item() - d["loss"].item()) <= 3e-2 and e_z <= 2e-2 and e_l <= 2e-2    # no sep_idx, yet fusion still runs
    for name, ref_norm in d["grad_norm"].items():
        if name.startswith("cls.predictions.decoder"):
            continue
        n = eng.grad_view(name).norm().item()
        if ref_norm is None:
            assert n == 0.0, (name, n)
        elif ref_norm >= 1e-7:
            assert abs(n - ref_norm) <= 0.10 * ref_norm + 1e-7, (name, n, ref_norm)


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
    # rotated labels put the scored entity anywhere in the ranking (2 .. 2063): in the bulk of 2,000 untrained candidates the
    # logits are ~3e-4 apart, so rank equality is not decidable in bf16 there (reported, not asserted row by row)
    dev = (ranks_s - s["ranks_shifted"].to(ranks_s.dtype)).abs()
    rel_dev = dev.float() / s["ranks_shifted"].float()
    print(f"rotated labels: reference ranks {s['ranks_shifted'].min().item()}..{s['ranks_shifted'].max().item()}, equal on "
          f"{(dev == 0).float().mean().item():.3f} of rows, median deviation {dev.median().item()} places, median relative "
          f"{rel_dev.median().item():.3f}")
    assert dev.median().item() <= 4


def test_compat_route_reference_training_step_at_base_size(base):
    """The exact drop-in route at the benchmarked size: `MKGformerKGC` (dense (B, L, 42007) logits out, dense gradient in) driven
    by the body of the reference `TransformerLitModel.training_step` (lit_models/transformer.py:92-109, restated in torch ops:
    the reference file cannot travel to the GPU box) gives the loss the reference itself recorded in the fixture, and
    `loss.backward()` leaves gradients whose norms follow the reference's."""
    from mkg_analogy_b200.modeling import MKGformerKGC
    from oracle import mkgformer_oracle as O
    g, params, _ = base
    d = g["c1"]
    batch = to_cuda(_case_batch(d))
    model = MKGformerKGC(BASE, device="cuda")
    model.eval()                                            # the fixture is the deterministic (eval-mode) graph
    model.load_state_dict({**model.state_dict(), **params})
    out = model(**{k: batch[k] for k in MODEL_KEYS}, return_dict=True)
    logits, trans = out[0].logits, out[1]
    assert logits.shape == (d["B"], d["L"], BASE.vocab)
    bs = logits.shape[0]
    _, mask_idx = (batch["input_ids"] == 103).nonzero(as_tuple=True)                      # :94
    ids = torch.as_tensor(MARS_VOCAB.analogy_entity_ids(), device="cuda")
    mask_logits = logits[torch.arange(bs), mask_idx][:, ids]                              # :95
    loss = O.label_smooth_ce(mask_logits, batch["label"], 0.1) + 0.43 * O.relaxation_loss(trans, batch["rel_idx"], batch["q_head_idx"],
                                                                                         batch["a_head_idx"])   # :103-109
    print(f"compat route at base size: loss {loss.item():.6f} reference {d['loss'].item():.6f}")
    assert abs(loss.item() - d["loss"].item()) <= 3e-2
    loss.backward()
    named = dict(model.named_parameters())
    assert named["unimo.text_pooler.dense.weight"].grad is None          # never reached by the reference either
    for name in ("unimo.encoder.text_layer.0.attention.self.query.weight", "unimo.encoder.vision_layers.11.mlp.fc2.weight",
                 "cls.predictions.transform.dense.weight", "unimo.text_embeddings.position_embeddings.weight"):
        n, ref = named[name].grad.norm().item(), d["grad_norm"][name]
        assert abs(n - ref) <= 0.5 * ref, (name, n, ref)                 # chaotic random-weight model: see the module docstring
