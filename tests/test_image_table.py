"""Entity-image table (collator replacement): host index logic on CPU, device gather on the GPU.
Reference: DataCollatorForSeq2Seq.__call__, MarT/data/data_module.py:127-161."""
import pytest
import torch

from mkg_analogy_b200.data import EntityImageTable
from oracle import mkgformer_oracle as O

ENTITIES = ["Q1", "Q2", "Q3", "Q2", "Q5"]          # a duplicate name: list.index returns the first occurrence
HEADS = ["Q1", "Q2", None, "", "Q5", "Q3", None]
TAILS = ["Q3", None, "Q5", "Q1", "", "Q2", None]   # both / head only / tail only / empty head (falsy, not None) / empty tail / ...


def _features(n=5, s=8, seed=0):
    return torch.randn(n, 3, s, s, generator=torch.Generator().manual_seed(seed))


def test_indices_follow_the_reference_branches():
    feats = _features()
    tab = EntityImageTable(feats, ENTITIES)
    idx = tab.indices(HEADS, TAILS)
    # rebuild pixel_values from the indices on the CPU and compare with the reference collator's own stacking
    ref = O.collate_images(feats, ENTITIES, HEADS, TAILS)
    mine = torch.stack([torch.stack([feats[i] if i >= 0 else torch.zeros_like(feats[0]) for i in row.tolist()]) for row in idx])
    assert torch.equal(mine, ref)
    assert idx.tolist() == [[0, 2], [1, -1], [4, -1], [-1, -1], [4, -1], [2, 1], [-1, -1]]
    assert idx.dtype == torch.int64 and idx.shape == (7, 2)


def test_unknown_entity_raises_like_list_index():
    tab = EntityImageTable(_features(), ENTITIES)
    with pytest.raises(ValueError):
        tab.indices(["Q9"], ["Q1"])
    with pytest.raises(ValueError):
        O.collate_images(_features(), ENTITIES, ["Q9"], ["Q1"])
    with pytest.raises(ValueError):
        EntityImageTable(_features(4), ENTITIES)                 # one row per entity name


def test_gather_has_no_cpu_fallback():
    from mkg_analogy_b200._lib import B200MartError
    tab = EntityImageTable(_features(), ENTITIES)
    with pytest.raises(B200MartError):
        tab.gather(torch.zeros(1, 2, dtype=torch.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_device_gather_is_bit_exact(dtype):
    feats = _features(n=5, s=224, seed=3)
    tab = EntityImageTable(feats, ENTITIES, device="cuda", dtype=dtype)
    px = tab.collate_images(HEADS, TAILS)
    torch.cuda.synchronize()
    ref = O.collate_images(feats.to(dtype).float(), ENTITIES, HEADS, TAILS)      # bf16 storage rounds the table, not the gather
    assert px.shape == (7, 2, 3, 224, 224) and px.dtype == torch.float32
    assert torch.equal(px.cpu(), ref)
    assert int(tab.last_status.item()) == 0
    # out-of-range row: flagged, zeros, no fault
    bad = tab.gather(torch.tensor([[1, 99]], device="cuda"))
    torch.cuda.synchronize()
    assert int(tab.last_status.item()) == 16
    assert torch.equal(bad[0, 0].cpu(), ref[1, 0]) and float(bad[0, 1].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------ whole collator
import os  # noqa: E402

from mkg_analogy_b200.data import MarsCollator  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "collator.pt")


def _check_against(out, ref, keys, feats_img):
    got_keys = [k if k != "pixel_idx" else "pixel_values" for k in out.keys()]
    assert got_keys == keys                                      # same fields, same order
    for k in keys:
        if k == "pixel_values":
            idx = out["pixel_idx"]
            px = torch.stack([torch.stack([feats_img[i] if i >= 0 else torch.zeros_like(feats_img[0]) for i in row.tolist()]) for row in idx])
            assert torch.equal(px, ref[k])
        else:
            assert out[k].dtype == ref[k].dtype and torch.equal(out[k], ref[k]), k


def test_collator_matches_reference_golden():
    """Fixture = outputs of the reference's own DataCollatorForSeq2Seq (oracle/make_golden.py --collator): fine-tune and
    pre-train field sets, ragged lengths, pad_to_multiple_of None / 8, a single-sample batch, every head / tail combination."""
    g = torch.load(GOLD, weights_only=False)
    tab = EntityImageTable(g["image_features"], g["entities"])
    for tag, c in g["cases"].items():
        coll = MarsCollator(tab, pad_token_id=g["pad_token_id"], pad_to_multiple_of=c["pad_to_multiple_of"], images="index")
        feats = [dict(f) for f in c["features"]]
        out = coll(feats)
        _check_against(out, c["out"], c["keys"], g["image_features"])
        assert feats == c["features"]                            # the caller's features are left intact
        if c["pad_to_multiple_of"]:
            assert out["input_ids"].shape[1] % c["pad_to_multiple_of"] == 0


def test_collator_passes_unknown_fields_through_and_pins():
    g = torch.load(GOLD, weights_only=False)
    tab = EntityImageTable(g["image_features"], g["entities"])
    feats = [dict(f, note=f"s{i}") for i, f in enumerate(g["cases"]["finetune"]["features"])]
    out = MarsCollator(tab, images="index")(feats)
    assert out["note"] == [f"s{i}" for i in range(len(feats))] and list(out.keys())[-1] == "note"
    with pytest.raises(ValueError):
        MarsCollator(tab, images="host")


@pytest.mark.reference
def test_collator_matches_reference_live():
    """The reference collator executed here (other seed than the fixture) against MarsCollator on the same features."""
    import copy
    import tempfile
    from oracle import make_golden as MG
    from oracle import reference_loader as RL
    DM = RL.load_collator()
    tok = RL.make_bert_tokenizer(tempfile.mkdtemp())
    ents = [f"Q{i}" for i in range(7)]
    img = torch.randn(7, 3, 4, 4, generator=torch.Generator().manual_seed(9))
    DM.entities, DM.entity2visual_features = ents, img
    tab = EntityImageTable(img, ents)
    for seed, pretrain, mult in ((101, False, None), (102, True, 8), (103, False, 8)):
        feats = MG.collator_features(seed, 7, pretrain)
        ref = DM.DataCollatorForSeq2Seq(tok, model=None, label_pad_token_id=tok.pad_token_id, pad_to_multiple_of=mult,
                                        padding="longest", max_length=128, num_labels=7)(copy.deepcopy(feats))
        out = MarsCollator(tab, pad_token_id=tok.pad_token_id, pad_to_multiple_of=mult, images="index")(feats)
        _check_against(out, ref, list(ref.keys()), img)


@pytest.mark.gpu
def test_collator_device_images_match_golden():
    g = torch.load(GOLD, weights_only=False)
    tab = EntityImageTable(g["image_features"], g["entities"], device="cuda")
    c = g["cases"]["finetune"]
    out = MarsCollator(tab, pad_token_id=g["pad_token_id"], pin_memory=True)(c["features"])
    torch.cuda.synchronize()
    assert list(out.keys()) == c["keys"] and out["input_ids"].is_pinned()
    assert torch.equal(out["pixel_values"].cpu(), c["out"]["pixel_values"])
    for k in c["keys"]:
        if k != "pixel_values":
            assert torch.equal(out[k], c["out"][k])
