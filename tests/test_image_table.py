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
