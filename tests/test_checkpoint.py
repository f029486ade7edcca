"""Checkpoint I/O in the reference's layout (MarT/main.py:133-134, :144-159) — CPU only (the flat parameter buffer lives wherever
the engine does; no kernel is involved)."""
import pytest
import torch

from mkg_analogy_b200 import checkpoint as CK
from mkg_analogy_b200.config import TINY, init_params
from mkg_analogy_b200.engine import Engine


def _engine(seed):
    e = Engine(TINY, torch.device("cpu"))
    e.load(init_params(TINY, seed))
    return e


def test_round_trip_through_a_lightning_style_file(tmp_path):
    a, b = _engine(1), _engine(2)
    path = str(tmp_path / "epoch=3.ckpt")
    CK.save_checkpoint(a, path, epoch=3)
    blob = torch.load(path, weights_only=False)
    assert blob["epoch"] == 3 and all(k.startswith("model.") for k in blob["state_dict"])
    assert "model.cls.predictions.decoder.weight" in blob["state_dict"]          # tied aliases, as the reference saves them
    # the two position_ids buffers too: a strict load into the reference module (main.py:159) needs the full key set
    assert blob["state_dict"]["model.unimo.text_embeddings.position_ids"].shape == (1, a.cfg.max_position)
    assert blob["state_dict"]["model.unimo.vision_embeddings.position_ids"].shape == (1, a.cfg.patches + 1)
    missing, unexpected = CK.load_checkpoint(b, path, strict=True)
    assert missing == [] and unexpected == [] and b.mirror_stale
    for n in a.by_name:
        assert torch.equal(a.view(n), b.view(n)), n


def test_load_follows_nn_module_semantics():
    a, b = _engine(1), _engine(2)
    sd = CK.lightning_state_dict(a)
    keep = b.view("unimo.text_pooler.dense.weight").clone()
    del sd["model.unimo.text_pooler.dense.weight"]
    sd["model.some.new.head.weight"] = torch.zeros(3)
    missing, unexpected = CK.load_state_dict(b, sd, strict=False)                 # main.py:134 passes strict=False
    assert missing == ["unimo.text_pooler.dense.weight"] and unexpected == ["some.new.head.weight"]
    assert torch.equal(b.view("unimo.text_pooler.dense.weight"), keep)            # untouched
    assert torch.equal(b.view("cls.predictions.bias"), a.view("cls.predictions.bias"))
    with pytest.raises(RuntimeError):
        CK.load_state_dict(_engine(3), sd, strict=True)
    bad = CK.lightning_state_dict(a)
    bad["model.cls.predictions.bias"] = torch.zeros(7)
    bad.pop("model.cls.predictions.decoder.bias")
    with pytest.raises(RuntimeError, match="size mismatch"):
        CK.load_state_dict(_engine(3), bad)
    untied = CK.lightning_state_dict(a)
    untied["model.cls.predictions.decoder.weight"] = untied["model.cls.predictions.decoder.weight"] + 1
    with pytest.raises(RuntimeError, match="tied"):
        CK.load_state_dict(_engine(3), untied)
    # keys without the "model." prefix (a bare UnimoForMaskedLM.state_dict()) are accepted too
    bare = {k[len("model."):]: v for k, v in CK.lightning_state_dict(a).items()}
    c = _engine(4)
    assert CK.load_state_dict(c, bare, strict=True) == ([], [])
    assert torch.equal(c.view(CK.WORD), a.view(CK.WORD))


@pytest.mark.reference
def test_checkpoints_move_between_the_reference_and_the_engine(tmp_path):
    """A checkpoint written the way the reference's ModelCheckpoint writes it (lit_model.state_dict(): "model." + the reference
    module's own state_dict, buffers and tied aliases included) loads into the engine; one written by the engine loads into the
    reference module with nothing missing and nothing unexpected."""
    from oracle import reference_loader as RL
    P = init_params(TINY, 5)
    ref = RL.build_reference_model(TINY, P)
    blob = {"state_dict": {"model." + k: v.clone() for k, v in ref.state_dict().items()}}
    path = str(tmp_path / "ref.ckpt")
    torch.save(blob, path)
    e = _engine(6)
    missing, unexpected = CK.load_checkpoint(e, path, strict=True)
    assert missing == [] and unexpected == []
    for n in e.by_name:
        assert torch.equal(e.view(n), P[n].float()), n
    # engine -> reference
    e2 = _engine(7)
    sd = {k[len("model."):]: v for k, v in CK.lightning_state_dict(e2, max_position=TINY.max_position,
                                                                     n_vision_positions=(TINY.image_size // TINY.patch) ** 2 + 1).items()}
    ref2 = RL.build_reference_model(TINY)
    res = ref2.load_state_dict(sd, strict=True)
    assert list(res.missing_keys) == [] and list(res.unexpected_keys) == []
    got = dict(ref2.named_parameters())
    for n in e2.by_name:
        assert torch.equal(got[n].detach(), e2.view(n)), n
