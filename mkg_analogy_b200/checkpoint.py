"""Checkpoint I/O in the reference's layout for the stand-alone trainer (first half).

The reference saves with `pl.callbacks.ModelCheckpoint(save_weights_only=True)` (MarT/main.py:144-149) and restores with
`lit_model.load_state_dict(torch.load(path)["state_dict"], strict=False)` (:133-134, :159): a file holding
`{"state_dict": {"model." + <UnimoForMaskedLM state_dict name>: tensor}}`.  Besides the 379+ parameters that state_dict carries
two `position_ids` buffers and the two aliases of the tied decoder (`cls.predictions.decoder.weight` = the word-embedding table,
`cls.predictions.decoder.bias` = `cls.predictions.bias`, modeling_unimo.py:904-926).  The drop-in `MKGformerKGC` module goes
through torch's own state_dict machinery; these helpers do the same for a bare `Engine` (flat parameter buffer), in both
directions, so that checkpoints move freely between the reference, the drop-in module and `MartTrainer`."""
import torch

PREFIX = "model."
WORD = "unimo.text_embeddings.word_embeddings.weight"
DEC_W, DEC_B, CLS_B = "cls.predictions.decoder.weight", "cls.predictions.decoder.bias", "cls.predictions.bias"
BUFFERS = ("unimo.vision_embeddings.position_ids", "unimo.text_embeddings.position_ids")


def lightning_state_dict(engine, max_position=None, n_vision_positions=None):
    """{"model." + name: CPU tensor} for every parameter of `engine`, plus the tied-decoder aliases and (when their lengths are
    given) the two `position_ids` buffers, i.e. the keys a reference checkpoint has."""
    sd = {PREFIX + n: t.detach().to("cpu", copy=True) for n, t in engine.state().items()}
    if PREFIX + WORD in sd:
        sd[PREFIX + DEC_W] = sd[PREFIX + WORD]
    if PREFIX + CLS_B in sd:
        sd[PREFIX + DEC_B] = sd[PREFIX + CLS_B]
    if n_vision_positions is not None:
        sd[PREFIX + BUFFERS[0]] = torch.arange(n_vision_positions).expand((1, -1)).clone()
    if max_position is not None:
        sd[PREFIX + BUFFERS[1]] = torch.arange(max_position).expand((1, -1)).clone()
    return sd


def save_checkpoint(engine, path, **extra):
    """Write a file `main.py:134` can restore (`extra` keys, e.g. epoch / global_step, are stored next to "state_dict")."""
    blob = dict(extra)
    # the full reference key set, position_ids buffers included, so that a strict load (main.py:159) succeeds
    blob["state_dict"] = lightning_state_dict(engine, max_position=engine.cfg.max_position, n_vision_positions=engine.cfg.patches + 1)
    torch.save(blob, path)


def load_state_dict(engine, state_dict, strict=False):
    """Copy a reference-layout state_dict (keys with or without the "model." prefix) into `engine`.  Follows
    `nn.Module.load_state_dict`: returns (missing, unexpected); shape mismatches always raise; with strict=True so do missing
    / unexpected keys.  The decoder aliases must agree with the tensors they alias (a checkpoint whose tie was broken is
    rejected instead of silently picking one side); `position_ids` buffers are ignored."""
    sd = {(k[len(PREFIX):] if k.startswith(PREFIX) else k): v for k, v in state_dict.items()}
    names = list(engine.by_name)
    missing = [n for n in names if n not in sd]
    known_extra = set(BUFFERS) | {DEC_W, DEC_B}
    unexpected = [k for k in sd if k not in engine.by_name and k not in known_extra]
    errors = []
    for n in names:
        if n in sd and tuple(sd[n].shape) != tuple(engine.by_name[n].shape):
            errors.append(f"size mismatch for {n}: checkpoint {tuple(sd[n].shape)} vs model {tuple(engine.by_name[n].shape)}")
    for alias, src in ((DEC_W, WORD), (DEC_B, CLS_B)):
        if alias in sd and src in sd and (sd[alias].shape != sd[src].shape or not torch.equal(sd[alias], sd[src])):
            errors.append(f"{alias} differs from {src}: the decoder is tied to it (modeling_unimo.py:904-926)")
    if strict and (missing or unexpected):
        errors.append(f"missing keys {missing[:5]}{'...' if len(missing) > 5 else ''}, unexpected keys {unexpected[:5]}")
    if errors:
        raise RuntimeError("Error(s) in loading state_dict for b200mart Engine:\n\t" + "\n\t".join(errors))
    with torch.no_grad():
        for n in names:
            if n in sd:
                engine.view(n).copy_(sd[n].to(engine.device, torch.float32))
        # a checkpoint that carries only the alias (e.g. exported from a head-only fine-tune) still defines the tied tensor
        for alias, src in ((DEC_W, WORD), (DEC_B, CLS_B)):
            if alias in sd and src not in sd and src in engine.by_name:
                engine.view(src).copy_(sd[alias].to(engine.device, torch.float32))
                missing.remove(src)
    engine.mirror_stale = True            # the bf16 mirror is refreshed before the next forward
    return missing, unexpected


def load_checkpoint(engine, path, strict=False, map_location="cpu"):
    """`lit_model.load_state_dict(torch.load(path)["state_dict"], strict=False)` (main.py:134) for a bare engine."""
    blob = torch.load(path, map_location=map_location, weights_only=False)
    return load_state_dict(engine, blob["state_dict"] if "state_dict" in blob else blob, strict=strict)
