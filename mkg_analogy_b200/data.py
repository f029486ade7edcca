"""Seeded synthetic MARS-shaped batches.

Host-side logic only.  The batch schema is the one the reference collator produces
(MarT/data/data_module.py:48-182) and `TransformerLitModel.training_step` consumes
(MarT/lit_models/transformer.py:59-109): input_ids / attention_mask / token_type_ids (B, L) int64,
pixel_values (B, 2, 3, S, S) fp32, sep_idx (B, 6), label, rel_idx (B, 2), q_head_idx, a_head_idx, pre_type.
"""
from dataclasses import dataclass

import torch

CLS, SEP, MASK, PAD = 101, 102, 103, 0


@dataclass(frozen=True)
class VocabLayout:
    """Token-id layout after the reference adds its special tokens (data_module.py:195-206, transformer.py:41-54):
    [0, word_vocab) word pieces, then entities, then relations, then "[R]" as the last id."""
    word_vocab: int = 30522
    n_entities: int = 11292
    n_relations: int = 192
    n_analogy_entities: int = 2063
    n_analogy_relations: int = 27
    filler_lo: int = 1000
    filler_hi: int = 30000

    @property
    def entity_id_st(self):
        return self.word_vocab

    @property
    def entity_id_ed(self):
        return self.word_vocab + self.n_entities

    @property
    def relation_id_st(self):
        return self.entity_id_ed

    @property
    def relation_id_ed(self):
        return self.entity_id_ed + self.n_relations

    @property
    def r_token(self):
        return self.relation_id_ed

    @property
    def vocab(self):
        return self.relation_id_ed + 1

    def analogy_entity_ids(self, seed=7):
        g = torch.Generator().manual_seed(seed)
        perm = torch.randperm(self.n_entities, generator=g)[: self.n_analogy_entities]
        return (perm.sort().values + self.entity_id_st).tolist()

    def analogy_relation_ids(self, seed=7):
        g = torch.Generator().manual_seed(seed + 1)
        perm = torch.randperm(self.n_relations, generator=g)[: self.n_analogy_relations]
        return (perm.sort().values + self.relation_id_st).tolist()


MARS_VOCAB = VocabLayout()
TINY_VOCAB = VocabLayout(word_vocab=500, n_entities=150, n_relations=40, n_analogy_entities=60,
                         n_analogy_relations=9, filler_lo=110, filler_hi=500)


def _composition(total, parts, g):
    """Random non-negative integers summing to `total`."""
    if total <= 0:
        return [0] * parts
    cuts = torch.randint(0, total + 1, (parts - 1,), generator=g).sort().values.tolist()
    edges = [0] + cuts + [total]
    return [edges[i + 1] - edges[i] for i in range(parts)]


def make_region_features(B, regions_per_entity=36, dim=2048, seed=77, single_every=3):
    """Synthetic stand-in for the Faster-RCNN region features VisualBERT / ViLBERT consume (the collator stacks two entities of
    36 x 2048 features each, MarT/data/data_module.py:129-157): non-negative (post-ReLU) features; every `single_every`-th sample has
    only ONE entity — its second block is zeros with visual_attention_mask 0, as the collator builds it (:144-152).
    Returns (pixel_values (B, 72, dim) fp32, visual_attention_mask (B, 72) fp32)."""
    g = torch.Generator().manual_seed(seed)
    f = torch.relu(torch.randn(B, 2 * regions_per_entity, dim, generator=g))
    m = torch.ones(B, 2 * regions_per_entity)
    for b in range(B):
        if single_every and b % single_every == single_every - 1:
            f[b, regions_per_entity:] = 0
            m[b, regions_per_entity:] = 0
    return f, m


def make_finetune_batch(B, L, image_size=224, vocab: VocabLayout = MARS_VOCAB, seed=1234, pad=True, min_len=24):
    """One fine-tune batch.  Prompt structure per sample (positions):
    [CLS] e_h ... [SEP]0 ... [R] ... [SEP]1 e_t ... [SEP]2 e_q ... [SEP]3 ... [R] ... [SEP]4 ... [MASK] ... [SEP]5 [PAD]*
    s = sep_idx[:, 2] is the split used by the adaptive attention scaling (modeling_unimo.py:342-349)."""
    g = torch.Generator().manual_seed(seed)
    assert L >= 14
    min_len = max(14, min(min_len, L))
    ids = torch.zeros(B, L, dtype=torch.long)
    att = torch.zeros(B, L, dtype=torch.long)
    tt = torch.zeros(B, L, dtype=torch.long)
    sep_idx = torch.zeros(B, 6, dtype=torch.long)
    rel_idx = torch.zeros(B, 2, dtype=torch.long)
    q_head = torch.ones(B, dtype=torch.long)
    a_head = torch.zeros(B, dtype=torch.long)
    for b in range(B):
        n = int(torch.randint(min_len, L + 1, (1,), generator=g)) if pad else L
        s_hi = min(40, n - 8)
        s_lo = min(max(6, 8), s_hi)
        s = int(torch.randint(s_lo, s_hi + 1, (1,), generator=g))
        a, bb, c = _composition(s - 6, 3, g)
        d, e, f = _composition(n - 7 - s, 3, g)
        row = torch.randint(vocab.filler_lo, vocab.filler_hi, (n,), generator=g)
        ent = torch.randint(vocab.entity_id_st, vocab.entity_id_ed, (3,), generator=g)
        row[0] = CLS
        row[1] = ent[0]
        sep0 = 2 + a
        sep1 = sep0 + 2 + bb
        sep2 = sep1 + 2 + c
        sep3 = sep2 + 2 + d
        sep4 = sep3 + 2 + e
        sep5 = sep4 + 2 + f
        assert sep2 == s and sep5 == n - 1, (sep2, s, sep5, n)
        for k, p in enumerate((sep0, sep1, sep2, sep3, sep4, sep5)):
            row[p] = SEP
            sep_idx[b, k] = p
        r0 = sep0 + 1 + int(torch.randint(0, bb + 1, (1,), generator=g))
        r1 = sep3 + 1 + int(torch.randint(0, e + 1, (1,), generator=g))
        row[r0] = vocab.r_token
        row[r1] = vocab.r_token
        rel_idx[b, 0], rel_idx[b, 1] = r0, r1
        row[sep1 + 1] = ent[1]
        row[s + 1] = ent[2]
        a_head[b] = s + 1
        mpos = sep4 + 1 + int(torch.randint(0, f + 1, (1,), generator=g))
        row[mpos] = MASK
        ids[b, :n] = row
        att[b, :n] = 1
        tt[b, s + 1:n] = 1
    px = torch.randn(B, 2, 3, image_size, image_size, generator=g)
    mode = torch.rand(B, generator=g)
    px[mode < 0.4, 1] = 0.0   # "[PAD] image" = zeros (data_module.py:134-142); MARS mode-0 share is 40 %
    label = torch.randint(0, vocab.n_analogy_entities, (B,), generator=g)
    return {
        "input_ids": ids, "attention_mask": att, "token_type_ids": tt, "pixel_values": px, "sep_idx": sep_idx,
        "label": label, "rel_idx": rel_idx, "q_head_idx": q_head, "a_head_idx": a_head,
        "rel_label": torch.randint(0, vocab.n_analogy_relations, (B,), generator=g),
        "pre_type": torch.zeros(B, dtype=torch.long),
    }


def analogy_classes(vocab: VocabLayout = MARS_VOCAB, n_classes=16, tokens_per_class=4, seed=3):
    """Class table of the decidable synthetic split: `n_classes` gold entities (as indices into
    `vocab.analogy_entity_ids()`, i.e. label values) and, per class, a private set of filler word-pieces and three private
    context entities (never gold entities)."""
    g = torch.Generator().manual_seed(seed)
    label_idx = torch.randperm(vocab.n_analogy_entities, generator=g)[:n_classes].sort().values
    toks = vocab.filler_lo + torch.randperm(vocab.filler_hi - vocab.filler_lo, generator=g)[: n_classes * tokens_per_class]
    gold = set((torch.as_tensor(vocab.analogy_entity_ids())[label_idx]).tolist())
    ctx = [e for e in (vocab.entity_id_st + torch.randperm(vocab.n_entities, generator=g)).tolist() if e not in gold][: 3 * n_classes]
    return label_idx, toks.view(n_classes, tokens_per_class), torch.as_tensor(ctx).view(n_classes, 3)


def make_decidable_batch(B, L, image_size=224, vocab: VocabLayout = MARS_VOCAB, seed=1234, n_classes=16, class_seed=3, min_len=24):
    """A fine-tune batch whose label can be read off its inputs (the "learnable synthetic task" of SURVEY.md §8 c5, used to
    build an evaluation split with real top-1 margins): the prompt structure (lengths, separator / [MASK] / [R] positions,
    padding, zero second images) varies per sample exactly as in `make_finetune_batch`, but a sample's filler word-pieces come
    from the private 4-token set of its class (cyclically), its three context entities are the class's, its images are the
    class's seeded image plus N(0, 0.1) noise, and label = the class's gold entity.  Gold entities never occur as input tokens, so their
    embedding rows only act through the tied decoder."""
    batch = make_finetune_batch(B, L, image_size, vocab, seed=seed, min_len=min_len)
    label_idx, toks, ctx = analogy_classes(vocab, n_classes, seed=class_seed)
    g = torch.Generator().manual_seed(seed + 7919)
    cls = torch.randint(0, n_classes, (B,), generator=g)
    ids, px = batch["input_ids"], batch["pixel_values"]
    for b in range(B):
        row, c = ids[b], int(cls[b])
        fill = (row >= vocab.filler_lo) & (row < vocab.filler_hi)
        row[fill] = toks[c][torch.arange(int(fill.sum())) % toks.shape[1]]
        ent = ((row >= vocab.entity_id_st) & (row < vocab.entity_id_ed)).nonzero().flatten()
        row[ent] = ctx[c][: ent.numel()]
        base = torch.randn(2, 3, image_size, image_size, generator=torch.Generator().manual_seed(class_seed * 100003 + c))
        zero2 = bool((px[b, 1] == 0).all())
        px[b] = base + 0.1 * px[b]
        if zero2:
            px[b, 1] = 0.0
    batch["label"] = label_idx[cls].clone()
    batch["class"] = cls
    return batch


def make_pretrain_batch(B, L, image_size=224, vocab: VocabLayout = MARS_VOCAB, seed=4321, min_len=12):
    """Pre-train batch (transformer.py:72-90): no sep_idx; pre_type 1 = entity prediction over
    [entity_id_st, entity_id_ed), 2 = relation prediction over [relation_id_st, relation_id_ed)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros(B, L, dtype=torch.long)
    att = torch.zeros(B, L, dtype=torch.long)
    tt = torch.zeros(B, L, dtype=torch.long)
    pre_type = torch.randint(1, 3, (B,), generator=g)
    label = torch.zeros(B, dtype=torch.long)
    for b in range(B):
        n = int(torch.randint(min(min_len, L), L + 1, (1,), generator=g))
        row = torch.randint(vocab.filler_lo, vocab.filler_hi, (n,), generator=g)
        row[0] = CLS
        row[1] = int(torch.randint(vocab.entity_id_st, vocab.entity_id_ed, (1,), generator=g))
        row[n - 1] = SEP
        mpos = int(torch.randint(2, n - 1, (1,), generator=g))
        row[mpos] = MASK
        ids[b, :n] = row
        att[b, :n] = 1
        if pre_type[b] == 2:
            label[b] = int(torch.randint(0, vocab.n_relations, (1,), generator=g))
        else:
            label[b] = int(torch.randint(0, vocab.n_entities, (1,), generator=g))
    px = torch.randn(B, 2, 3, image_size, image_size, generator=g)
    px[torch.rand(B, generator=g) < 0.4, 1] = 0.0
    return {"input_ids": ids, "attention_mask": att, "token_type_ids": tt, "pixel_values": px,
            "label": label, "pre_type": pre_type}


class EntityImageTable:
    """GPU-resident replacement of the image half of the reference collator.

    Reference (MarT/data/data_module.py:127-160, MKGformer branch): for every sample the collator looks the head / tail
    entity up with `entities.index(name)` (a linear scan of an ~11 k-element list, twice per sample), stacks the two
    (3, S, S) fp32 feature tensors on the host — an all-zero "[PAD] image" for a missing entity — and the batch's
    (B, 2, 3, S, S) pixel_values (38.5 MB at B = 32, S = 224) cross PCIe every step.  Here the table is uploaded once
    (fp32, or bf16 to halve its 6.8 GB), names resolve through a dict, and a batch's pixel_values are gathered on the
    device by `b200mart_gather_images` from a (B, 2) index tensor (-1 = zero image): 512 bytes per step instead of 38.5 MB.
    """

    def __init__(self, features, entities, device=None, dtype=torch.float32):
        if features.dim() != 4 or features.shape[0] != len(entities):
            raise ValueError("features must be (n_entities, C, S, S) with one row per entity name")
        if dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("table dtype must be float32 or bfloat16")
        self.image_shape = tuple(features.shape[1:])
        self.n_entities = features.shape[0]
        # first occurrence wins, like list.index (data_module.py:131-132)
        self.index = {}
        for i, name in enumerate(entities):
            self.index.setdefault(name, i)
        self.device = torch.device(device) if device is not None else None
        self.table = features.to(device=self.device, dtype=dtype).contiguous() if self.device is not None else features.to(dtype).contiguous()

    def indices(self, head_ent, tail_ent):
        """(B, 2) int64 entity rows of the two images of every sample, -1 = all-zero image; the branch structure of
        data_module.py:127-154: both entities given -> (head, tail); otherwise the one that is not None (head first),
        or a zero image when that one is empty, followed by a zero image."""
        rows = []
        for head, tail in zip(head_ent, tail_ent):
            if head and tail:
                rows.append((self._row(head), self._row(tail)))
            else:
                entity = head if head is not None else tail
                rows.append((self._row(entity) if entity else -1, -1))
        return torch.tensor(rows, dtype=torch.int64).reshape(-1, 2)

    def _row(self, name):
        try:
            return self.index[name]
        except KeyError:
            raise ValueError(f"{name!r} is not in list") from None      # what entities.index(name) raises

    def gather(self, idx, out=None):
        """pixel_values (B, 2, C, S, S) fp32 on the table's device for a (B, 2) int64 index tensor on that device."""
        from . import _lib
        if self.device is None or self.device.type != "cuda":
            raise _lib.B200MartError("EntityImageTable.gather needs the table on a CUDA device (no CPU fallback)")
        idx = idx.to(device=self.device, dtype=torch.int64).contiguous()
        n = idx.numel()
        if out is None:
            out = torch.empty((idx.shape[0], 2) + self.image_shape, device=self.device, dtype=torch.float32)
        elems = 1
        for d in self.image_shape:
            elems *= d
        status = torch.zeros(1, device=self.device, dtype=torch.int32)
        _lib.check(_lib.lib().b200mart_gather_images(self.table.data_ptr(), int(self.table.dtype == torch.bfloat16), idx.data_ptr(),
                                                     out.data_ptr(), n, elems, self.n_entities, status.data_ptr(), _lib.cur_stream()))
        self.last_status = status
        return out

    def collate_images(self, head_ent, tail_ent):
        """pixel_values for a batch of (head, tail) entity names: host index lookup + device gather."""
        return self.gather(self.indices(head_ent, tail_ent).to(self.device, non_blocking=True))


class MarsCollator:
    """Batch assembly of the reference collator (`DataCollatorForSeq2Seq.__call__`, MarT/data/data_module.py:91-182) with the
    image half served by an `EntityImageTable`.

    Same output contract: the text fields right-padded to the longest sequence of the batch (`tokenizer.pad(padding=
    "longest")`, :117-123), rounded up to `pad_to_multiple_of` (8 under `--precision 16`, :215), int64 tensors; `label`,
    `pre_type`, `rel_idx`, `sep_idx`, `rel_label`, `q_head_idx`, `a_head_idx` as int64 tensors when present (:162-178);
    any other per-sample field passed through as a list (:107-110, :180); keys in the reference's order.  Differences, on
    purpose: the caller's feature dicts are not mutated (the reference pops from them); tensors can be allocated in pinned
    memory (`pin_memory=True`; the reference's DataLoader uses pin_memory=False, :268) so that the host -> device copy of the
    few KB of ids is asynchronous; and `images="index"` returns `pixel_idx` (B, 2) int64 rows of the table (-1 = zero image)
    instead of `pixel_values`, for `table.gather(batch.pop("pixel_idx"))` on the training device."""

    TEXT_KEYS = ("input_ids", "attention_mask", "token_type_ids")
    TENSOR_KEYS = ("pre_type", "rel_idx", "sep_idx", "rel_label", "q_head_idx", "a_head_idx")      # order of :164-178
    POPPED = ("label", "rel_label", "head_ent", "tail_ent", "pre_type", "rel_idx", "sep_idx", "q_head_idx", "a_head_idx")

    def __init__(self, image_table, pad_token_id=0, pad_token_type_id=0, pad_to_multiple_of=None, pin_memory=False, images="device"):
        if images not in ("device", "index"):
            raise ValueError("images must be 'device' or 'index'")
        self.table, self.pad_token_id, self.pad_token_type_id = image_table, pad_token_id, pad_token_type_id
        self.pad_to_multiple_of, self.pin_memory, self.images = pad_to_multiple_of, pin_memory, images

    def _tensor(self, rows):
        t = torch.tensor(rows, dtype=torch.int64)
        return t.pin_memory() if self.pin_memory else t

    def __call__(self, features):
        first = features[0]
        col = {k: [f[k] for f in features] for k in self.POPPED if k in first}
        extra = {k: [f[k] for f in features] for k in first if k not in self.TEXT_KEYS and k not in self.POPPED}
        # ---- text: right-pad to the longest, rounded up to the multiple
        longest = max(len(f["input_ids"]) for f in features)
        m = self.pad_to_multiple_of
        if m is not None and longest % m != 0:
            longest = (longest // m + 1) * m
        out = {}
        text_keys = [k for k in first if k in self.TEXT_KEYS]
        if "attention_mask" not in text_keys:
            text_keys.insert(1, "attention_mask")               # tokenizer.pad creates it for BERT-family tokenizers
        fill = {"input_ids": self.pad_token_id, "attention_mask": 0, "token_type_ids": self.pad_token_type_id}
        for k in text_keys:
            rows = []
            for f in features:
                seq = list(f[k]) if k in f else [1] * len(f["input_ids"])
                rows.append(seq + [fill[k]] * (longest - len(seq)))
            out[k] = self._tensor(rows)
        # ---- images
        idx = self.table.indices(col.get("head_ent", [None] * len(features)), col.get("tail_ent", [None] * len(features)))
        if self.images == "index":
            out["pixel_idx"] = idx.pin_memory() if self.pin_memory else idx
        else:
            out["pixel_values"] = self.table.gather(idx.to(self.table.device, non_blocking=True))
        # ---- index / label tensors, then pass-through lists
        out["label"] = self._tensor(col["label"])
        for k in self.TENSOR_KEYS:
            if col.get(k):
                out[k] = self._tensor(col[k])
        out.update(extra)
        return out
