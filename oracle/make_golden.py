"""ORACLE TOOLING — generate tests/golden/*.pt by executing the REFERENCE (build container only).

    python oracle/make_golden.py

Runs /root/reference/MarT code (modeling_unimo.py via oracle/reference_loader.py, the reference
TransformerLitModel.training_step / _eval, and LabelSmoothSoftmaxCEV1) on seeded synthetic inputs and stores
the outputs.  Inputs and weights are NOT stored: they are regenerated from the same seeds
(mkg_analogy_b200.config.init_params, mkg_analogy_b200.data.make_*_batch); a checksum of both is stored so
that RNG drift is detected instead of silently comparing different problems.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mkg_analogy_b200.config import TINY, init_params          # noqa: E402
from mkg_analogy_b200.data import TINY_VOCAB, make_finetune_batch, make_pretrain_batch  # noqa: E402
from oracle import reference_loader as RL                      # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
GRAD_FULL = (  # small tensors stored in full; every other parameter contributes its grad norm
    "adaptive_weight", "LayerNorm", "layer_norm", "layrnorm", ".bias", "class_embedding", "position_embedding",
    "token_type_embeddings",
)


def checksum(tensors):
    return float(sum(t.double().abs().sum().item() for t in tensors))


def tiny_case(seed_w=11, seed_b=21, B=4, L=24):
    cfg, voc = TINY, TINY_VOCAB
    assert cfg.vocab == voc.vocab
    params = init_params(cfg, seed_w)
    model = RL.build_reference_model(cfg, params)
    out = {"cfg": cfg.to_dict(), "seed_w": seed_w, "seed_b": seed_b, "B": B, "L": L,
           "param_checksum": checksum(params.values())}

    # ---- fine-tune: reference LitModel.training_step (loss) + backward, eval mode (dropout off)
    batch = make_finetune_batch(B, L, cfg.image_size, voc, seed=seed_b, min_len=16)
    out["batch_checksum"] = checksum([batch["input_ids"], batch["pixel_values"], batch["sep_idx"]])
    lit = RL.build_reference_lit(model, voc, alpha=0.43, label_smoothing=0.1, pretrain=0)
    lit.eval()
    with torch.no_grad():
        mo = model(**{k: batch[k] for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "sep_idx")},
                   return_dict=True)
        logits, trans = mo[0].logits, mo[1]
    _, mpos = (batch["input_ids"] == 103).nonzero(as_tuple=True)
    out["ft_mask_logits_full"] = logits[torch.arange(B), mpos].clone()          # (B, V)
    out["ft_logits_row0"] = logits[0].clone()                                   # (L, V) full-logits check
    out["ft_trans"] = trans.clone()
    model.zero_grad()
    loss = lit.training_step({k: v.clone() for k, v in batch.items()}, 1)
    loss.backward()
    out["ft_loss"] = loss.detach().clone()
    gn, gf = {}, {}
    for name, p in model.named_parameters():
        if p.grad is None:
            gn[name] = None
            continue
        gn[name] = p.grad.norm().item()
        if any(t in name for t in GRAD_FULL) and p.numel() <= 4096:
            gf[name] = p.grad.clone()
    gf["unimo.text_embeddings.word_embeddings.weight"] = model.unimo.text_embeddings.word_embeddings.weight.grad.clone()
    gf["unimo.encoder.text_layer.9.attention.self.query.weight"] = \
        model.unimo.encoder.text_layer[9].attention.self.query.weight.grad.clone()
    gf["unimo.encoder.vision_layers.8.self_attn.k_proj.weight"] = \
        model.unimo.encoder.vision_layers[8].self_attn.k_proj.weight.grad.clone()
    gf["unimo.vision_embeddings.patch_embedding.weight"] = \
        model.unimo.vision_embeddings.patch_embedding.weight.grad.clone()[:8]
    out["ft_grad_norm"], out["ft_grad_full"] = gn, gf
    with torch.no_grad():
        out["ft_ranks"] = torch.as_tensor(lit._eval({k: v.clone() for k, v in batch.items()}, 1)["entity_ranks"])

    # ---- pre-train branch (no sep_idx)
    pb = make_pretrain_batch(B, L, cfg.image_size, voc, seed=seed_b + 1)
    lit_p = RL.build_reference_lit(model, voc, label_smoothing=0.1, pretrain=1)
    lit_p.eval()
    model.zero_grad()
    ploss = lit_p.training_step({k: v.clone() for k, v in pb.items()}, 1)
    ploss.backward()
    out["pt_loss"] = ploss.detach().clone()
    out["pt_grad_norm"] = {n: (None if p.grad is None else p.grad.norm().item()) for n, p in model.named_parameters()}
    with torch.no_grad():
        r = lit_p._eval({k: v.clone() for k, v in pb.items()}, 1)
        out["pt_ranks"] = {k: torch.as_tensor(v) for k, v in r.items()}
    return out


def loss_case():
    """LabelSmoothSoftmaxCEV1 (lit_models/utils.py:42-66) known answers incl. ignore_index and its gradient."""
    U = RL.load_module("ref_lit_utils", "lit_models/utils.py")
    g = torch.Generator().manual_seed(5)
    out = {}
    for C in (7, 192, 2063):
        x = (torch.randn(6, C, generator=g) * 3).requires_grad_(True)
        y = torch.randint(0, C, (6,), generator=g)
        y2 = y.clone()
        y2[2] = -100
        for tag, lab in (("plain", y), ("ignore", y2)):
            loss = U.LabelSmoothSoftmaxCEV1(lb_smooth=0.1)(x, lab)
            (gx,) = torch.autograd.grad(loss, x)
            out[f"C{C}_{tag}"] = dict(x=x.detach().clone(), y=lab.clone(), loss=loss.detach().clone(), grad=gx.clone())
    return out


def stack_case():
    """Reference FlavaEncoder (text-encoder use: mask + adaptive scaling on rows 1:s) and ViltEncoder (rows img_length+1:s')
    on seeded inputs and weights: outputs, input gradient and a few parameter gradients."""
    from transformers import FlavaTextConfig, ViltConfig
    Fm, Vm = RL.load_flava(), RL.load_vilt()
    out = {}
    for tag, mod, cfg_cls, T, row_lo in (("flava", Fm.FlavaEncoder, FlavaTextConfig, 40, 1), ("vilt", Vm.ViltEncoder, ViltConfig, 56, 17)):
        cfg = cfg_cls(hidden_size=128, num_hidden_layers=3, num_attention_heads=2, intermediate_size=256,
                      hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        torch.manual_seed(31)
        enc = mod(cfg).eval()
        with torch.no_grad():
            for n, prm in enc.named_parameters():
                if prm.dim() >= 2:
                    prm.normal_(0, 0.05)
                elif "adaptive_weight.0" in n:
                    prm.fill_(0.3)
                elif "adaptive_weight.1" in n:
                    prm.fill_(0.5)
                elif "layernorm" in n and n.endswith("weight"):
                    prm.normal_(1.0, 0.1)
                else:
                    prm.normal_(0, 0.05)
        g = torch.Generator().manual_seed(32)
        B = 3
        x = torch.randn(B, T, 128, generator=g).requires_grad_(True)
        lens = torch.tensor([T, T - 7, T - 13])
        mask = (torch.arange(T)[None] < lens[:, None]).long()
        sep = torch.zeros(B, 6, dtype=torch.long)
        sep[:, 2] = torch.tensor([row_lo + 6, row_lo + 11, row_lo + 4])
        ext = (1.0 - mask[:, None, None, :].float()) * -10000.0
        kw = dict(attention_mask=ext, sep_idx=sep.clone(), return_dict=True)
        if tag == "vilt":
            kw["img_length"] = row_lo - 1
        y = enc(x, **kw).last_hidden_state
        gy = torch.randn(y.shape, generator=g)
        y.backward(gy)
        names = ["layer.0.attention.attention.query.weight", "layer.2.output.dense.bias", "layer.1.attention.attention.adaptive_weight.0",
                 "layer.1.attention.attention.adaptive_weight.1", "layer.2.layernorm_before.weight", "layer.0.intermediate.dense.weight"]
        sd = {k: v.detach().clone() for k, v in enc.state_dict().items()}
        out[tag] = dict(params=sd, x=x.detach().clone(), mask=mask, sep=sep, row_lo=row_lo, y=y.detach().clone(), gy=gy,
                        dx=x.grad.clone(), grads={n: dict(enc.named_parameters())[n].grad.clone() for n in names},
                        grad_norm={n: p.grad.norm().item() for n, p in enc.named_parameters()})
    return out


def vilt_case():
    """Reference ViltForMaskedLM (modeling_vilt.py:896-968) on a tiny config: logits / trans_hidden of the text rows (the
    only rows MarT reads; they do not depend on the reference's random patch permutation) and a few gradients."""
    from transformers import ViltConfig
    Vm = RL.load_vilt()
    cfg = ViltConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, vocab_size=300,
                     image_size=64, patch_size=32, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    torch.manual_seed(41)
    ref = Vm.ViltForMaskedLM(cfg).eval()
    with torch.no_grad():
        for n, prm in ref.named_parameters():
            if prm.dim() >= 2:
                prm.normal_(0, 0.05)
            elif "adaptive_weight.0" in n:
                prm.fill_(0.25)
            elif "adaptive_weight.1" in n:
                prm.fill_(0.5)
            elif "ayer" in n and "norm" in n.lower() and n.endswith("weight"):
                prm.normal_(1.0, 0.1)
            else:
                prm.normal_(0, 0.05)
    ref.mlm_score.decoder.weight = ref.vilt.embeddings.text_embeddings.word_embeddings.weight   # tie (main.py:121; 4.19 did it)
    g = torch.Generator().manual_seed(42)
    B, L = 3, 48                                       # longer than max_position_embeddings = 40 on purpose
    ids = torch.randint(5, 300, (B, L), generator=g)
    lens = torch.tensor([48, 41, 30])
    mask = (torch.arange(L)[None] < lens[:, None]).long()
    ids = ids * mask
    tt = (torch.arange(L)[None] > torch.tensor([20, 18, 9])[:, None]).long() * mask
    px = torch.randn(B, 2, 3, 64, 64, generator=g)
    sep = torch.zeros(B, 6, dtype=torch.long)
    sep[:, 2] = torch.tensor([20, 18, 9])
    out = ref(input_ids=ids, attention_mask=mask, token_type_ids=tt, pixel_values=px, sep_idx=sep.clone(), return_dict=True)
    logits, trans = out[0].logits, out[1]
    gl = torch.randn(logits.shape, generator=g) * 1e-2
    gt = torch.randn(trans.shape, generator=g) * 1e-1
    ((logits * gl).sum() + (trans * gt).sum()).backward()
    names = ["vilt.embeddings.text_embeddings.word_embeddings.weight", "vilt.embeddings.patch_embeddings.projection.weight",
             "vilt.embeddings.position_embeddings", "vilt.embeddings.cls_token", "vilt.embeddings.token_type_embeddings.weight",
             "vilt.encoder.layer.1.attention.attention.value.weight", "vilt.encoder.layer.0.attention.attention.adaptive_weight.1",
             "vilt.layernorm.weight", "mlm_score.transform.dense.weight", "mlm_score.bias"]
    prm = dict(ref.named_parameters())
    return dict(cfg=dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256, vocab_size=300,
                         image_size=64, patch_size=32),
                params={k: v.detach().clone() for k, v in ref.state_dict().items()}, ids=ids, mask=mask, tt=tt, px=px, sep=sep,
                logits=logits.detach().clone(), trans=trans.detach().clone(), gl=gl, gt=gt,
                grads={n: prm[n].grad.clone() for n in names})


def flava_case():
    """Reference FlavaForMaskedLM (modeling_flava.py:2127-2204) on a tiny config: logits / trans_hidden_states (text rows of
    the multimodal encoder) and a few gradients from each of the three towers."""
    from transformers import FlavaConfig
    Fm = RL.load_flava()
    small = dict(hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256)
    cfg = dict(image_config=dict(small, image_size=32, patch_size=16),
               text_config=dict(small, vocab_size=300, max_position_embeddings=64),
               multimodal_config=dict(small), hidden_size=128, projection_dim=64)
    torch.manual_seed(51)
    ref = Fm.FlavaForMaskedLM(FlavaConfig(**cfg)).eval()
    with torch.no_grad():
        for n, prm in ref.named_parameters():
            if prm.dim() >= 2 and "token" not in n:
                prm.normal_(0, 0.05)
            elif "adaptive_weight.0" in n:
                prm.fill_(0.25)
            elif "adaptive_weight.1" in n:
                prm.fill_(0.5)
            elif "ayer" in n and "norm" in n.lower() and n.endswith("weight"):
                prm.normal_(1.0, 0.1)
            elif n != "flava.logit_scale":
                prm.normal_(0, 0.05)
    ref.cls.decoder.weight = ref.flava.text_model.embeddings.word_embeddings.weight      # tie (4.19 did it in post_init)
    g = torch.Generator().manual_seed(52)
    B, L = 3, 40
    ids = torch.randint(5, 300, (B, L), generator=g)
    lens = torch.tensor([40, 33, 25])
    mask = (torch.arange(L)[None] < lens[:, None]).long()
    ids = ids * mask
    tt = (torch.arange(L)[None] > torch.tensor([17, 15, 8])[:, None]).long() * mask
    px = torch.randn(B, 2, 3, 32, 32, generator=g)
    sep = torch.zeros(B, 6, dtype=torch.long)
    sep[:, 2] = torch.tensor([17, 15, 8])
    out = ref(input_ids=ids, attention_mask=mask, token_type_ids=tt, pixel_values=px, sep_idx=sep.clone(), return_dict=True)
    logits, trans = out[0].logits, out[1]
    gl = torch.randn(logits.shape, generator=g) * 1e-2
    gt = torch.randn(trans.shape, generator=g) * 1e-1
    ((logits * gl).sum() + (trans * gt).sum()).backward()
    names = ["flava.text_model.embeddings.word_embeddings.weight", "flava.text_model.embeddings.position_embeddings.weight",
             "flava.image_model.embeddings.patch_embeddings.projection.weight", "flava.image_model.embeddings.position_embeddings",
             "flava.image_model.embeddings.cls_token", "flava.multimodal_model.cls_token",
             "flava.text_model.encoder.layer.1.attention.attention.value.weight",
             "flava.text_model.encoder.layer.0.attention.attention.adaptive_weight.1",
             "flava.image_model.encoder.layer.0.intermediate.dense.weight",
             "flava.multimodal_model.encoder.layer.1.output.dense.weight", "flava.multimodal_model.layernorm.weight",
             "flava.image_to_mm_projection.weight", "flava.text_to_mm_projection.bias", "cls.transform.dense.weight", "cls.bias"]
    prm = dict(ref.named_parameters())
    unused = sorted(n for n, v in prm.items() if v.grad is None)
    return dict(cfg=cfg, params={k: v.detach().clone() for k, v in ref.state_dict().items()}, ids=ids, mask=mask, tt=tt, px=px,
                sep=sep, logits=logits.detach().clone(), trans=trans.detach().clone(), gl=gl, gt=gt,
                grads={n: prm[n].grad.clone() for n in names}, unused=unused)


def collator_features(seed, n, pretrain):
    """Synthetic per-sample features of the shape `get_dataset` hands the collator (processor.py): ragged token lists plus
    the scalar / list fields the collator pops (data_module.py:96-105)."""
    import random
    rng = random.Random(seed)
    ents = [f"Q{i}" for i in range(7)]
    feats = []
    for i in range(n):
        ln = rng.randint(9, 29)
        f = dict(input_ids=[2] + [rng.randint(5, 60) for _ in range(ln - 2)] + [3], attention_mask=[1] * ln, token_type_ids=[0] * ln,
                 label=rng.randint(0, 6))
        mode = i % 4
        f["head_ent"] = ents[rng.randint(0, 6)] if mode in (0, 1) else (None if mode == 2 else "")
        f["tail_ent"] = ents[rng.randint(0, 6)] if mode in (0, 2) else (None if mode == 1 else ents[rng.randint(0, 6)])
        if pretrain:
            f["pre_type"] = rng.randint(0, 2)
            f["rel_label"] = rng.randint(0, 5)
        else:
            f["rel_idx"] = [rng.randint(1, 3), rng.randint(4, 6)]
            f["sep_idx"] = sorted(rng.sample(range(1, ln), 6))
            f["q_head_idx"] = 1
            f["a_head_idx"] = rng.randint(2, ln - 1)
        feats.append(f)
    return feats


def collator_case():
    """Outputs of the reference collator itself (DataCollatorForSeq2Seq.__call__, data_module.py:91-182) on synthetic features:
    fine-tune and pre-train field sets, with and without pad_to_multiple_of = 8 (`--precision 16`, :215)."""
    import copy
    import tempfile
    DM = RL.load_collator()
    tok = RL.make_bert_tokenizer(tempfile.mkdtemp())
    ents = [f"Q{i}" for i in range(7)]
    feats_img = torch.randn(7, 3, 8, 8, generator=torch.Generator().manual_seed(5))
    DM.entities = ents
    DM.entity2visual_features = feats_img
    cases = {}
    for tag, pretrain, mult, n in (("finetune", False, None, 6), ("finetune_mult8", False, 8, 5), ("pretrain", True, None, 4),
                                   ("single", False, 8, 1)):
        feats = collator_features(17 + len(cases), n, pretrain)
        coll = DM.DataCollatorForSeq2Seq(tok, model=None, label_pad_token_id=tok.pad_token_id, pad_to_multiple_of=mult,
                                         padding="longest", max_length=128, num_labels=len(ents))
        out = coll(copy.deepcopy(feats))
        cases[tag] = dict(features=feats, pad_to_multiple_of=mult, out={k: v.clone() for k, v in out.items()}, keys=list(out.keys()))
    return dict(entities=ents, image_features=feats_img, pad_token_id=tok.pad_token_id, cases=cases)


if __name__ == "__main__":
    assert RL.available(), "reference tree not found"
    if "--collator" in sys.argv:     # only the collator fixture (the model fixtures are left untouched)
        os.makedirs(GOLD, exist_ok=True)
        torch.save(collator_case(), os.path.join(GOLD, "collator.pt"))
        print("collator.pt", os.path.getsize(os.path.join(GOLD, "collator.pt")))
        sys.exit(0)
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.save(tiny_case(), os.path.join(GOLD, "mkgformer_tiny.pt"))
    torch.save(loss_case(), os.path.join(GOLD, "ls_ce.pt"))
    torch.save(stack_case(), os.path.join(GOLD, "preln_stack.pt"))
    torch.save(vilt_case(), os.path.join(GOLD, "vilt_tiny.pt"))
    torch.save(flava_case(), os.path.join(GOLD, "flava_tiny.pt"))
    torch.save(collator_case(), os.path.join(GOLD, "collator.pt"))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))
