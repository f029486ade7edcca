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


if __name__ == "__main__":
    assert RL.available(), "reference tree not found"
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.save(tiny_case(), os.path.join(GOLD, "mkgformer_tiny.pt"))
    torch.save(loss_case(), os.path.join(GOLD, "ls_ce.pt"))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))
