"""Fused-route LightningModule shim: `--litmodel_class` is loaded by name (MarT/main.py:31,46,74), so
a subclass of the reference `TransformerLitModel` that overrides only `training_step` / `_eval` plugs in with no reference
file edited.  The subclass is built against whatever base class the host environment provides:

    # MarT/lit_models/__init__.py  (one added line)
    from mkg_analogy_b200.lit import make_fused_litmodel; B200TransformerLitModel = make_fused_litmodel(TransformerLitModel)

`training_step` scores the [MASK] rows only and leaves the gradients in the module's flat gradient buffer; it returns a
loss tensor whose `backward()` is a no-op hook-up, so Lightning's automatic optimisation keeps working unchanged.
"""
import numpy as np
import torch


class _GradsReady(torch.autograd.Function):
    """loss.backward() for the fused route: gradients were already produced by b200mart_backward inside training_step;
    backward only attaches them to the parameters (scaled by the incoming factor, 1 under plain automatic optimisation)."""

    @staticmethod
    def forward(ctx, anchor, loss, module):
        ctx.module = module
        return loss.clone()

    @staticmethod
    def backward(ctx, g):
        m = ctx.module
        if float(g) != 1.0:
            m.engine.grads32.mul_(g)
        m._attach_grads()
        return None, None, None


def make_fused_litmodel(base_cls):
    class B200TransformerLitModel(base_cls):
        def _split(self, batch):
            label = batch.pop("label")
            extra = {k: batch.pop(k) if k in batch else None for k in ("rel_label", "pre_type", "rel_idx", "q_head_idx", "a_head_idx")}
            return label, extra

        def _head(self, label, extra):
            dev = self.model.engine.device
            if self.args.pretrain:
                return dict(mode="pretrain", entity_range=(self.entity_id_st, self.entity_id_ed),
                            relation_range=(self.relation_id_st, self.relation_id_ed), label=label, pre_type=extra["pre_type"],
                            label_smoothing=self.args.label_smoothing, mask_token_id=self.tokenizer.mask_token_id)
            if not hasattr(self, "_col_ids"):
                self._col_ids = torch.as_tensor(self.analogy_entity_ids, device=dev, dtype=torch.int64)
            return dict(mode="finetune", col_ids=self._col_ids, label=label, rel_idx=extra["rel_idx"].contiguous(),
                        q_head_idx=extra["q_head_idx"].contiguous(), a_head_idx=extra["a_head_idx"].contiguous(),
                        alpha=self.alpha, label_smoothing=self.args.label_smoothing, mask_token_id=self.tokenizer.mask_token_id)

        def _inputs(self, batch):
            eng = self.model.engine
            b = {k: batch[k].to(eng.device).contiguous() for k in ("input_ids", "attention_mask", "pixel_values")}
            b["attention_mask"] = b["attention_mask"].to(torch.int64)
            for k in ("token_type_ids", "sep_idx"):
                if batch.get(k) is not None:
                    b[k] = batch[k].to(eng.device).contiguous()
            return b

        def training_step(self, batch, batch_idx):        # reference: lit_models/transformer.py:59-113
            label, extra = self._split(batch)
            m, eng = self.model, self.model.engine
            mb = self._inputs(batch)
            eng.mirror_stale = True                        # a torch optimizer may have stepped the fp32 masters
            eng.train(True)
            m._prepare_grads()
            eng.forward(mb, want_z=False)
            loss, _, _ = eng.head_loss(mb, self._head(label, extra), compute_grad=True)
            eng.backward(mb)
            return _GradsReady.apply(m._anchor, loss.reshape(()), m)

        @torch.no_grad()
        def _eval(self, batch, batch_idx):                # reference: lit_models/transformer.py:115-166
            label, extra = self._split(batch)
            eng = self.model.engine
            mb = self._inputs(batch)
            eng.mirror_stale = True
            eng.train(False)
            eng.forward(mb, want_z=False)
            _, _, ranks = eng.head_loss(mb, self._head(label, extra), compute_grad=False, want_ranks=True)
            ranks = ranks.cpu().numpy()
            if not self.args.pretrain:
                return dict(entity_ranks=np.array(ranks))
            rel = (extra["pre_type"] == 2).cpu().numpy()
            out = {}
            if (~rel).any():
                out["entity_ranks"] = np.array(ranks[~rel])
            if rel.any():
                out["relation_ranks"] = np.array(ranks[rel])
            if not out:
                raise ValueError("entity and relation cannot be None at the same time.")
            return out

    return B200TransformerLitModel
