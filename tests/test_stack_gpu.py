"""The pre-LayerNorm encoder stack (FLAVA / ViLT block) through the C ABI against golden outputs of the reference
FlavaEncoder / ViltEncoder (tests/golden/preln_stack.pt) and the oracle restatement.  Tolerances as for the MKGformer path
(bf16 operands, fp32 accumulate): outputs 2e-2 relative, gradients 6e-2 per tensor."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _stack_from(c, **kw):
    from mkg_analogy_b200.stack import EncoderStack
    st = EncoderStack(128, 3, 2, 256, eps=1e-12, act="gelu", device="cuda", **kw)
    missing, unexpected = st.load_state_dict(c["params"], strict=True)
    return st


@pytest.mark.parametrize("tag", ["flava", "vilt"])
def test_stack_matches_reference_encoder(tag):
    c = torch.load(os.path.join(GOLD, "preln_stack.pt"), weights_only=False)[tag]
    st = _stack_from(c).eval()
    x = c["x"].cuda().requires_grad_(True)
    y = st(x, attention_mask=c["mask"].cuda(), sep_idx=c["sep"].cuda(), row_lo=c["row_lo"])
    y.backward(c["gy"].cuda())
    torch.cuda.synchronize()
    print(f"{tag}: y rel err {rel(y.detach().cpu(), c['y']):.3e}, dx rel err {rel(x.grad.cpu(), c['dx']):.3e}")
    assert rel(y.detach().cpu(), c["y"]) <= 2e-2
    assert rel(x.grad.cpu(), c["dx"]) <= 6e-2
    named = dict(st.named_parameters())
    for n, gr in c["grads"].items():
        e = rel(named[n].grad.cpu(), gr)
        assert e <= (0.25 if "adaptive_weight" in n else 6e-2), (n, e)
    for n, ref_norm in c["grad_norm"].items():
        if ref_norm < 1e-7:
            continue
        got = named[n].grad.norm().item()
        assert abs(got - ref_norm) <= (0.25 if "adaptive_weight" in n else 0.06) * ref_norm + 1e-6, (n, got, ref_norm)


def test_stack_without_mask_or_scaling_and_autograd_around_it():
    """Image-encoder use (no mask, no sep_idx) inside a torch graph: an nn.Linear before and a loss after the stack train
    through it; adaptive weights keep .grad = None as in the reference."""
    from oracle import mkgformer_oracle as O
    c = torch.load(os.path.join(GOLD, "preln_stack.pt"), weights_only=False)["flava"]
    st = _stack_from(c).eval()
    torch.manual_seed(0)
    pre = torch.nn.Linear(64, 128).cuda()
    inp = torch.randn(2, 50, 64, device="cuda")
    y = st(pre(inp))
    (y ** 2).mean().backward()
    P = {k: v.cuda().clone().requires_grad_(True) for k, v in c["params"].items()}
    pre2 = torch.nn.Linear(64, 128).cuda()
    pre2.load_state_dict(pre.state_dict())
    yr = O.preln_stack(P, pre2(inp), 2)
    (yr ** 2).mean().backward()
    assert rel(y.detach(), yr.detach()) <= 2e-2
    assert rel(pre.weight.grad, pre2.weight.grad) <= 6e-2
    named = dict(st.named_parameters())
    assert named["layer.0.attention.attention.adaptive_weight.0"].grad is None
    assert rel(named["layer.1.output.dense.weight"].grad, P["layer.1.output.dense.weight"].grad) <= 6e-2


def test_stack_training_dropout_matches_oracle_masks():
    from oracle import mkgformer_oracle as O
    c = torch.load(os.path.join(GOLD, "preln_stack.pt"), weights_only=False)["vilt"]
    st = _stack_from(c, hidden_dropout=0.1, attn_dropout=0.1).train()
    x = c["x"].cuda().requires_grad_(True)
    y = st(x, attention_mask=c["mask"].cuda(), sep_idx=c["sep"].cuda(), row_lo=c["row_lo"])
    y.backward(c["gy"].cuda())
    drop = O.hash_dropout(st._last_seed, 0.1, 0.1)
    P = {k: v.cuda().clone().requires_grad_(True) for k, v in c["params"].items()}
    xr = c["x"].cuda().requires_grad_(True)
    yr = O.preln_stack(P, xr, 2, attention_mask=c["mask"].cuda(), sep_split=c["sep"][:, 2].cuda(), row_lo=c["row_lo"], drop=drop)
    yr.backward(c["gy"].cuda())
    assert rel(yr.detach().cpu(), c["y"]) > 1e-2                 # the masks matter
    assert rel(y.detach(), yr.detach()) <= 2e-2 and rel(x.grad, xr.grad) <= 6e-2
    named = dict(st.named_parameters())
    for n in ("layer.0.attention.attention.query.weight", "layer.2.output.dense.bias", "layer.1.attention.output.dense.bias",
              "layer.0.intermediate.dense.bias", "layer.2.layernorm_after.weight"):
        assert rel(named[n].grad, P[n].grad) <= 6e-2, n


@pytest.mark.parametrize("tag", ["flava", "vilt"])
def test_encoder_adapter_keeps_the_reference_call_signature(tag):
    """EncoderAdapter is called exactly as FlavaTextModel (modeling_flava.py:1100-1108) / ViltModel (modeling_vilt.py:850-859)
    call their encoders: additive 4-D mask, sep_idx, img_length; result indexable and with hidden_states[-1]."""
    from mkg_analogy_b200.accelerate import EncoderAdapter
    c = torch.load(os.path.join(GOLD, "preln_stack.pt"), weights_only=False)[tag]
    ad = EncoderAdapter(128, 3, 2, 256, eps=1e-12, act="gelu", device="cuda").eval()
    ad.load_state_dict(c["params"], strict=True)
    ext = (1.0 - c["mask"][:, None, None, :].float().cuda()) * -10000.0
    kw = dict(attention_mask=ext, head_mask=[None] * 3, output_hidden_states=True, return_dict=True, sep_idx=c["sep"].cuda())
    if tag == "vilt":
        kw["img_length"] = c["row_lo"] - 1
    with torch.no_grad():
        out = ad(c["x"].cuda(), **kw)
    assert rel(out[0].cpu(), c["y"]) <= 2e-2 and out.hidden_states[-1] is out.last_hidden_state
    with torch.no_grad():
        tup = ad(c["x"].cuda(), attention_mask=ext, return_dict=False, sep_idx=c["sep"].cuda(),
                 img_length=(c["row_lo"] - 1) if tag == "vilt" else None)
    assert isinstance(tup, tuple) and rel(tup[0].cpu(), c["y"]) <= 2e-2
    with pytest.raises(NotImplementedError):
        ad(c["x"].cuda(), output_attentions=True)


def test_vilt_dropin_model_vs_reference_golden():
    """ViltKGC end to end on the GPU (CUDA encoder stack + torch plumbing) against the reference ViltForMaskedLM fixture:
    text-row logits, trans_hidden_states and parameter gradients."""
    from mkg_analogy_b200.modeling_vilt import ViltKGC
    c = torch.load(os.path.join(GOLD, "vilt_tiny.pt"), weights_only=False)
    m = ViltKGC(device="cuda", **c["cfg"])
    m.load_state_dict(c["params"])
    m.tie_weights()
    m.eval()
    out = m(input_ids=c["ids"].cuda(), attention_mask=c["mask"].cuda(), token_type_ids=c["tt"].cuda(), pixel_values=c["px"].cuda(),
            sep_idx=c["sep"].cuda(), return_dict=True)
    logits, trans = out[0].logits, out[1]
    print(f"vilt: logits rel err {rel(logits.detach().cpu(), c['logits']):.3e}, trans {rel(trans.detach().cpu(), c['trans']):.3e}")
    assert rel(logits.detach().cpu(), c["logits"]) <= 2e-2 and rel(trans.detach().cpu(), c["trans"]) <= 2e-2
    ((logits * c["gl"].cuda()).sum() + (trans * c["gt"].cuda()).sum()).backward()
    prm = dict(m.named_parameters())
    for n, gr in c["grads"].items():
        e = rel(prm[n].grad.cpu(), gr)
        assert e <= (0.25 if "adaptive_weight" in n else 6e-2), (n, e)


def test_flava_dropin_model_vs_reference_golden():
    """FlavaKGC end to end on the GPU (three CUDA encoder stacks + torch plumbing) against the reference FlavaForMaskedLM
    fixture: logits, trans_hidden_states and parameter gradients from every tower."""
    from mkg_analogy_b200.modeling_flava import FlavaKGC
    c = torch.load(os.path.join(GOLD, "flava_tiny.pt"), weights_only=False)
    m = FlavaKGC(device="cuda", **c["cfg"])
    m.load_state_dict(c["params"])
    m.tie_weights()
    m.eval()
    out = m(input_ids=c["ids"].cuda(), attention_mask=c["mask"].cuda(), token_type_ids=c["tt"].cuda(), pixel_values=c["px"].cuda(),
            sep_idx=c["sep"].cuda(), return_dict=True)
    logits, trans = out[0].logits, out[1]
    print(f"flava: logits rel err {rel(logits.detach().cpu(), c['logits']):.3e}, trans {rel(trans.detach().cpu(), c['trans']):.3e}")
    assert rel(logits.detach().cpu(), c["logits"]) <= 2e-2 and rel(trans.detach().cpu(), c["trans"]) <= 2e-2
    ((logits * c["gl"].cuda()).sum() + (trans * c["gt"].cuda()).sum()).backward()
    prm = dict(m.named_parameters())
    for n, gr in c["grads"].items():
        e = rel(prm[n].grad.cpu(), gr)
        assert e <= (0.25 if "adaptive_weight" in n else 6e-2), (n, e)
