"""Kernel-level parity (through the C ABI) of the non-GEMM operators against fp32 torch restatements of the
reference ops (file:line in each test).  Inputs are bf16-rounded where the kernel consumes bf16, so the only
differences are accumulation order, fast-math exp, the bf16 rounding of bf16 outputs (2^-8 relative) and, in the
tensor-core attention, the bf16 rounding of the P / dS / dO operands of the second product (2^-9 relative each)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def _rand(shape, seed, scale=1.0, bf16=True):
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(shape, generator=g) * scale
    return (t.to(torch.bfloat16) if bf16 else t).cuda()


def _ref_attention(q, k, v, heads, scale, key_mask=None, sep=None, w0=None, w1=None, pk=None, pv=None):
    """BertSelfAttention.forward modeling_unimo.py:317-377 / CLIPAttention.forward :212-272 in fp32."""
    B, Lq, HD = q.shape

    def hs(x):
        return x.float().view(B, x.shape[1], heads, 64).permute(0, 2, 1, 3)
    kk, vv = k, v
    n1 = 0
    if pk is not None:
        n1 = pk.shape[1]
        kk, vv = torch.cat([pk, k], 1), torch.cat([pv, v], 1)
    S = (hs(q) @ hs(kk).transpose(-1, -2)) * scale
    if sep is not None:
        n = S.shape[-1]
        ar_c = torch.arange(n, device=q.device).view(1, 1, 1, n) - n1
        ar_r = torch.arange(Lq, device=q.device).view(1, 1, Lq, 1)
        s = sep[:, 2].view(B, 1, 1, 1)
        f = torch.where((ar_c >= s) & (ar_r < s), torch.clamp(w0, 0, 0.5),
                        torch.where((ar_c >= s) & (ar_r >= s), torch.clamp(w1, 0.5, 1), torch.ones((), device=q.device)))
        S = S * f
    if key_mask is not None:
        m = (1.0 - key_mask.float()) * -10000.0
        if n1:
            m = torch.cat([torch.zeros(B, n1, device=q.device), m], 1)
        S = S + m[:, None, None, :]
    P = torch.softmax(S, -1)
    return (P @ hs(vv)).permute(0, 2, 1, 3).reshape(B, Lq, HD), torch.logsumexp(S, -1)


@pytest.mark.parametrize("B,L,heads", [(3, 24, 2), (4, 64, 12), (2, 128, 4), (2, 200, 2), (1, 418, 3)])   # > 128: the tiled kernels
def test_text_attention_fwd_bwd(B, L, heads):
    from mkg_analogy_b200 import ops
    HD = heads * 64
    qkv = _rand((B, L, 3 * HD), 1, 1.0)
    q, k, v = qkv[..., :HD], qkv[..., HD:2 * HD], qkv[..., 2 * HD:]
    g = torch.Generator().manual_seed(2)
    lens = torch.randint(L // 2, L + 1, (B,), generator=g)
    mask = (torch.arange(L)[None] < lens[:, None]).long().cuda()
    sep = torch.zeros(B, 6, dtype=torch.long)
    sep[:, 2] = torch.randint(3, L // 2, (B,), generator=g)
    sep = sep.cuda()
    w0 = torch.tensor([0.3], device="cuda", requires_grad=True)
    w1 = torch.tensor([0.5], device="cuda", requires_grad=True)      # exactly at the clamp bound (reference init)
    out, out32, lse = ops.attention_fwd(q, k, v, heads, 0.125, key_mask=mask, sep_idx=sep, w0=w0.detach(), w1=w1.detach(), want_f32=True)
    qf, kf, vf = (t.float().clone().requires_grad_(True) for t in (q, k, v))
    ref, ref_lse = _ref_attention(qf, kf, vf, heads, 0.125, mask, sep, w0, w1)
    print('attn fwd rel', rel(out32, ref), rel(out.float(), ref))
    assert rel(out32, ref) < 5e-3 and rel(out.float(), ref) < 7e-3
    assert (lse - ref_lse).abs().max().item() < 1e-3
    dout = _rand((B, L, HD), 3, 1.0, bf16=False)
    ref.backward(dout)
    dk_add = _rand((B, L, HD), 4, 0.1, bf16=False)
    dv_add = _rand((B, L, HD), 5, 0.1, bf16=False)
    dq, dk, dv, _, _, dw0, dw1 = ops.attention_bwd(dout, out, q, k, v, lse, heads, 0.125, key_mask=mask, sep_idx=sep,
                                                   w0=w0.detach(), w1=w1.detach(), dk_add=dk_add, dv_add=dv_add)
    # the key pass rounds dO to bf16 -> ~2^-9 relative noise on dK/dV
    print('attn bwd rel', rel(dq.float(), qf.grad), rel(dk.float(), kf.grad + dk_add), rel(dv.float(), vf.grad + dv_add),
          dw0.item(), w0.grad.item(), dw1.item(), w1.grad.item())
    assert rel(dq.float(), qf.grad) < 1.2e-2
    assert rel(dk.float(), kf.grad + dk_add) < 1.2e-2
    assert rel(dv.float(), vf.grad + dv_add) < 1.2e-2
    assert abs(dw0.item() - w0.grad.item()) <= 2e-2 * abs(w0.grad.item()) + 1e-2
    assert abs(dw1.item() - w1.grad.item()) <= 2e-2 * abs(w1.grad.item()) + 1e-2


def test_adaptive_weight_clamp_blocks_gradient():
    """clamp passes the gradient only inside [lo, hi] (inclusive): w0 = 0.7 > 0.5 gets none."""
    from mkg_analogy_b200 import ops
    B, L, heads = 2, 32, 2
    HD = heads * 64
    q, k, v = (_rand((B, L, HD), s) for s in (1, 2, 3))
    sep = torch.zeros(B, 6, dtype=torch.long)
    sep[:, 2] = 10
    sep = sep.cuda()
    mask = torch.ones(B, L, dtype=torch.long, device="cuda")
    w0 = torch.tensor([0.7], device="cuda")
    w1 = torch.tensor([0.8], device="cuda")
    out, _, lse = ops.attention_fwd(q, k, v, heads, 0.125, key_mask=mask, sep_idx=sep, w0=w0, w1=w1)
    dout = _rand((B, L, HD), 4, bf16=False)
    *_, dw0, dw1 = ops.attention_bwd(dout, out, q, k, v, lse, heads, 0.125, key_mask=mask, sep_idx=sep, w0=w0, w1=w1)
    assert dw0.item() == 0.0 and dw1.item() != 0.0


@pytest.mark.parametrize("B,Nv,L,heads", [(2, 9, 24, 2), (3, 99, 64, 12), (1, 513, 64, 2)])
def test_vision_attention_with_text_prefix(B, Nv, L, heads):
    """CLIPAttention with the text K/V prefix (:227-229): no mask, prefix keys first."""
    from mkg_analogy_b200 import ops
    HD = heads * 64
    qkv = _rand((B, Nv, 3 * HD), 11)
    tqkv = _rand((B, L, 3 * HD), 12)
    q, k, v = qkv[..., :HD], qkv[..., HD:2 * HD], qkv[..., 2 * HD:]
    pk, pv = tqkv[..., HD:2 * HD], tqkv[..., 2 * HD:]
    out, _, lse = ops.attention_fwd(q, k, v, heads, 0.125, prefix_k=pk, prefix_v=pv)
    qf, kf, vf, pkf, pvf = (t.float().clone().requires_grad_(True) for t in (q, k, v, pk, pv))
    ref, _ = _ref_attention(qf, kf, vf, heads, 0.125, pk=pkf, pv=pvf)
    assert rel(out.float(), ref) < 7e-3
    dout = _rand((B, Nv, HD), 13, bf16=False)
    ref.backward(dout)
    dq, dk, dv, dpk, dpv, _, _ = ops.attention_bwd(dout, out, q, k, v, lse, heads, 0.125, prefix_k=pk, prefix_v=pv)
    print('vis attn bwd rel', rel(dq.float(), qf.grad), rel(dk.float(), kf.grad), rel(dv.float(), vf.grad), rel(dpk, pkf.grad), rel(dpv, pvf.grad))
    assert rel(dq.float(), qf.grad) < 1.2e-2 and rel(dk.float(), kf.grad) < 1.2e-2 and rel(dv.float(), vf.grad) < 1.2e-2
    assert rel(dpk, pkf.grad) < 1.2e-2 and rel(dpv, pvf.grad) < 1.2e-2


@pytest.mark.parametrize("B,L,Nv,H", [(2, 24, 9, 128), (3, 64, 99, 768), (2, 30, 513, 1024)])
def test_fusion_fwd_bwd(B, L, Nv, H):
    """BertFusion.forward modeling_unimo.py:400-414: softmax(ctx vh^T) vh, unscaled, unmasked, fp32."""
    from mkg_analogy_b200 import ops
    ctx = _rand((B, L, H), 21, 0.5, bf16=False).requires_grad_(True)
    vh = _rand((B, Nv, H), 22, 1.0, bf16=False).requires_grad_(True)
    out, lse = ops.fusion_fwd(ctx.detach(), vh.detach())
    S = ctx @ vh.transpose(1, 2)
    ref = torch.softmax(S, -1) @ vh
    assert rel(out.float(), ref) < 6e-3
    assert (lse - torch.logsumexp(S, -1)).abs().max().item() < 2e-3
    dg = _rand((B, L, H), 23, bf16=False)
    ref.backward(dg)
    dvh_in = _rand((B, Nv, H), 24, bf16=False)
    dctx, dvh = ops.fusion_bwd(dg, ctx.detach(), vh.detach(), lse, dvh_in=dvh_in)
    # scores are O(sqrt(H)) and fast-math exp has ~1e-6 * |x| relative error -> few 1e-4 on the gradients
    assert rel(dctx, ctx.grad) < 2e-3
    assert rel(dvh, vh.grad + dvh_in) < 2e-3


@pytest.mark.parametrize("M,H,eps", [(50, 128, 1e-12), (2048, 768, 1e-12), (3168, 768, 1e-5), (100, 1024, 1e-5)])
def test_layernorm_fwd_bwd(M, H, eps):
    from mkg_analogy_b200 import ops
    x = _rand((M, H), 31, 2.0, bf16=False).requires_grad_(True)
    gamma = (_rand((H,), 32, 0.2, bf16=False) + 1).requires_grad_(True)
    beta = _rand((H,), 33, 0.2, bf16=False).requires_grad_(True)
    y32, y16, mean, rstd = ops.layernorm_fwd(x.detach(), gamma.detach(), beta.detach(), eps)
    ref = torch.nn.functional.layer_norm(x, (H,), gamma, beta, eps)
    assert (y32 - ref).abs().max().item() < 2e-5
    assert rel(y16.float(), ref) < 4e-3
    dy = _rand((M, H), 34, bf16=False)
    add = _rand((M, H), 35, bf16=False)
    ref.backward(dy)
    dx, dgamma, dbeta = ops.layernorm_bwd(dy, x.detach(), mean, rstd, gamma.detach(), add=add)
    assert rel(dx, x.grad + add) < 1e-5
    assert rel(dgamma, gamma.grad) < 1e-4 and rel(dbeta, beta.grad) < 1e-4


def test_ls_ce_against_reference_golden():
    """LabelSmoothSoftmaxCEV1 (lit_models/utils.py:42-66) known answers generated by the reference itself."""
    import os
    from mkg_analogy_b200 import ops
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ls_ce.pt"), weights_only=False)
    for tag, c in g.items():
        if tag.endswith("ignore"):
            continue   # ignore_index never occurs on this path; the kernel takes valid labels only
        loss, dl = ops.ls_ce(c["x"].cuda(), c["y"].cuda(), 0.1)
        assert abs(loss.item() - c["loss"].item()) < 2e-5 * max(1.0, abs(c["loss"].item())), tag
        assert (dl.cpu() - c["grad"]).abs().max().item() < 2e-6, tag
    # eps = 0 -> nn.CrossEntropyLoss (transformer.py:25)
    x = _rand((5, 192), 41, 3.0, bf16=False).requires_grad_(True)
    y = torch.randint(0, 192, (5,)).cuda()
    ref = torch.nn.functional.cross_entropy(x, y)
    ref.backward()
    loss, dl = ops.ls_ce(x.detach(), y, 0.0)
    assert abs(loss.item() - ref.item()) < 1e-5 and (dl - x.grad).abs().max().item() < 1e-6
