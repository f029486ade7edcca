"""mkg_analogy_b200.nn: the dense plumbing of the FLAVA / ViLT drop-ins on the library's kernels (tcgen05 GEMM, LayerNorm, fused
LS-CE) against plain PyTorch fp32 (`F.linear`, `nn.Conv2d`, `nn.LayerNorm`, LabelSmoothSoftmaxCEV1 restated): values and every
gradient.  Tolerance: bf16 operands / fp32 accumulation -> 1e-2 relative (Frobenius)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


@pytest.mark.parametrize("M,K,N", [(300, 768, 768), (67, 128, 300), (5, 768, 2063)])
def test_linear_forward_backward(M, K, N):
    from mkg_analogy_b200 import nn as B
    g = torch.Generator(device="cuda").manual_seed(M + N)
    x = torch.randn(M, K, device="cuda", generator=g, requires_grad=True)
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).requires_grad_(True)
    b = torch.randn(N, device="cuda", generator=g).requires_grad_(True)
    dy = torch.randn(M, N, device="cuda", generator=g)
    y = B.linear(x, w, b)
    y.backward(dy)
    got = (y.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone())
    x.grad = w.grad = b.grad = None
    yr = F.linear(x, w, b)
    yr.backward(dy)
    for name, a, r in zip(("y", "dx", "dw", "db"), got, (yr.detach(), x.grad, w.grad, b.grad)):
        assert rel(a, r) <= 1e-2, (name, rel(a, r))


def test_patch_embed_and_layer_norm():
    from mkg_analogy_b200 import nn as B
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(3, 128, kernel_size=16, stride=16).cuda()
    ln = torch.nn.LayerNorm(128, eps=1e-12).cuda()
    with torch.no_grad():
        ln.weight.normal_(1.0, 0.1)
        ln.bias.normal_(0.0, 0.1)
    px = torch.randn(3, 3, 64, 64, device="cuda", requires_grad=True)
    dy = torch.randn(3, 16, 128, device="cuda")
    y = B.layer_norm(B.patch_embed(px, conv), ln)
    y.backward(dy)
    got = (y.detach(), px.grad.clone(), conv.weight.grad.clone(), conv.bias.grad.clone(), ln.weight.grad.clone(), ln.bias.grad.clone())
    px.grad = None
    for p in list(conv.parameters()) + list(ln.parameters()):
        p.grad = None
    torch.backends.cudnn.allow_tf32 = False
    yr = ln(conv(px).flatten(2).transpose(1, 2))
    yr.backward(dy)
    ref = (yr.detach(), px.grad, conv.weight.grad, conv.bias.grad, ln.weight.grad, ln.bias.grad)
    for name, a, r in zip(("y", "dpx", "dW", "db", "dgamma", "dbeta"), got, ref):
        assert rel(a, r) <= 1e-2, (name, rel(a, r))


def test_mask_row_loss_matches_label_smooth_ce():
    from mkg_analogy_b200 import nn as B
    g = torch.Generator(device="cuda").manual_seed(9)
    Bn, L, H, V = 6, 24, 128, 700
    trans = torch.randn(Bn, L, H, device="cuda", generator=g, requires_grad=True)
    table = (torch.randn(V, H, device="cuda", generator=g) * 0.1).requires_grad_(True)
    bias = (torch.randn(V, device="cuda", generator=g) * 0.1).requires_grad_(True)
    cols = torch.randperm(V, device="cuda", generator=g)[:203].sort().values
    mpos = torch.randint(0, L, (Bn,), device="cuda", generator=g)
    label = torch.randint(0, 203, (Bn,), device="cuda", generator=g)
    for columns in (cols, (100, 292)):
        for t in (trans, table, bias):
            t.grad = None
        loss, _ = B.mask_row_loss(trans, mpos, table, bias, columns, label if not isinstance(columns, tuple) else label % 192)
        loss.backward()
        got = (loss.detach(), trans.grad.clone(), table.grad.clone(), bias.grad.clone())
        for t in (trans, table, bias):
            t.grad = None
        lab = label if not isinstance(columns, tuple) else label % 192
        E, bs = (table[columns], bias[columns]) if not isinstance(columns, tuple) else (table[columns[0]:columns[1]], bias[columns[0]:columns[1]])
        x = trans[torch.arange(Bn, device="cuda"), mpos] @ E.t() + bs
        lp = F.log_softmax(x, dim=-1)
        q = torch.full_like(lp, 0.1 / lp.shape[-1]).scatter_(1, lab[:, None], 0.9)        # lit_models/utils.py:42-66
        ref_loss = -(q * lp).sum(-1).mean()
        ref_loss.backward()
        assert abs(got[0].item() - ref_loss.item()) <= 2e-3
        for name, a, r in zip(("dtrans", "dtable", "dbias"), got[1:], (trans.grad, table.grad, bias.grad)):
            assert rel(a, r) <= 2e-2, (name, rel(a, r))
