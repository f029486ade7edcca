"""Kernel-level parity of the tcgen05 GEMM (through the C ABI) against torch fp32 matmul on the same
bf16-rounded operands. Tolerance: fp32 accumulation of exact bf16 products -> only summation-order
error: |err| <= 2e-3 * sqrt(K)-scaled bound below; bf16 outputs add one rounding (rel 2^-8)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(shape, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * 0.5).to(torch.bfloat16).cuda()


def _check(out, ref, K, bf16_out=False):
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    tol = 1e-5 * K ** 0.5 * scale + (scale * 2 ** -8 if bf16_out else 0)
    assert err <= tol, f"max err {err} > tol {tol} (scale {scale})"


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 128, 128), (2048, 768, 768), (3168, 2304, 768),
                                   (100, 200, 192), (32, 2063, 768), (2048, 768, 3072)])
@pytest.mark.parametrize("bn", [64, 128, 256, 512])     # 512 = 256x256 CTA-pair (cta_group::2) kernel
def test_gemm_nt(M, N, K, bn):
    from mkg_analogy_b200 import _lib
    A, B = _mk((M, K), 1), _mk((N, K), 2)
    out = torch.full((M, N), float("nan"), device="cuda")
    if bn == 512 and N % 2:      # odd N needs the scalar GENERIC epilogue, which has no CTA-pair instantiation: loud error
        with pytest.raises(_lib.B200MartError):
            _lib.gemm_bf16(A, B, out_f32=out, bn=bn)
        return
    _lib.gemm_bf16(A, B, out_f32=out, bn=bn)
    torch.cuda.synchronize()
    _check(out, A.float() @ B.float().t(), K)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (2048, 768, 768), (3168, 768, 3072), (100, 200, 192)])
def test_gemm_nn_b_mn(M, N, K):
    """dX = dY @ W with W stored [K][N] (MN-major B)."""
    from mkg_analogy_b200 import _lib
    A, B = _mk((M, K), 3), _mk((K, N), 4)
    for bn in (0, 512):
        out = torch.full((M, N), float("nan"), device="cuda")
        _lib.gemm_bf16(A, B, b_mn=True, out_f32=out, bn=bn)
        torch.cuda.synchronize()
        _check(out, A.float() @ B.float(), K)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (768, 768, 2048), (3072, 768, 3168), (200, 136, 100)])
@pytest.mark.parametrize("split", [1, 4])
def test_gemm_tn_both_mn(M, N, K, split):
    """dW = dY^T @ X with both operands stored [K][rows]; accumulate (+=) semantics, split-K."""
    from mkg_analogy_b200 import _lib
    A, B = _mk((K, M), 5), _mk((K, N), 6)
    for bn in (0, 128, 512):
        out = torch.ones((M, N), device="cuda")
        _lib.gemm_bf16(A, B, a_mn=True, b_mn=True, out_f32=out, accumulate=True, split_k=split, bn=bn)
        torch.cuda.synchronize()
        _check(out, 1.0 + A.float().t() @ B.float(), K)


def test_gemm_a_mn_only():
    from mkg_analogy_b200 import _lib
    M, N, K = 256, 192, 320
    A, B = _mk((K, M), 7), _mk((N, K), 8)
    out = torch.full((M, N), float("nan"), device="cuda")
    _lib.gemm_bf16(A, B, a_mn=True, out_f32=out)
    torch.cuda.synchronize()
    _check(out, A.float().t() @ B.float().t(), K)


@pytest.mark.parametrize("act", [0, 1, 2])
def test_gemm_epilogue_fwd(act):
    """bias + activation + residual, dual output, saved pre-activation, dual operand pair (K1+K2)."""
    from mkg_analogy_b200 import _lib
    M, N, K1, K2 = 300, 384, 256, 128
    A, B, A2, B2 = _mk((M, K1), 9), _mk((N, K1), 10), _mk((M, K2), 11), _mk((N, K2), 12)
    g = torch.Generator().manual_seed(13)
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).cuda()
    o32 = torch.full((M, N), float("nan"), device="cuda")
    o16 = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16)
    pre = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16)
    _lib.gemm_bf16(A, B, A2=A2, B2=B2, bias=bias, act=act, residual=res, out_f32=o32, out_bf16=o16,
                   out_preact=pre, alpha=0.5)
    torch.cuda.synchronize()
    z = 0.5 * (A.float() @ B.float().t() + A2.float() @ B2.float().t()) + bias
    if act == 1:
        y = torch.nn.functional.gelu(z)
    elif act == 2:
        y = z * torch.sigmoid(1.702 * z)
    else:
        y = z
    ref = y + res
    assert (o32 - ref).abs().max().item() < 2e-3   # --use_fast_math exp in quick-gelu
    assert (o16.float() - ref).abs().max().item() < 2e-3 + ref.abs().max().item() * 2 ** -8
    assert (pre.float() - z).abs().max().item() < 1e-4 + z.abs().max().item() * 2 ** -8


@pytest.mark.parametrize("act", [1, 2])
def test_gemm_saved_derivative_roundtrip(act):
    """Model path: the forward epilogue saves act'(pre) (save_grad), the backward epilogue multiplies by it (dact=3)."""
    from mkg_analogy_b200 import _lib
    M, N, K = 3168, 3072, 768
    A, B = _mk((M, K), 21), _mk((N, K), 22)
    bias = torch.randn(N, generator=torch.Generator().manual_seed(23)).cuda()
    y16 = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16)
    g16 = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16)
    _lib.gemm_bf16(A, B, bias=bias, act=act, out_preact=g16, save_grad=True, out_bf16=y16)
    z = (A.float() @ B.float().t() + bias).requires_grad_(True)
    y = torch.nn.functional.gelu(z) if act == 1 else z * torch.sigmoid(1.702 * z)
    (gz,) = torch.autograd.grad(y.sum(), z)
    assert (y16.float() - y.detach()).abs().max().item() < 2e-3 + y.abs().max().item() * 2 ** -8
    assert (g16.float() - gz).abs().max().item() < 2e-3 + gz.abs().max().item() * 2 ** -8
    dY, W = _mk((M, K), 24), _mk((K, N), 25)       # dX = (dY @ W) * act'(z), W stored [K][N]
    d16 = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16)
    _lib.gemm_bf16(dY, W, b_mn=True, dact_src=g16, dact=3, out_bf16=d16)
    torch.cuda.synchronize()
    ref = (dY.float() @ W.float()) * g16.float()
    assert (d16.float() - ref).abs().max().item() < 2e-3 + ref.abs().max().item() * 2 ** -7


@pytest.mark.parametrize("kind", ["res", "f32", "acc", "bf16"])
@pytest.mark.parametrize("M,N,K", [(3168, 768, 768), (2048, 2304, 768), (100, 202, 192)])
def test_gemm_specialised_epilogues(kind, M, N, K):
    """Every fused epilogue kind on ragged and model-sized shapes, automatic N-tile choice."""
    from mkg_analogy_b200 import _lib
    A, B = _mk((M, K), 31), _mk((N, K), 32)
    g = torch.Generator().manual_seed(33)
    bias = torch.randn(N, generator=g).cuda()
    ref = A.float() @ B.float().t()
    if kind == "res":
        res = torch.randn(M, N, generator=g).cuda()
        out = torch.full((M, N), float("nan"), device="cuda")
        o16 = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16)
        _lib.gemm_bf16(A, B, bias=bias, residual=res, out_f32=out, out_bf16=o16)
        _check(out, ref + bias + res, K)
        _check(o16, ref + bias + res, K, bf16_out=True)
    elif kind == "f32":
        out = torch.full((M, N), float("nan"), device="cuda")
        _lib.gemm_bf16(A, B, bias=bias, out_f32=out)
        _check(out, ref + bias, K)
    elif kind == "acc":
        out = torch.ones((M, N), device="cuda")
        _lib.gemm_bf16(A, B, out_f32=out, accumulate=True, split_k=3)
        _check(out, ref + 1.0, K)
    else:
        o16 = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16)
        _lib.gemm_bf16(A, B, bias=bias, out_bf16=o16)
        _check(o16, ref + bias, K, bf16_out=True)
    torch.cuda.synchronize()


@pytest.mark.parametrize("dact", [1, 2])
def test_gemm_epilogue_dact(dact):
    from mkg_analogy_b200 import _lib
    M, N, K = 256, 256, 128
    A, B = _mk((M, K), 14), _mk((K, N), 15)
    x = _mk((M, N), 16)
    o16 = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16)
    _lib.gemm_bf16(A, B, b_mn=True, dact_src=x, dact=dact, out_bf16=o16)
    torch.cuda.synchronize()
    xf = x.float().requires_grad_(True)
    y = torch.nn.functional.gelu(xf) if dact == 1 else xf * torch.sigmoid(1.702 * xf)
    (gx,) = torch.autograd.grad(y.sum(), xf)
    ref = (A.float() @ B.float()) * gx
    assert (o16.float() - ref).abs().max().item() < 2e-3 + ref.abs().max().item() * 2 ** -7
