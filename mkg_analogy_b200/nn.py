"""The dense plumbing around the encoder stacks of the FLAVA / ViLT drop-ins on the library's own kernels:
`nn.Linear`, the patch-embedding convolution (kernel = stride = patch: an im2col view + one product), `nn.LayerNorm` and the
`[MASK]`-row score head of the fused route, each as a `torch.autograd.Function` over the C ABI — tcgen05 GEMMs with bf16
operands / fp32 accumulation (`b200mart_gemm_bf16`: forward, dX and dW), `b200mart_layernorm_fwd/bwd`, `b200mart_ls_ce`
(LabelSmoothSoftmaxCEV1 with its softmax-backward fused).  No cuBLAS / cuDNN call is left on the CUDA path of the two families;
what stays in PyTorch are index gathers (`nn.Embedding`), concatenations, casts and the GELU between the two head products.

On a CPU device (parameter storage, and the plumbing tests of the CPU suite, which inject the oracle's encoder) the same
functions evaluate with the plain PyTorch expressions they replace; the product path is CUDA.
"""
import torch
import torch.nn.functional as F

from . import _lib, ops


def _bf16(t):
    return t.to(torch.bfloat16).contiguous()


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b.  x (M, K) fp32, W (N, K) fp32, K % 8 == 0 (TMA rows of 16 bytes)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        M, K = x.shape
        N = weight.shape[0]
        x16, w16 = _bf16(x), _bf16(weight)
        y = torch.empty(M, N, device=x.device, dtype=torch.float32)
        _lib.gemm_bf16(x16, w16, bias=None if bias is None else bias.float().contiguous(), out_f32=y)
        ctx.save_for_backward(x16, w16)
        ctx.has_bias = bias is not None
        ctx.needs = (x.requires_grad, weight.requires_grad)
        return y

    @staticmethod
    def backward(ctx, dy):
        x16, w16 = ctx.saved_tensors
        M, K = x16.shape
        N = w16.shape[0]
        dy = dy.contiguous()
        Np = (N + 7) // 8 * 8                               # MN-major / K-major rows of dY must be multiples of 16 bytes
        if Np != N:
            dy16 = torch.zeros(M, Np, device=dy.device, dtype=torch.bfloat16)
            dy16[:, :N] = dy
        else:
            dy16 = _bf16(dy)
        dx = dw = db = None
        if ctx.needs[0]:                                    # dX (M, K) = dY (M, N) W (N, K): B operand stored [K_c = N][N_out = K]
            dx = torch.empty(M, K, device=dy.device, dtype=torch.float32)
            wp = w16
            if Np != N:
                wp = torch.zeros(Np, K, device=dy.device, dtype=torch.bfloat16)
                wp[:N] = w16
            _lib.gemm_bf16(dy16, wp, b_mn=True, out_f32=dx)
        if ctx.needs[1]:                                    # dW (N, K) = dY^T X: A stored [K_c = M][M_out = N], B stored [K_c = M][N_out = K]
            dwp = torch.empty(Np, K, device=dy.device, dtype=torch.float32)
            _lib.gemm_bf16(dy16, x16, a_mn=True, b_mn=True, out_f32=dwp)
            dw = dwp[:N]
        if ctx.has_bias:
            db = dy.sum(0)
        return dx, dw, db


def linear(x, weight, bias=None):
    """F.linear on the tcgen05 GEMM (any leading shape)."""
    if not x.is_cuda:
        return F.linear(x, weight, bias)
    K = x.shape[-1]
    if K % 8 != 0:
        raise _lib.B200MartError(f"linear: inner dimension {K} must be a multiple of 8")
    y = _LinearFn.apply(x.reshape(-1, K).float(), weight, bias)
    return y.reshape(*x.shape[:-1], weight.shape[0])


def patch_embed(pixels, conv):
    """`nn.Conv2d(C, hidden, kernel_size = stride = patch)` (CLIP / FLAVA / ViLT patch embeddings) as im2col + one product:
    (B, C, S, S) -> (B, (S / patch)^2, hidden), patches in raster order."""
    if not pixels.is_cuda:
        return conv(pixels).flatten(2).transpose(1, 2)
    p = conv.kernel_size[0]
    B, Cc, S, _ = pixels.shape
    g = S // p
    cols = pixels.reshape(B, Cc, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B * g * g, Cc * p * p)
    y = linear(cols, conv.weight.reshape(conv.out_channels, -1), conv.bias)
    return y.reshape(B, g * g, conv.out_channels)


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        y32, _, mean, rstd = ops.layernorm_fwd(x, weight, bias, eps)
        ctx.save_for_backward(x, mean, rstd, weight)
        return y32

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, weight = ctx.saved_tensors
        dx, dgamma, dbeta = ops.layernorm_bwd(dy.contiguous(), x, mean, rstd, weight)
        return dx, dgamma, dbeta, None


def layer_norm(x, ln):
    """`nn.LayerNorm` module `ln` applied with the library's row kernels (hidden % 128 == 0)."""
    if not x.is_cuda:
        return ln(x)
    H = x.shape[-1]
    y = _LayerNormFn.apply(x.reshape(-1, H).float().contiguous(), ln.weight.float().contiguous(), ln.bias.float().contiguous(), float(ln.eps))
    return y.reshape(x.shape)


class _LsCeFn(torch.autograd.Function):
    """LabelSmoothSoftmaxCEV1 (MarT/lit_models/utils.py:42-66) with the gradient produced by the same kernel."""

    @staticmethod
    def forward(ctx, logits, label, eps):
        loss, dl = ops.ls_ce(logits.contiguous(), label.contiguous(), eps)
        ctx.save_for_backward(dl)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dl,) = ctx.saved_tensors
        return dl * g, None, None


def mask_row_loss(trans, mask_pos, table, bias, columns, label, label_smoothing=0.1):
    """The fused route's score head: logits of the [MASK] rows over `columns` (an index tensor, or a
    (start, end) range of the tied table) = trans[b, mask_pos[b]] . table[columns]^T + bias[columns], then LS-CE."""
    B = trans.shape[0]
    zm = trans[torch.arange(B, device=trans.device), mask_pos]
    if isinstance(columns, tuple):
        E, bsel = table[columns[0]:columns[1]], bias[columns[0]:columns[1]]
    else:
        E, bsel = table[columns], bias[columns]
    logits = linear(zm, E, bsel)
    if not logits.is_cuda:
        lp = F.log_softmax(logits.float(), dim=-1)
        q = torch.full_like(lp, label_smoothing / lp.shape[-1]).scatter_(1, label[:, None], 1.0 - label_smoothing)
        return -(q * lp).sum(-1).mean(), logits
    return _LsCeFn.apply(logits, label, label_smoothing), logits
