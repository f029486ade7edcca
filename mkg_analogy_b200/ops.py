"""Stand-alone operators of the path (thin wrappers of the C ABI; CUDA tensors in, CUDA tensors out).
Used by the kernel-level parity tests; the model path calls the same kernels from inside the library."""
import ctypes as C

import torch

from . import _lib
from ._lib import ptr


def attention_fwd(q, k, v, heads, scale, *, prefix_k=None, prefix_v=None, key_mask=None, sep_idx=None, w0=None, w1=None,
                  want_f32=False):
    """q: (B, Lq, heads*64) bf16 (may be a strided view with unit inner stride), k/v: (B, n, heads*64).
    Returns (ctx bf16 (B, Lq, heads*64), ctx fp32 or None, lse (B, heads, Lq))."""
    B, Lq, HD = q.shape
    n = k.shape[1]
    d = _lib.AttnDesc()
    d.B, d.Lq, d.heads, d.n_keys, d.scale = B, Lq, heads, n, scale
    d.q, d.ldq = q.data_ptr(), q.stride(1)
    d.k, d.v, d.ldkv = k.data_ptr(), v.data_ptr(), k.stride(1)
    if prefix_k is not None:
        d.prefix_k, d.prefix_v, d.ld_prefix, d.n_prefix = prefix_k.data_ptr(), prefix_v.data_ptr(), prefix_k.stride(1), prefix_k.shape[1]
    d.key_mask, d.sep_idx, d.w0, d.w1 = ptr(key_mask), ptr(sep_idx), ptr(w0), ptr(w1)
    out = torch.empty(B, Lq, HD, device=q.device, dtype=torch.bfloat16)
    out32 = torch.empty(B, Lq, HD, device=q.device) if want_f32 else None
    lse = torch.empty(B, heads, Lq, device=q.device)
    d.out, d.ldo, d.out_f32, d.lse = out.data_ptr(), HD, ptr(out32), lse.data_ptr()
    _lib.check(_lib.lib().b200mart_attention_fwd(C.byref(d), _lib.cur_stream()))
    return out, out32, lse


def attention_bwd(dout, out, q, k, v, lse, heads, scale, *, prefix_k=None, prefix_v=None, key_mask=None, sep_idx=None, w0=None,
                  w1=None, dk_add=None, dv_add=None):
    """Returns dq, dk, dv (bf16), dprefix_k, dprefix_v (fp32 or None), dw0, dw1 (fp32 scalars or None)."""
    B, Lq, HD = q.shape
    n = k.shape[1]
    d = _lib.AttnDesc()
    d.B, d.Lq, d.heads, d.n_keys, d.scale = B, Lq, heads, n, scale
    d.q, d.ldq = q.data_ptr(), q.stride(1)
    d.k, d.v, d.ldkv = k.data_ptr(), v.data_ptr(), k.stride(1)
    dpk = dpv = None
    if prefix_k is not None:
        d.prefix_k, d.prefix_v, d.ld_prefix, d.n_prefix = prefix_k.data_ptr(), prefix_v.data_ptr(), prefix_k.stride(1), prefix_k.shape[1]
        dpk = torch.empty(B, prefix_k.shape[1], HD, device=q.device)
        dpv = torch.empty_like(dpk)
        d.dprefix_k, d.dprefix_v, d.ldd_prefix = dpk.data_ptr(), dpv.data_ptr(), HD
    d.key_mask, d.sep_idx, d.w0, d.w1 = ptr(key_mask), ptr(sep_idx), ptr(w0), ptr(w1)
    d.lse = lse.data_ptr()
    d.out, d.ldo = out.data_ptr(), HD      # forward context (bf16): delta = rowsum(dO * O)
    dout = dout.contiguous()
    d.dout, d.lddo = dout.data_ptr(), HD
    dq = torch.empty(B, Lq, HD, device=q.device, dtype=torch.bfloat16)
    dk = torch.empty(B, n, HD, device=q.device, dtype=torch.bfloat16)
    dv = torch.empty_like(dk)
    d.dq, d.dk, d.dv, d.ldd = dq.data_ptr(), dk.data_ptr(), dv.data_ptr(), HD
    if dk_add is not None:
        d.dk_add, d.dv_add, d.ldadd = dk_add.data_ptr(), dv_add.data_ptr(), HD
    dw0 = dw1 = None
    if sep_idx is not None:
        dw0 = torch.zeros(1, device=q.device)
        dw1 = torch.zeros(1, device=q.device)
        d.dw0, d.dw1 = dw0.data_ptr(), dw1.data_ptr()
    delta = torch.empty(B, heads, Lq, device=q.device)
    d.delta = delta.data_ptr()
    _lib.check(_lib.lib().b200mart_attention_bwd(C.byref(d), _lib.cur_stream()))
    return dq, dk, dv, dpk, dpv, dw0, dw1


def fusion_fwd(ctx, vh):
    """BertFusion: ctx (B, L, H) fp32, vh (B, Nv, H) fp32 -> (out bf16 (B, L, H), lse (B, L))."""
    B, L, H = ctx.shape
    Nv = vh.shape[1]
    out = torch.empty(B, L, H, device=ctx.device, dtype=torch.bfloat16)
    lse = torch.empty(B, L, device=ctx.device)
    _lib.check(_lib.lib().b200mart_fusion_fwd(ctx.data_ptr(), vh.data_ptr(), out.data_ptr(), lse.data_ptr(), B, L, Nv, H,
                                              _lib.cur_stream()))
    return out, lse


def fusion_bwd(dg, ctx, vh, lse, dvh_in=None):
    """Returns (dctx, dvh) fp32."""
    B, L, H = ctx.shape
    Nv = vh.shape[1]
    pbuf = torch.empty(B, L, Nv, device=ctx.device)
    dsbuf = torch.empty_like(pbuf)
    dctx = torch.zeros_like(ctx)
    dvh = torch.empty_like(vh)
    _lib.check(_lib.lib().b200mart_fusion_bwd(dg.data_ptr(), ctx.data_ptr(), vh.data_ptr(), lse.data_ptr(), pbuf.data_ptr(),
                                              dsbuf.data_ptr(), dctx.data_ptr(), ptr(dvh_in), dvh.data_ptr(), B, L, Nv, H,
                                              _lib.cur_stream()))
    return dctx, dvh


def layernorm_fwd(x, gamma, beta, eps):
    M, H = x.shape
    y32 = torch.empty_like(x)
    y16 = torch.empty(M, H, device=x.device, dtype=torch.bfloat16)
    mean = torch.empty(M, device=x.device)
    rstd = torch.empty(M, device=x.device)
    _lib.check(_lib.lib().b200mart_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, y32.data_ptr(),
                                                 y16.data_ptr(), mean.data_ptr(), rstd.data_ptr(), M, H, _lib.cur_stream()))
    return y32, y16, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, add=None):
    M, H = x.shape
    dx = torch.empty_like(x)
    dgamma = torch.zeros(H, device=x.device)
    dbeta = torch.zeros(H, device=x.device)
    _lib.check(_lib.lib().b200mart_layernorm_bwd(dy.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                                 ptr(add), dx.data_ptr(), None, dgamma.data_ptr(), dbeta.data_ptr(), M, H,
                                                 _lib.cur_stream()))
    return dx, dgamma, dbeta


def ls_ce(logits, label, eps=0.1):
    """LabelSmoothSoftmaxCEV1 + gradient: returns (loss (1,), dlogits (B, C))."""
    B, Cn = logits.shape
    loss = torch.zeros(1, device=logits.device)
    dl = torch.empty_like(logits)
    _lib.check(_lib.lib().b200mart_ls_ce(logits.data_ptr(), B, Cn, label.data_ptr(), eps, loss.data_ptr(), dl.data_ptr(),
                                         _lib.cur_stream()))
    return loss, dl
