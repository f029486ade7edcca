"""ctypes binding of libb200mart.so (the C ABI declared in include/b200mart.h).

There is deliberately NO fallback: if the shared library is missing or a symbol cannot be
resolved, importing the product path raises.  The CPU oracle under oracle/ is test infrastructure
and is never reachable from here.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200mart.so")

vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float


class B200MartError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", i32), ("N", i32), ("K", i32), ("K2", i32),
        ("a_mn", i32), ("b_mn", i32),
        ("A", vp), ("lda", i64),
        ("B", vp), ("ldb", i64),
        ("A2", vp), ("lda2", i64),
        ("B2", vp), ("ldb2", i64),
        ("alpha", f32),
        ("bias", vp),
        ("bias2", vp),
        ("out_preact", vp), ("ldp", i64),
        ("act", i32), ("save_grad", i32),
        ("dact_src", vp), ("ldx", i64), ("dact", i32),
        ("residual", vp), ("ldr", i64),
        ("out_f32", vp), ("ldo32", i64),
        ("out_bf16", vp), ("ldo16", i64),
        ("accumulate", i32), ("split_k", i32), ("bn", i32),
    ]


class ModelDesc(C.Structure):
    _fields_ = [
        ("hidden", i32), ("layers", i32), ("heads", i32), ("intermediate", i32),
        ("vocab", i32), ("max_position", i32), ("type_vocab", i32),
        ("image_size", i32), ("patch", i32), ("fuse_from", i32),
        ("text_eps", f32), ("vision_eps", f32),
    ]


class Buffers(C.Structure):
    _fields_ = [("params32", vp), ("params16", vp), ("grads32", vp), ("workspace", vp), ("workspace_bytes", i64)]


class Batch(C.Structure):
    _fields_ = [("B", i32), ("L", i32), ("input_ids", vp), ("attention_mask", vp), ("token_type_ids", vp),
                ("pixel_values", vp), ("sep_idx", vp)]


class Head(C.Structure):
    _fields_ = [("mode", i32), ("col_ids", vp), ("n_cols", i32),
                ("ent_st", i32), ("ent_ed", i32), ("rel_st", i32), ("rel_ed", i32),
                ("label", vp), ("pre_type", vp), ("rel_idx", vp), ("q_head_idx", vp), ("a_head_idx", vp),
                ("label_smoothing", f32), ("alpha", f32), ("mask_token_id", i32), ("grad_scale", vp)]


class StackDesc(C.Structure):
    _fields_ = [("hidden", i32), ("layers", i32), ("heads", i32), ("intermediate", i32), ("act", i32), ("eps", f32), ("post_ln", i32)]


class StackIO(C.Structure):
    _fields_ = [("B", i32), ("T", i32), ("x", vp), ("key_mask", vp), ("split", vp), ("row_lo", i32), ("y", vp),
                ("p_hidden", f32), ("p_attn", f32), ("seed", C.c_uint32)]


class AttnDesc(C.Structure):
    _fields_ = [
        ("B", i32), ("Lq", i32), ("heads", i32), ("n_prefix", i32), ("n_keys", i32),
        ("scale", f32),
        ("q", vp), ("ldq", i64),
        ("prefix_k", vp), ("prefix_v", vp), ("ld_prefix", i64),
        ("k", vp), ("v", vp), ("ldkv", i64),
        ("key_mask", vp), ("sep_idx", vp), ("w0", vp), ("w1", vp),
        ("out", vp), ("ldo", i64), ("out_f32", vp), ("lse", vp),
        ("dout", vp), ("lddo", i64),
        ("dq", vp), ("dk", vp), ("dv", vp), ("ldd", i64),
        ("dk_add", vp), ("dv_add", vp), ("ldadd", i64),
        ("dprefix_k", vp), ("dprefix_v", vp), ("ldd_prefix", i64),
        ("dw0", vp), ("dw1", vp), ("delta", vp),
    ]


# every symbol include/b200mart.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "b200mart_last_error": (C.c_char_p, []),
    "b200mart_version": (i32, []),
    "b200mart_launch_count": (i64, []),
    "b200mart_profile_begin": (i32, [vp]),
    "b200mart_profile_end": (i32, [vp, vp, C.POINTER(C.c_double * 3), C.POINTER(i64 * 3), C.POINTER(C.c_double)]),
    "b200mart_profile_gemm_bytes": (C.c_double, [vp]),
    "b200mart_gemm_timeline": (i64, [vp, i64]),
    "b200mart_gemm_bf16": (i32, [C.POINTER(GemmDesc), vp]),
    "b200mart_create": (i32, [C.POINTER(ModelDesc), C.POINTER(vp)]),
    "b200mart_destroy": (None, [vp]),
    "b200mart_param_count": (i32, [vp]),
    "b200mart_param_total": (i64, [vp]),
    "b200mart_param_info": (i32, [vp, i32, C.POINTER(C.c_char_p), C.POINTER(i64), C.POINTER(i32), C.POINTER(i64 * 4),
                                  C.POINTER(i32), C.POINTER(i32)]),
    "b200mart_cast_params": (i32, [vp, vp, vp, vp]),
    "b200mart_cast_params_range": (i32, [vp, vp, vp, i64, i64, vp]),
    "b200mart_workspace_bytes": (i64, [vp, i32, i32]),
    "b200mart_forward": (i32, [vp, C.POINTER(Buffers), C.POINTER(Batch), vp, vp, vp]),
    "b200mart_head_loss": (i32, [vp, C.POINTER(Buffers), C.POINTER(Batch), C.POINTER(Head), vp, vp, vp, i32, vp]),
    "b200mart_backward": (i32, [vp, C.POINTER(Buffers), C.POINTER(Batch), i32, vp, vp, vp, C.POINTER(vp), vp]),
    "b200mart_layer_extent": (i32, [vp, i32, C.POINTER(i64), C.POINTER(i64)]),
    "b200mart_set_option": (i32, [vp, C.c_char_p, i32]),
    "b200mart_set_dropout": (i32, [vp, f32, f32, C.c_uint32]),
    "b200mart_last_dropout_seed": (C.c_uint32, [vp]),
    "b200mart_adamw": (i32, [vp, vp, vp, vp, vp, vp, i32, f32, f32, f32, f32, f32, i32, i64, i64, i32, vp]),
    "b200mart_status": (i32, [vp, C.POINTER(Buffers), vp, C.POINTER(i32)]),
    "b200mart_read_hidden": (i32, [vp, C.POINTER(Buffers), i32, i32, vp, vp]),
    "b200mart_multimem_add": (i32, [vp, vp, i64, f32, vp]),
    "b200mart_set_step_block": (i32, [vp, vp]),
    "b200mart_adamw_multicast": (i32, [vp, vp, vp, vp, vp, vp, vp, i32, f32, f32, f32, f32, f32, i32, i64, i64, i32, i32, vp]),
    "b200mart_stack_create": (i32, [C.POINTER(StackDesc), C.POINTER(vp)]),
    "b200mart_stack_destroy": (None, [vp]),
    "b200mart_stack_param_count": (i32, [vp]),
    "b200mart_stack_param_total": (i64, [vp]),
    "b200mart_stack_param_info": (i32, [vp, i32, C.POINTER(C.c_char_p), C.POINTER(i64), C.POINTER(i32), C.POINTER(i64 * 4)]),
    "b200mart_stack_workspace_bytes": (i64, [vp, i32, i32]),
    "b200mart_stack_forward": (i32, [vp, C.POINTER(Buffers), C.POINTER(StackIO), vp]),
    "b200mart_stack_backward": (i32, [vp, C.POINTER(Buffers), C.POINTER(StackIO), vp, vp, vp]),
    "b200mart_attention_fwd": (i32, [C.POINTER(AttnDesc), vp]),
    "b200mart_attention_bwd": (i32, [C.POINTER(AttnDesc), vp]),
    "b200mart_fusion_fwd": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "b200mart_fusion_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "b200mart_layernorm_fwd": (i32, [vp, vp, vp, f32, vp, vp, vp, vp, i32, i32, vp]),
    "b200mart_layernorm_bwd": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "b200mart_ls_ce": (i32, [vp, i32, i32, vp, f32, vp, vp, vp]),
    "b200mart_gather_images": (i32, [vp, i32, vp, vp, i32, i64, i64, vp, vp]),
}

_lib = None


def lib():
    """Load (once) and return the ctypes handle; raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200MartError(
            f"{LIB_PATH} not found: build it with `python -m mkg_analogy_b200.build` "
            "(there is no CPU or PyTorch fallback for this path)")
    l = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(l, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def check(rc):
    if rc != 0:
        raise B200MartError(f"b200mart error {rc}: {lib().b200mart_last_error().decode()}")


def ptr(t):
    """Device (or host) address of a tensor, None for None."""
    return None if t is None else t.data_ptr()


def cur_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def gemm_bf16(A, B, *, a_mn=False, b_mn=False, A2=None, B2=None, alpha=1.0, bias=None, bias2=None, act=0,
              out_preact=None, save_grad=False, dact_src=None, dact=0, residual=None, out_f32=None, out_bf16=None,
              accumulate=False, split_k=1, bn=0, stream=None):
    """D = epilogue(A @ B^T) on the tcgen05 kernel. Shapes: A [M,K] (or [K,M] when a_mn), B [N,K] (or [K,N])."""
    M, K = (A.shape[1], A.shape[0]) if a_mn else A.shape
    N = B.shape[1] if b_mn else B.shape[0]
    d = GemmDesc()
    d.M, d.N, d.K, d.K2 = M, N, K, 0
    d.a_mn, d.b_mn = int(a_mn), int(b_mn)
    d.A, d.lda, d.B, d.ldb = A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0)
    if A2 is not None:
        d.K2 = A2.shape[0] if a_mn else A2.shape[1]
        d.A2, d.lda2, d.B2, d.ldb2 = A2.data_ptr(), A2.stride(0), B2.data_ptr(), B2.stride(0)
    d.alpha = alpha
    d.bias = ptr(bias)
    d.bias2 = ptr(bias2)
    if out_preact is not None:
        d.out_preact, d.ldp = out_preact.data_ptr(), out_preact.stride(0)
    d.act = act
    d.save_grad = int(save_grad)
    if dact_src is not None:
        d.dact_src, d.ldx, d.dact = dact_src.data_ptr(), dact_src.stride(0), dact
    if residual is not None:
        d.residual, d.ldr = residual.data_ptr(), residual.stride(0)
    if out_f32 is not None:
        d.out_f32, d.ldo32 = out_f32.data_ptr(), out_f32.stride(0)
    if out_bf16 is not None:
        d.out_bf16, d.ldo16 = out_bf16.data_ptr(), out_bf16.stride(0)
    d.accumulate, d.split_k, d.bn = int(accumulate), split_k, bn
    s = stream if stream is not None else cur_stream()
    check(lib().b200mart_gemm_bf16(C.byref(d), s))
