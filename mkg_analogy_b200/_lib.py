"""ctypes binding of libb200mart.so (the C ABI declared in include/b200mart.h).

There is deliberately NO fallback: if the shared library is missing or a symbol cannot be
resolved, importing the product path raises.  The CPU oracle under oracle/ is test infrastructure
and is never reachable from here.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200mart.so")


class B200MartError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("K2", C.c_int32),
        ("a_mn", C.c_int32), ("b_mn", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64),
        ("B", C.c_void_p), ("ldb", C.c_int64),
        ("A2", C.c_void_p), ("lda2", C.c_int64),
        ("B2", C.c_void_p), ("ldb2", C.c_int64),
        ("alpha", C.c_float),
        ("bias", C.c_void_p),
        ("out_preact", C.c_void_p), ("ldp", C.c_int64),
        ("act", C.c_int32),
        ("dact_src", C.c_void_p), ("ldx", C.c_int64), ("dact", C.c_int32),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("out_f32", C.c_void_p), ("ldo32", C.c_int64),
        ("out_bf16", C.c_void_p), ("ldo16", C.c_int64),
        ("accumulate", C.c_int32), ("split_k", C.c_int32), ("bn", C.c_int32),
    ]


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
    l.b200mart_last_error.restype = C.c_char_p
    l.b200mart_version.restype = C.c_int32
    l.b200mart_gemm_bf16.restype = C.c_int32
    l.b200mart_gemm_bf16.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
    _lib = l
    return l


def check(rc):
    if rc != 0:
        raise B200MartError(f"b200mart error {rc}: {lib().b200mart_last_error().decode()}")


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def gemm_bf16(A, B, *, a_mn=False, b_mn=False, A2=None, B2=None, alpha=1.0, bias=None, act=0,
              out_preact=None, dact_src=None, dact=0, residual=None, out_f32=None, out_bf16=None,
              accumulate=False, split_k=1, bn=0, stream=None):
    """D = epilogue(A @ B^T) on the tcgen05 kernel. Shapes: A [M,K] (or [K,M] when a_mn), B [N,K] (or [K,N])."""
    import torch
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
    d.bias = bias.data_ptr() if bias is not None else None
    if out_preact is not None:
        d.out_preact, d.ldp = out_preact.data_ptr(), out_preact.stride(0)
    d.act = act
    if dact_src is not None:
        d.dact_src, d.ldx, d.dact = dact_src.data_ptr(), dact_src.stride(0), dact
    if residual is not None:
        d.residual, d.ldr = residual.data_ptr(), residual.stride(0)
    if out_f32 is not None:
        d.out_f32, d.ldo32 = out_f32.data_ptr(), out_f32.stride(0)
    if out_bf16 is not None:
        d.out_bf16, d.ldo16 = out_bf16.data_ptr(), out_bf16.stride(0)
    d.accumulate, d.split_k, d.bn = int(accumulate), split_k, bn
    s = stream if stream is not None else torch.cuda.current_stream().cuda_stream
    check(lib().b200mart_gemm_bf16(C.byref(d), C.c_void_p(s)))
