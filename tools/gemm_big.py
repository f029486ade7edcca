"""Steady-state throughput of the tcgen05 GEMM on a large problem (overheads amortised), next to torch.matmul (cuBLAS)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkg_analogy_b200 import _lib

def timeit(fn, iters=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for (M, N, K) in [(8192, 8192, 8192), (4096, 4096, 4096), (3168, 3072, 768), (3168, 768, 3072), (16384, 3072, 768)]:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16); B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    res = []
    for bn in (128, 256):
        t = timeit(lambda: _lib.gemm_bf16(A, B, bias=bias, out_bf16=out, bn=bn))
        res.append(f"bn={bn}: {t*1e3:8.1f} us {2.0*M*N*K/t/1e9:7.0f} TF/s")
    t = timeit(lambda: torch.matmul(A, B.t()))
    print(f"M={M} N={N} K={K}  " + "  ".join(res) + f"  | cuBLAS {t*1e3:8.1f} us {2.0*M*N*K/t/1e9:7.0f} TF/s", flush=True)
