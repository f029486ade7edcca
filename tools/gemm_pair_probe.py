"""Bring-up probe of the CTA-pair (cta_group::2) GEMM: correctness per layout / epilogue, then timing vs 1-CTA tiles."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkg_analogy_b200 import _lib


def mk(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * 0.5).to(torch.bfloat16).cuda()


def check(name, M, N, K, a_mn, b_mn, kind):
    A = mk((K, M) if a_mn else (M, K), 1); B = mk((K, N) if b_mn else (N, K), 2)
    Af = A.float().t() if a_mn else A.float(); Bf = B.float() if b_mn else B.float().t()
    ref = Af @ Bf
    kw = {}
    if kind == "f32":
        out = torch.full((M, N), float("nan"), device="cuda"); kw["out_f32"] = out
    elif kind == "bf16":
        out = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16); kw["out_bf16"] = out; kw["bias"] = torch.zeros(N, device="cuda")
    elif kind == "acc":
        out = torch.ones((M, N), device="cuda"); kw["out_f32"] = out; kw["accumulate"] = True; kw["split_k"] = 2; ref = ref + 1
    elif kind == "res":
        res = torch.randn(M, N, device="cuda"); out = torch.full((M, N), float("nan"), device="cuda")
        kw["out_f32"] = out; kw["residual"] = res; kw["bias"] = torch.zeros(N, device="cuda"); ref = ref + res
    try:
        _lib.gemm_bf16(A, B, a_mn=bool(a_mn), b_mn=bool(b_mn), bn=512, **kw)
        torch.cuda.synchronize()
    except Exception as e:
        print(name, "EXC", repr(e)[:300]); return False
    err = torch.nan_to_num((out.float() - ref).abs(), nan=1e9)
    tol = 1e-2 * K ** 0.5 if kind == "bf16" else 2e-3 * K ** 0.5
    ok = err.max().item() < tol
    print(f"{name} M{M} N{N} K{K} a_mn={a_mn} b_mn={b_mn} {kind}: max_err={err.max().item():.4g} {'OK' if ok else 'WRONG'}", flush=True)
    if not ok:
        for i in range((M + 127) // 128):
            print("   ", "".join("." if err[i * 128:(i + 1) * 128, j * 64:(j + 1) * 64].max().item() < tol else "X" for j in range((N + 63) // 64)))
    return ok


def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "check"
    if which == "check":
        ok = True
        for (M, N, K) in [(256, 256, 64), (256, 256, 256), (512, 512, 768), (3168, 768, 768), (200, 300, 192)]:
            ok &= check("nt", M, N, K, 0, 0, "f32")
        ok &= check("nt", 3168, 2304, 768, 0, 0, "bf16")
        ok &= check("nt", 2048, 768, 3072, 0, 0, "res")
        ok &= check("nn", 3168, 768, 3072, 0, 1, "f32")
        ok &= check("tn", 3072, 768, 3168, 1, 1, "acc")
        print("ALL OK" if ok else "FAILURES")
    else:
        for (M, N, K) in [(8192, 8192, 8192), (3168, 3072, 768), (3168, 768, 3072), (3168, 2304, 768), (2048, 3072, 768), (2048, 768, 768)]:
            A = mk((M, K), 1); B = mk((N, K), 2)
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); bias = torch.zeros(N, device="cuda")
            r = []
            for bn in (128, 256, 512):
                t = timeit(lambda: _lib.gemm_bf16(A, B, bias=bias, out_bf16=out, bn=bn))
                r.append(f"bn={bn}: {t * 1e3:7.1f} us {2.0 * M * N * K / t / 1e9:6.0f} TF/s")
            t = timeit(lambda: torch.matmul(A, B.t()))
            print(f"M={M} N={N} K={K}  " + "  ".join(r) + f"  | cuBLAS {t * 1e3:7.1f} us {2.0 * M * N * K / t / 1e9:6.0f} TF/s", flush=True)
