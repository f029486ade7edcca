"""Diagnostic (not a test): per-64x64-block error map of the tcgen05 GEMM for each operand layout."""
import sys
import torch
from mkg_analogy_b200 import _lib


def probe(name, M, N, K, a_mn, b_mn, bn):
    g = torch.Generator().manual_seed(0)
    A = (torch.randn((K, M) if a_mn else (M, K), generator=g)).to(torch.bfloat16).cuda()
    B = (torch.randn((K, N) if b_mn else (N, K), generator=g)).to(torch.bfloat16).cuda()
    out = torch.full((M, N), float("nan"), device="cuda")
    try:
        _lib.gemm_bf16(A, B, a_mn=a_mn, b_mn=b_mn, out_f32=out, bn=bn)
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        print(name, "EXC", repr(e)[:300])
        return False
    Af = A.float().t() if a_mn else A.float()
    Bf = B.float() if b_mn else B.float().t()
    ref = Af @ Bf
    err = (out - ref).abs()
    err = torch.nan_to_num(err, nan=1e9)
    ok = err.max().item() < 1e-2 * K ** 0.5
    print(f"{name} M{M} N{N} K{K} a_mn={int(a_mn)} b_mn={int(b_mn)} bn={bn}: max_err={err.max().item():.4g} {'OK' if ok else 'WRONG'}")
    if not ok:
        mb, nb = (M + 63) // 64, (N + 63) // 64
        for i in range(mb):
            row = []
            for j in range(nb):
                e = err[i * 64:(i + 1) * 64, j * 64:(j + 1) * 64].max().item()
                row.append("." if e < 1e-2 * K ** 0.5 else "X")
            print("   ", "".join(row))
    return ok


if __name__ == "__main__":
    which = sys.argv[1]
    cfg = {"nt": (False, False), "nn": (False, True), "tn": (True, True), "tk": (True, False)}[which]
    for (M, N, K) in [(128, 128, 64), (128, 128, 256), (256, 256, 512), (200, 136, 100)]:
        for bn in (64, 128, 256):
            probe(which, M, N, K, cfg[0], cfg[1], bn)
