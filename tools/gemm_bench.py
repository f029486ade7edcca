"""Times the tcgen05 GEMM on the shapes of the MKGformer-base step (B=32, L=64) for every N-tile."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkg_analogy_b200 import _lib

Mt, Mv, H, F = 2048, 3168, 768, 3072
# (name, M, N, K, a_mn, b_mn, epilogue kind, split)
shapes = [
    ("v.qkv fwd", Mv, 3 * H, H, 0, 0, "bf16", 1), ("v.out fwd", Mv, H, H, 0, 0, "res", 1),
    ("v.fc1 fwd", Mv, F, H, 0, 0, "act", 1), ("v.fc2 fwd", Mv, H, F, 0, 0, "res", 1),
    ("t.qkv fwd", Mt, 3 * H, H, 0, 0, "bf16", 1), ("t.out fwd", Mt, H, H, 0, 0, "res", 1),
    ("t.int fwd", Mt, F, H, 0, 0, "act", 1), ("t.o fwd", Mt, H, F, 0, 0, "res", 1),
    ("v.fc2 dX", Mv, F, H, 0, 1, "dact", 1), ("v.fc1 dX", Mv, H, F, 0, 1, "f32", 1), ("v.qkv dX", Mv, H, 3 * H, 0, 1, "f32", 1),
    ("v.out dX", Mv, H, H, 0, 1, "f32", 1),
    ("t.o dX", Mt, F, H, 0, 1, "dact", 1), ("t.int dX", Mt, H, F, 0, 1, "res", 1),
    ("v.fc2 dW", H, F, Mv, 1, 1, "acc", 1), ("v.fc1 dW", F, H, Mv, 1, 1, "acc", 1), ("v.qkv dW", 3 * H, H, Mv, 1, 1, "acc", 2),
    ("v.out dW", H, H, Mv, 1, 1, "acc", 4), ("t.qkv dW", 3 * H, H, Mt, 1, 1, "acc", 2), ("t.out dW", H, H, Mt, 1, 1, "acc", 4),
    ("v.fc1 dW s2", F, H, Mv, 1, 1, "acc", 2), ("v.fc1 dW s4", F, H, Mv, 1, 1, "acc", 4), ("v.qkv dW s4", 3 * H, H, Mv, 1, 1, "acc", 4),
    ("v.out dW s8", H, H, Mv, 1, 1, "acc", 8), ("t.int dW s2", F, H, Mt, 1, 1, "acc", 2),
]


def run(name, M, N, K, a_mn, b_mn, kind, split, bn, iters=20):
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn((K, M) if a_mn else (M, K), device="cuda", generator=g).to(torch.bfloat16)
    B = torch.randn((K, N) if b_mn else (N, K), device="cuda", generator=g).to(torch.bfloat16)
    kw = {}
    if kind in ("bf16", "act", "dact"):
        kw["out_bf16"] = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        kw["bias"] = torch.zeros(N, device="cuda")
    if kind == "act":
        kw["act"] = 2; kw["out_preact"] = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    if kind == "dact":
        kw["dact"] = 3; kw["dact_src"] = torch.randn(M, N, device="cuda").to(torch.bfloat16); kw.pop("bias")
    if kind in ("res", "f32", "acc"):
        kw["out_f32"] = torch.zeros(M, N, device="cuda")
    if kind == "res":
        kw["residual"] = torch.randn(M, N, device="cuda"); kw["bias"] = torch.zeros(N, device="cuda")
    if kind == "acc":
        kw["accumulate"] = True; kw["split_k"] = split
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        _lib.gemm_bf16(A, B, a_mn=bool(a_mn), b_mn=bool(b_mn), bn=bn, **kw)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.gemm_bf16(A, B, a_mn=bool(a_mn), b_mn=bool(b_mn), bn=bn, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    t = ts[len(ts) // 2]
    return t * 1e3, 2.0 * M * N * K / (t * 1e-3) / 1e12


if __name__ == "__main__":
    print(f"{'shape':12s} {'M':>5s} {'N':>5s} {'K':>5s}  " + "  ".join(f"bn={b:3d} us / TF" for b in (64, 128, 256, 0)))
    if "--few" in sys.argv:
        shapes = [s for s in shapes if "dW" in s[0] or s[0] in ("t.qkv fwd", "v.out fwd")]
    for sh in shapes:
        res = [run(*sh, bn) for bn in (64, 128, 256, 0)]
        print(f"{sh[0]:12s} {sh[1]:5d} {sh[2]:5d} {sh[3]:5d}  " + "  ".join(f"{us:8.1f} / {tf:5.0f}" for us, tf in res), flush=True)
