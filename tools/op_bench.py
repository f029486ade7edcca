"""Isolated kernel durations (CUPTI through torch.profiler, warm L2, mean of 10) of the non-GEMM operators at MKGformer-base
step shapes.  Run with B200MART_PDL=0: under programmatic dependent launch a kernel starts while its predecessor drains and
its CUPTI duration includes that wait.
    B200MART_PDL=0 python tools/op_bench.py [tag]   -> gpurun_out/op_bench_<tag>.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkg_analogy_b200 import ops  # noqa: E402


def timeit(fn, n=10, warm=3):
    """Kernel durations from CUPTI (torch.profiler): {kernel: mean us} over n calls, plus their sum."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
    per = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA and "Memset" not in e.name and "Memcpy" not in e.name:
            name = e.name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0].split("<")[0].split("::")[-1][:40]
            per[name] = per.get(name, 0.0) + e.device_time / n
    return per


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    B, L, Nv, H, nh = 32, 64, 99, 768, 12
    lines = []

    def rnd(*shape, scale=1.0, dtype=torch.float32):
        return (torch.randn(*shape, device=dev, generator=g) * scale).to(dtype)

    # ---- fusion
    ctx, vh, dg = rnd(B, L, H, scale=0.5), rnd(B, Nv, H), rnd(B, L, H)
    out, lse = ops.fusion_fwd(ctx, vh)
    lines.append(("fusion_fwd", timeit(lambda: ops.fusion_fwd(ctx, vh))))
    lines.append(("fusion_bwd (rows+keys)", timeit(lambda: ops.fusion_bwd(dg, ctx, vh, lse))))
    # ---- attention: text (64 x 64, mask + adaptive scaling), vision (99 x 163 with prefix)
    bf = torch.bfloat16
    tq = rnd(B, L, 3 * H, dtype=bf)
    q, k, v = tq[..., :H], tq[..., H:2 * H], tq[..., 2 * H:]
    mask = torch.ones(B, L, dtype=torch.int64, device=dev)
    sep = torch.full((B,), 20, dtype=torch.int64, device=dev)
    w0, w1 = torch.full((1,), 0.3, device=dev), torch.full((1,), 0.5, device=dev)
    o, _, l2 = ops.attention_fwd(q, k, v, nh, 0.125, key_mask=mask, sep_idx=sep, w0=w0, w1=w1)
    do = rnd(B, L, H)
    lines.append(("attn_fwd text", timeit(lambda: ops.attention_fwd(q, k, v, nh, 0.125, key_mask=mask, sep_idx=sep, w0=w0, w1=w1))))
    lines.append(("attn_bwd text", timeit(lambda: ops.attention_bwd(do, o, q, k, v, l2, nh, 0.125, key_mask=mask, sep_idx=sep, w0=w0, w1=w1))))
    vq = rnd(B, Nv, 3 * H, dtype=bf)
    q2, k2, v2 = vq[..., :H], vq[..., H:2 * H], vq[..., 2 * H:]
    o2, _, l3 = ops.attention_fwd(q2, k2, v2, nh, 0.125, prefix_k=k, prefix_v=v)
    do2 = rnd(B, Nv, H)
    lines.append(("attn_fwd vision+prefix", timeit(lambda: ops.attention_fwd(q2, k2, v2, nh, 0.125, prefix_k=k, prefix_v=v))))
    lines.append(("attn_bwd vision+prefix", timeit(lambda: ops.attention_bwd(do2, o2, q2, k2, v2, l3, nh, 0.125, prefix_k=k, prefix_v=v))))
    o3, _, l4 = ops.attention_fwd(q2, k2, v2, nh, 0.125)
    lines.append(("attn_fwd vision", timeit(lambda: ops.attention_fwd(q2, k2, v2, nh, 0.125))))
    lines.append(("attn_bwd vision", timeit(lambda: ops.attention_bwd(do2, o3, q2, k2, v2, l4, nh, 0.125))))
    # ---- layernorm
    for M in (B * L, B * Nv):
        x, dy, add = rnd(M, H, scale=2.0), rnd(M, H), rnd(M, H)
        gamma, beta = rnd(H, scale=0.2) + 1, rnd(H, scale=0.2)
        y32, y16, mean, rstd = ops.layernorm_fwd(x, gamma, beta, 1e-5)
        lines.append((f"layernorm_fwd M={M}", timeit(lambda: ops.layernorm_fwd(x, gamma, beta, 1e-5))))
        lines.append((f"layernorm_bwd M={M}", timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, gamma, add=add))))
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/op_bench_{tag}.txt", "w") as f:
        for name, per in lines:
            s = f"{name:28s} total {sum(per.values()):7.1f} us   " + "  ".join(f"{k}={v:.1f}" for k, v in sorted(per.items(), key=lambda kv: -kv[1]))
            print(s)
            f.write(s + "\n")


if __name__ == "__main__":
    main()
