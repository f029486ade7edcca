"""Where does a GEMM launch spend its time?  Registers a b200mart_gemm_timeline buffer, runs one steady-state training step
(all streams, as benchmarked) and prints, per (shape, kernel) group, the median duration of each phase of a CTA's life:
prologue | first TMA issue | operands landed | main loop of tile 0 | MMA drain | epilogue of tile 0 | remaining tiles | exit."""
import os
import sys
from collections import defaultdict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkg_analogy_b200 import _lib  # noqa: E402
from mkg_analogy_b200.config import BASE  # noqa: E402
from mkg_analogy_b200.data import MARS_VOCAB, make_finetune_batch  # noqa: E402
from mkg_analogy_b200.trainer import MartTrainer, init_weights_  # noqa: E402

EPI = {0: "GEN", 1: "BF16", 2: "ACT", 3: "RES", 4: "F32", 5: "DACT", 6: "ACC"}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    tr = MartTrainer(BASE, MARS_VOCAB, total_steps=1000)
    init_weights_(tr.engine, 0)
    batch = {k: v.cuda() for k, v in make_finetune_batch(32, 64, BASE.image_size, MARS_VOCAB, seed=1).items()}
    for _ in range(5):
        tr.train_step(batch)
    torch.cuda.synchronize()
    R = 400
    buf = torch.zeros(R * 148 * 16, dtype=torch.int64, device="cuda")
    lib = _lib.lib()
    lib.b200mart_gemm_timeline(buf.data_ptr(), R)
    tr.train_step(batch)
    torch.cuda.synchronize()
    n = lib.b200mart_gemm_timeline(None, 0)
    rec = buf.cpu().numpy().astype(np.uint64).reshape(R, 148, 16)[:n]
    # SM clock from the longest CTA lives
    dclk = (rec[..., 9] - rec[..., 1]).astype(np.float64)
    dgt = (rec[..., 10] - rec[..., 0]).astype(np.float64)
    ok = (rec[..., 0] > 0) & (dgt > 5000)
    ghz = float(np.median(dclk[ok] / dgt[ok]))
    groups = defaultdict(list)
    for i in range(n):
        r = rec[i]
        live = r[:, 0] > 0
        if not live.any():
            continue
        r0 = r[0]
        M, N = int(r0[12] & 0xFFFFFFFF), int(r0[12] >> 32)
        K, grid = int(r0[13] & 0xFFFFFFFF) * 64, int(r0[13] >> 32)
        kind, bn = int(r0[14] & 0xFFFFFFFF), int(r0[14] >> 32)
        groups[(M, N, K, bn, kind, grid)].append(r[live])
    lines = [f"{n} GEMM launches in one step; SM clock {ghz:.3f} GHz (clock64 / globaltimer)",
             "times in us, medians over CTAs x launches; span = max(gt_end) - min(gt_start) of a launch",
             f"{'M':>5} {'N':>5} {'K':>5} {'tile':>4} {'epi':>5} {'lay':>3} {'grid':>4} {'n':>3} {'t/cta':>5} | {'span':>6} {'cta':>6} | "
             f"{'prolog':>6} {'tma0':>5} {'land':>5} {'main0':>6} {'drain':>5} {'epi0':>5} {'rest':>6} {'exit':>5} | {'us/kb':>6} {'TF/s':>5}"]
    tot_span = 0.0
    rows = []
    for key, recs in groups.items():
        M, N, K, bn, kind, grid = key
        a = np.concatenate(recs).astype(np.float64)
        tiles = (a[:, 11].astype(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.int64)
        us = lambda x: float(np.median(x)) / ghz / 1e3
        span = float(np.median([float(r[:, 10].max() - r[:, 0].min()) for r in recs])) / 1e3
        multi = tiles > 1
        rest = us((a[multi, 8] - a[multi, 7])) if multi.any() else 0.0
        last = np.where(multi, a[:, 8], a[:, 7])
        kb = K // 64
        lead = a[a[:, 4] > 0]            # CTAs that issue MMAs (pair kernels: the leader of each pair)
        main0 = us(lead[:, 5] - lead[:, 4])
        flops = 2.0 * M * N * K
        rows.append((span * len(recs), f"{M:5d} {N:5d} {K:5d} {bn:4d} {EPI[kind & 15]:>5} {('T' if kind & 16 else 'N') + ('T' if kind & 32 else 'N'):>3} "
                     f"{grid:4d} {len(recs):3d} {int(tiles.max()):5d} | {span:6.1f} {us(a[:, 9] - a[:, 1]):6.1f} | "
                     f"{us(a[:, 2] - a[:, 1]):6.2f} {us(a[:, 3] - a[:, 2]):5.2f} {us(lead[:, 4] - lead[:, 3]):5.2f} {main0:6.2f} "
                     f"{us(lead[:, 6] - lead[:, 5]):5.2f} {us(a[:, 7] - a[:, 6]):5.2f} {rest:6.2f} {us(a[:, 9] - last):5.2f} | "
                     f"{main0 / max(kb, 1):6.3f} {flops / span / 1e6:5.0f}"))
        tot_span += span * len(recs)
    for _, line in sorted(rows, key=lambda t: -t[0]):
        lines.append(line)
    lines.append(f"sum of spans {tot_span / 1e3:.3f} ms")
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/gemm_timeline_{tag}.txt", "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
