"""In-situ kernel timeline of the training step (CUPTI through torch.profiler; nsys is not in the image).
Writes gpurun_out/timeline_<tag>.json.gz (chrome trace of the profiled steps) and prints per-stream / per-kernel-family
statistics of ONE steady-state step: busy time per stream, time with 0 / 1 / >= 2 kernels resident, in-situ durations."""
import gzip
import json
import os
import re
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkg_analogy_b200.config import BASE  # noqa: E402
from mkg_analogy_b200.data import MARS_VOCAB, make_finetune_batch  # noqa: E402
from mkg_analogy_b200.trainer import MartTrainer, init_weights_  # noqa: E402


def family(name):
    name = re.sub(r"^void\s+", "", name)
    m = re.match(r"(?:\(anonymous namespace\)::|<unnamed>::|gemmk::)?([A-Za-z0-9_]+)(<[^>]*>)?", name)
    base = m.group(1) if m else name
    if base.startswith("gemm_bf16_tcgen05"):
        return base + (m.group(2) or "")
    return base


def analyse(events, out):
    ks = sorted((e for e in events if e.get("cat") == "kernel"), key=lambda e: e["ts"])
    t0, t1 = ks[0]["ts"], max(e["ts"] + e["dur"] for e in ks)
    span = t1 - t0
    print(f"{len(ks)} kernels, span {span / 1e3:.3f} ms", file=out)
    by_stream = defaultdict(list)
    for e in ks:
        by_stream[e["args"].get("stream")].append(e)
    for s, es in sorted(by_stream.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
        busy = sum(e["dur"] for e in es)
        print(f"  stream {s}: {len(es):4d} kernels, busy {busy / 1e3:7.3f} ms ({100 * busy / span:5.1f} % of span), "
              f"first +{(es[0]['ts'] - t0) / 1e3:.3f} ms, last end +{(max(e['ts'] + e['dur'] for e in es) - t0) / 1e3:.3f} ms", file=out)
    pts = []
    for e in ks:
        pts.append((e["ts"], 1))
        pts.append((e["ts"] + e["dur"], -1))
    pts.sort()
    depth, last, hist = 0, t0, defaultdict(float)
    for t, d in pts:
        hist[min(depth, 4)] += t - last
        depth += d
        last = t
    print("  resident kernels over the span: " + ", ".join(f"{k}{'+' if k == 4 else ''}: {100 * v / span:.1f} %" for k, v in sorted(hist.items())), file=out)
    fam = defaultdict(lambda: [0, 0.0])
    for e in ks:
        f = fam[family(e["name"])]
        f[0] += 1
        f[1] += e["dur"]
    tot = sum(v[1] for v in fam.values())
    print(f"  sum of kernel durations {tot / 1e3:.3f} ms (overlap factor {tot / span:.2f})", file=out)
    for name, (n, d) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"    {d / 1e3:7.3f} ms {100 * d / tot:5.1f} %  x{n:<4d} avg {d / n:7.1f} us  {name}", file=out)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    # under torchrun (N > 1): every rank steps, rank 0 is profiled (its kernels + the exchange kernels that run on its GPU)
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    tr = MartTrainer(BASE, MARS_VOCAB, total_steps=1000)
    init_weights_(tr.engine, 0)
    batch = {k: v.cuda() for k, v in make_finetune_batch(32, 64, BASE.image_size, MARS_VOCAB, seed=1 + 100 * rank).items()}

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(5):
        tr.train_step(batch)
    sync()
    if rank != 0:
        for i in range(3):
            tr.train_step(batch)
            sync()
        dist.destroy_process_group()
        return
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        for i in range(3):
            tr.train_step(batch)
            sync()                        # step boundaries: kernels of step i end before step i+1 starts
    os.makedirs("gpurun_out", exist_ok=True)
    path = f"gpurun_out/timeline_{tag}.json"
    prof.export_chrome_trace(path)
    with open(path) as f:
        trace = json.load(f)
    ev = [e for e in trace["traceEvents"] if e.get("cat") == "kernel"]
    ev.sort(key=lambda e: e["ts"])
    # split into steps at gaps > 200 us (the synchronize between steps)
    steps, cur = [], [ev[0]]
    for e in ev[1:]:
        if e["ts"] - max(x["ts"] + x["dur"] for x in cur[-8:]) > 200:
            steps.append(cur)
            cur = []
        cur.append(e)
    steps.append(cur)
    slim = [dict(name=e["name"], ts=e["ts"], dur=e["dur"], cat="kernel", ph="X", pid=0, tid=e["args"].get("stream"),
                 args=dict(stream=e["args"].get("stream"), grid=e["args"].get("grid"), block=e["args"].get("block")))
            for e in steps[-1]]
    with gzip.open(f"gpurun_out/timeline_{tag}.json.gz", "wt") as f:
        json.dump({"traceEvents": slim}, f)
    os.remove(path)
    with open(f"gpurun_out/timeline_{tag}.txt", "w") as out:
        print(f"{len(steps)} steps found; analysing the last", file=out)
        analyse(steps[-1], out)
    print(open(f"gpurun_out/timeline_{tag}.txt").read())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
