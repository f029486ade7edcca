#!/usr/bin/env python
"""MARS analogy-examples/sec of the MKGformer-base fine-tune step (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one optimizer step on one 32-example synthetic MARS-shaped batch per GPU (224x224 image pair, 64 text
tokens, 2,063-entity score head over the 42,007-token vocabulary): forward, label-smoothed CE + relaxation loss on the
[MASK] rows, backward, (N>1: NCCL all-reduce of the flat fp32 gradient buffer), AdamW.  Arithmetic: bf16 tensor-core
operands, fp32 accumulation / residual streams / LayerNorm / softmax / loss / master weights; training-mode dropout
(p = 0.1 at the four BERT-stream sites) on.

Prints ONE JSON line (rank 0): value = whole-job examples/s with inputs resident in HBM (CUDA events, max over ranks);
e2e = the same step through MartTrainer.train_step_host_async with pinned host inputs (every step's H2D input copy and loss
D2H read inside the timed region, pipelined behind the previous step's compute);
roofline = the tcgen05 GEMM kernel (dominant: all dense contractions) timed per launch with CUDA events in a separate
profiled pass of the same steps; cpu_baseline = the CPU oracle port of the reference step on this box's host cores.

--impl reference times the reference's CPU implementation (oracle port: /root/reference does not exist on the GPU box)
on the same metric/config, each step a bounded sample (8 examples) of the workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MARS analogy-examples/sec (MKGformer-base, bs=32, 224px/64tok)"
B_PER_GPU, SEQ_LEN = 32, 64
# algorithmic FLOPs per example, forward + backward, decoder on the [MASK] row only
GF_PER_EXAMPLE_FWD_BWD = 90.50


def config_dict(n_gpus):
    return {"workload": "MKGformer-base MARS fine-tune step (fwd + LS-CE/relaxation loss + bwd + AdamW), synthetic batch",
            "batch_per_gpu": B_PER_GPU, "global_batch": B_PER_GPU * n_gpus, "seq_len": SEQ_LEN, "image": "2x3x224x224",
            "score_head": "2063 analogy entities of a 42007-token vocab", "params_M": 234.7,
            "parallelism": f"dp{n_gpus}" + ("" if n_gpus == 1 else " (flat fp32 gradient mean all-reduce per layer extent, pipelined behind backward)"),
            "optimizer": "AdamW per layer extent, pipelined behind the backward of the layers below",
            "dropout": "hidden 0.1 / attention 0.1 (BERT stream, as the reference trains)",
            "l2": "per-step working set (0.47 GB bf16 weights + >1 GB activations) exceeds the 126 MB L2; 4 rotating batches"}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.samples, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def window(self, t0, t1):
        rows = [l.split(", ") for t, l in self.samples if t0 <= t <= t1]
        rows = [r for r in rows if len(r) >= 6]
        if not rows:
            rows = [l.split(", ") for _, l in self.samples[-3:]]
            rows = [r for r in rows if len(r) >= 6]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].strip().lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(float(r[0]) for r in rows), "sm_max_mhz": float(rows[0][1]), "reasons": reasons,
                "samples": len(rows)}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ------------------------------------------------------------------------------------------------ CPU oracle legs
def cpu_oracle_rate(sample_examples, steps, warmup, threads=None):
    """examples/s of the reference step (oracle port, fp32, torch CPU) on this box's host cores."""
    import torch
    from mkg_analogy_b200.config import BASE, init_params
    from mkg_analogy_b200.data import MARS_VOCAB, make_finetune_batch
    from oracle import mkgformer_oracle as O
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    P = {k: v.requires_grad_(True) for k, v in init_params(BASE, 0).items()}
    ids = MARS_VOCAB.analogy_entity_ids()
    state = {}
    times = []
    for it in range(warmup + steps):
        batch = make_finetune_batch(sample_examples, SEQ_LEN, BASE.image_size, MARS_VOCAB, seed=1234 + it)
        t0 = time.perf_counter()
        for p in P.values():
            p.grad = None
        loss, _, _ = O.finetune_loss(P, BASE, batch, ids, alpha=0.43, label_smoothing=0.1)
        loss.backward()
        with torch.no_grad():
            O.adamw_step({k: v for k, v in P.items()}, {k: v.grad for k, v in P.items() if v.grad is not None}, state,
                         it + 1, 5e-5)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sample_examples * len(times) / sum(times), sum(times) / len(times), threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 8
    rate, step_s, threads = cpu_oracle_rate(sample, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "examples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_dict(args.gpus),
            "cpu_baseline": {"value": rate, "unit": "examples/s", "cores": threads, "kind": "port",
                             "sample": f"{sample}-example slices of the 32-example batch per step, fwd+loss+bwd+AdamW, torch CPU fp32"},
            "e2e": {"value": rate, "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ CUDA arm
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from mkg_analogy_b200.config import BASE
    from mkg_analogy_b200.data import MARS_VOCAB, make_finetune_batch
    from mkg_analogy_b200.trainer import MartTrainer, init_weights_, BATCH_KEYS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its banner ("NCCL version ...") to stdout when the communicator
        # is created, so file descriptor 1 points at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    dev = torch.device("cuda", local)

    K, W = args.steps, max(args.warmup, 3)
    trainer = MartTrainer(BASE, MARS_VOCAB, total_steps=10 * (K + W) * 4, device=dev, overlap_comm=not args.no_overlap)
    init_weights_(trainer.engine, seed=0)                     # identical replicas on every rank
    nb = 4
    host = [make_finetune_batch(B_PER_GPU, SEQ_LEN, BASE.image_size, MARS_VOCAB, seed=1234 + rank * 100 + i) for i in range(nb)]
    host = [{k: v.pin_memory() for k, v in b.items() if k in BATCH_KEYS} for b in host]
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    # ---------------- device-resident timing
    for i in range(W):
        trainer.train_step(devb[i % nb])
    barrier()
    l0 = trainer.engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(K):
        loss = trainer.train_step(devb[i % nb])
    e1.record()
    barrier()
    t1 = time.perf_counter()
    launches = trainer.engine.launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    loss_val = float(loss.item())
    clocks = sampler.window(t0, t1) if sampler else None

    # ---------------- end-to-end through the public API with host inputs
    # every step: pinned host batch -> device (38.6 MB) and one loss read back; copies ride a copy stream behind the previous
    # step's compute and the loss is read one step late (asynchronous logging), see MartTrainer.train_step_host_async
    for i in range(3):
        trainer.train_step_host_async(host[i % nb])
    trainer.drain()
    barrier()
    t2 = time.perf_counter()
    for i in range(K):
        trainer.train_step_host_async(host[i % nb])
    trainer.drain()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t2], device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_rate = B_PER_GPU * world * K / e2e_s.item()

    # ---------------- per-kernel-family device time (separate profiled pass, same steps)
    roof = prof = None
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except Exception:
            pass
        peak_tf, peak_src = (peaks["bf16_tflops_sustained"], "measured (sustained)") if "bf16_tflops_sustained" in peaks else (1400.0, "fallback")
    np_steps = 3
    trainer.engine.set_option("streams", 0)        # serial launches: clean per-kernel durations
    trainer.engine.profile_begin()
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    for i in range(np_steps):
        trainer.train_step(devb[i % nb])
    pe1.record()
    cats, gemm_flops = trainer.engine.profile_end()
    trainer.engine.set_option("streams", 1)
    if rank == 0:
        gemm_ms, gemm_n = cats["gemm"]
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
        step_ms = ms_total / K
        roof = {"bound": "tensor", "kernel": "gemm_bf16_tcgen05", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "peak_source": peak_src,
                # DRAM bytes per launch from the committed ncu --set full capture of 40 GEMM launches of this step
                # (profiles/gemm_full_r1f.txt); well below the algorithmic bytes because activations stay in the 126 MB L2
                "traffic": 19614874.0, "traffic_source": "profiles/gemm_full_r1f.txt (ncu dram__bytes_read+write.sum, per-launch mean)",
                "algorithmic_bytes_per_launch": trainer.engine.last_profile_gemm_bytes / gemm_n,
                "launches_per_step": gemm_n // np_steps, "avg_launch_us": gemm_ms * 1e3 / gemm_n,
                "gflop_per_launch": gemm_flops / gemm_n / 1e9,
                "gemm_ms_per_step": gemm_ms / np_steps,
                "attention_ms_per_step": cats["attention"][0] / np_steps, "fusion_ms_per_step": cats["fusion"][0] / np_steps,
                "profiled_step_ms": pe0.elapsed_time(pe1) / np_steps,
                "step_frac_of_peak": GF_PER_EXAMPLE_FWD_BWD * 1e9 * B_PER_GPU / (step_ms * 1e-3) / 1e12 / peak_tf}

    # ---------------- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        rate, step_s, threads = cpu_oracle_rate(8, 1, 1)
        cpu = {"value": rate, "unit": "examples/s", "cores": threads, "kind": "port",
               "sample": "one 8-example slice of the 32-example batch (after one warm-up), fwd+loss+bwd+AdamW, torch CPU fp32"}
    if sampler:
        sampler.stop()
    if rank == 0:
        h2d = MartTrainer.h2d_bytes(host[0])
        line = {"metric": METRIC, "value": B_PER_GPU * world * K / (ms_total * 1e-3), "unit": "examples/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": config_dict(world),
                "clocks": clocks, "gpu_launches": int(launches), "final_loss": loss_val,
                "e2e": {"value": e2e_rate, "unit": "examples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "api": "MartTrainer.train_step_host_async (pinned host batch in, float loss out, wall clock)"},
                "roofline": roof, "cpu_baseline": cpu}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-overlap", action="store_true", help="run the gradient exchange / AdamW after backward instead of pipelined per layer")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
