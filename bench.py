#!/usr/bin/env python
"""MARS analogy-examples/sec of the MarT training step (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config base|large|flava|vilt|visualbert]

Default (--config base, BASELINE.json configs[1], the configuration the metric is quoted on): one "step" = one optimizer step on
one 32-example synthetic MARS-shaped batch per GPU (224x224 image pair, 64 text tokens, 2,063-entity score head over the
42,007-token vocabulary): forward, label-smoothed CE + relaxation loss on the [MASK] rows, backward, (N>1: gradient exchange over
NCCL), AdamW.  Arithmetic: bf16 tensor-core operands, fp32 accumulation / residual streams / LayerNorm / softmax / loss / master
weights; training-mode dropout (p = 0.1 at the four BERT-stream sites) on.  Other configs: `large` = BASELINE.json configs[4]
(MKGformer-large, 24 x 1024, ViT-L/14 patches, bs=16/GPU), `flava` = configs[2] (FLAVA-base fine-tune, bs=64/GPU), `vilt` =
configs[3] (ViLT-base pre-train, 384 px, L=128) - the last two run their three / one encoder stacks in the library and keep
embeddings / projections / head transform as PyTorch plumbing (stated in `config`).

Prints ONE JSON line (rank 0): value = whole-job examples/s with inputs resident in HBM (CUDA events, max over ranks);
e2e = the same step through MartTrainer.train_step_host_async with pinned HOST inputs (every step's H2D input copy and loss D2H
read inside the timed region): token / index tensors plus (B, 2) rows of a device-resident entity-image table, as the input path
of SURVEY.md §8 f2 feeds a real run (`e2e.pixels_from_host` repeats the leg with the 38.5 MB of pixels crossing PCIe per step);
roofline = the tcgen05 GEMM kernel (dominant: all dense contractions) timed per launch with CUDA events in a separate
profiled pass of the same steps; cpu_baseline = the CPU oracle port of the reference step on this box's host cores;
gpu_eager_baseline = the same oracle port (the reference graph in stock PyTorch ops) on this GPU, fp32 and bf16 autocast.

--impl reference times the reference's CPU implementation (oracle port: /root/reference does not exist on the GPU box)
on the same metric/config, each step a bounded sample (8 examples) of the workload; plus BASELINE.json configs[0]
(forward only, 4 examples) as `config1_forward`.
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEQ_LEN = 64
# algorithmic GFLOP per example, forward + backward (= 3 x forward), decoder on the [MASK] row only
WORKLOADS = {
    "base": dict(metric="MARS analogy-examples/sec (MKGformer-base, bs=32, 224px/64tok)", B=32, L=64, gf=90.50, params_M=234.7,
                 workload="MKGformer-base MARS fine-tune step (fwd + LS-CE/relaxation loss + bwd + AdamW), synthetic batch"),
    "large": dict(metric="MARS analogy-examples/sec (MKGformer-large, bs=16/GPU, 224px/64tok)", B=16, L=64, gf=1165.5, params_M=None,
                  workload="MKGformer-large (24 x 1024, ViT-L/14 patches, 513 vision tokens) MARS fine-tune step, synthetic batch"),
    "flava": dict(metric="MARS analogy-examples/sec (FLAVA-base, bs=64/GPU, 224px/64tok)", B=64, L=64, gf=3 * 127.4, params_M=None,
                  workload="FLAVA-base MarT fine-tune step (image 12 + text 12 + multimodal 6 blocks), synthetic batch"),
    "vilt": dict(metric="MarKG pre-train examples/sec (ViLT-base, 384px/128tok)", B=32, L=128, gf=3 * 79.0, params_M=None,
                 workload="ViLT-base MarT pre-train step on synthetic MarKG triples (12 blocks over 128 + 290 tokens)"),
    # not a BASELINE.json config: the reference's scripts/run_finetune_visualbert.sh shape (batch 64, max_seq_length 128, 2 x 36 region features)
    "visualbert": dict(metric="MARS analogy-examples/sec (VisualBERT-base, bs=64/GPU, 72 regions/128tok)", B=64, L=128, gf=3 * 36.0, params_M=None,
                       workload="VisualBERT-base MarT fine-tune step (12 post-LN blocks over 128 text tokens + 72 region features), synthetic batch"),
}


def config_dict(name, n_gpus, extra=None):
    w = WORKLOADS[name]
    d = {"workload": w["workload"], "name": name, "batch_per_gpu": w["B"], "global_batch": w["B"] * n_gpus, "seq_len": w["L"],
         "image": "2x3x384x384" if name == "vilt" else "72 region features x 2048 (Faster-RCNN stand-in)" if name == "visualbert" else "2x3x224x224",
         "score_head": "entity (11292) + relation (192) column groups of a 42007-token vocab" if name == "vilt"
         else "2063 analogy entities of a 42007-token vocab",
         "parallelism": f"dp{n_gpus}",
         "optimizer": "AdamW per layer extent, pipelined behind the backward of the layers below",
         "dropout": "hidden 0.1 / attention 0.1 (BERT stream, as the reference trains)" if name in ("base", "large", "visualbert") else "off (HF defaults of the family)",
         "l2": "per-step working set (bf16 weights + >1 GB activations) exceeds the 126 MB L2; 4 rotating batches"}
    if w["params_M"]:
        d["params_M"] = w["params_M"]
    if extra:
        d.update(extra)
    return d


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


# ------------------------------------------------------------------------------------------------ baselines (oracle port)
def oracle_rate(sample_examples, steps, warmup, device="cpu", autocast=False, threads=None, forward_only=False):
    """examples/s of the reference step (oracle port = the reference graph in stock PyTorch ops, MKGformer-base) on the host
    cores (device 'cpu', fp32) or on the GPU in eager mode (device 'cuda', fp32 or bf16 autocast)."""
    import torch
    from mkg_analogy_b200.config import BASE, init_params
    from mkg_analogy_b200.data import MARS_VOCAB, make_finetune_batch
    from oracle import mkgformer_oracle as O
    if device == "cpu":
        threads = threads or os.cpu_count() or 1
        torch.set_num_threads(threads)
    P = {k: v.to(device).requires_grad_(not forward_only) for k, v in init_params(BASE, 0).items()}
    ids = MARS_VOCAB.analogy_entity_ids()
    opt = None
    if not forward_only:
        nd = [v for k, v in P.items() if "bias" in k or "LayerNorm.weight" in k]
        dc = [v for k, v in P.items() if not ("bias" in k or "LayerNorm.weight" in k)]
        opt = torch.optim.AdamW([{"params": dc, "weight_decay": 0.01}, {"params": nd, "weight_decay": 0.0}], lr=5e-5, eps=1e-8)
    times = []
    for it in range(warmup + steps):
        batch = {k: v.to(device) for k, v in make_finetune_batch(sample_examples, SEQ_LEN, BASE.image_size, MARS_VOCAB, seed=1234 + it).items()}
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        if forward_only:
            with torch.no_grad():
                O.forward(P, BASE, batch, full_logits=True)
        else:
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast and device != "cpu"):
                loss, _, _ = O.finetune_loss(P, BASE, batch, ids, alpha=0.43, label_smoothing=0.1)
            loss.float().backward()
            opt.step()
        if device != "cpu":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return sample_examples * len(times) / sum(times), sum(times) / len(times), threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 8
    rate, step_s, threads = oracle_rate(sample, args.steps, args.warmup)
    f_rate, f_s, _ = oracle_rate(4, 2, 1, forward_only=True)            # BASELINE.json configs[0]
    line = {"impl": "reference", "metric": WORKLOADS["base"]["metric"], "value": rate, "unit": "examples/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_dict("base", args.gpus, {"batch_per_step": sample, "graph": "eval mode (no dropout), torch CPU fp32",
                                                      "note": "one host process whatever --gpus is; each step is an 8-example sample of the 32-example batch"}),
            "cpu_baseline": {"value": rate, "unit": "examples/s", "cores": threads, "kind": "port",
                             "sample": f"{sample}-example slices of the 32-example batch per step, fwd+loss+bwd+AdamW, torch CPU fp32"},
            "config1_forward": {"value": f_rate, "unit": "examples/s", "ms": f_s * 1e3,
                                "what": "BASELINE.json configs[0]: MKGformer-base single forward on 4 synthetic MARS examples, full (4, 64, 42007) logits, torch CPU fp32"},
            "e2e": {"value": rate, "unit": "examples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ helpers
def init_dist():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
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
    return world, rank, local


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return p["bf16_tflops_sustained"], "measured (sustained)"
    except Exception:
        return 1400.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """Mean DRAM bytes per GEMM launch from the newest committed `ncu --set full` summary (profiles/gemm_full_*.txt)."""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for f in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        if f.startswith("gemm_full_") and f.endswith(".txt"):
            m = re.search(r"mean DRAM traffic per launch (\d+) bytes", open(os.path.join(pdir, f)).read())
            if m:
                rnd = re.search(r"_r(\d+)", f)
                key = (int(rnd.group(1)) if rnd else 0, f)
                if best is None or key > best[0]:
                    best = (key, float(m.group(1)), f)
    return (best[1], f"profiles/{best[2]} (ncu dram__bytes_read+write.sum, per-launch mean)") if best else (None, "no committed ncu summary")


# ------------------------------------------------------------------------------------------------ MKGformer arm (base / large)
def run_mkgformer(args):
    import torch
    import torch.distributed as dist
    from mkg_analogy_b200.config import BASE, LARGE
    from mkg_analogy_b200.data import MARS_VOCAB, EntityImageTable, make_finetune_batch
    from mkg_analogy_b200.trainer import MartTrainer, init_weights_, BATCH_KEYS

    world, rank, local = init_dist()
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    dev = torch.device("cuda", local)
    name = args.config
    w = WORKLOADS[name]
    cfg = LARGE if name == "large" else BASE
    Bp, L = w["B"], w["L"]
    K, W = args.steps, max(args.warmup, 3)
    trainer = MartTrainer(cfg, MARS_VOCAB, total_steps=10 * (K + W) * 4, device=dev, overlap_comm=not args.no_overlap,
                          shard_optimizer=args.shard, multicast=False if args.no_multicast else None)
    exchange = "none (1 GPU)" if world == 1 else (
        "NVSwitch multicast: per layer extent ONE kernel = multimem.ld_reduce of the rank's gradient shard + AdamW + multimem.st of the "
        "new fp32 / bf16 parameters to every replica (no collective library on the path)" if trainer._mc is not None else
        ("NCCL reduce-scatter -> AdamW on the rank's shard -> NCCL all-gather, per layer extent" if trainer.shard_optimizer else
         "NCCL mean all-reduce per layer extent, replicated AdamW"))
    init_weights_(trainer.engine, seed=0)                     # identical replicas on every rank
    nb = 4
    host = [make_finetune_batch(Bp, L, cfg.image_size, MARS_VOCAB, seed=1234 + rank * 100 + i) for i in range(nb)]
    host = [{k: v.pin_memory() for k, v in b.items() if k in BATCH_KEYS} for b in host]
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    # ---------------- device-resident timing
    graphed = not args.no_graph and world == 1 and not args.no_overlap
    per_step_launches = None
    graph_note = None
    if graphed:      # two eager steps per shape (plans, workspace), then ONE captured graph per step
        try:
            trainer.enable_graphs(buckets=(L,), warmup=2)
            lc = trainer.engine.launch_count()
            trainer.train_step(devb[0])
            per_step_launches = trainer.engine.launch_count() - lc   # kernels one replay re-issues (the counter ticks at capture only)
            trainer.train_step(devb[1])
            trainer.train_step(devb[2])                              # third step of the shape: capture + first replay
            torch.cuda.synchronize()
        except Exception as e:      # noqa: BLE001 - a box whose driver refuses the capture still gets the (eager) GPU measurement
            graph_note = f"eager (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
            graphed = False
            trainer._graphs = None
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
    if graphed:
        launches = per_step_launches * K
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    loss_val = float(loss.item())
    clocks = sampler.window(t0, t1) if sampler else None

    # ---------------- replicas agree? (N > 1: max |param on rank 0 - param on this rank| after the timed steps)
    replica = None
    if world > 1:
        p0 = trainer.engine.params32.clone()
        dist.broadcast(p0, src=0)
        dmax = (p0 - trainer.engine.params32).abs().max().reshape(1)
        dist.all_reduce(dmax, op=dist.ReduceOp.MAX)
        replica = {"max_abs_param_diff_vs_rank0": float(dmax.item()), "after_steps": W + K}
        del p0

    # ---------------- end-to-end through the public API with HOST inputs
    def e2e_leg(batches):
        for i in range(3):
            trainer.train_step_host_async(batches[i % nb])
        trainer.drain()
        barrier()
        t2 = time.perf_counter()
        for i in range(K):
            trainer.train_step_host_async(batches[i % nb])
        trainer.drain()
        barrier()
        s = torch.tensor([time.perf_counter() - t2], device=dev)
        if world > 1:
            dist.all_reduce(s, op=dist.ReduceOp.MAX)
        return Bp * world * K / s.item()

    # (a) input path of SURVEY.md §8 f2: a device-resident entity-image table (synthetic entities, bf16), host batches carry
    #     (B, 2) table rows (-1 = zero image) instead of pixels; the gather runs on the device inside the timed region
    n_ent = 512
    g = torch.Generator().manual_seed(99)
    feats = torch.randn(n_ent, 3, cfg.image_size, cfg.image_size, generator=g)
    table = EntityImageTable(feats, [f"Q{i}" for i in range(n_ent)], device=dev, dtype=torch.bfloat16)
    del feats
    trainer.attach_image_table(table)
    host_idx = []
    for i, b in enumerate(host):
        hb = {k: v for k, v in b.items() if k != "pixel_values"}
        idx = torch.randint(0, n_ent, (Bp, 2), generator=g)
        idx[torch.rand(Bp, generator=g) < 0.4, 1] = -1
        hb["pixel_idx"] = idx.pin_memory()
        host_idx.append(hb)
    e2e_rate = e2e_leg(host_idx)
    h2d = MartTrainer.h2d_bytes(host_idx[0])
    # (b) the pixels themselves crossing PCIe every step (what the reference's collator ships)
    e2e_px_rate = e2e_leg(host)
    h2d_px = MartTrainer.h2d_bytes(host[0])

    # ---------------- per-kernel-family device time (separate profiled pass, same steps, chains serialised)
    roof = None
    np_steps = 3
    trainer._graphs = None                         # the profiled pass issues its launches eagerly
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
        peak_tf, peak_src = measured_peaks()
        gemm_ms, gemm_n = cats["gemm"]
        achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12
        step_ms = ms_total / K
        traffic, traffic_src = ncu_traffic() if name == "base" else (None, "not captured for this config")
        roof = {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_pair", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": trainer.engine.last_profile_gemm_bytes / gemm_n,
                "launches_per_step": gemm_n // np_steps, "avg_launch_us": gemm_ms * 1e3 / gemm_n,
                "gflop_per_launch": gemm_flops / gemm_n / 1e9, "gemm_ms_per_step": gemm_ms / np_steps,
                "attention_ms_per_step": cats["attention"][0] / np_steps, "fusion_ms_per_step": cats["fusion"][0] / np_steps,
                "profiled_step_ms": pe0.elapsed_time(pe1) / np_steps,
                "step_frac_of_peak": w["gf"] * 1e9 * Bp / (step_ms * 1e-3) / 1e12 / peak_tf,
                "how": "per GEMM launch, CUDA events in a pass with the chains serialised; step_frac_of_peak = algorithmic FLOPs of the "
                       "whole step / concurrent step time / peak"}

    # ---------------- baselines (rank 0, N = 1 only): bounded samples of the same workload
    cpu = eager = None
    if rank == 0 and world == 1 and not args.no_cpu and name == "base":
        del trainer, table, devb
        torch.cuda.empty_cache()
        eager = {}
        for tag, ac in (("fp32", False), ("bf16_autocast", True)):
            r, s, _ = oracle_rate(Bp, 3, 1, device="cuda", autocast=ac)
            eager[tag] = {"value": r, "unit": "examples/s", "ms_per_step": s * 1e3}
        eager["what"] = ("the reference graph in stock PyTorch ops (oracle port of modeling_unimo.py + training_step, eval-mode graph) on this "
                         "GPU, same 32-example batch shape, fwd + loss + bwd + torch AdamW, eager, wall clock around synchronised steps")
        rate, step_s, threads = oracle_rate(8, 1, 1)
        cpu = {"value": rate, "unit": "examples/s", "cores": threads, "kind": "port",
               "sample": "one 8-example slice of the 32-example batch (after one warm-up), fwd+loss+bwd+AdamW, torch CPU fp32"}
    if sampler:
        sampler.stop()
    if rank == 0:
        line = {"metric": w["metric"], "value": Bp * world * K / (ms_total * 1e-3), "unit": "examples/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": config_dict(name, world, {"parallelism": f"dp{world}", "gradient_exchange": exchange,
                                                    "launch": "one CUDA graph per step (captured after 2 eager steps; per-step scalars in a device block)"
                                                    if graphed else (graph_note or "eager: every kernel launched from the host each step")}),
                "clocks": clocks, "gpu_launches": int(launches), "final_loss": loss_val,
                "e2e": {"value": e2e_rate, "unit": "examples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "api": "MartTrainer.train_step_host_async (pinned host batch in: ids / masks / labels + (B, 2) rows of the "
                               "device-resident entity-image table; float loss out; wall clock)",
                        "pixels_from_host": {"value": e2e_px_rate, "unit": "examples/s", "h2d_bytes_per_step": h2d_px}},
                "roofline": roof, "cpu_baseline": cpu, "gpu_eager_baseline": eager, "replica_check": replica,
                "parity": "tests/test_base_gpu.py: Hits@1 / ranks equal on all 256 rows of the held-out split vs the reference _eval; "
                          "trans / logits rel err 4e-2 at base size with random weights (torch.autocast(bf16) of the reference graph: 1.1e-1)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ FLAVA / ViLT arm
def run_family(args):
    """BASELINE.json configs[2] / [3]: the FLAVA / ViLT drop-ins (encoder stacks in the library, embeddings / projections /
    head transform in PyTorch), fused-route loss on the [MASK] rows, torch autograd + fused torch AdamW, DDP-style mean
    all-reduce of the gradients when N > 1."""
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F
    from mkg_analogy_b200.data import MARS_VOCAB, make_finetune_batch, make_pretrain_batch, make_region_features

    world, rank, local = init_dist()
    dev = torch.device("cuda", local)
    name = args.config
    w = WORKLOADS[name]
    Bp, L = w["B"], w["L"]
    K, W = args.steps, max(args.warmup, 3)
    torch.manual_seed(0)
    voc = MARS_VOCAB
    if name == "flava":
        from mkg_analogy_b200.modeling_flava import FlavaKGC
        model, S = FlavaKGC(device=dev), 224
    elif name == "visualbert":
        from mkg_analogy_b200.modeling_visual_bert import VisualBertKGC
        model, S = VisualBertKGC(device=dev), 32
    else:
        from mkg_analogy_b200.modeling_vilt import ViltKGC
        model, S = ViltKGC(device=dev), 384
    model.resize_token_embeddings(voc.vocab)
    model.train()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=5e-5, eps=1e-8, weight_decay=0.01, fused=True)
    cols = torch.as_tensor(voc.analogy_entity_ids(), device=dev)
    E, bias = model.get_input_embeddings().weight, model.get_output_embeddings().bias
    nb = 4
    mk = (lambda i: make_pretrain_batch(Bp, L, S, voc, seed=4321 + rank * 100 + i, min_len=64)) if name == "vilt" else \
         (lambda i: make_finetune_batch(Bp, L, S, voc, seed=1234 + rank * 100 + i))
    batches = [mk(i) for i in range(nb)]
    if name == "visualbert":
        for i, b in enumerate(batches):
            b["pixel_values"], b["visual_attention_mask"] = make_region_features(Bp, seed=99 + rank * 100 + i)
    host = [{k: v.pin_memory() for k, v in b.items()} for b in batches]
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]

    from mkg_analogy_b200 import nn as BN      # [MASK]-row score product on the tcgen05 GEMM + the fused LS-CE kernel

    def step(b):
        opt.zero_grad(set_to_none=True)
        kw = {k: b[k] for k in ("input_ids", "attention_mask", "token_type_ids", "pixel_values", "visual_attention_mask") if k in b}
        if "sep_idx" in b:
            kw["sep_idx"] = b["sep_idx"]
        _, trans = model(**kw, output_logits=False)
        B = trans.shape[0]
        _, mpos = (b["input_ids"] == 103).nonzero(as_tuple=True)
        # [MASK] rows only (fused route)
        if name == "vilt":                                                  # pre-train branch, transformer.py:72-90
            ent = b["pre_type"] != 2
            loss = trans.new_zeros(())
            if ent.any():
                loss = loss + BN.mask_row_loss(trans[ent], mpos[ent], E, bias, (voc.entity_id_st, voc.entity_id_ed), b["label"][ent])[0]
            if (~ent).any():
                loss = loss + BN.mask_row_loss(trans[~ent], mpos[~ent], E, bias, (voc.relation_id_st, voc.relation_id_ed), b["label"][~ent])[0]
        else:                                                               # fine-tune branch, transformer.py:92-109
            loss = BN.mask_row_loss(trans, mpos, E, bias, cols, b["label"])[0]
            ar = torch.arange(B, device=dev)
            cos = F.cosine_similarity
            sh = b["pixel_values"].shape[1] if name == "visualbert" else 0     # transformer.py:98-102 shifts the indices for VisualBertKGC
            sim = torch.relu(cos(trans[ar, b["q_head_idx"] + sh], trans[ar, b["a_head_idx"] + sh], dim=-1)) + 1 - \
                cos(trans[ar, b["rel_idx"][:, 0] + sh], trans[ar, b["rel_idx"][:, 1] + sh], dim=-1)
            loss = loss + 0.43 * sim.mean()
        loss.backward()
        if world > 1:      # DDP semantics: mean of the per-rank gradients, a few large buckets
            gs = [p.grad for p in params if p.grad is not None]
            bucket, size = [], 0
            for gr in gs + [None]:
                if gr is not None:
                    bucket.append(gr)
                    size += gr.numel()
                if bucket and (gr is None or size >= 32 * 1024 * 1024):
                    flat = torch.cat([t.reshape(-1) for t in bucket])
                    dist.all_reduce(flat, op=dist.ReduceOp.AVG)
                    torch._foreach_copy_(bucket, [c.view_as(t) for c, t in zip(flat.split([t.numel() for t in bucket]), bucket)])
                    bucket, size = [], 0
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from mkg_analogy_b200 import _lib
    sampler = ClockSampler(local) if rank == 0 else None
    for i in range(W):
        step(devb[i % nb])
    barrier()
    l0 = _lib.lib().b200mart_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(K):
        loss = step(devb[i % nb])
    e1.record()
    barrier()
    t1 = time.perf_counter()
    launches = _lib.lib().b200mart_launch_count() - l0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    clocks = sampler.window(t0, t1) if sampler else None
    # end to end: pinned host batch -> device every step, loss read back
    barrier()
    t2 = time.perf_counter()
    for i in range(K):
        b = {k: v.to(dev, non_blocking=True) for k, v in host[i % nb].items()}
        lv = float(step(b).item())
    barrier()
    s = torch.tensor([time.perf_counter() - t2], device=dev)
    if world > 1:
        dist.all_reduce(s, op=dist.ReduceOp.MAX)
    if sampler:
        sampler.stop()
    if args.timeline and rank == 0 and world == 1:      # CUPTI kernel timeline of two steps (profiling runs; after the timed regions)
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import timeline as TL
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            for i in range(2):
                step(devb[i % nb])
                torch.cuda.synchronize()
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        tp = os.path.join(ROOT, "gpurun_out", f"timeline_{name}.json")
        prof.export_chrome_trace(tp)
        with open(tp) as f:
            ev = [e for e in json.load(f)["traceEvents"] if e.get("cat") == "kernel"]
        os.remove(tp)
        with open(os.path.join(ROOT, "gpurun_out", f"timeline_{name}.txt"), "w") as out:
            print(f"bench.py --config {name} --timeline: kernels of two steps (bs {Bp})", file=out)
            TL.analyse(ev, out)
    if rank == 0:
        peak_tf, peak_src = measured_peaks()
        step_ms = ms_total / K
        h2d = int(sum(v.numel() * v.element_size() for v in host[0].values()))
        line = {"metric": w["metric"], "value": Bp * world * K / (ms_total * 1e-3), "unit": "examples/s", "n_gpus": world, "steps": K,
                "warmup": W, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": config_dict(name, world, {
                    "library_part": "every transformer block (attention + MLP, forward and backward) of the encoder stacks: b200mart_stack_*",
                    "library_plumbing": "patch embedding (im2col view + tcgen05 GEMM), mm projections (FLAVA), LayerNorms, head transform, "
                                        "[MASK]-row score product and LS-CE loss: mkg_analogy_b200/nn.py over the C ABI (no cuBLAS / cuDNN call)",
                    "torch_part": "embedding gathers, concatenations, casts, GELU of the head, relaxation loss, fused torch AdamW, bucketed NCCL all-reduce",
                    "parallelism": f"dp{world}" + ("" if world == 1 else " (mean all-reduce of every parameter gradient after backward)")}),
                "clocks": clocks, "gpu_launches": int(launches), "final_loss": lv,
                "e2e": {"value": Bp * world * K / s.item(), "unit": "examples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                        "api": "model(**batch, output_logits=False) + [MASK]-row loss + backward + AdamW, pinned host batch in, float loss out"},
                "roofline": {"bound": "tensor", "kernel": "gemm_bf16_tcgen05_pair", "unit": "TFLOP/s", "peak": peak_tf, "peak_source": peak_src,
                             "achieved": w["gf"] * 1e9 * Bp / (step_ms * 1e-3) / 1e12, "frac": w["gf"] * 1e9 * Bp / (step_ms * 1e-3) / 1e12 / peak_tf,
                             "traffic": None, "how": "whole step: algorithmic FLOPs (SURVEY.md §8 d4, 3 x forward) / step time; no per-launch pass for this config"},
                "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="base", choices=list(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true", help="N = 1: launch every kernel from the host each step instead of replaying the "
                                                            "step as one CUDA graph (MartTrainer.enable_graphs)")
    ap.add_argument("--timeline", action="store_true", help="flava / vilt / visualbert, N = 1: also write a CUPTI kernel summary of two steps to gpurun_out/")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / gpu_eager_baseline legs (profiling runs)")
    ap.add_argument("--no-overlap", action="store_true", help="run the gradient exchange / AdamW after backward instead of pipelined per layer")
    ap.add_argument("--shard", action="store_true", help="N > 1, NCCL route: reduce-scatter + sharded AdamW + all-gather instead of all-reduce + replicated AdamW")
    ap.add_argument("--no-multicast", action="store_true", help="N > 1: exchange through NCCL instead of the fused NVSwitch-multicast optimizer kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config in ("base", "large"):
        run_mkgformer(args)
    else:
        run_family(args)


if __name__ == "__main__":
    main()
