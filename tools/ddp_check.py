"""Run under torchrun (N >= 2): overlapped per-layer gradient exchange == plain post-backward mean all-reduce ==
mean of the per-rank gradients, on the TINY model; and a few optimizer steps keep the replicas bit-identical."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkg_analogy_b200.config import TINY, init_params  # noqa: E402
from mkg_analogy_b200.data import TINY_VOCAB, make_finetune_batch  # noqa: E402
from mkg_analogy_b200.trainer import MartTrainer, shard_seed  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
params = init_params(TINY, 11)
batch = {k: v.cuda() for k, v in make_finetune_batch(4, 24, TINY.image_size, TINY_VOCAB, seed=shard_seed(21, rank), min_len=16).items()}

grads = {}
for overlap in (False, True):
    tr = MartTrainer(TINY, TINY_VOCAB, total_steps=100, overlap_comm=overlap, dropout=False, multicast=False)
    tr.engine.load(params)
    eng = tr.engine
    mb = tr._model_inputs(batch)
    eng.zero_grad()
    eng.forward(mb, want_z=False)
    eng.head_loss(mb, tr._head(batch), compute_grad=True)
    if overlap:
        tr._backward_overlapped(mb)          # exchange only (no update): gradients stay in the buffer
    else:
        local_g = None
        eng.backward(mb)
        local_g = eng.grads32.clone()
        tr.allreduce_grads()
    torch.cuda.synchronize()
    grads[overlap] = eng.grads32.clone()
# reference: explicit mean of the per-rank local gradients
gathered = [torch.empty_like(local_g) for _ in range(world)]
dist.all_gather(gathered, local_g)
mean = torch.stack(gathered).mean(0)
scale = mean.abs().max().item()
e_plain = (grads[False] - mean).abs().max().item()
e_over = (grads[True] - mean).abs().max().item()
assert e_plain <= 1e-5 * scale and e_over <= 2e-4 * scale, (e_plain, e_over, scale)   # atomics: summation order only
# replicas stay identical through optimizer steps, with the sharded optimizer (reduce-scatter -> shard AdamW -> all-gather) and
# with the replicated one (all-reduce -> full AdamW); both give the same parameters (summation order of the collective aside)
final = {}
modes = {"multicast": dict(multicast=True), True: dict(multicast=False, shard_optimizer=True), False: dict(multicast=False, shard_optimizer=False)}
for shard, kw in modes.items():
    tr = MartTrainer(TINY, TINY_VOCAB, total_steps=100, lr=1e-3, overlap_comm=True, dropout=False, **kw)
    if shard == "multicast" and tr._mc is None:
        if rank == 0:
            print("multicast exchange unavailable on this box: skipped")
        continue
    tr.engine.load(params)
    for step in range(3):
        b = {k: v.cuda() for k, v in make_finetune_batch(4, 24, TINY.image_size, TINY_VOCAB, seed=shard_seed(50, rank, step), min_len=16).items()}
        tr.train_step(b)
    torch.cuda.synchronize()
    p0 = tr.engine.params32.clone()
    dist.broadcast(p0, src=0)
    assert torch.equal(p0, tr.engine.params32), f"replicas diverged (shard_optimizer={shard})"
    assert torch.equal(tr.engine.params16.float(), tr.engine.params32.bfloat16().float()), "bf16 mirror out of date"
    assert tr.engine.grads32.abs().max().item() == 0.0, "gradients not cleared"
    final[shard] = tr.engine.params32.clone()
d = (final[True] - final[False]).abs()
assert d.max().item() <= 2 * 3 * 1e-3 + 1e-6 and d.mean().item() <= 2e-5, (d.max().item(), d.mean().item())
if "multicast" in final:
    dm = (final["multicast"] - final[False]).abs()
    assert dm.max().item() <= 2 * 3 * 1e-3 + 1e-6 and dm.mean().item() <= 2e-5, (dm.max().item(), dm.mean().item())
    if rank == 0:
        print(f"multicast (multimem.ld_reduce + AdamW + multimem.st) vs replicated NCCL: max |dp| {dm.max().item():.2e}, mean {dm.mean().item():.2e}")
if rank == 0:
    print(f"ddp_check ok: world {world}, plain err {e_plain:.2e}, overlapped err {e_over:.2e} (scale {scale:.2e}); "
          f"sharded vs replicated optimizer: max |dp| {d.max().item():.2e}, mean {d.mean().item():.2e}")
dist.destroy_process_group()
