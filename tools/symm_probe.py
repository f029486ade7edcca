"""torchrun probe: does NVSwitch multicast (torch symmetric memory + multimem.red) work on this box?"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mkg_analogy_b200 import _lib  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import torch.distributed._symmetric_memory as symm_mem  # noqa: E402

n = 1 << 20
t = symm_mem.empty(n, dtype=torch.float32, device=torch.device("cuda", local))
t.zero_()
hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
print(f"rank {rank}: multicast_ptr {hdl.multicast_ptr:#x} buffer_ptrs {[hex(p) for p in hdl.buffer_ptrs]}", flush=True)
hdl.barrier(channel=0)
src = torch.full((n,), float(rank + 1), device="cuda")
if hdl.multicast_ptr:
    _lib.check(_lib.lib().b200mart_multimem_add(src.data_ptr(), hdl.multicast_ptr, n, 0.5, _lib.cur_stream()))
    hdl.barrier(channel=0)
    torch.cuda.synchronize()
    want = 0.5 * sum(range(1, world + 1))
    print(f"rank {rank}: after multimem.red: min {t.min().item()} max {t.max().item()} (expected {want})", flush=True)
    # timing: 64 MB
    big = symm_mem.empty(16 << 20, dtype=torch.float32, device=torch.device("cuda", local))
    big.zero_()
    h2 = symm_mem.rendezvous(big, group=dist.group.WORLD)
    s2 = torch.ones(16 << 20, device="cuda")
    h2.barrier(channel=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _lib.check(_lib.lib().b200mart_multimem_add(s2.data_ptr(), h2.multicast_ptr, 16 << 20, 1.0, _lib.cur_stream()))
    e1.record()
    h2.barrier(channel=0)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"rank {rank}: multimem.red of 64 MB: {ms:.3f} ms = {64e6 * 1.048576 / ms / 1e6:.0f} GB/s out per GPU; value {big[0].item()} (expected {10.0 * world})", flush=True)
dist.destroy_process_group()
