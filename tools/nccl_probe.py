"""Micro-benchmark of the per-iteration exchange pattern (run under torchrun): how much of it is the 2.5 MB
all-reduce and how much the small all-gathers, and what one packed all-gather would cost instead."""
import os, sys
import torch, torch.distributed as dist
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
P, n_loc = 158568, 500
big = torch.zeros(4 * P, device=dev)
rp = torch.zeros(world * n_loc, dtype=torch.float64, device=dev)
tb = [torch.zeros(world * n_loc, device=dev) for _ in range(6)]
packed = torch.zeros(world * n_loc * 8, device=dev)   # 32 bytes per entry: one double + six floats
def timeit(fn, n=200):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def ag(t): dist.all_gather_into_tensor(t, t[rank * (t.numel() // world):(rank + 1) * (t.numel() // world)])
def f_ar(): dist.all_reduce(big)
def f_small():
    with dist._coalescing_manager(device=dev):
        ag(rp)
        for t in tb: ag(t)
def f_all():
    with dist._coalescing_manager(device=dev):
        dist.all_reduce(big); ag(rp)
        for t in tb: ag(t)
def f_packed():
    with dist._coalescing_manager(device=dev):
        dist.all_reduce(big); ag(packed)
def f_one_small(): ag(rp)
res = {k: timeit(f) for k, f in [("allreduce 2.5MB", f_ar), ("one small all-gather", f_one_small), ("7 small all-gathers (group)", f_small),
                                 ("allreduce + 7 all-gathers (group) = today", f_all), ("allreduce + 1 packed all-gather (group)", f_packed)]}
if rank == 0:
    for k, v in res.items(): print(f"N={world}  {k:45s} {v:7.1f} us", flush=True)
dist.destroy_process_group()
