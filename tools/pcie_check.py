"""Platform check: pinned host<->device copy bandwidth per rank when all ranks copy at once
(torchrun, one rank per GPU).  Explains what bounds bench.py's `e2e` at N > 1."""
import os
import time

import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl")
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.ones(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=8):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        if up:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * n / dt / 1e9


for name, up, down in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
    run(up, down, 2)
    g = run(up, down)
    t = torch.tensor([g], device="cuda")
    if world > 1:
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        dist.all_reduce(t)
        if rank == 0:
            print(f"{name}: per direction, per rank min {lo.item():.1f} max {hi.item():.1f} GB/s, sum over {world} ranks {t.item():.1f} GB/s", flush=True)
    elif rank == 0:
        print(f"{name}: {g:.1f} GB/s per direction", flush=True)
