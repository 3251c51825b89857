"""2+ ranks (torchrun): lpx_comm_* over NCCL — the incumbent max-allreduce and the small all-gather —
checked against torch.distributed on the same values.  Also shards a batch of IP instances over
the ranks and shares the incumbents."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linear_programming_solver_lpr381_b200 import _ffi as F  # noqa: E402
from linear_programming_solver_lpr381_b200 import api, sharding, workloads  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
F.check(F.lib().lpx_init(local))
uid = (C.c_byte * 128)()
if rank == 0:
    F.check(F.lib().lpx_comm_unique_id(uid))
obj = [bytes(uid)]
dist.broadcast_object_list(obj, src=0)
uid = (C.c_byte * 128).from_buffer_copy(obj[0])
F.check(F.lib().lpx_comm_init(world, rank, uid))
vals = np.array([float(rank), -float(rank), 100.0 - rank, float("-inf") if rank else 7.5])
mine = vals.copy()
F.check(F.lib().lpx_comm_allreduce_max(F.ptr(mine), mine.size))
want = sharding.share_incumbent(vals, device=dev)
assert np.array_equal(mine, want), (mine, want)
send = np.arange(6, dtype=np.int32) + 10 * rank
recv = np.zeros(6 * world, dtype=np.int32)
F.check(F.lib().lpx_comm_allgather(F.ptr(send), F.ptr(recv), send.nbytes))
assert recv.tolist() == [v + 10 * r for r in range(world) for v in range(6)], recv
# instance sharding + incumbent sharing on a small IP batch
count = 2 * world
As, bs, cs = zip(*[workloads.ip_c4(m=12, n=18, seed=70 + k) for k in range(count)])
lo, hi = sharding.shard_range(count, rank, world)
r = api.bnb_simplex_batched(np.stack(As[lo:hi]), np.stack(bs[lo:hi]), np.stack(cs[lo:hi]))
best = np.full(count, -np.inf)
best[lo:hi] = np.where(r["found"], r["best_z"], -np.inf)
F.check(F.lib().lpx_comm_allreduce_max(F.ptr(best), best.size))
if rank == 0:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc_ffi
    ref = [orc_ffi.bnb_simplex(As[k], bs[k], cs[k]) for k in range(count)]
    assert best.tolist() == [q["best_z"] if q["found"] else -np.inf for q in ref], best
    print("comm_check ok: world", world, "incumbents", best.tolist())
F.lib().lpx_comm_destroy()
dist.destroy_process_group()
