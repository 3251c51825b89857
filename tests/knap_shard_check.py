"""2+ ranks (torchrun): ONE knapsack search tree over all ranks (lpx_options.knap_shard_tree).

Every rank plans and commits the same best-first search; per round each rank evaluates the
speculative subtrees it owns (round robin), the relaxations are merged by an NCCL all-reduce
(lpx_comm), foreign nodes get their assignment vectors, and the incumbents are max-all-reduced.
Checked against the unsharded solve of the same instance and against the oracle; rank 0 prints one
JSON line with both timings."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from linear_programming_solver_lpr381_b200 import _ffi as F  # noqa: E402
from linear_programming_solver_lpr381_b200 import api, workloads  # noqa: E402

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

out = []
for kind, seed in (("uncorrelated", 13), ("weak", 14), ("fractional", 15)):
    p, w, cap = workloads.knapsack_c5(seed=seed, kind=kind)
    alone = api.bnb_knapsack(p, w, cap)
    dist.barrier()
    t0 = time.perf_counter()
    alone = api.bnb_knapsack(p, w, cap)
    t_alone = time.perf_counter() - t0
    api.bnb_knapsack(p, w, cap, shard_tree=True)
    dist.barrier()
    t0 = time.perf_counter()
    shard = api.bnb_knapsack(p, w, cap, shard_tree=True)
    t_shard = time.perf_counter() - t0
    same = (shard["found"] == alone["found"] and shard["n_evals"] == alone["n_evals"] and shard["n_pops"] == alone["n_pops"]
            and np.float64(shard["best"]).view(np.uint64) == np.float64(alone["best"]).view(np.uint64)
            and list(shard["best_x"]) == list(alone["best_x"]))
    flags = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import orc_ffi
        ref = orc_ffi.knapsack(p, w, cap)
        ok_ref = (ref["found"] == shard["found"] and ref["n_evals"] == shard["n_evals"] and ref["n_pops"] == shard["n_pops"]
                  and list(ref["best_x"]) == list(shard["best_x"]))
        out.append(dict(kind=kind, n_evals=int(shard["n_evals"]), all_ranks_match_unsharded=bool(flags.item()),
                        matches_oracle=bool(ok_ref), unsharded_s=t_alone, sharded_s=t_shard,
                        unsharded_nodes_per_s=shard["n_evals"] / t_alone, sharded_nodes_per_s=shard["n_evals"] / t_shard))
if rank == 0:
    print(json.dumps(dict(check="knapsack sharded tree", world=world, cases=out)))
    assert all(c["all_ranks_match_unsharded"] and c["matches_oracle"] for c in out)
F.lib().lpx_comm_destroy()
dist.destroy_process_group()
