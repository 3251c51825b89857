"""2+ ranks (torchrun): ONE pooled-tree Branch & Bound (Mode B, lpx_bnb_pooled) over all ranks.

Every rank runs the same deterministic control loop; node j of a round is solved by rank j mod world, and a
child whose parent was solved elsewhere reads the parent's tableau from that GPU's node pool over NVLink
(CUDA IPC mapping).  Checked: every rank returns the single-rank result bit for bit (nodes, outcomes, pivots,
z of every node, incumbent), which itself equals the oracle's.  Rank 0 prints one JSON line with the timings."""
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

CASES = [dict(m=60, n=120, seed=11, batch=64), dict(m=60, n=120, seed=12, batch=256), dict(m=40, n=80, seed=5, batch=1024)]
alone = []
for cs in CASES:  # before lpx_comm_init: the single-rank search
    A, b, c = workloads.ip_c4(m=cs["m"], n=cs["n"], seed=cs["seed"])
    api.bnb_pooled(A, b, c, batch=cs["batch"])
    t0 = time.perf_counter()
    r = api.bnb_pooled(A, b, c, batch=cs["batch"])
    alone.append((r, time.perf_counter() - t0))

uid = (C.c_byte * 128)()
if rank == 0:
    F.check(F.lib().lpx_comm_unique_id(uid))
obj = [bytes(uid)]
dist.broadcast_object_list(obj, src=0)
uid = (C.c_byte * 128).from_buffer_copy(obj[0])
F.check(F.lib().lpx_comm_init(world, rank, uid))

out = []
for cs, (one, t_one) in zip(CASES, alone):
    A, b, c = workloads.ip_c4(m=cs["m"], n=cs["n"], seed=cs["seed"])
    api.bnb_pooled(A, b, c, batch=cs["batch"])
    dist.barrier()
    t0 = time.perf_counter()
    sh = api.bnb_pooled(A, b, c, batch=cs["batch"])
    t_sh = time.perf_counter() - t0
    same = (sh["found"] == one["found"] and sh["n_nodes"] == one["n_nodes"] and sh["rounds"] == one["rounds"]
            and sh["node_id"].tolist() == one["node_id"].tolist() and sh["outcome"].tolist() == one["outcome"].tolist()
            and sh["pivots"].tolist() == one["pivots"].tolist()
            and sh["z"].view(np.uint64).tolist() == one["z"].view(np.uint64).tolist()
            and np.float64(sh["best_z"]).view(np.uint64) == np.float64(one["best_z"]).view(np.uint64)
            and sh["best_x"].tolist() == one["best_x"].tolist())
    flags = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    tmax = torch.tensor([t_sh], dtype=torch.float64, device=dev)
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import orc_ffi
        ref = orc_ffi.bnb_pooled(A, b, c, batch=cs["batch"])
        ok_ref = (ref["n_nodes"] == sh["n_nodes"] and ref["node_id"].tolist() == sh["node_id"].tolist()
                  and ref["z"].view(np.uint64).tolist() == sh["z"].view(np.uint64).tolist()
                  and ref["best_x"].tolist() == sh["best_x"].tolist())
        out.append(dict(case=cs, nodes=int(sh["n_nodes"]), rounds=int(sh["rounds"]), best_z=sh["best_z"],
                        all_ranks_match_single_rank=bool(flags.item()), matches_oracle=bool(ok_ref),
                        single_rank_s=t_one, sharded_s=float(tmax.item()),
                        single_rank_nodes_per_s=sh["n_nodes"] / t_one, sharded_nodes_per_s=sh["n_nodes"] / float(tmax.item())))
if rank == 0:
    print(json.dumps(dict(check="pooled-tree B&B over all ranks (Mode B)", world=world, cases=out)))
    assert all(c["all_ranks_match_single_rank"] and c["matches_oracle"] for c in out)
F.lib().lpx_comm_destroy()
dist.destroy_process_group()
