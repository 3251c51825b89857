"""CPU, world_size 2, gloo: the multi-rank host logic (sharding, gather, incumbent sharing).
The per-unit solves are done by the oracle here — it only stands in as the evaluator so that the
sharded result can be compared with the unsharded one; the GPU path is covered by -m gpu tests."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import orc_ffi
    from linear_programming_solver_lpr381_b200 import sharding, workloads
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    count = 13
    A, b, c = workloads.batch_c2(count=count, m=12, n=20, seed=4)
    lo, hi = sharding.shard_range(count, rank, world)
    r = orc_ffi.primal_batch(A[lo:hi], b[lo:hi], c[lo:hi])
    z_all = sharding.gather_shards(r["z"], count)
    x_all = sharding.gather_shards(r["x"], count)
    piv = sharding.all_sum(r["total_pivots"])
    slow = sharding.all_max(1.0 + rank)
    inc = sharding.share_incumbent(np.array([10.0 * rank, -5.0 + rank, float("-inf")]))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), z=z_all, x=x_all, piv=piv, slow=slow, inc=inc, lo=lo, hi=hi)
    dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc_ffi
    from linear_programming_solver_lpr381_b200 import sharding, workloads
    world, port = 2, 29731
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    A, b, c = workloads.batch_c2(count=13, m=12, n=20, seed=4)
    want = orc_ffi.primal_batch(A, b, c)
    got = [np.load(os.path.join(tmp_path, f"rank{r}.npz")) for r in range(world)]
    assert (got[0]["lo"], got[0]["hi"], got[1]["lo"], got[1]["hi"]) == (0, 7, 7, 13)
    for g in got:
        assert g["z"].tobytes() == want["z"].tobytes() and g["x"].tobytes() == want["x"].tobytes()
        assert float(g["piv"]) == want["total_pivots"] and float(g["slow"]) == 2.0
        assert g["inc"].tolist() == [10.0, -4.0, float("-inf")]


def test_shard_range_partitions():
    from linear_programming_solver_lpr381_b200 import sharding
    for count in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            cuts = [sharding.shard_range(count, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == count
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
