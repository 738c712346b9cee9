"""world_size-2 gloo test of the N>1 host logic: batch partitioning + the allgather of result
records give exactly what one process computes for the whole batch (the solve itself has no
collective).  The per-rank solver here is the CPU oracle; on GPUs bench.py uses the same helpers."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import trajopt_b200 as to
    import oracle_py
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = to.problems.cartpole(constrained=True)
    b0, b1 = to.sharding.shard_range(B, rank, world)
    x0 = to.problems.batch_x0("cartpole", b1 - b0, offset=b0)
    r = oracle_py.solve(prob, to.ALTROSolverOptions(), x0=x0, B=b1 - b0, inner_cap=0, outer_cap=0)
    allr = to.sharding.allgather_results(r["results"], dist)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), allr)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_allgather(tmp_path, oracle, to):
    B, world = 5, 2  # ragged: 3 + 2
    assert [to.sharding.shard_range(B, r, world) for r in range(world)] == [(0, 3), (3, 5)]
    assert to.sharding.shard_range(8, 3, 4) == (6, 8) and to.sharding.shard_range(2, 3, 4) == (2, 2)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(tmp_path, "gathered.npy"))
    prob = to.problems.cartpole(constrained=True)
    ref = oracle.solve(prob, to.ALTROSolverOptions(), x0=to.problems.batch_x0("cartpole", B), B=B, inner_cap=0, outer_cap=0)["results"]
    assert got.dtype == ref.dtype and got.tobytes() == ref.tobytes()
