"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank owns shard_id % world == rank, applies its own stream
through an oracle-backed stand-in for the engine (the GPU engine is per-rank and identical), and the whole-job
rate is sum(units) / max(time).  No data-path collective: ranks never exchange keys or values."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from rocksplicator_b200 import partition


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_shards, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    from oracle import okv
    from rocksplicator_b200 import synth
    mine = partition.shards_of_rank(n_shards, rank, world)
    dbs = {int(s): okv.Okv(okv.load_port()) for s in mine}
    idx = np.arange(2000, dtype=np.uint64)
    sh = (idx % np.uint64(n_shards)).astype(np.int64)
    keep = partition.owner_of(sh, world) == rank
    b = synth.single_put_batches(synth.keys16(1, idx[keep]), synth.values(1, sh[keep], idx[keep], 0), idx[keep])
    for row, s, t in zip(b, sh[keep], idx[keep]):
        assert dbs[int(s)].apply(row.tobytes(), int(t)) == 0
    units = int(keep.sum())
    # every rank's shards hold exactly their keys; no other rank touched them
    assert sum(d.latest_seq() for d in dbs.values()) == units
    rate = partition.whole_job_rate(units, 0.5 + 0.5 * rank, dist)  # rank 1 is "slower": max time = 1.0 s
    import torch
    tot = torch.tensor([float(units)], dtype=torch.float64)
    dist.all_reduce(tot)
    if rank == 0:
        out.put((rate, float(tot.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_partition_and_aggregate():
    world, n_shards = 2, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_shards, q)) for r in range(world)]
    for p in procs:
        p.start()
    rate, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == 2000
    assert abs(rate - 2000 / 1.0) < 1e-6  # sum of units / max over ranks


def test_partition_and_route():
    assert list(partition.shards_of_rank(10, 1, 4)) == [1, 5, 9]
    assert partition.db_name("segment", 7) == "segment00007"
    ids = np.array([5, 0, 3, 8, 1, 4])
    order, counts = partition.route(ids, 4)
    assert list(counts) == [3, 2, 0, 1]
    assert list(ids[order]) == [0, 8, 4, 5, 1, 3]  # grouped by owner, original order kept within a group
