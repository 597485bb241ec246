"""shard -> GPU partitioning: shards are independent DBs ("segment%05d",
common/segment_utils.cpp:26-29), so ranks own disjoint shard sets and no data-path collective exists.
Host logic only (tested on CPU with gloo, world_size 2)."""
import numpy as np


def owner_of(shard_id, world_size):
    """gpu = shard_id % n_gpus"""
    return np.asarray(shard_id) % world_size


def shards_of_rank(n_shards_total, rank, world_size):
    return np.arange(rank, n_shards_total, world_size)


def db_name(segment, shard_id):
    return "%s%05d" % (segment, shard_id)


def route(shard_ids, world_size):
    """Bucket a cross-shard request by owning rank: returns (order, counts) with `order` a stable permutation
    grouping the positions by rank — what a router does before calling rsp_multi_get once per GPU."""
    owners = owner_of(shard_ids, world_size)
    order = np.argsort(owners, kind="stable")
    counts = np.bincount(owners, minlength=world_size)
    return order, counts


def whole_job_rate(units_this_rank, seconds_this_rank, dist=None):
    """bench.py's aggregation: all ranks' units / max over ranks' time."""
    if dist is None or not dist.is_initialized():
        return units_this_rank / seconds_this_rank
    import torch
    t = torch.tensor([float(seconds_this_rank)], dtype=torch.float64)
    u = torch.tensor([float(units_this_rank)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item() / t.item())
