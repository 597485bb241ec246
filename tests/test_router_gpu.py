"""One process, several engines: the router (include/rsp_b200.h rsp_router_*, SURVEY §8e / counter_router.cpp:36-66).
Cross-shard batches are bucketed by engine, every engine runs its part on its own device, results come back in the
caller's order — checked against one oracle per shard.  Needs two CUDA devices (`gpurun --gpus 2`); on the CPU suite the
same test body runs against the emulation with two emulated devices (tests/test_emul_cpu.py)."""
import ctypes as C
import random
import struct

import numpy as np
import pytest

from oracle import okv
from rocksplicator_b200.write_batch import WriteBatch
from streams import bench_key, bench_value

pytestmark = pytest.mark.gpu


def _device_count(lib_path):
    import os
    if os.environ.get("RSP_TEST_EMUL_LIB"):
        return int(os.environ.get("RSP_EMUL_DEVICES", "1"))
    import torch
    return torch.cuda.device_count()


def test_two_engines_behind_one_router(port_lib):
    from rocksplicator_b200 import engine
    if _device_count(engine.SO_PATH) < 2:
        pytest.skip("needs two CUDA devices")
    engs = [engine.Engine(0), engine.Engine(1)]
    router = engine.Router(engs)
    n_shards, n_keys = 10, 3000
    shards, oracles = [], []
    for g in range(n_shards):  # shard_id % n_gpus -> GPU
        s = engs[g % 2].open_shard("segment%05d" % g, merge_op=engine.MERGE_COUNTER)
        router.add_shard(g, s)
        shards.append(s)
        oracles.append(okv.Okv(port_lib, merge_op=okv.MERGE_COUNTER))
    rnd = random.Random(7)
    for rnd_no in range(3):
        ids, batches, ts = [], [], []
        for i in range(n_keys):
            g = rnd.randrange(n_shards)
            wb = WriteBatch().put(bench_key(2, i), bench_value(2, g, i, rnd_no))
            if i % 4 == 0:
                wb.merge(b"ctr%d" % (i % 40), struct.pack("<q", i + rnd_no))
            if i % 9 == 0:
                wb.delete(bench_key(2, (i * 5) % n_keys))
            ids.append(g); batches.append(wb.data()); ts.append(1000 + i)
        st = router.apply_many(ids, batches, ts)
        assert not st.any(), st
        for g, b, t in zip(ids, batches, ts):
            assert oracles[g].apply(b, t) == 0
        # maintenance on BOTH devices in one process: flush, then a full compaction (k_compact_sort's shared-memory
        # opt-in is per device)
        if rnd_no == 0:
            for e in engs:
                assert e.flush_all() == 0
        if rnd_no == 1:
            for e in engs:
                assert e.compact_all() == 0
    for g in range(n_shards):
        assert shards[g].latest_seq() == oracles[g].latest_seq()
    # one cross-shard MultiGet: every key asked of every shard, interleaved; unknown shard id -> InvalidArgument
    keys = [bench_key(2, i) for i in range(0, n_keys, 7)] + [b"ctr%d" % i for i in range(40)] + [b"absent"]
    q_ids, q_keys = [], []
    for k in keys:
        for g in range(n_shards):
            q_ids.append(g); q_keys.append(k)
    q_ids.append(777); q_keys.append(b"x")
    got, rc = router.multi_get(q_ids, q_keys, stride=64)
    want = {g: dict(zip(keys, oracles[g].multi_get(keys))) for g in range(n_shards)}
    for (g, k), res in zip(zip(q_ids[:-1], q_keys[:-1]), got[:-1]):
        assert res == want[g][k], (g, k, res, want[g][k])
    assert got[-1][0] == engine.INVALID_ARGUMENT
    # fixed-shape form (the fast kernel on each device)
    idx = np.arange(0, n_keys, 3)
    qk = np.frombuffer(b"".join(bench_key(2, int(i)) for i in idx for _ in range(n_shards)), dtype=np.uint8).copy()
    qi = np.array([g for _ in idx for g in range(n_shards)], dtype=np.uint32)
    nq = qi.size
    vals = np.zeros((nq, 64), dtype=np.uint8); vlen = np.zeros(nq, dtype=np.uint32); st = np.full(nq, -1, dtype=np.int32)
    assert router.multi_get_fixed(qi, qk, 16, vals.reshape(-1), 64, vlen, st) == 0
    j = 0
    for i in idx:
        k = bench_key(2, int(i))
        for g in range(n_shards):
            wst, wv = oracles[g].get(k)
            assert int(st[j]) == wst, (g, i, st[j], wst)
            if wst == 0:
                assert vals[j, :vlen[j]].tobytes() == wv
            j += 1
    router.close()
    for e in engs:
        e.close()
