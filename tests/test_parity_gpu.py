"""GPU parity: the CUDA engine, called through the C ABI, against
  (1) fixtures generated from the reference's own RocksDB binary (tests/golden, oracle/gen_golden.py),
  (2) the oracle port on fresh seeded streams, incl. flush/compaction at arbitrary points,
  (3) batched multi-shard ticks (rsp_apply_many / rsp_multi_get) vs per-shard oracle replay.
Bit-exact: byte/integer work only.
"""
import random
import struct

import numpy as np
import pytest

import golden_util as G
from oracle import okv
from rocksplicator_b200.write_batch import WriteBatch
from streams import bench_key, bench_value, corrupt_cases, random_stream

pytestmark = pytest.mark.gpu
u64 = lambda x: struct.pack("<Q", x)  # noqa: E731


@pytest.fixture(scope="module")
def eng():
    from rocksplicator_b200 import engine
    e = engine.Engine(0)
    yield e
    e.close()


_n = [0]


def new_shard(eng, merge_op=0, **kw):
    _n[0] += 1
    return eng.open_shard("t%05d" % _n[0], merge_op=merge_op, **kw)


@pytest.mark.parametrize("case", G.load("streams.json"), ids=lambda c: c["name"])
def test_golden_streams(eng, case):
    s = new_shard(eng, G.MERGE_IDS[case["merge"]])
    G.replay_stream_case(s, case)
    s.close()


@pytest.mark.parametrize("case", G.load("streams.json")[:6], ids=lambda c: c["name"])
def test_golden_streams_memtable_only(eng, case):
    """same streams, never flushed: reads served by the hash memtable + version chains"""
    s = new_shard(eng, G.MERGE_IDS[case["merge"]])
    keys = [bytes.fromhex(k) for k in case["keys"]]
    for st in case["steps"]:
        assert s.apply(bytes.fromhex(st["batch"]), st["ts"]) == st["rc"]
        assert s.latest_seq() == st["seq"]
    probe = keys + [b"zz-missing"] + keys[:3]
    for (khex, rc, vhex), k in zip(case["final"]["get"], probe):
        assert s.get(k) == (rc, G.unhex(vhex)), khex
    assert s.multi_get(probe) == [(rc, G.unhex(v)) for rc, v in case["final"]["multi_get"]]
    s.close()


def test_golden_corrupt(eng):
    pre = WriteBatch().put(b"pre", b"x").data()
    for c in G.load("corrupt.json"):
        s = new_shard(eng, okv.MERGE_UINT64ADD)
        assert s.apply(pre, 1) == 0
        assert s.apply(bytes.fromhex(c["batch"]), 0x1122334455667788) == c["rc"], c["name"]
        if c["rc"]:
            assert s.last_error == c["msg"], c["name"]
        assert s.latest_seq() == c["seq"], c["name"]
        assert s.apply(pre, 2) == c["rc_after"], c["name"]  # the error latch
        assert s.latest_seq() == c["seq_after"], c["name"]
        assert [[k.hex(), v.hex()] for k, v in s.scan()] == c["scan"], c["name"]
        s.close()


def test_known_answers(eng):
    want = {r[0]: r[1:] for r in G.load("known_answers.json")}
    s = new_shard(eng, okv.MERGE_UINT64ADD)
    assert [s.latest_seq()] == want["fresh_seq"]
    b = WriteBatch().put(b"k1", b"v1").delete(b"k2").merge(b"c", u64(5)).put_log_data(u64(1234)).set_sequence(999)
    assert [s.apply(b.data(), 5)] == want["row2_rc"]
    assert [s.latest_seq()] == want["row2_seq"]
    s.apply(WriteBatch().put_log_data(u64(1)).data(), 5)
    assert [s.latest_seq()] == want["row3_seq"]
    s.apply(bytes(12), 5)
    assert [s.latest_seq()] == want["row4_seq"]
    s.apply(WriteBatch().merge(b"c", u64(7)).data(), 5)
    assert [s.get(b"c")[1].hex()] == want["row5_get_c"]
    s.apply(WriteBatch().merge(b"k1", u64(1)).data(), 5)
    assert [s.get(b"k1")[1].hex()] == want["row6_get_k1"]
    s.apply(WriteBatch().put(b"z", u64(100)).delete(b"z").merge(b"z", u64(3)).merge(b"z", u64(4)).data(), 5)
    assert [s.latest_seq()] == want["row7_seq"]
    assert [s.get(b"z")[1].hex()] == want["row7_get_z"]
    s.apply(WriteBatch().put(b"x", b"1").put(b"x", b"2").delete(b"x").put(b"x", b"3").delete(b"y").data(), 5)
    assert [s.latest_seq()] == want["row8_seq"]
    assert [s.get(b"x")[1].hex()] == want["row8_get_x"]
    assert [s.get(b"y")[0]] == want["row8_get_y_rc"]
    s.apply(WriteBatch().put(b"", b"").put(b"ev", b"").data(), 5)
    assert [s.get(b"")[0], s.get(b"")[1].hex()] == want["row9_get_empty"]
    assert [[[k.hex(), v.hex()] for k, v in s.scan()]] == want["row10_scan"]
    assert [[[k.hex(), v.hex()] for k, v in s.scan(start=b"k", limit=1)]] == want["row10_seek_k"]
    mg = [[rc, v.hex() if v is not None else None] for rc, v in s.multi_get([b"c", b"zz", b"c", b"k2", b"ev", b""])]
    assert [mg] == want["row11_multi_get"]
    s.flush()
    assert [[[k.hex(), v.hex()] for k, v in s.scan()]] == want["row12_scan_after_flush"]
    s.apply(WriteBatch().single_delete(b"ev").data(), 5)
    assert [s.latest_seq()] == want["row18_seq"]
    assert [s.get(b"ev")[0]] == want["row18_get_ev_rc"]
    s.close()
    s = new_shard(eng, okv.MERGE_NONE)
    assert [s.apply(WriteBatch().merge(b"m", b"1").data(), 5)] == want["row13_write_rc"]
    assert [s.latest_seq()] == want["row13_seq"]
    rc, _ = s.get(b"m")
    assert [rc, s.last_error] == want["row13_get_rc"]
    s.close()


def compare_all(s, o, keys, tag):
    assert s.latest_seq() == o.latest_seq(), tag
    probe = keys + [b"zz-missing"]
    for k in probe:
        assert s.get(k) == o.get(k), (tag, k)
    assert s.multi_get(probe + probe[:5]) == o.multi_get(probe + probe[:5]), tag
    assert s.scan() == o.scan(), tag
    a, b = s.iterator(), o.iterator()
    a.seek_to_last()
    b.seek_to_last()
    while True:
        assert a.valid() == b.valid(), tag
        if not a.valid():
            break
        assert (a.key(), a.value()) == (b.key(), b.value()), tag
        a.prev()
        b.prev()
    rng = random.Random(5)
    for _ in range(12):
        k = rng.choice(keys) if rng.random() < 0.6 else bytes(rng.getrandbits(8) for _ in range(rng.randint(0, 4)))
        a.seek(k)
        b.seek(k)
        for _step in range(5):
            assert a.valid() == b.valid(), (tag, "seek", k)
            if not a.valid():
                break
            assert (a.key(), a.value()) == (b.key(), b.value()), (tag, "seek", k)
            if rng.random() < 0.5:
                a.next(), b.next()
            else:
                a.prev(), b.prev()
    a.close()
    b.close()


@pytest.mark.parametrize("merge,mname", [(okv.MERGE_COUNTER, "counter"), (okv.MERGE_APPEND, "append"),
                                         (okv.MERGE_UINT64ADD, "counter"), (okv.MERGE_NONE, None)])
def test_random_streams_vs_oracle(eng, port_lib, merge, mname):
    """rocksdb_assumption_test.cpp:329-432 restated: the same shuffled stream through the oracle and the
    engine gives equal sequence numbers and equal reads, with flush/compaction at arbitrary points."""
    for seed in range(4):
        rng = random.Random(seed)
        keys, stream = random_stream(500 + seed * 13 + merge, 150, merge=mname, bad_operands=(seed == 3))
        s = new_shard(eng, merge, write_buffer_bytes=(4096 if seed == 1 else 0))
        o = okv.Okv(port_lib, merge_op=merge)
        for i, (bt, ts) in enumerate(stream):
            assert s.apply(bt, ts) == o.apply(bt, ts), (seed, i)
            r = rng.random()
            if r < 0.04:
                s.flush()
            elif r < 0.06:
                s.compact()
            if i % 50 == 49:
                compare_all(s, o, keys, (merge, seed, i))
        s.compact()
        compare_all(s, o, keys, (merge, seed, "final"))
        st = s.stats()
        assert st["latest_seq"] == o.latest_seq()
        s.close()
        o.close()


def test_replicator_test_vectors(eng):
    """rocksdb_replicator_test.cpp:146-208 (100 x 2 Puts -> seq 200) and
    rocksdb_assumption_test.cpp:136-187 (each op +1, batch of n +n)."""
    s = new_shard(eng, okv.MERGE_APPEND)
    for i in range(100):
        t = str(i).encode()
        assert s.apply(WriteBatch().put(t + b"key", t + b"value").put(t + b"key2", t + b"value2").data(), i) == 0
        assert s.latest_seq() == 2 * (i + 1)
    for i in range(100):
        t = str(i).encode()
        assert s.get(t + b"key") == (0, t + b"value")
        assert s.get(t + b"key2") == (0, t + b"value2")
    seq = s.latest_seq()
    s.apply(WriteBatch().delete(b"a").put(b"b", b"1").put(b"c", b"2").merge(b"b", b"3").data(), 0)
    assert s.latest_seq() == seq + 4
    assert s.get(b"b") == (0, b"13")
    s.close()


def test_apply_many_multi_shard(eng, port_lib):
    """one tick = many batches for many shards (the batching front-end), vs per-shard oracle replay;
    includes a corrupt batch mid-tick: it and every later batch of THAT shard fail, others proceed."""
    n_shards = 37
    shards = [new_shard(eng, okv.MERGE_COUNTER) for _ in range(n_shards)]
    oracles = [okv.Okv(port_lib, merge_op=okv.MERGE_COUNTER) for _ in range(n_shards)]
    rng = random.Random(99)
    all_keys = [[] for _ in range(n_shards)]
    for tick in range(6):
        six, batches, ts = [], [], []
        for _ in range(400):
            j = rng.randrange(n_shards)
            wb = WriteBatch()
            for _ in range(rng.randint(1, 4)):
                k = bench_key(7, rng.randrange(200))
                all_keys[j].append(k)
                r = rng.random()
                if r < 0.6:
                    wb.put(k, bench_value(7, j, rng.randrange(200), tick))
                elif r < 0.8:
                    wb.merge(k, struct.pack("<q", rng.randint(-9, 9)))
                else:
                    wb.delete(k)
            b = wb.data()
            if tick == 3 and j == 5 and rng.random() < 0.3:
                b = b[:-2]  # corrupt: shard 5 latches
            six.append(shards[j].index)
            batches.append(b)
            ts.append(rng.getrandbits(40))
        st = eng.apply_many(six, batches, ts)
        want = [oracles[[s.index for s in shards].index(ix)].apply(b, t) for ix, b, t in zip(six, batches, ts)]
        assert list(st) == want, tick
    for j in range(n_shards):
        assert shards[j].latest_seq() == oracles[j].latest_seq()
        keys = sorted(set(all_keys[j]))
        assert shards[j].multi_get(keys) == oracles[j].multi_get(keys)
    # cross-shard MultiGet in one call
    six, keys = [], []
    for j in range(n_shards):
        for k in sorted(set(all_keys[j]))[:20]:
            six.append(shards[j].index)
            keys.append(k)
    got = eng.multi_get(six, keys, stride=64)
    want = []
    for ix, k in zip(six, keys):
        j = [s.index for s in shards].index(ix)
        want.append(oracles[j].get(k))
    assert got == want
    for s in shards:
        s.close()


def test_scale_properties(eng, port_lib):
    """size-independent properties at a scale the oracle replays in seconds: load N keys over S shards
    through the apply path, overwrite a third, delete a tenth; every shard compacted; then
    (a) every live key reads its newest value, deleted keys NotFound, (b) scans are sorted, complete and
    equal to the oracle's, (c) sequence numbers equal the op counts, (d) MultiGet == Get."""
    S, N = 16, 40000
    shards = [new_shard(eng, 0, write_buffer_bytes=1 << 18) for _ in range(S)]
    oracles = [okv.Okv(port_lib) for _ in range(S)]
    seed = 0x5EED0001
    idx = np.arange(N)
    for rnd, sel in enumerate((idx, idx[::3])):
        six, batches, ts = [], [], []
        for i in sel:
            j = int(i) % S
            six.append(shards[j].index)
            batches.append(WriteBatch().put(bench_key(seed, int(i)), bench_value(seed, j, int(i), rnd)).data())
            ts.append(1000 + int(i))
        for lo in range(0, len(six), 8192):
            st = eng.apply_many(six[lo:lo + 8192], batches[lo:lo + 8192], ts[lo:lo + 8192])
            assert not st.any()
        for ix, b, t in zip(six, batches, ts):
            oracles[[s.index for s in shards].index(ix)].apply(b, t)
    dels = idx[::10]
    six = [shards[int(i) % S].index for i in dels]
    batches = [WriteBatch().delete(bench_key(seed, int(i))).data() for i in dels]
    assert not eng.apply_many(six, batches, [0] * len(six)).any()
    for i, b in zip(dels, batches):
        oracles[int(i) % S].apply(b, 0)
    for j in range(S):
        assert shards[j].latest_seq() == oracles[j].latest_seq()
    eng.compact_all()
    for j in range(S):
        got = shards[j].scan()
        assert got == oracles[j].scan()
        assert all(got[i][0] < got[i + 1][0] for i in range(len(got) - 1))
        assert shards[j].stats()["n_runs"] == 1
    keys = [bench_key(seed, int(i)) for i in idx[::7]]
    six = [shards[int(i) % S].index for i in idx[::7]]
    got = eng.multi_get(six, keys, stride=64)
    for (rc, v), i in zip(got, idx[::7]):
        want = oracles[int(i) % S].get(bench_key(seed, int(i)))
        assert (rc, v) == want
    # batched scans: Seek + 128 x Next
    starts = [bench_key(seed, int(i)) for i in idx[5::5000]]
    six = [shards[int(i) % S].index for i in idx[5::5000]]
    res = eng.multi_scan(six, starts, 128, 128 * (8 + 16 + 64))
    for (st, recs), i, k in zip(res, idx[5::5000], starts):
        assert st == 0
        assert recs == oracles[int(i) % S].scan(start=k, limit=128)
    for s in shards:
        s.close()


def test_hot_keys_one_tick(eng, port_lib):
    """thousands of batches in ONE tick hammering the same few keys (Put / Merge / Delete mixed): the
    lock-free, sequence-ordered version-chain insert must leave every key exactly as serial replay does."""
    s = new_shard(eng, okv.MERGE_COUNTER, write_buffer_bytes=4 << 20)
    o = okv.Okv(port_lib, merge_op=okv.MERGE_COUNTER)
    rng = random.Random(7)
    keys = [b"hot%d" % i for i in range(5)] + [bench_key(3, i) for i in range(3)]
    for rnd in range(3):
        batches = []
        for _ in range(3000):
            wb = WriteBatch()
            for _ in range(rng.randint(1, 3)):
                k = rng.choice(keys)
                r = rng.random()
                if r < 0.3:
                    wb.put(k, struct.pack("<q", rng.randint(-5, 5)))
                elif r < 0.9:
                    wb.merge(k, struct.pack("<q", rng.randint(-5, 5)))
                else:
                    wb.delete(k)
            batches.append(wb.data())
        st = eng.apply_many([s.index] * len(batches), batches, list(range(len(batches))))
        assert not st.any()
        for i, b in enumerate(batches):
            assert o.apply(b, i) == 0
        assert s.latest_seq() == o.latest_seq()
        for k in keys:
            assert s.get(k) == o.get(k), (rnd, k)
        if rnd == 1:
            s.flush()
    assert s.scan() == o.scan()
    s.close()
    o.close()


def test_concurrent_callers(eng, port_lib):
    """the C ABI is thread-safe (rocksdb_replicator.h:80-82): 8 threads apply to their own shards and read
    concurrently; each shard ends equal to its oracle."""
    import threading
    n = 8
    shards = [new_shard(eng, okv.MERGE_COUNTER) for _ in range(n)]
    streams = [random_stream(800 + i, 120, merge="counter") for i in range(n)]
    errs = []

    def work(i):
        try:
            keys, stream = streams[i]
            for j, (bt, ts) in enumerate(stream):
                shards[i].apply(bt, ts)
                if j % 10 == 0:
                    shards[i].get(keys[j % len(keys)])
                if j == 60:
                    shards[i].flush()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for i in range(n):
        o = okv.Okv(port_lib, merge_op=okv.MERGE_COUNTER)
        keys, stream = streams[i]
        for bt, ts in stream:
            o.apply(bt, ts)
        assert shards[i].latest_seq() == o.latest_seq()
        assert shards[i].multi_get(keys) == o.multi_get(keys)
        assert shards[i].scan() == o.scan()
        o.close()
        shards[i].close()


def _mgf(eng, six, keys, stride):
    """rsp_multi_get_fixed (the 16-byte-key kernel k_multi_get16 + pending list) -> [(st, value|None)]"""
    n = len(keys)
    k = np.frombuffer(b"".join(keys), dtype=np.uint8).copy()
    s = np.ascontiguousarray(six, dtype=np.uint32)
    vals = np.zeros(n * stride, dtype=np.uint8)
    vlen = np.zeros(n, dtype=np.uint32)
    st = np.zeros(n, dtype=np.int32)
    assert eng.multi_get_fixed(s, k, 16, vals, stride, vlen, st) == 0
    return [(int(st[i]), vals[i * stride:i * stride + vlen[i]].tobytes() if st[i] == 0 else None) for i in range(n)], vlen


def test_fixed_key_kernel_shapes(eng, port_lib):
    """the 2-lane 16-byte-key kernel on everything but the benchmark shape: 256-byte and odd-sized values,
    Deletes, counter Merges, versions still in the memtable, several runs, misses, too-small output stride."""
    s = new_shard(eng, okv.MERGE_COUNTER, write_buffer_bytes=1 << 20)
    o = okv.Okv(port_lib, merge_op=okv.MERGE_COUNTER)
    rng = random.Random(21)
    keys = [bench_key(9, i) for i in range(600)]

    def tick(sel, mk):
        bs = [mk(k) for k in sel]
        st = eng.apply_many([s.index] * len(bs), bs, [1] * len(bs))
        assert not st.any()
        for b in bs:
            assert o.apply(b, 1) == 0

    def check(tag):
        probe = keys + [bench_key(9, 10_000 + i) for i in range(50)]
        got, _ = _mgf(eng, [s.index] * len(probe), probe, 512)
        assert got == o.multi_get(probe), tag

    tick(keys, lambda k: WriteBatch().put(k, bytes(rng.getrandbits(8) for _ in range(256))).data())
    check("memtable 256B")
    s.flush()
    check("one run 256B")
    tick(keys[::3], lambda k: WriteBatch().put(k, bytes(rng.getrandbits(8) for _ in range(rng.choice([0, 1, 10, 64, 100, 300])))).data())
    check("run + memtable, odd sizes")
    s.flush()
    tick(keys[::5], lambda k: WriteBatch().delete(k).data())
    tick(keys[::7], lambda k: WriteBatch().merge(k, struct.pack("<q", 5)).merge(k, struct.pack("<q", 7)).data())
    tick(keys[1::7], lambda k: WriteBatch().put(k, struct.pack("<q", 100)).merge(k, struct.pack("<q", -1)).data())
    check("two runs + memtable, deletes and merges")
    s.compact()
    check("compacted")
    # stride smaller than some values: those report INCOMPLETE with the size needed
    got, vlen = _mgf(eng, [s.index] * len(keys), keys, 64)
    want = o.multi_get(keys)
    for (st, v), (wst, wv), vl in zip(got, want, vlen):
        if wst == 0 and len(wv) > 64:
            assert st == 7 and vl == len(wv)
        else:
            assert (st, v) == (wst, wv)
    # an unknown shard id answers InvalidArgument, not a crash
    got, _ = _mgf(eng, [60000, s.index], [keys[0], keys[1]], 512)
    assert got[0][0] == 4 and got[1] == o.get(keys[1])
    s.close()
    o.close()


def test_packed_tick_virtual_trailer(eng, port_lib):
    """>= 1024 batches already grouped by shard take the packed path (descriptors derived on the device, the
    follower's LogData(timestamp) record is virtual).  Includes a batch whose last value legally SWALLOWS the
    first bytes of that record (the rest of the timestamp parses as Noop tags), and corrupt batches that latch."""
    from rocksplicator_b200.write_batch import varint32
    a, b, c = (new_shard(eng, okv.MERGE_COUNTER) for _ in range(3))
    oa, ob, oc = (okv.Okv(port_lib, merge_op=okv.MERGE_COUNTER) for _ in range(3))
    rng = random.Random(77)
    six, batches, ts = [], [], []

    def add(sh, data, t):
        six.append(sh.index)
        batches.append(data)
        ts.append(t)

    for i in range(700):
        wb = WriteBatch().put(bench_key(5, i), bench_value(5, 0, i, 0))
        if i % 7 == 0:
            wb.merge(b"ctr", struct.pack("<q", i))
        add(a, wb.data(), 1000 + i)
    swallow_ts = int.from_bytes(bytes([0x41] + [0x0D] * 7), "little")
    for i in range(400):
        if i == 200:
            v = b"tail-swallows-"
            raw = bytes(8) + struct.pack("<I", 1) + b"\x01" + varint32(3) + b"swk" + varint32(len(v) + 3) + v
            add(b, raw, swallow_ts)
        else:
            add(b, WriteBatch().put(b"k%d" % (i % 50), bytes(rng.getrandbits(8) for _ in range(rng.choice([0, 5, 64, 200])))).data(), 5 + i)
    good = WriteBatch().put(b"x", b"1").data()
    for i in range(300):
        if i == 250:
            add(c, good[:-1], 9)        # truncated: latches shard c
        else:
            add(c, WriteBatch().put(b"c%d" % i, b"v").delete(b"c%d" % (i - 1)).data(), 9)
    st = eng.apply_many(six, batches, ts)
    want = []
    for ix, bt, t in zip(six, batches, ts):
        o = oa if ix == a.index else (ob if ix == b.index else oc)
        want.append(o.apply(bt, t))
    assert list(st) == want
    assert want[700 + 200] == 0 and want[-1] != 0
    for s, o in ((a, oa), (b, ob), (c, oc)):
        assert s.latest_seq() == o.latest_seq()
        assert s.scan() == o.scan()
    assert b.get(b"swk") == ob.get(b"swk") == (0, b"tail-swallows-" + bytes([0x03, 0x08, 0x41]))
    assert c.last_error == oc.last_error
    for s in (a, b, c):
        s.close()


def test_iterator_status_rises_when_the_failing_key_is_reached(eng, port_lib):
    """DBIter's status_ is sticky and is set when the iterator lands on a key whose merge fails (the key stays, with an
    empty value) — not earlier, although the engine fetches entries ahead in chunks.  Found by the emulation fuzzer."""
    s = new_shard(eng, okv.MERGE_COUNTER)
    o = okv.Okv(port_lib, merge_op=okv.MERGE_COUNTER)
    wb = WriteBatch()
    for i in range(40):
        wb.put(b"k%03d" % i, struct.pack("<q", i))
    wb.merge(b"k020", b"xyz")          # operand of the wrong size: the counter operator refuses it
    wb.put(b"k021", b"")               # a legitimately empty value right behind it
    for db in (s, o):
        assert db.apply(wb.data(), 5) == 0
    for flushed in (False, True):
        if flushed:
            s.flush()
        a, b = s.iterator(), o.iterator()
        a.seek_to_first(), b.seek_to_first()
        steps = 0
        while b.valid():
            assert a.valid() and (a.key(), a.value(), a.status()) == (b.key(), b.value(), b.status()), (flushed, steps)
            a.next(), b.next()
            steps += 1
        assert not a.valid() and a.status() == b.status() != 0 and steps == 40
        a.close(), b.close()
        a, b = s.iterator(), o.iterator()
        a.seek_to_last(), b.seek_to_last()
        for _ in range(25):
            assert (a.valid(), a.key(), a.value(), a.status()) == (b.valid(), b.key(), b.value(), b.status()), flushed
            a.prev(), b.prev()
        a.seek(b"k030"), b.seek(b"k030")   # the status is sticky: a later Seek does not clear it
        assert (a.key(), a.status()) == (b.key(), b.status())
        a.close(), b.close()
        # the batched scan reports the failure for the scan and an empty value for the key
        st, recs = eng.multi_scan([s.index], [b"k018"], 5, 4096)[0]
        assert st == 2 and recs == [(b"k018", struct.pack("<q", 18)), (b"k019", struct.pack("<q", 19)), (b"k020", b""),
                                    (b"k021", b""), (b"k022", struct.pack("<q", 22))]
    s.close()
    o.close()


@pytest.mark.parametrize("merge", [okv.MERGE_APPEND, okv.MERGE_COUNTER])
def test_long_scan_over_merged_keys_with_large_values(eng, port_lib, merge):
    """A fetch-ahead chunk that runs out of room while it already carries another status (a key to fold on the host, a
    failed merge) must still be continued: the iterator once stopped there.  Found by the emulation fuzzer."""
    s = new_shard(eng, merge)
    o = okv.Okv(port_lib, merge_op=merge)
    rnd = random.Random(11)
    for c in range(8):
        wb = WriteBatch()
        for i in range(60):
            k = b"key%04d" % (c * 60 + i)
            wb.put(k, rnd.randbytes(rnd.choice([8, 300, 900])))
            if i % 7 == 0:
                wb.merge(k, b"tail" if merge == okv.MERGE_APPEND else b"xyz")   # append folds on the host; counter refuses
        for db in (s, o):
            assert db.apply(wb.data(), c) == 0
        if c == 3:
            s.flush()
    a, b = s.iterator(), o.iterator()
    a.seek_to_first(), b.seek_to_first()
    n = 0
    while b.valid():
        assert a.valid() and (a.key(), a.value(), a.status()) == (b.key(), b.value(), b.status()), n
        a.next(), b.next()
        n += 1
    assert not a.valid() and n == 480
    a.seek_to_last(), b.seek_to_last()
    n = 0
    while b.valid():
        assert a.valid() and (a.key(), a.value()) == (b.key(), b.value()), n
        a.prev(), b.prev()
        n += 1
    assert not a.valid() and n == 480
    a.close(), b.close()
    s.close()
    o.close()


def test_range_deletion_is_refused(eng, port_lib):
    """The one deliberate difference from RocksDB on this path: a well-formed DeleteRange record
    — which the reference never issues — is refused with NotSupported, whole batch, nothing applied; RocksDB would
    apply a range tombstone.  Malformed ones fail with RocksDB's own error (tests/golden/corrupt.json)."""
    for bt in (WriteBatch().put(b"k", b"v").data()[:8] + struct.pack("<I", 2) + b"\x01\x01k\x01v" + b"\x0f\x01a\x01z",
               bytes(8) + struct.pack("<I", 1) + b"\x0e\x00\x01a\x01z"):
        s = new_shard(eng)
        o = okv.Okv(port_lib)
        for db in (s, o):
            assert db.apply(WriteBatch().put(b"pre", b"x").data(), 1) == 0
            assert db.apply(bt, 2) == 3
            assert db.last_error == "Not implemented: WriteBatch tag outside the replicated hot path"
            assert db.latest_seq() == 1 and db.scan() == [(b"pre", b"x")]
        assert s.apply(WriteBatch().put(b"q", b"y").data(), 3) == o.apply(WriteBatch().put(b"q", b"y").data(), 3)
        assert s.latest_seq() == o.latest_seq() and s.scan() == o.scan()
        s.close()
        o.close()


@pytest.mark.parametrize("shape", ["distinct_prefixes", "shared_prefix", "short_and_padded", "versions"])
def test_flush_sort_paths(eng, port_lib, shape):
    """The flush sorts the memtable by an LSD radix sort over the keys' 8-byte prefixes (k_flush_sort) and leaves a
    shard to the comparison sort when distinct keys share their prefix: both paths, the boundary between them (keys
    shorter than 8 bytes vs the same bytes zero-extended) and many versions per key (stability = newest first), each
    flushed at several sizes and read back through scans and MultiGet against the oracle."""
    rnd = random.Random(hash(shape) & 0xffff)
    for n in (1, 2, 31, 33, 700, 5000):
        s = new_shard(eng, 0)
        o = okv.Okv(port_lib)
        if shape == "distinct_prefixes":
            keys = [struct.pack(">Q", rnd.getrandbits(40)) + b"tail%d" % i for i in range(n)]
        elif shape == "shared_prefix":
            keys = [b"user_profile_%06d" % rnd.randrange(10 ** 6) for _ in range(n)]
        elif shape == "short_and_padded":
            base = [bytes([rnd.randrange(97, 100) for _ in range(rnd.randrange(1, 8))]) for _ in range(n)]
            keys = base + [k + b"\x00" * rnd.randrange(1, 4) for k in base[::2]]
        else:
            pool = [struct.pack(">Q", rnd.getrandbits(24)) + b"k" for _ in range(max(1, n // 8))]
            keys = [rnd.choice(pool) for _ in range(n)]
        six, batches, ts = [], [], []
        for i, k in enumerate(keys):
            wb = WriteBatch()
            if shape == "versions" and i % 5 == 4:
                wb.delete(k)
            else:
                wb.put(k, b"v%d-" % i + k[:6])
            six.append(s.index); batches.append(wb.data()); ts.append(i)
        for lo in range(0, len(six), 4096):
            assert not eng.apply_many(six[lo:lo + 4096], batches[lo:lo + 4096], ts[lo:lo + 4096]).any()
        for b, t in zip(batches, ts):
            assert o.apply(b, t) == 0
        assert s.flush() == 0
        assert s.latest_seq() == o.latest_seq()
        generic = s.stats()["flush_comparison_sorts"]
        if shape in ("distinct_prefixes", "versions"):
            assert generic == 0, "the radix sort should have handled this memtable"
        elif shape == "shared_prefix" and n >= 31:
            assert generic == 1, "distinct keys share their 8-byte prefix: the comparison sort must take over"
        got = s.scan()
        assert got == o.scan(), (shape, n)
        probe = list(dict.fromkeys(keys))[:300] + [b"zz-missing"]
        assert s.multi_get(probe) == o.multi_get(probe)
        s.close()


@pytest.mark.parametrize("grouped", [False, True])
def test_long_groups_chunked_tick(eng, port_lib, grouped):
    """Ticks whose groups are longer than one chunk (k_tick_chunks: a CTA per chunk of <= 128 batches / 16 KB, the
    sequencing state handed from chunk to chunk): batch sizes that cut chunks by bytes as well as by count, multi-op
    batches, a corrupt batch deep inside one shard's group (it and everything after it on THAT shard fail with the
    latched status, chunks later), a second tick on the latched shard; per-batch statuses, sequence numbers and
    contents against the oracle.  grouped: the caller's batches already sit shard by shard (the packed tick)."""
    rnd = random.Random(4242 + grouped)
    n_shards = 3
    shards = [new_shard(eng, okv.MERGE_COUNTER, write_buffer_bytes=8 << 20) for _ in range(n_shards)]
    oracles = [okv.Okv(port_lib, merge_op=okv.MERGE_COUNTER) for _ in range(n_shards)]
    keys_of = [set() for _ in range(n_shards)]
    for tick in range(2):
        per = [[] for _ in range(n_shards)]
        for j in range(n_shards):
            for i in range(1500 if j else 700):
                wb = WriteBatch()
                for _ in range(1 if rnd.random() < 0.8 else rnd.randint(2, 5)):
                    k = bench_key(11, rnd.randrange(3000))
                    keys_of[j].add(k)
                    r = rnd.random()
                    if r < 0.7:
                        wb.put(k, bytes([rnd.randrange(256)]) * (rnd.choice((8, 64, 64, 64, 300, 3000)) if r < 0.05 else 64))
                    elif r < 0.85:
                        wb.merge(k, struct.pack("<q", rnd.randint(-5, 5)))
                    else:
                        wb.delete(k)
                b = wb.data()
                if tick == 0 and j == 1 and i == 777:
                    b = b[:-3]  # corrupt: shard 1 latches in the middle of its group
                per[j].append(b)
        if grouped:
            order = [(j, b) for j in range(n_shards) for b in per[j]]
        else:
            order = []
            cursors = [0] * n_shards
            while any(cursors[j] < len(per[j]) for j in range(n_shards)):
                j = rnd.choice([x for x in range(n_shards) if cursors[x] < len(per[x])])
                order.append((j, per[j][cursors[j]]))
                cursors[j] += 1
        six = [shards[j].index for j, _ in order]
        batches = [b for _, b in order]
        ts = [1000 + i for i in range(len(order))]
        st = eng.apply_many(six, batches, ts)
        want = [oracles[j].apply(b, t) for (j, b), t in zip(order, ts)]
        assert list(st) == want, tick
        assert any(want) and not all(want)
    for j in range(n_shards):
        assert shards[j].latest_seq() == oracles[j].latest_seq()
        keys = sorted(keys_of[j])
        assert shards[j].multi_get(keys, stride=4096) == oracles[j].multi_get(keys)
        assert shards[j].scan() == oracles[j].scan()
    for s in shards:
        s.close()


def test_memtable_filter_across_flush_cycles(eng, port_lib):
    """The 16-byte-key MultiGet kernels consult a per-shard filter (one bit per inserted key hash) before they probe a
    memtable, and the flush clears it: a false negative would serve a stale value from a run (or NotFound) for a key whose
    newest version sits in the memtable.  Several rounds of overwrite-some / insert-some / delete-some WITHOUT flushing,
    every key of the shard looked up through the fixed-shape entry point after each round, then a flush (answers must
    not change), single-run and two-run shapes, both kernels (k_multi_get16, k_multi_get16m)."""
    rnd = random.Random(777)
    S, N = 3, 3000
    shards = [new_shard(eng, 0) for _ in range(S)]
    oracles = [okv.Okv(port_lib) for _ in range(S)]
    universe = [[bench_key(5, j * 100000 + i) for i in range(N)] for j in range(S)]

    def tick(ops):
        six, batches, ts = [], [], []
        for j, wb in ops:
            six.append(shards[j].index); batches.append(wb.data()); ts.append(len(six))
        for lo in range(0, len(six), 4096):
            assert not eng.apply_many(six[lo:lo + 4096], batches[lo:lo + 4096], ts[lo:lo + 4096]).any()
        for (j, wb), t in zip(ops, ts):
            assert oracles[j].apply(wb.data(), t) == 0

    def check(tag):
        for j in range(S):
            keys = universe[j] + [bench_key(6, i) for i in range(50)]  # (never written: NotFound)
            n = len(keys)
            vals = np.zeros((n, 64), dtype=np.uint8)
            vlen = np.zeros(n, dtype=np.uint32)
            st = np.full(n, -1, dtype=np.int32)
            six = np.full(n, shards[j].index, dtype=np.uint32)
            assert eng.multi_get_fixed(six, np.frombuffer(b"".join(keys), dtype=np.uint8), 16, vals.reshape(-1), 64, vlen, st) == 0
            want = oracles[j].multi_get(keys)
            for i, (wst, wv) in enumerate(want):
                assert int(st[i]) == wst, (tag, j, i, int(st[i]), wst)
                if wst == 0:
                    assert vals[i, :vlen[i]].tobytes() == wv, (tag, j, i)

    tick([(j, WriteBatch().put(k, bench_value(5, j, i, 0))) for j in range(S) for i, k in enumerate(universe[j][:N // 2])])
    assert eng.compact_all() == 0
    check("compacted")
    for rnd_no in range(1, 5):
        ops = []
        for j in range(S):
            for i in rnd.sample(range(N), 400):  # overwrites of run keys, first writes of new keys
                ops.append((j, WriteBatch().put(universe[j][i], bench_value(5, j, i, rnd_no))))
            for i in rnd.sample(range(N), 60):
                ops.append((j, WriteBatch().delete(universe[j][i])))
        tick(ops)
        check("round %d, memtable" % rnd_no)
        if rnd_no % 2 == 0:
            for s in shards:
                assert s.flush() == 0  # a second run: the multi-run kernel; the filter rows start over
            check("round %d, flushed" % rnd_no)
    assert eng.compact_all() == 0
    check("compacted again")
    for s in shards:
        s.close()
