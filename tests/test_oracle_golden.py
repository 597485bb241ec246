"""Pins the oracle port (oracle/kv_oracle.c) to the reference:
  1. known-answer table measured on the reference's shipped librocksdb.so.5.4,
  2. tests/golden/*.json generated from that binary by oracle/gen_golden.py,
  3. live differential runs against the binary when oracle/_ref is present (build container).
CPU only.
"""
import struct

import pytest

import golden_util as G
from oracle import okv
from rocksplicator_b200.write_batch import WriteBatch
from streams import corrupt_cases, random_stream

u64 = lambda x: struct.pack("<Q", x)  # noqa: E731


def test_wire_format_sample():
    # SURVEY §9 "Wire format (confirmed byte-for-byte)" and §8 a6 probe sample
    b = WriteBatch().put(b"k1", b"v1").delete(b"k2").merge(b"c", u64(5)).put_log_data(u64(1234))
    assert b.data().hex() == ("0000000000000000" "03000000" "01026b31027631" "00026b32"
                              "020163080500000000000000" "0308d204000000000000")
    assert WriteBatch().put(b"key", b"value").data().hex() == "00000000000000000100000001036b65790576616c7565"
    assert WriteBatch().put(b"k" * 300, b"").data()[13:15] == bytes([0xAC, 0x02])


def test_port_known_answers(port_lib):
    """§9 rows 1-13/18 — the expectations are the committed outputs of the reference binary."""
    want = {r[0]: r[1:] for r in G.load("known_answers.json")}
    db = okv.Okv(port_lib, merge_op=okv.MERGE_UINT64ADD)
    assert [db.latest_seq()] == want["fresh_seq"]
    b = WriteBatch().put(b"k1", b"v1").delete(b"k2").merge(b"c", u64(5)).put_log_data(u64(1234)).set_sequence(999)
    assert [db.apply(b.data(), 5)] == want["row2_rc"]
    assert [db.latest_seq()] == want["row2_seq"] == [3]
    assert [db.apply(WriteBatch().put_log_data(u64(1)).data(), 5)] == want["row3_rc"]
    assert [db.latest_seq()] == want["row3_seq"]
    assert [db.apply(bytes(12), 5)] == want["row4_rc"]
    assert [db.latest_seq()] == want["row4_seq"]
    db.apply(WriteBatch().merge(b"c", u64(7)).data(), 5)
    assert [db.get(b"c")[1].hex()] == want["row5_get_c"]
    db.apply(WriteBatch().merge(b"k1", u64(1)).data(), 5)
    assert [db.get(b"k1")[1].hex()] == want["row6_get_k1"]
    db.apply(WriteBatch().put(b"z", u64(100)).delete(b"z").merge(b"z", u64(3)).merge(b"z", u64(4)).data(), 5)
    assert [db.latest_seq()] == want["row7_seq"]
    assert [db.get(b"z")[1].hex()] == want["row7_get_z"]
    db.apply(WriteBatch().put(b"x", b"1").put(b"x", b"2").delete(b"x").put(b"x", b"3").delete(b"y").data(), 5)
    assert [db.latest_seq()] == want["row8_seq"]
    assert [db.get(b"x")[1].hex()] == want["row8_get_x"]
    assert [db.get(b"y")[0]] == want["row8_get_y_rc"]
    db.apply(WriteBatch().put(b"", b"").put(b"ev", b"").data(), 5)
    assert [db.get(b"")[0], db.get(b"")[1].hex()] == want["row9_get_empty"]
    assert [[[k.hex(), v.hex()] for k, v in db.scan()]] == want["row10_scan"]
    assert [[[k.hex(), v.hex()] for k, v in db.scan(start=b"k", limit=1)]] == want["row10_seek_k"]
    mg = [[rc, v.hex() if v is not None else None] for rc, v in db.multi_get([b"c", b"zz", b"c", b"k2", b"ev", b""])]
    assert [mg] == want["row11_multi_get"]
    db.flush()
    assert [[[k.hex(), v.hex()] for k, v in db.scan()]] == want["row12_scan_after_flush"]
    assert [db.latest_seq()] == want["row12_seq"]
    db.apply(WriteBatch().single_delete(b"ev").data(), 5)
    assert [db.latest_seq()] == want["row18_seq"]
    assert [db.get(b"ev")[0]] == want["row18_get_ev_rc"]
    db.close()
    db = okv.Okv(port_lib, merge_op=okv.MERGE_NONE)
    assert [db.apply(WriteBatch().merge(b"m", b"1").data(), 5)] == want["row13_write_rc"]
    assert [db.latest_seq()] == want["row13_seq"]
    rc, _ = db.get(b"m")
    assert [rc, db.last_error] == want["row13_get_rc"]
    db.close()


@pytest.mark.parametrize("case", G.load("streams.json"), ids=lambda c: c["name"])
def test_port_golden_streams(port_lib, case):
    db = okv.Okv(port_lib, merge_op=G.MERGE_IDS[case["merge"]])
    G.replay_stream_case(db, case)
    db.close()


def test_port_golden_corrupt(port_lib):
    pre = WriteBatch().put(b"pre", b"x").data()
    for c in G.load("corrupt.json"):
        db = okv.Okv(port_lib, merge_op=okv.MERGE_UINT64ADD)
        assert db.apply(pre, 1) == 0
        assert db.apply(bytes.fromhex(c["batch"]), 0x1122334455667788) == c["rc"], c["name"]
        assert db.last_error == c["msg"], c["name"]
        assert db.latest_seq() == c["seq"], c["name"]
        assert db.apply(pre, 2) == c["rc_after"], c["name"]
        assert db.last_error == c["msg_after"], c["name"]
        assert db.latest_seq() == c["seq_after"], c["name"]
        assert [[k.hex(), v.hex()] for k, v in db.scan()] == c["scan"], c["name"]
        db.close()


def test_replicator_test_vectors(port_lib):
    """rocksdb_replicator/tests/rocksdb_replicator_test.cpp:146-208: 100 batches x 2 Puts -> follower seq
    200 and every key reads back; rocksdb_assumption_test.cpp:136-187: Put/Delete/Merge +1, batch of n +n."""
    db = okv.Okv(port_lib, merge_op=okv.MERGE_APPEND)
    for i in range(100):
        s = str(i).encode()
        wb = WriteBatch().put(s + b"key", s + b"value").put(s + b"key2", s + b"value2")
        assert db.apply(wb.data(), i) == 0
        assert db.latest_seq() == 2 * (i + 1)
    for i in range(100):
        s = str(i).encode()
        assert db.get(s + b"key") == (0, s + b"value")
        assert db.get(s + b"key2") == (0, s + b"value2")
    seq = db.latest_seq()
    db.apply(WriteBatch().delete(b"a").put(b"b", b"1").put(b"c", b"2").merge(b"b", b"3").data(), 0)
    assert db.latest_seq() == seq + 4
    assert db.get(b"b") == (0, b"13")
    db.close()


def test_port_vs_reference_live(port_lib, ref_lib):
    """Differential: fresh random streams (not in the fixtures) through both."""
    for mop, mname in ((okv.MERGE_COUNTER, "counter"), (okv.MERGE_APPEND, "append")):
        for seed in (91, 92):
            keys, stream = random_stream(seed + mop, 80, merge=mname)
            a = okv.Okv(port_lib, merge_op=mop)
            b = okv.Okv(ref_lib, merge_op=mop)
            for i, (bt, ts) in enumerate(stream):
                assert a.apply(bt, ts) == b.apply(bt, ts)
                if i == 40:
                    b.flush()
            b.compact()
            assert a.latest_seq() == b.latest_seq()
            for k in keys:
                assert a.get(k) == b.get(k)
            assert a.multi_get(keys) == b.multi_get(keys)
            assert a.scan() == b.scan()
            a.close()
            b.close()
    for name, bt in corrupt_cases():
        a = okv.Okv(port_lib, merge_op=okv.MERGE_UINT64ADD)
        b = okv.Okv(ref_lib, merge_op=okv.MERGE_UINT64ADD)
        assert a.apply(bt, 7) == b.apply(bt, 7), name
        assert a.last_error == b.last_error, name
        assert a.latest_seq() == b.latest_seq(), name
        a.close()
        b.close()


def test_port_vs_reference_fuzz(port_lib, ref_lib):
    """oracle/fuzz_port_vs_ref.py on a few seeds: flush/compaction at random points, MultiGet, iterator walks."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "fuzz_port_vs_ref.py")
    spec = importlib.util.spec_from_file_location("fuzz_port_vs_ref", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(200, 212, port_lib, ref_lib) == 0


def test_port_vs_reference_corruption_fuzz(port_lib, ref_lib):
    """oracle/fuzz_corrupt_port_vs_ref.py on a few seeds: mutated batches, error class + text + latch + contents"""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "fuzz_corrupt_port_vs_ref.py")
    spec = importlib.util.spec_from_file_location("fuzz_corrupt_port_vs_ref", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(0, 25, port_lib, ref_lib) == 0
