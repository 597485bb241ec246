"""SST interchange on the host, CPU only.
  * the reference's golden file rocksdb_admin/tests/old_sst_data.sst (copied byte-for-byte to tests/golden/: a 978-byte
    test fixture, Snappy-compressed block) must read as key0..key9 -> value0..value9 (sst_binary.cpp:43-58);
  * files written by the reference's own SstFileWriter (librocksdb.so.5.4) must read back exactly;
  * files written by OUR writer must be accepted by the reference's DB::IngestExternalFile and scan back exactly,
    with the sequence-number rule of rocksdb_assumption_test.cpp:248-262 (ingest into an empty DB: seq stays 0)."""
import ctypes as C
import os
import random
import tempfile

import pytest

from rocksplicator_b200 import sst

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "old_sst_data.sst")


def test_golden_old_sst_data():
    entries, props = sst.read_sst(open(GOLDEN, "rb").read())
    assert [(k, t, v) for k, _, t, v in entries] == [(b"key%d" % i, 1, b"value%d" % i) for i in range(10)]
    assert props["num_entries"] == 10
    assert all(seq == 0 for _, seq, _, _ in entries)


def test_rejects_garbage():
    with pytest.raises(ValueError):
        sst.read_sst(b"not an sst file at all, but long enough to have a footer ...............")
    data = bytearray(open(GOLDEN, "rb").read())
    data[10] ^= 0xFF  # flip a byte inside the first data block: the CRC32C must catch it
    with pytest.raises(ValueError):
        sst.read_sst(bytes(data))


def _kv(n, seed, max_v=300):
    rng = random.Random(seed)
    keys = sorted({bytes(rng.getrandbits(8) for _ in range(rng.randint(1, 40))) for _ in range(n)})
    return [(k, bytes(rng.getrandbits(8) for _ in range(rng.choice([0, 1, 8, 64, max_v])))) for k in keys]


def test_roundtrip_own_writer_reader():
    for n, seed in ((1, 1), (10, 2), (3000, 3)):
        kv = _kv(n, seed)
        entries, props = sst.read_sst(sst.write_sst(kv, block_size=512 if n > 100 else 4096))
        assert [(k, v) for k, _, _, v in entries] == kv
        assert props["external_version"] == 2 and props["num_entries"] == len(kv)
    with pytest.raises(ValueError):
        sst.write_sst([(b"b", b"1"), (b"a", b"2")])  # "Keys must be added in order"


def _ref_extras(ref_lib):
    ref_lib.okv_write_sst.restype = C.c_int
    ref_lib.okv_write_sst.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_char_p, C.c_size_t]
    ref_lib.okv_ingest_sst.restype = C.c_int
    ref_lib.okv_ingest_sst.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
    return ref_lib


def test_reads_files_written_by_the_reference(ref_lib):
    import numpy as np
    lib = _ref_extras(ref_lib)
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    for n, seed in ((1, 11), (50, 12), (4000, 13)):
        kv = _kv(n, seed, max_v=2000)
        koff = np.zeros(len(kv) + 1, dtype=np.uint64)
        voff = np.zeros(len(kv) + 1, dtype=np.uint64)
        np.cumsum([len(k) for k, _ in kv], out=koff[1:])
        np.cumsum([len(v) for _, v in kv], out=voff[1:])
        path = os.path.join(tmp, "ref%d.sst" % n)
        err = C.create_string_buffer(256)
        rc = lib.okv_write_sst(path.encode(), len(kv), b"".join(k for k, _ in kv) + b"\0", koff.ctypes.data,
                               b"".join(v for _, v in kv) + b"\0", voff.ctypes.data, err, 256)
        assert rc == 0, err.value
        entries, props = sst.read_sst(open(path, "rb").read())
        assert [(k, t, v) for k, _, t, v in entries] == [(k, 1, v) for k, v in kv]
        assert props["external_version"] == 2 and props["global_seqno"] == 0 and props["num_entries"] == len(kv)


def test_reference_ingests_files_written_here(ref_lib):
    from oracle import okv
    lib = _ref_extras(ref_lib)
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    for n, seed in ((1, 21), (200, 22), (5000, 23)):
        kv = _kv(n, seed, max_v=1500)
        path = os.path.join(tmp, "ours%d.sst" % n)
        open(path, "wb").write(sst.write_sst(kv))
        db = okv.Okv(lib)
        err = C.create_string_buffer(256)
        rc = lib.okv_ingest_sst(db.h, path.encode(), 1, err, 256)
        assert rc == 0, err.value
        assert db.latest_seq() == 0  # ingest into an empty DB does not consume sequence numbers
        assert db.scan() == kv
        for k, v in kv[:: max(1, len(kv) // 50)]:
            assert db.get(k) == (0, v)
        # an overlapping second ingest is given global sequence number last+1 (rocksdb_assumption_test.cpp:272-283)
        kv2 = [(k, b"second:" + v[:5]) for k, v in kv[::3]]
        path2 = os.path.join(tmp, "ours%d_b.sst" % n)
        open(path2, "wb").write(sst.write_sst(kv2))
        assert lib.okv_ingest_sst(db.h, path2.encode(), 1, err, 256) == 0, err.value
        assert db.latest_seq() == 1
        want = dict(kv)
        want.update(kv2)
        assert db.scan() == sorted(want.items())
        # (move_files = false: RocksDB stamps the sequence number into ITS copy; the file written here is untouched)
        entries, props = sst.read_sst(open(path2, "rb").read())
        assert props["global_seqno"] == 0 and [(k, v) for k, _, _, v in entries] == kv2
        db.close()


def test_shard_file_helpers_host_half():
    """sst.export_file / sst.ingest_file around a stand-in shard: the file goes through the real writer and reader"""
    class Standin:
        def __init__(self, kv=()): self.kv = list(kv)
        def ingest(self, kv, allow_global_seqno=True): self.kv = list(kv); return 0
        def scan(self): return self.kv
    kv = _kv(300, 31)
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    path = os.path.join(tmp, "shard.sst")
    assert sst.export_file(Standin(kv), path) == len(kv)
    dst = Standin()
    assert sst.ingest_file(dst, path) == 0 and dst.kv == kv
    # the reference's own golden file (SstFileWriter output, Snappy block) loads the same way
    assert sst.ingest_file(dst, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "old_sst_data.sst")) == 0
    assert dst.kv == [(b"key%d" % i, b"value%d" % i) for i in range(10)]
