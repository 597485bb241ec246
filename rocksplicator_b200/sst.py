"""ctypes binding of the host-side SST reader / writer (rocksplicator_b200/host/sst/sst_format.h in librsp_host.so):
RocksDB block-based table files in (bulk ingest) and out (spill).  SURVEY.md §8(f) rank 2."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_SO = os.path.join(HERE, "librsp_host.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(HOST_SO):
            from . import build
            build.build_host()
        lib = C.CDLL(HOST_SO)
        lib.rsp_sst_read.restype = C.c_int
        lib.rsp_sst_read.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                     C.POINTER(C.c_size_t), C.POINTER(C.c_uint64), C.c_char_p, C.c_size_t]
        lib.rsp_sst_write.restype = C.c_int
        lib.rsp_sst_write.argtypes = [C.c_size_t, C.c_char_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint32,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_char_p, C.c_size_t]
        lib.rsp_host_free.restype = None
        lib.rsp_host_free.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def read_sst(data: bytes):
    """-> ([(user_key, seq, type, value)], {"num_entries", "external_version", "global_seqno"})"""
    lib = _load()
    out, out_len, n = C.c_void_p(), C.c_size_t(), C.c_size_t()
    props = (C.c_uint64 * 3)()
    err = C.create_string_buffer(256)
    rc = lib.rsp_sst_read(data, len(data), C.byref(out), C.byref(out_len), C.byref(n), props, err, 256)
    if rc != 0:
        raise ValueError(err.value.decode())
    raw = C.string_at(out.value, out_len.value)
    lib.rsp_host_free(out)
    entries, at = [], 0
    for _ in range(n.value):
        kl, vl = int.from_bytes(raw[at:at + 4], "little"), int.from_bytes(raw[at + 4:at + 8], "little")
        seq, typ = int.from_bytes(raw[at + 8:at + 16], "little"), raw[at + 16]
        entries.append((raw[at + 17:at + 17 + kl], seq, typ, raw[at + 17 + kl:at + 17 + kl + vl]))
        at += 17 + kl + vl
    return entries, {"num_entries": props[0], "external_version": props[1], "global_seqno": props[2]}


def write_sst(sorted_kv, block_size=4096) -> bytes:
    """sorted_kv: [(key, value)] in strictly increasing key order -> an SST image IngestExternalFile accepts"""
    lib = _load()
    n = len(sorted_kv)
    koff = np.zeros(n + 1, dtype=np.uint64)
    voff = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(np.fromiter((len(k) for k, _ in sorted_kv), dtype=np.uint64, count=n), out=koff[1:])
    np.cumsum(np.fromiter((len(v) for _, v in sorted_kv), dtype=np.uint64, count=n), out=voff[1:])
    keys = b"".join(k for k, _ in sorted_kv) + b"\0"
    vals = b"".join(v for _, v in sorted_kv) + b"\0"
    out, out_len = C.c_void_p(), C.c_size_t()
    err = C.create_string_buffer(256)
    rc = lib.rsp_sst_write(n, keys, koff.ctypes.data, vals, voff.ctypes.data, block_size, C.byref(out), C.byref(out_len), err, 256)
    if rc != 0:
        raise ValueError(err.value.decode())
    data = C.string_at(out.value, out_len.value)
    lib.rsp_host_free(out)
    return data


def ingest_file(shard, path, allow_global_seqno=True) -> int:
    """DB::IngestExternalFile(path) on an engine shard (rocksdb_admin/admin_handler.cpp:1820-1845): parse on the host,
    install as one sorted run on the device (rsp_ingest_sorted).  Returns a rocksdb::Status code."""
    with open(path, "rb") as f:
        entries, props = read_sst(f.read())
    if props["external_version"] == 0:
        raise ValueError("External file version not found")
    if any(t != 1 or seq != 0 for _, seq, t, _ in entries):
        raise ValueError("external file holds a record that is not a Put with sequence number 0")
    return shard.ingest([(k, v) for k, _, _, v in entries], allow_global_seqno)


def export_file(shard, path) -> int:
    """The shard's visible contents (merges folded, tombstones dropped) as one ingestible SST file; -> entries"""
    kv = shard.scan()
    with open(path, "wb") as f:
        f.write(write_sst(kv))
    return len(kv)
