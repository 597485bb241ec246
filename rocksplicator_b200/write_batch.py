"""Host-side WriteBatch builder/reader: the byte contract between producers and the apply path.

Wire format (RocksDB db/write_batch.cc, the contract `rocksdb_replicator/rocksdb_wrapper.cpp:17-18`
consumes and `examples/counter_service/counter_handler.cpp:152-156,212-216` produce):

    [fixed64 sequence LE][fixed32 count LE] { tag, varint32-length-prefixed slices }*

    0x00 Delete(key)        0x01 Put(key, value)     0x02 Merge(key, value)
    0x03 LogData(blob)      0x07 SingleDelete(key)   0x0D Noop
    0x04/0x05/0x06/0x08 = column-family forms (varint32 cf id first)

LogData is not counted and consumes no sequence number.
"""
import struct

T_DELETE, T_PUT, T_MERGE, T_LOGDATA, T_SINGLE_DELETE, T_NOOP = 0, 1, 2, 3, 7, 13
HEADER = 12


def varint32(n: int) -> bytes:
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def _lp(b: bytes) -> bytes:
    return varint32(len(b)) + b


class WriteBatch:
    """Mirror of rocksdb::WriteBatch's mutators (Put/Delete/SingleDelete/Merge/PutLogData/Data/Count)."""

    def __init__(self, rep: bytes = None):
        self._rep = bytearray(rep) if rep is not None else bytearray(HEADER)

    def _bump(self):
        c = struct.unpack_from("<I", self._rep, 8)[0]
        struct.pack_into("<I", self._rep, 8, c + 1)

    def put(self, key: bytes, value: bytes):
        self._bump()
        self._rep += bytes([T_PUT]) + _lp(key) + _lp(value)
        return self

    def delete(self, key: bytes):
        self._bump()
        self._rep += bytes([T_DELETE]) + _lp(key)
        return self

    def single_delete(self, key: bytes):
        self._bump()
        self._rep += bytes([T_SINGLE_DELETE]) + _lp(key)
        return self

    def merge(self, key: bytes, value: bytes):
        self._bump()
        self._rep += bytes([T_MERGE]) + _lp(key) + _lp(value)
        return self

    def put_log_data(self, blob: bytes):
        self._rep += bytes([T_LOGDATA]) + _lp(blob)
        return self

    def set_sequence(self, seq: int):
        struct.pack_into("<Q", self._rep, 0, seq)
        return self

    def count(self) -> int:
        return struct.unpack_from("<I", self._rep, 8)[0]

    def data(self) -> bytes:
        return bytes(self._rep)

    def data_size(self) -> int:
        return len(self._rep)

    def clear(self):
        self._rep = bytearray(HEADER)


def single_put(key: bytes, value: bytes, ts_ms: int = None) -> bytes:
    """The benchmark's replicated unit: one Put (+ the leader's 8-byte timestamp LogData,
    replicated_db.cpp:115-117)."""
    wb = WriteBatch().put(key, value)
    if ts_ms is not None:
        wb.put_log_data(struct.pack("<Q", ts_ms))
    return wb.data()
