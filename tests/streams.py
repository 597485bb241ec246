"""Deterministic WriteBatch stream generators shared by the parity tests (host logic, no GPU)."""
import random
import struct

from rocksplicator_b200.write_batch import WriteBatch, varint32

MASK = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK
    return z ^ (z >> 31)


def bench_key(seed, idx):
    """16 B = big-endian u64 index ‖ u64 splitmix64(seed ^ index)."""
    return struct.pack(">QQ", idx, splitmix64(seed ^ idx))


def bench_value(seed, shard, idx, version, vlen=64):
    out = bytearray()
    s = splitmix64(seed ^ (shard << 40) ^ (idx << 8) ^ version)
    while len(out) < vlen:
        s = splitmix64(s)
        out += struct.pack("<Q", s)
    return bytes(out[:vlen])


def random_stream(seed, n_batches, n_keys=40, merge="counter", max_ops=6, var_len=True,
                  bad_operands=False):
    """Mixed Put/Delete/SingleDelete/Merge/LogData batches over a small key space (many overwrites)."""
    rng = random.Random(seed)
    keys = []
    for i in range(n_keys):
        if var_len:
            kl = rng.choice([0, 1, 3, 4, 8, 15, 16, 17, 31, 32, 33, 64, 130, 300]) if i else 0
            keys.append(bytes(rng.getrandbits(8) for _ in range(kl)) if kl else b"")
        else:
            keys.append(bench_key(seed, i))
    keys = list(dict.fromkeys(keys))
    out = []
    pending_sd = []
    for b in range(n_batches):
        wb = WriteBatch()
        for _ in range(rng.randint(0, max_ops)):
            k = rng.choice(keys)
            r = rng.random()
            if r < 0.45:
                if merge == "counter" and (not bad_operands or rng.random() < 0.6):
                    v = struct.pack("<q", rng.randint(-1000, 1000))
                else:
                    vl = rng.choice([0, 1, 5, 8, 16, 63, 64, 65, 200, 1000]) if var_len else 64
                    v = bytes(rng.getrandbits(8) for _ in range(vl))
                wb.put(k, v)
            elif r < 0.60:
                wb.delete(k)
            elif r < 0.65:
                # SingleDelete's contract (exactly one Put before it, no overwrite) is honoured with
                # dedicated keys; outside the contract RocksDB's own result depends on compaction timing
                sd = b"sd-%d-%d" % (b, len(wb.data()))
                if rng.random() < 0.5:
                    wb.put(sd, b"once")
                    pending_sd.append(sd)
                elif pending_sd:
                    wb.single_delete(pending_sd.pop(rng.randrange(len(pending_sd))))
            elif r < 0.95 and merge:
                if merge == "counter":
                    v = b"xyz" if (bad_operands and rng.random() < 0.05) else struct.pack("<q", rng.randint(-50, 50))
                else:
                    v = bytes(rng.getrandbits(8) for _ in range(rng.choice([0, 1, 8, 20])))
                wb.merge(k, v)
            else:
                wb.put_log_data(bytes(rng.getrandbits(8) for _ in range(rng.choice([0, 8, 9]))))
        wb.set_sequence(rng.getrandbits(48))  # header seq is ignored by the follower (§9 row 2)
        out.append((wb.data(), rng.getrandbits(40)))
    return keys, out


def corrupt_cases():
    """(name, bytes) for the failure rows of SURVEY §9 (each must run in its own shard: the error latches)."""
    good = WriteBatch().put(b"k1", b"v1").delete(b"k2").merge(b"c", struct.pack("<Q", 5))
    g = good.data()
    cases = [
        ("count_high", g[:8] + struct.pack("<I", 4) + g[12:]),
        ("count_low", g[:8] + struct.pack("<I", 2) + g[12:]),
        ("trunc_merge", g[:-3]),
        ("trunc_merge_1", g[:-1]),
        ("trunc_put", g[:15]),
        ("trunc_in_klen", g[:13]),
        ("too_small", g[:7]),
        ("too_small_11", g[:11]),
        ("empty", b""),
        ("unknown_tag", g + b"\x40"),
        ("cf_put_0", g[:8] + struct.pack("<I", 4) + g[12:] + b"\x05\x00" + b"\x01a\x01b"),
        ("cf_put_9", g[:8] + struct.pack("<I", 4) + g[12:] + b"\x05\x09" + b"\x01a\x01b"),
        ("cf_del_0", g[:8] + struct.pack("<I", 4) + g[12:] + b"\x04\x00" + b"\x02k1"),
        ("cf_merge_0", g[:8] + struct.pack("<I", 4) + g[12:] + b"\x06\x00" + b"\x01c\x08" + struct.pack("<Q", 2)),
        ("cf_sdel_0", g[:8] + struct.pack("<I", 4) + g[12:] + b"\x08\x00" + b"\x02k1"),
        ("noop", g + b"\x0d"),
        ("header_only", bytes(12)),
        ("logdata_only", WriteBatch().put_log_data(b"12345678").data()),
        ("bad_varint", g[:12] + b"\x01\xff\xff\xff\xff\xff\x01"),
        ("varint5", WriteBatch().put(b"a" * 300, b"b" * 70000).data()),
        ("klen_over", g[:12] + b"\x01" + varint32(1000) + b"abc"),
        ("swallow_logdata", WriteBatch().put(b"k", b"").data()[:-1] + varint32(4)),
        # two-phase-commit markers: RocksDB parses them and, outside WAL recovery, ignores them (not counted)
        ("2pc_markers_ok", g + b"\x09" + b"\x0a\x03abc" + b"\x0b\x00" + b"\x0c\x01x"),
        ("2pc_only", bytes(12) + b"\x09\x0a\x01x\x0b\x01x"),
        ("2pc_begin_is_not_counted", g[:8] + struct.pack("<I", 4) + g[12:] + b"\x09"),
        ("2pc_end_bad", g + b"\x0a\x7f"),
        ("2pc_commit_bad", g + b"\x0b\x7f"),
        ("2pc_rollback_bad", g + b"\x0c\x7f"),
        # range deletions: only malformed ones here (RocksDB's error class); a well-formed one is refused by this
        # engine and applied by RocksDB — the one deliberate difference, tested in test_range_deletion_is_refused
        ("delrange_bad_begin", g[:8] + struct.pack("<I", 4) + g[12:] + b"\x0f\x7f"),
        ("delrange_bad_end", g[:8] + struct.pack("<I", 4) + g[12:] + b"\x0f\x01a\x7f"),
        ("delrange_cf_bad_varint", g[:8] + struct.pack("<I", 4) + g[12:] + b"\x0e\xff\xff\xff\xff\xff\xff"),
        ("delrange_cf9", g[:8] + struct.pack("<I", 4) + g[12:] + b"\x0e\x09\x01a\x01z"),
        ("delrange_then_unknown_tag", g[:8] + struct.pack("<I", 4) + g[12:] + b"\x0f\x01a\x01z\x40"),
        ("delrange_wrong_count", g + b"\x0f\x01a\x01z"),
    ]
    return cases
