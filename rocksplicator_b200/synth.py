"""Synthetic shard data of SURVEY.md §8(d), vectorised with numpy (bench + tests).

keys   : 16 B = big-endian u64 index ‖ big-endian u64 splitmix64(seed ^ index); shard = index % n_shards
values : vlen bytes from a splitmix64 stream keyed by (seed, shard, index, version)
batches: the replicated unit — one Put + the leader's 8-byte timestamp LogData = 105 wire bytes at 16/64
seeds  : data 0x5EED0001, queries 0x5EED0002, zipf 0x5EED0003
"""
import numpy as np

SEED_DATA, SEED_QUERY, SEED_ZIPF = 0x5EED0001, 0x5EED0002, 0x5EED0003


def splitmix64(x):
    x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def keys16(seed, idx):
    """(n, 16) uint8"""
    idx = np.asarray(idx, dtype=np.uint64)
    out = np.empty((idx.size, 2), dtype=">u8")
    out[:, 0] = idx
    out[:, 1] = splitmix64(np.uint64(seed) ^ idx)
    return out.view(np.uint8).reshape(idx.size, 16)


def values(seed, shard, idx, version, vlen=64):
    """(n, vlen) uint8"""
    idx = np.asarray(idx, dtype=np.uint64)
    shard = np.asarray(shard, dtype=np.uint64)
    s = splitmix64(np.uint64(seed) ^ (shard << np.uint64(40)) ^ (idx << np.uint64(8)) ^ np.uint64(version))
    nw = (vlen + 7) // 8
    out = np.empty((idx.size, nw), dtype="<u8")
    for w in range(nw):
        s = splitmix64(s)
        out[:, w] = s
    return out.view(np.uint8).reshape(idx.size, nw * 8)[:, :vlen]


def _varint(n):
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def single_put_batches(keys, vals, ts_ms):
    """(n, L) uint8 wire bytes: header(seq=0,count=1) Put(key,val) LogData(ts) — what the leader serves
    (replicated_db.cpp:115-117, 527-530)."""
    n, klen = keys.shape
    vlen = vals.shape[1]
    kv, vv = _varint(klen), _varint(vlen)
    L = 12 + 1 + len(kv) + klen + len(vv) + vlen + 10
    b = np.zeros((n, L), dtype=np.uint8)
    b[:, 8] = 1
    at = 12
    b[:, at] = 1
    at += 1
    b[:, at:at + len(kv)] = np.frombuffer(kv, dtype=np.uint8)
    at += len(kv)
    b[:, at:at + klen] = keys
    at += klen
    b[:, at:at + len(vv)] = np.frombuffer(vv, dtype=np.uint8)
    at += len(vv)
    b[:, at:at + vlen] = vals
    at += vlen
    b[:, at] = 3
    b[:, at + 1] = 8
    b[:, at + 2:at + 10] = np.asarray(ts_ms, dtype="<u8").reshape(n, 1).view(np.uint8)
    return b


def zipf_ranks(rng, n_items, theta, size):
    """YCSB-style zipfian ranks in [0, n_items): P(rank r) ~ 1 / (r+1)^theta (Gray et al. generator)."""
    i = np.arange(1, n_items + 1, dtype=np.float64)
    zetan = float(np.sum(1.0 / np.power(i, theta)))
    zeta2 = 1.0 + 0.5 ** theta
    alpha = 1.0 / (1.0 - theta)
    eta = (1.0 - (2.0 / n_items) ** (1.0 - theta)) / (1.0 - zeta2 / zetan)
    u = rng.random(size)
    uz = u * zetan
    r = (n_items * np.power(eta * u - eta + 1.0, alpha)).astype(np.int64)
    r[uz < zeta2] = 1
    r[uz < 1.0] = 0
    return np.clip(r, 0, n_items - 1).astype(np.uint64)


def scatter_ranks(ranks, n_items):
    """spread popularity ranks over the key space (hot keys land on many shards)"""
    return (ranks * np.uint64(2654435761) + np.uint64(12345)) % np.uint64(n_items)


# ---- the same generator on a torch device (tools/stretch.py: a billion keys would take numpy minutes of CPU time) ----
def _t_splitmix64(x):
    """splitmix64 on int64 torch tensors (two's complement arithmetic wraps; logical shifts are masked)"""
    import torch

    def lsr(v, s):
        return (v >> s) & ((1 << (64 - s)) - 1)

    def c(u):  # a uint64 constant as the int64 with the same bits
        return u - (1 << 64) if u >= (1 << 63) else u
    x = x + c(0x9E3779B97F4A7C15)
    z = (x ^ lsr(x, 30)) * c(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * c(0x94D049BB133111EB)
    return z ^ lsr(z, 31)


def torch_single_put_batches(seed, shard, idx, version, ts_ms, vlen=64):
    """== single_put_batches(keys16(seed, idx), values(seed, shard, idx, version, vlen), ts_ms) for 16-byte keys and
    vlen < 128, computed on idx's device: (n, 105) uint8 at vlen 64.  shard / idx / ts_ms: int64 tensors."""
    import torch
    assert vlen < 128 and vlen % 8 == 0
    n = idx.numel()
    L = 12 + 2 + 16 + 1 + vlen + 10
    dev = idx.device
    b = torch.zeros((n, L), dtype=torch.uint8, device=dev)
    b[:, 8] = 1
    b[:, 12] = 1
    b[:, 13] = 16

    def be_bytes(v):  # (n,) int64 -> (n, 8) uint8 big-endian
        return v.view(torch.uint8).reshape(n, 8).flip(1)

    def le_bytes(v):
        return v.view(torch.uint8).reshape(n, 8)
    idx = idx.contiguous()
    b[:, 14:22] = be_bytes(idx)
    b[:, 22:30] = be_bytes(_t_splitmix64(idx ^ seed).contiguous())
    b[:, 30] = vlen
    s = _t_splitmix64(((shard << 40) ^ (idx << 8) ^ version) ^ seed)
    at = 31
    for _ in range(vlen // 8):
        s = _t_splitmix64(s)
        b[:, at:at + 8] = le_bytes(s.contiguous())
        at += 8
    b[:, at] = 3
    b[:, at + 1] = 8
    b[:, at + 2:at + 10] = le_bytes(ts_ms.contiguous())
    return b
