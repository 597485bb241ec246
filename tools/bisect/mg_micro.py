"""A/B of the MultiGet kernel between two builds of librsp_b200.so on the SAME GPU in the same call: loads config 2
(1024 shards x 10 M KV), compacts, times K device-resident MultiGet launches with CUDA events.  Only entry points both
builds export are used.   python tools/bisect/mg_micro.py <path.so> [label]"""
import ctypes as C, sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from rocksplicator_b200 import synth
so, label = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
lib = C.CDLL(so)
vp = C.c_void_p
lib.rsp_engine_create.argtypes = [C.c_int, vp, C.POINTER(vp)]
lib.rsp_shard_open.argtypes = [vp, C.c_char_p, vp, C.POINTER(vp)]
lib.rsp_shard_index.restype = C.c_uint32; lib.rsp_shard_index.argtypes = [vp]
lib.rsp_apply_many.argtypes = [vp, C.c_size_t, vp, vp, vp, vp, vp]
lib.rsp_compact_all.argtypes = [vp]
lib.rsp_engine_stream.restype = vp; lib.rsp_engine_stream.argtypes = [vp]
lib.rsp_multi_get_device.argtypes = [vp, C.c_size_t, vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp, vp]
class Opts(C.Structure):
    _fields_ = [("merge_op", C.c_uint32), ("reserved", C.c_uint32), ("write_buffer_bytes", C.c_uint64), ("merge_fn", vp), ("merge_state", vp)]
eng = vp(); assert lib.rsp_engine_create(0, None, C.byref(eng)) == 0
S, NKV, Q, K, W = 1024, 10_000_000, 8388608, 10, 3
six_of = np.zeros(S, dtype=np.uint32)
for i in range(S):
    h = vp(); o = Opts(write_buffer_bytes=2 << 20)
    assert lib.rsp_shard_open(eng, b"segment%05d" % i, C.byref(o), C.byref(h)) == 0
    six_of[i] = lib.rsp_shard_index(h)
seed = synth.SEED_DATA
for lo in range(0, NKV, 1 << 20):
    idx = np.arange(lo, min(NKV, lo + (1 << 20)), dtype=np.uint64)
    sh = (idx % np.uint64(S)).astype(np.int64)
    b = synth.single_put_batches(synth.keys16(seed, idx), synth.values(seed, sh, idx, 0), 1000 + idx)
    off = np.arange(idx.size + 1, dtype=np.uint64) * np.uint64(b.shape[1])
    st = np.zeros(idx.size, dtype=np.int32); six = six_of[sh]; ts = (1000 + idx).astype(np.uint64); bb = b.reshape(-1)
    lib.rsp_apply_many(eng, idx.size, six.ctypes.data, bb.ctypes.data, off.ctypes.data, ts.ctypes.data, st.ctypes.data)
    assert not st.any()
assert lib.rsp_compact_all(eng) == 0
stream = torch.cuda.ExternalStream(lib.rsp_engine_stream(eng))
rng = np.random.default_rng(synth.SEED_QUERY)
with torch.cuda.stream(stream):
    qs = [rng.integers(0, NKV, size=Q, dtype=np.uint64) for _ in range(W + K)]
    d_keys = [torch.from_numpy(synth.keys16(seed, q).reshape(-1)).cuda() for q in qs]
    d_six = [torch.from_numpy(six_of[(q % np.uint64(S)).astype(np.int64)].astype(np.int32)).cuda() for q in qs]
    d_vals = torch.empty(Q * 64, dtype=torch.uint8, device="cuda"); d_vlen = torch.empty(Q, dtype=torch.int32, device="cuda"); d_st = torch.empty(Q, dtype=torch.int32, device="cuda")
sp = vp(stream.cuda_stream)
def mg(i):
    assert lib.rsp_multi_get_device(eng, Q, d_six[i].data_ptr(), d_keys[i].data_ptr(), 16, d_vals.data_ptr(), 64, d_vlen.data_ptr(), d_st.data_ptr(), sp) == 0
for rep in range(2):
    for i in range(W): mg(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(K): mg(W + k)
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    assert int(d_st.count_nonzero().item()) == 0
    print("%s rep %d: %.3f ms per launch -> %.2f G lookups/s, frac %.3f" % (label, rep, ms, Q / ms / 1e6, 168 * Q / (ms * 1e-3) / 1e9 / 6570.3), flush=True)
free, total = torch.cuda.mem_get_info()
print(label, "device memory in use %.2f GB" % ((total - free) / 1e9))
