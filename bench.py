#!/usr/bin/env python
"""bench.py — the hot path of BASELINE.json on B200: MultiGet lookups/s + replicated applies/s per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1], SURVEY.md §8d "config 2"): 1024 shards on one GPU, 10 M KV total,
16 B keys / 64 B values, loaded through the apply path and fully compacted.  One step =
  * one MultiGet pass: 256 concurrent MultiGet(4096) calls = 1,048,576 uniform lookups in one launch, and
  * one apply tick: 1024 shards x 50 replicated single-Put WriteBatches (pull-sized, 105 wire bytes each).
`value` is MultiGet lookups/s with queries and results resident in HBM (CUDA events on the engine's
stream); `applies` carries the apply-side numbers; `e2e` is the same through the host-buffer C ABI
(rsp_multi_get_fixed / rsp_apply_many) with pinned host memory, H2D + D2H inside the timed region.
N > 1: one process per GPU (torchrun), each rank an independent engine with its own 1024 shards —
shards partition shard_id -> GPU, no collective on the data path ("scaling": "weak").
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

A_GET = 16 + (16 + 64 + 8) + 64          # algorithmic bytes per MultiGet hit
A_PUT = 83 + 22 + 88                     # per single-Put apply


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def run_cpu(lib_kind, threads, shards, kv, apply_cap, get_secs, batch=4096, wal=1):
    """times oracle/okv_cpu_bench (the reference's RocksDB binary when oracle/_ref is present, else the port)"""
    from oracle import okv
    okv.build(ref=os.path.isdir("/root/reference"))
    lib = okv.REF_SO if (lib_kind == "reference" and okv.ref_available()) else okv.PORT_SO
    base = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    d = os.path.join(base, "okv_bench_%d" % os.getpid())
    subprocess.call(["rm", "-rf", d])
    os.makedirs(d)
    try:
        out = subprocess.check_output([os.path.join(ROOT, "oracle", "okv_cpu_bench"), lib, str(threads), str(shards),
                                       str(kv), "64", str(wal), str(apply_cap), str(get_secs), str(batch), "1", d],
                                      text=True)
    finally:
        subprocess.call(["rm", "-rf", d])
    return json.loads(out.strip().splitlines()[-1])


class SeamCfg(C.Structure):
    _fields_ = [("device", C.c_int32), ("shards", C.c_uint32), ("kv_total", C.c_uint64), ("value_len", C.c_uint32),
                ("executor_threads", C.c_uint32), ("updates_per_response", C.c_uint32), ("update_rounds", C.c_uint32),
                ("multiget_threads", C.c_uint32), ("multiget_batch", C.c_uint32), ("multiget_secs", C.c_double),
                ("get_threads", C.c_uint32), ("get_secs", C.c_double), ("seed", C.c_uint64),
                ("first_shard_id", C.c_uint32), ("reserved", C.c_uint32)]


class SeamResult(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("load_s", "load_applies_per_s", "resp_p50_ms", "resp_p99_ms", "compact_s",
                                          "mget_lookups_per_s", "mget_p50_ms", "mget_p99_ms")] + \
               [("mget_calls", C.c_uint64)] + \
               [(n, C.c_double) for n in ("get_per_s", "get_p50_us", "get_p99_us", "mixed_applies_per_s",
                                          "mixed_lookups_per_s", "mixed_resp_p50_ms", "mixed_resp_p99_ms")] + \
               [(n, C.c_uint64) for n in ("applied_total", "parity_errors", "status_errors", "engine_launches")]


def run_seams(device, shards, kv, rank, world, secs=2.0, value_len=64, update_rounds=20):
    """The hot path through the reference's own seams (rocksplicator_b200/host/bench/seam_bench.cpp): followers pull
    from a synthetic leader through RocksDBReplicator -> DbWrapper; readers call ApplicationDB::MultiGet(4096) / Get from
    many threads.  Host buffers, H2D + D2H inside; every value checked against the generator."""
    from rocksplicator_b200 import build
    host_so = os.environ.get("RSP_TEST_EMUL_HOST_LIB")  # tests/emul/bench_dryrun.py only (CPU emulation, no timings)
    if not host_so:
        build.build_host()
        host_so = build.HOST_SO
    lib = C.CDLL(host_so)
    lib.rsp_seam_bench.restype = C.c_int
    lib.rsp_seam_bench.argtypes = [C.POINTER(SeamCfg), C.POINTER(SeamResult)]
    cores = os.cpu_count() or 8
    share = max(1, cores // max(1, world))
    cfg = SeamCfg(device=device, shards=shards, kv_total=kv, value_len=value_len,
                  executor_threads=max(16, min(32, share // 2)), updates_per_response=50, update_rounds=update_rounds,
                  multiget_threads=max(4, min(64, share // 2)), multiget_batch=4096, multiget_secs=secs,
                  get_threads=max(8, min(256, share * 2)), get_secs=secs / 2, seed=0x5EED0001 + rank,
                  first_shard_id=rank * shards)
    res = SeamResult()
    rc = lib.rsp_seam_bench(C.byref(cfg), C.byref(res))
    out = {n: getattr(res, n) for n, _ in SeamResult._fields_}
    out["rc"] = rc
    out["threads"] = {"executor": cfg.executor_threads, "multiget": cfg.multiget_threads, "get": cfg.get_threads}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--kv", type=int, default=10_000_000)
    ap.add_argument("--shards", type=int, default=1024)
    ap.add_argument("--mg-batches", type=int, default=2048, help="concurrent MultiGet(4096) calls per launch")
    ap.add_argument("--tick", type=int, default=50, help="replicated updates per shard per apply tick")
    ap.add_argument("--cpu-kv", type=int, default=2_000_000)
    ap.add_argument("--cpu-get-secs", type=float, default=6.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-seams", action="store_true", help="skip the through-the-seams phase (librsp_host.so)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = args.steps, max(args.warmup, 0)
    ncores = os.cpu_count() or 1
    workload = "%d shards x %d KV total, 16B key/64B value, uniform MultiGet batch=4096 x %d in flight, apply tick %d shards x %d single-Put WriteBatch" % (
        args.shards, args.kv, args.mg_batches, args.shards, args.tick)

    if args.impl == "reference":
        if rank != 0:
            return
        r = run_cpu("reference", ncores, args.shards, args.cpu_kv, 60.0, max(2.0, min(20.0, 1.0 * K)))
        line = {
            "impl": "reference", "metric": "multiget_lookups_per_s", "value": r["lookups_per_s"], "unit": "lookups/s",
            "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * r["get_s"] / max(K, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload, "sample": "%d KV applied (WAL on, default WriteOptions), flush+compact, %.0f s of MultiGet(4096) split per shard" % (r["applied"], r["get_s"])},
            "applies": {"value": r["applies_per_s"], "unit": "applies/s"},
            "cpu_baseline": {"value": r["lookups_per_s"], "unit": "lookups/s", "cores": r["threads"], "kind": r["kind"],
                             "sample": "%d KV over %d shards, %d threads" % (r["applied"], r["shards"], r["threads"]),
                             "applies_per_s": r["applies_per_s"]},
            "e2e": {"value": r["lookups_per_s"], "unit": "lookups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from rocksplicator_b200 import build, engine, synth
    if not os.path.exists(engine.SO_PATH):
        build.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    lib = engine.load_library()
    eng = engine.Engine(local_rank, max_shards=max(16384, args.shards))
    stream = torch.cuda.ExternalStream(lib.rsp_engine_stream(eng.h), device=torch.device("cuda", local_rank))
    S, NKV = args.shards, args.kv
    shards = [eng.open_shard("segment%05d" % (rank * S + i), write_buffer_bytes=2 << 20) for i in range(S)]
    six_of = np.array([s.index for s in shards], dtype=np.uint32)
    seed = synth.SEED_DATA + rank

    # ---- load through the apply path, then fully compact -------------------------------------------
    t_load = time.perf_counter()
    CH = 1 << 20
    for lo in range(0, NKV, CH):
        idx = np.arange(lo, min(NKV, lo + CH), dtype=np.uint64)
        sh = (idx % np.uint64(S)).astype(np.int64)
        b = synth.single_put_batches(synth.keys16(seed, idx), synth.values(seed, sh, idx, 0), 1000 + idx)
        off = (np.arange(idx.size + 1, dtype=np.uint64) * np.uint64(b.shape[1]))
        st = eng.apply_packed(six_of[sh], b.reshape(-1), off, 1000 + idx)
        assert not st.any(), "load failed"
    eng.compact_all()
    t_load = time.perf_counter() - t_load
    assert sum(s.latest_seq() for s in shards) == NKV

    # ---- MultiGet: device-resident queries -----------------------------------------------------------
    Q = args.mg_batches * 4096
    n_sets = W + K
    rng = np.random.default_rng(synth.SEED_QUERY + rank)
    with torch.cuda.stream(stream):
        q_idx = [rng.integers(0, NKV, size=Q, dtype=np.uint64) for _ in range(n_sets)]
        d_keys = [torch.from_numpy(synth.keys16(seed, qi).reshape(-1)).cuda() for qi in q_idx]
        d_six = [torch.from_numpy(six_of[(qi % np.uint64(S)).astype(np.int64)].astype(np.int32)).cuda() for qi in q_idx]
        d_vals = torch.empty(Q * 64, dtype=torch.uint8, device="cuda")
        d_vlen = torch.empty(Q, dtype=torch.int32, device="cuda")
        d_st = torch.empty(Q, dtype=torch.int32, device="cuda")
    sp = C.c_void_p(stream.cuda_stream)

    def mg(i):
        rc = lib.rsp_multi_get_device(eng.h, Q, d_six[i].data_ptr(), d_keys[i].data_ptr(), 16, d_vals.data_ptr(), 64,
                                      d_vlen.data_ptr(), d_st.data_ptr(), sp)
        assert rc == 0

    for i in range(W):
        mg(i)
    barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = eng.kernel_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record(stream)
    for k in range(K):
        ev[k][0].record(stream)
        mg(W + k)
        ev[k][1].record(stream)
    e_end.record(stream)
    barrier()
    mg_total_ms = max_over_ranks(e_start.elapsed_time(e_end))
    mg_kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    mg_launches = eng.kernel_launches() - launches0
    # parity at full size (size-independent property: every value is a pure function of its key)
    last = q_idx[W + K - 1] if K else q_idx[-1]
    assert int(d_st.count_nonzero().item()) == 0 and int((d_vlen != 64).count_nonzero().item()) == 0
    want = synth.values(seed, (last % np.uint64(S)).astype(np.int64), last, 0)
    got = d_vals.cpu().numpy().reshape(Q, 64)
    assert np.array_equal(got, want), "MultiGet parity failed at full size"

    # ---- zipf(0.99) MultiGet (config-4 access pattern on one GPU): hot keys are served from L2 ---------
    zrng = np.random.default_rng(synth.SEED_ZIPF + rank)
    z_idx = [synth.scatter_ranks(synth.zipf_ranks(zrng, NKV, 0.99, Q), NKV) for _ in range(3)]
    with torch.cuda.stream(stream):
        z_keys = [torch.from_numpy(synth.keys16(seed, zi).reshape(-1)).cuda() for zi in z_idx]
        z_six = [torch.from_numpy(six_of[(zi % np.uint64(S)).astype(np.int64)].astype(np.int32)).cuda() for zi in z_idx]

    def mgz(i):
        j = i % 3
        assert lib.rsp_multi_get_device(eng.h, Q, z_six[j].data_ptr(), z_keys[j].data_ptr(), 16, d_vals.data_ptr(), 64,
                                        d_vlen.data_ptr(), d_st.data_ptr(), sp) == 0

    for i in range(W):
        mgz(i)
    barrier()
    z0, z1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    z0.record(stream)
    for k in range(K):
        mgz(W + k)
    z1.record(stream)
    barrier()
    zipf_ms = max_over_ranks(z0.elapsed_time(z1))
    jz = (W + K - 1) % 3
    assert int(d_st.count_nonzero().item()) == 0
    assert np.array_equal(d_vals.cpu().numpy().reshape(Q, 64),
                          synth.values(seed, (z_idx[jz] % np.uint64(S)).astype(np.int64), z_idx[jz], 0)), "zipf MultiGet parity"

    # ---- MultiGet end to end: host (pinned) buffers through rsp_multi_get_fixed ------------------------
    h_keys = [torch.from_numpy(synth.keys16(seed, qi).reshape(-1)).pin_memory() for qi in q_idx[:min(n_sets, 4)]]
    h_six = [torch.from_numpy(six_of[(qi % np.uint64(S)).astype(np.int64)].astype(np.int32)).pin_memory() for qi in q_idx[:min(n_sets, 4)]]
    h_vals = torch.empty(Q * 64, dtype=torch.uint8).pin_memory()
    h_vlen = torch.empty(Q, dtype=torch.int32).pin_memory()
    h_st = torch.empty(Q, dtype=torch.int32).pin_memory()

    def mg_e2e(i):
        j = i % len(h_keys)
        rc = lib.rsp_multi_get_fixed(eng.h, Q, h_six[j].data_ptr(), h_keys[j].data_ptr(), 16, h_vals.data_ptr(), 64,
                                     h_vlen.data_ptr(), h_st.data_ptr())
        assert rc == 0

    for i in range(W):
        mg_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        mg_e2e(W + k)
    barrier()
    mg_e2e_s = max_over_ranks(time.perf_counter() - t0)
    assert int(h_st.count_nonzero().item()) == 0
    jl = (W + K - 1) % len(h_keys)
    assert np.array_equal(h_vals.numpy().reshape(Q, 64),
                          synth.values(seed, (q_idx[jl] % np.uint64(S)).astype(np.int64), q_idx[jl], 0))

    # ---- range scans: Seek(random existing key) + 128 x Next, device-resident (config-4 shape on one GPU) --
    NSC, LSC = 16384, 128
    REC = 8 + 16 + 64
    sc_idx = [rng.integers(0, NKV, size=NSC, dtype=np.uint64) for _ in range(n_sets)]
    with torch.cuda.stream(stream):
        d_sk = [torch.from_numpy(synth.keys16(seed, qi).reshape(-1)).cuda() for qi in sc_idx]
        d_ss = [torch.from_numpy(six_of[(qi % np.uint64(S)).astype(np.int64)].astype(np.int32)).cuda() for qi in sc_idx]
        d_sout = torch.empty(NSC * LSC * REC, dtype=torch.uint8, device="cuda")
        d_snout = torch.empty(NSC, dtype=torch.int32, device="cuda")
        d_sst = torch.empty(NSC, dtype=torch.int32, device="cuda")

    def scan(i):
        rc = lib.rsp_multi_scan_device(eng.h, NSC, d_ss[i].data_ptr(), d_sk[i].data_ptr(), 16, LSC, d_sout.data_ptr(),
                                       LSC * REC, d_snout.data_ptr(), d_sst.data_ptr(), sp)
        assert rc == 0

    for i in range(W):
        scan(i)
    barrier()
    s_start, s_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_start.record(stream)
    for k in range(K):
        scan(W + k)
    s_end.record(stream)
    barrier()
    sc_total_ms = max_over_ranks(s_start.elapsed_time(s_end))
    assert int(d_sst.count_nonzero().item()) == 0
    # parity: scan i of the last set returns the next keys of its shard in order, each with its value
    last_sc = sc_idx[W + K - 1] if K else sc_idx[-1]
    n_out = d_snout.cpu().numpy()
    entries_last = int(n_out.sum())
    per = NKV // S
    for qn in (0, 1, NSC // 2, NSC - 1):
        i0 = int(last_sc[qn]); sh0 = i0 % S; j0 = i0 // S
        cnt = int(n_out[qn])
        n_in_shard = len(range(sh0, NKV, S))
        assert cnt == min(LSC, n_in_shard - j0), (cnt, j0, n_in_shard)
        want_idx = np.arange(j0, j0 + cnt, dtype=np.uint64) * np.uint64(S) + np.uint64(sh0)
        got = d_sout[qn * LSC * REC: qn * LSC * REC + cnt * REC].cpu().numpy().reshape(cnt, REC)
        assert np.array_equal(got[:, 8:24], synth.keys16(seed, want_idx)), "scan keys"
        assert np.array_equal(got[:, 24:], synth.values(seed, np.full(cnt, sh0), want_idx, 0)), "scan values"
        assert (got[:, 0] == 16).all() and (got[:, 4] == 64).all()

    # ---- apply: replicated single-Put updates to existing keys, pull-sized groups per shard -----------
    T = S * args.tick
    ticks = []
    upd_idx = []
    for stp in range(n_sets * 2):
        # `tick` updates per shard: shard-local ordinals uniform, global index = shard + ordinal * S
        per = NKV // S
        ordn = rng.integers(0, per, size=T, dtype=np.uint64)
        sh = np.repeat(np.arange(S, dtype=np.uint64), args.tick)
        idx = sh + ordn * np.uint64(S)
        b = synth.single_put_batches(synth.keys16(seed, idx), synth.values(seed, sh.astype(np.int64), idx, stp + 1), 5000 + idx)
        ticks.append((six_of[sh.astype(np.int64)], b, (np.arange(T + 1, dtype=np.uint64) * np.uint64(b.shape[1])), 5000 + idx))
        upd_idx.append(idx)
    staged = []
    for stp in range(n_sets):
        six, b, off, ts = ticks[stp]
        h = C.c_void_p()
        rc = lib.rsp_stage_build(eng.h, T, six.ctypes.data, b.ctypes.data, off.ctypes.data, ts.ctypes.data, C.byref(h))
        assert rc == 0
        staged.append(h)
    st_out = np.zeros(T, dtype=np.int32)

    def apply_dev(i):
        assert lib.rsp_reserve(eng.h, staged[i]) == 0
        assert lib.rsp_apply_staged_device(eng.h, staged[i], sp) == 0
        assert lib.rsp_apply_staged_finish(eng.h, staged[i], st_out.ctypes.data) == 0

    for i in range(W):
        apply_dev(i)
    barrier()
    launches1 = eng.kernel_launches()
    a_start, a_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # the K timed ticks go to the device back to back (each reserved against the ones still in flight); their
    # per-batch statuses are read back and folded into the host mirrors, in order, before the region ends
    bad_status = 0
    in_flight = []

    def fold():
        nonlocal bad_status
        for j in in_flight:
            assert lib.rsp_apply_staged_finish(eng.h, staged[j], st_out.ctypes.data) == 0
            bad_status += int(st_out.any())
        del in_flight[:]

    a_start.record(stream)
    for k in range(K):
        rc = lib.rsp_reserve(eng.h, staged[W + k])
        if rc == 11:  # Busy: a memtable is full while ticks are in flight — fold them, then it can be flushed
            fold()
            rc = lib.rsp_reserve(eng.h, staged[W + k])
        assert rc == 0
        assert lib.rsp_apply_staged_device(eng.h, staged[W + k], sp) == 0
        in_flight.append(W + k)
    fold()
    a_end.record(stream)
    assert bad_status == 0
    barrier()
    ap_total_ms = max_over_ranks(a_start.elapsed_time(a_end))
    ap_kernel_ms = eng.last_kernel_ms("apply")
    ap_launches = eng.kernel_launches() - launches1
    assert not st_out.any()

    # e2e apply: pinned host buffers through rsp_apply_many (H2D + kernels + D2H of statuses inside)
    pinned_ticks = []
    for i in range(n_sets):
        six, b, off, ts = ticks[n_sets + i]
        pinned_ticks.append(tuple(torch.from_numpy(np.ascontiguousarray(x).reshape(-1)).pin_memory()
                                  for x in (six, b, off, ts.astype(np.uint64))))
    h_ast = torch.zeros(T, dtype=torch.int32).pin_memory()

    def apply_e2e(i):
        six_t, b_t, off_t, ts_t = pinned_ticks[i]
        rc = lib.rsp_apply_many(eng.h, T, six_t.data_ptr(), b_t.data_ptr(), off_t.data_ptr(), ts_t.data_ptr(),
                                h_ast.data_ptr())
        assert rc == 0 and int(h_ast.count_nonzero().item()) == 0

    for i in range(W):
        apply_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for k in range(K):
        apply_e2e(W + k)
    barrier()
    ap_e2e_s = max_over_ranks(time.perf_counter() - t0)
    clk = clocks.stop()

    # ---- config 3: apply ticks and MultiGet launches CONCURRENTLY (two streams, lock-free memtable) --------
    mticks = []
    for stp in range(2 * n_sets, 2 * n_sets + K):
        per = NKV // S
        ordn = rng.integers(0, per, size=T, dtype=np.uint64)
        sh = np.repeat(np.arange(S, dtype=np.uint64), args.tick)
        idx = sh + ordn * np.uint64(S)
        b = synth.single_put_batches(synth.keys16(seed, idx), synth.values(seed, sh.astype(np.int64), idx, stp + 1), 5000 + idx)
        six = six_of[sh.astype(np.int64)]
        off = np.arange(T + 1, dtype=np.uint64) * np.uint64(b.shape[1])
        ts = 5000 + idx
        h = C.c_void_p()
        assert lib.rsp_stage_build(eng.h, T, six.ctypes.data, b.ctypes.data, off.ctypes.data, ts.ctypes.data, C.byref(h)) == 0
        mticks.append(h)
        upd_idx.append(idx)
    stream_b = torch.cuda.Stream(device=torch.device("cuda", local_rank))
    spb = C.c_void_p(stream_b.cuda_stream)
    barrier()
    mb0, mb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ma0, ma1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stream_b.wait_stream(stream)
    t_mixed = time.perf_counter()
    mb0.record(stream_b)
    for k in range(K):  # reads: K launches queued on stream B
        assert lib.rsp_multi_get_device(eng.h, Q, d_six[W + k].data_ptr(), d_keys[W + k].data_ptr(), 16, d_vals.data_ptr(), 64,
                                        d_vlen.data_ptr(), d_st.data_ptr(), spb) == 0
    mb1.record(stream_b)
    ma0.record(stream)
    for k in range(K):  # writes: K ticks on the engine stream while B runs
        assert lib.rsp_reserve(eng.h, mticks[k]) == 0
        assert lib.rsp_apply_staged_device(eng.h, mticks[k], sp) == 0
        assert lib.rsp_apply_staged_finish(eng.h, mticks[k], st_out.ctypes.data) == 0
        assert not st_out.any()
    ma1.record(stream)
    barrier()
    mixed_wall_s = max_over_ranks(time.perf_counter() - t_mixed)
    mixed_get_ms = max_over_ranks(mb0.elapsed_time(mb1))
    mixed_apply_ms = max_over_ranks(ma0.elapsed_time(ma1))
    # a lookup that raced a tick may see the value before or after it: every result is SOME version of its key
    assert int(d_st.count_nonzero().item()) == 0
    lastq = q_idx[W + K - 1] if K else q_idx[-1]
    samp = np.arange(0, Q, 64)
    got_m = d_vals.cpu().numpy().reshape(Q, 64)[samp]
    idx_m = lastq[samp]
    ok_m = np.zeros(samp.size, dtype=bool)
    for ver in range(0, 2 * n_sets + K + 1):
        todo = ~ok_m
        if not todo.any():
            break
        cand = synth.values(seed, (idx_m[todo] % np.uint64(S)).astype(np.int64), idx_m[todo], ver)
        ok_m[np.flatnonzero(todo)[(cand == got_m[todo]).all(axis=1)]] = True
    assert ok_m.all(), "mixed-phase MultiGet returned bytes that are no version of the key"
    for h in mticks:
        lib.rsp_stage_free(h)
    n_ticks_total = 2 * n_sets + K

    # parity after the update ticks: the newest version wins for every updated key (last writer)
    newest = {}
    for stp in range(n_ticks_total):
        for i in upd_idx[stp][::97]:
            newest[int(i)] = stp + 1
    chk = np.fromiter(newest.keys(), dtype=np.uint64)
    ver = np.fromiter(newest.values(), dtype=np.int64)
    # a key sampled at step s may have been rewritten later by an unsampled update: resolve exactly
    lastver = {}
    for stp in range(n_ticks_total):
        for i in upd_idx[stp]:
            lastver[int(i)] = stp + 1
    ver = np.array([lastver[int(i)] for i in chk], dtype=np.int64)
    res = eng.multi_get(six_of[(chk % np.uint64(S)).astype(np.int64)], [k.tobytes() for k in synth.keys16(seed, chk)], stride=64)
    for (rc, v), i, vr in zip(res, chk, ver):
        w = synth.values(seed, np.array([int(i) % S]), np.array([i], dtype=np.uint64), int(vr))[0].tobytes()
        assert rc == 0 and v == w, "apply parity failed"
    assert sum(s.latest_seq() for s in shards) == NKV + n_ticks_total * T

    # ---- config-2 variant: MultiGet while the newest version of many keys is still in the memtables -----
    # (every update tick above landed in a memtable: nothing has been flushed since the load).  Informational: a
    # failure here is reported in the line, it does not void the phases above.
    def mg_phase_checked():
        """W + K MultiGet launches of the uniform query sets, K of them timed; status, length and (full size) values
        checked against the last update tick of every queried key.  Local to the rank: no collectives inside."""
        for i in range(W):
            mg(i)
        torch.cuda.synchronize()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record(stream)
        for k in range(K):
            mg(W + k)
        m1.record(stream)
        torch.cuda.synchronize()
        ms = float(m0.elapsed_time(m1))
        assert int(d_st.count_nonzero().item()) == 0 and int((d_vlen != 64).count_nonzero().item()) == 0, "status / length"
        uk = np.fromiter(lastver.keys(), dtype=np.uint64, count=len(lastver))
        uv = np.fromiter(lastver.values(), dtype=np.int64, count=len(lastver))
        order = np.argsort(uk)
        uk, uv = uk[order], uv[order]
        lastq = q_idx[W + K - 1] if K else q_idx[-1]
        pos = np.minimum(np.searchsorted(uk, lastq), len(uk) - 1)
        qver = np.where(uk[pos] == lastq, uv[pos], 0)
        got = d_vals.cpu().numpy().reshape(Q, 64)
        qsh = (lastq % np.uint64(S)).astype(np.int64)
        for v in np.unique(qver):
            m = qver == v
            assert np.array_equal(got[m], synth.values(seed, qsh[m], lastq[m], int(v))), "value parity (version %d)" % v
        return ms

    mt_ms, mt_err, mt_entries = -1.0, None, 0
    try:
        mt_entries = int(sum(s.stats()["memtable_entries"] for s in shards))
        mt_ms = mg_phase_checked()
    except Exception as ex:  # noqa: BLE001
        mt_err = "%s: %s" % (type(ex).__name__, str(ex)[:160])

    # ---- the same once more after a flush WITHOUT compaction: every shard now has two sorted runs (the state between a
    # flush and the merge at level0_file_num_compaction_trigger; config 5 lives there).  Informational, non-fatal.
    r2_ms, r2_err, r2_runs = -1.0, None, 0
    try:
        assert eng.flush_all() == 0
        r2_runs = int(max(s.stats()["n_runs"] for s in shards))
        r2_ms = mg_phase_checked()
    except Exception as ex:  # noqa: BLE001
        r2_err = "%s: %s" % (type(ex).__name__, str(ex)[:160])

    # ---- numbers ---------------------------------------------------------------------------------------
    peak, peak_src = peaks()
    # every collective is issued by every rank, before any rank-0-only code
    tot_lookups = sum_over_ranks(Q * K)
    mt_ms_all = max_over_ranks(mt_ms)
    mt_failed = sum_over_ranks(1.0 if mt_err else 0.0)
    mt_entries_all = sum_over_ranks(float(mt_entries))
    r2_ms_all = max_over_ranks(r2_ms)
    r2_failed = sum_over_ranks(1.0 if r2_err else 0.0)
    tot_applies = sum_over_ranks(T * K)
    tot_scans = sum_over_ranks(NSC * K)
    tot_scan_entries = sum_over_ranks(entries_last * K)
    lookups_per_s = tot_lookups / (mg_total_ms * 1e-3)
    applies_per_s = tot_applies / (ap_total_ms * 1e-3)
    ach = A_GET * Q / (mg_kernel_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "multiget_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    # ---- the same path through the reference's seams: DbWrapper (pull loop) and ApplicationDB (MultiGet / Get) ------
    seams = None
    if not args.no_seams:
        for h in staged:
            lib.rsp_stage_free(h)
        staged = []
        barrier()
        try:
            seams = run_seams(local_rank, S, NKV, rank, world)
        except Exception as ex:  # noqa: BLE001
            seams = {"rc": -1, "error": "%s: %s" % (type(ex).__name__, str(ex)[:200])}
        ok = 1.0 if (seams.get("rc") == 0 and seams.get("parity_errors") == 0 and seams.get("status_errors") == 0) else 0.0
        seams_all_ok = sum_over_ranks(ok) == world
        seams_sum = {k: sum_over_ranks(float(seams.get(k, 0.0))) for k in (
            "load_applies_per_s", "mget_lookups_per_s", "get_per_s", "mixed_applies_per_s", "mixed_lookups_per_s")}
        seams_max = {k: max_over_ranks(float(seams.get(k, 0.0))) for k in (
            "resp_p50_ms", "resp_p99_ms", "mget_p50_ms", "mget_p99_ms", "get_p50_us", "get_p99_us", "mixed_resp_p50_ms",
            "mixed_resp_p99_ms")}
    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu:
        r = run_cpu("reference", ncores, S, args.cpu_kv, 30.0, args.cpu_get_secs)
        cpu = {"value": r["lookups_per_s"], "unit": "lookups/s", "cores": r["threads"], "kind": r["kind"],
               "sample": "%d KV applied over %d shards with default WriteOptions (WAL on), flush+compact, then %.0f s of MultiGet(4096) split per shard; %d threads" % (
                   r["applied"], r["shards"], r["get_s"], r["threads"]),
               "applies_per_s": r["applies_per_s"]}
    if rank == 0:
        line = {
            "metric": "multiget_lookups_per_s", "value": lookups_per_s, "unit": "lookups/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": (mg_total_ms + ap_total_ms) / max(K, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload, "l2": "inputs larger than L2: %.2f GB entry heap per GPU, fresh uniform keys every step" % (NKV * 96 / 1e9),
                       "timing": "CUDA events on the engine stream, max over ranks", "load_s": round(t_load, 2),
                       "flags": {k: os.environ[k] for k in sorted(os.environ) if k.startswith("RSP_")}},
            "applies": {"value": applies_per_s, "unit": "applies/s", "ms_per_tick": ap_total_ms / max(K, 1),
                        "kernel_ms_last_tick": ap_kernel_ms, "batches_per_tick": T,
                        "issue": "the K ticks are launched back to back; statuses read back and folded, in order, inside the timed region",
                        "hbm_frac_of_peak": (A_PUT * T / (ap_kernel_ms * 1e-3) / 1e9 / peak) if ap_kernel_ms and ap_kernel_ms > 0 else None,
                        "e2e": {"value": tot_applies / ap_e2e_s, "unit": "applies/s", "h2d_bytes_per_step": int(ticks[0][1].size + 10 * T), "d2h_bytes_per_step": 4 * T + 24 * S}},
            "roofline": {"kernel": "k_multi_get16", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_lookup": A_GET,
                         "lookups_per_launch": Q, "launch_ms": mg_kernel_ms},
            "memtable": ({"error": mt_err or "failed on another rank"} if mt_failed else
                         {"what": "config-2 variant: the same uniform MultiGet with the update ticks still in the memtables (entries there = %.0f %% of the key count; nothing flushed since the load); values checked at full size" % (100.0 * mt_entries_all / max(1, NKV * world)),
                          "lookups_per_s": tot_lookups / (mt_ms_all * 1e-3), "memtable_entries": int(mt_entries_all),
                          "hbm_frac_of_peak_algorithmic": A_GET * Q / (mt_ms_all * 1e-3 / max(K, 1)) / 1e9 / peak}),
            "two_runs": ({"error": r2_err or "failed on another rank"} if r2_failed else
                         {"what": "the same MultiGet after a flush without compaction: %d sorted runs per shard (lookups that miss the newest run go on to the older one); values checked at full size" % r2_runs,
                          "lookups_per_s": tot_lookups / (r2_ms_all * 1e-3),
                          "hbm_frac_of_peak_algorithmic": A_GET * Q / (r2_ms_all * 1e-3 / max(K, 1)) / 1e9 / peak}),
            "zipf": {"theta": 0.99, "lookups_per_s": tot_lookups / (zipf_ms * 1e-3), "hbm_frac_of_peak_algorithmic": A_GET * Q / (zipf_ms * 1e-3 / max(K, 1)) / 1e9 / peak},
            "mixed": {"what": "config 3: %d apply ticks on the engine stream concurrent with %d MultiGet launches on a second stream" % (K, K),
                      "lookups_per_s": tot_lookups / (mixed_get_ms * 1e-3), "applies_per_s": tot_applies / (mixed_apply_ms * 1e-3),
                      "wall_ms": mixed_wall_s * 1e3},
            "scans": {"value": tot_scans / (sc_total_ms * 1e-3), "unit": "scans/s", "entries_per_s": tot_scan_entries / (sc_total_ms * 1e-3),
                      "scan_len": LSC, "scans_per_launch": NSC,
                      "hbm_frac_of_peak": (entries_last * 168 + 16 * NSC) / (sc_total_ms * 1e-3 / max(K, 1)) / 1e9 / peak},
            "e2e": {"value": tot_lookups / mg_e2e_s, "unit": "lookups/s", "h2d_bytes_per_step": Q * 20, "d2h_bytes_per_step": Q * 72},
            "seams": (None if seams is None else {
                "what": "the same path through the reference's seams (librsp_host.so): %d follower pull loops (RocksDBReplicator -> DbWrapper::HandleReplicateResponses, 50 updates per response, synthetic leader behind the Transport interface) load %d KV; ApplicationDB::MultiGet(4096 keys of one shard) and ApplicationDB::Get from caller threads; host std::string / Slice buffers, H2D + D2H inside; values checked against the generator" % (S * world, NKV * world),
                "ok": bool(seams_all_ok), "threads_per_rank": seams.get("threads"), "error": seams.get("error"),
                "applies_per_s": seams_sum["load_applies_per_s"], "response_to_next_pull_ms": {"p50": seams_max["resp_p50_ms"], "p99": seams_max["resp_p99_ms"]},
                "multiget_lookups_per_s": seams_sum["mget_lookups_per_s"], "multiget_call_ms": {"p50": seams_max["mget_p50_ms"], "p99": seams_max["mget_p99_ms"]},
                "get_per_s": seams_sum["get_per_s"], "get_call_us": {"p50": seams_max["get_p50_us"], "p99": seams_max["get_p99_us"]},
                "mixed": {"what": "config 3: replicated updates flowing through the pull loops while MultiGet runs",
                          "applies_per_s": seams_sum["mixed_applies_per_s"], "lookups_per_s": seams_sum["mixed_lookups_per_s"],
                          "response_to_next_pull_ms": {"p50": seams_max["mixed_resp_p50_ms"], "p99": seams_max["mixed_resp_p99_ms"]}},
                "rank0": {k: seams.get(k) for k in ("load_s", "compact_s", "mget_calls", "applied_total", "parity_errors", "status_errors", "engine_launches")}}),
            "cpu_baseline": cpu,
            "gpu_launches": int(mg_launches + ap_launches),
            "clocks": clk,
        }
        print(json.dumps(line))
    for h in staged:
        lib.rsp_stage_free(h)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
