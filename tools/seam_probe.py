"""Short runs of the seam bench (host/bench/seam_bench.cpp) with chosen thread counts, for tuning the combiners:
    python tools/seam_probe.py --shards 256 --kv 1000000 --get-threads 256 --get-secs 1 [--mget-threads 0] [--steady 0]
Environment knobs of the engine (RSP_WAIT_SPINS, RSP_COMPLETION_THREADS, ...) apply.  Prints one JSON line."""
import argparse, ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--shards", type=int, default=256)
ap.add_argument("--kv", type=int, default=1_000_000)
ap.add_argument("--executor", type=int, default=32)
ap.add_argument("--upr", type=int, default=50)
ap.add_argument("--mget-threads", type=int, default=0)
ap.add_argument("--mget-secs", type=float, default=1.0)
ap.add_argument("--get-threads", type=int, default=256)
ap.add_argument("--get-secs", type=float, default=1.0)
ap.add_argument("--steady", type=int, default=0)
ap.add_argument("--mixed", type=int, default=0)
ap.add_argument("--base", type=int, default=30000)
a = ap.parse_args()
from rocksplicator_b200 import build
build.build_host()
lib = C.CDLL(build.HOST_SO)
lib.rsp_seam_bench.restype = C.c_int
lib.rsp_seam_bench.argtypes = [C.POINTER(bench.SeamCfg), C.POINTER(bench.SeamResult)]
cfg = bench.SeamCfg(device=0, shards=a.shards, kv_total=a.kv, value_len=64, executor_threads=a.executor, updates_per_response=a.upr,
                    update_rounds=a.mixed, multiget_threads=a.mget_threads, multiget_batch=4096, multiget_secs=a.mget_secs,
                    get_threads=a.get_threads, get_secs=a.get_secs, seed=0x5EED0001, first_shard_id=a.base, steady_rounds=a.steady)
res = bench.SeamResult()
rc = lib.rsp_seam_bench(C.byref(cfg), C.byref(res))
out = {n: (list(getattr(res, n)) if n in ("trace_us", "apply_comb", "read_comb", "cpu_s") else getattr(res, n)) for n, _ in bench.SeamResult._fields_}
out["rc"] = rc
out["env"] = {k: v for k, v in os.environ.items() if k.startswith("RSP_")}
print(json.dumps(out))
