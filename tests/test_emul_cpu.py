"""The engine's host logic and the kernels' LOGIC on a machine without a GPU: the engine's own sources compiled by g++
against tests/emul (a CPU emulation of the CUDA slice they use — threads as fibers, warp collectives, block barriers,
a malloc-backed runtime; TEST INFRASTRUCTURE, see tests/emul/include/cuda_runtime.h), and the `-m gpu` parity tests
re-run against that build in a subprocess.

What this is for: catching logic and addressing bugs before GPU time is spent (tests/emul/build_emul.py --asan runs
the same under AddressSanitizer).  What it is NOT: a product path, a fallback, or a parity claim — librsp_b200.so is
CUDA-only and the parity tests proper are the `-m gpu` runs on a B200.
"""
import importlib.util
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul():
    spec = importlib.util.spec_from_file_location("build_emul", os.path.join(ROOT, "tests", "emul", "build_emul.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def _pytest_under_emulation(lib, extra_env, files, timeout=1500):
    env = dict(os.environ)
    env.update(extra_env)
    env["RSP_TEST_EMUL_LIB"] = lib
    env.setdefault("RSP_TEST_EMUL_ARENA", str(16 << 20))  # 16 MiB slabs: the emulated cudaMalloc poisons them (0xCD)
    p = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "-x", "-q", "-p", "no:cacheprovider"] + files,
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    print(p.stdout[-3000:], p.stderr[-2000:])
    assert p.returncode == 0, p.stdout[-3000:]
    return p.stdout


PARITY = ["tests/test_parity_gpu.py", "tests/test_zz_ingest_gpu.py"]


def test_parity_suite_under_emulation(emul):
    out = _pytest_under_emulation(emul[0], {"RSP_DIRECT_RUNS": "0", "RSP_DECODE_THREAD": "0"}, PARITY)
    assert " passed" in out and "failed" not in out


def test_parity_suite_under_emulation_experiments_on(emul):
    """functional check of the experiments that wait for GPU time: RSP_DIRECT_RUNS=1 (hash-addressed
    run heaps, k_multi_get16d), RSP_MG_PREFETCH (grid-level L2 prefetch) and RSP_DECODE_THREAD=1 (a thread per batch)"""
    out = _pytest_under_emulation(emul[0], {"RSP_DIRECT_RUNS": "1", "RSP_DECODE_THREAD": "1", "RSP_MG_PREFETCH": "5", "RSP_MG_MULTIRUN": "1"}, PARITY)
    assert " passed" in out and "failed" not in out


def test_direct_runs_are_what_the_experiment_serves(emul):
    """with the experiment on, a compacted fixed-shape shard really is a RUN_DIRECT run served by the direct kernel
    (nothing deferred to the generic path), and ordered access still works through the restart array"""
    code = r'''
import os, sys
import numpy as np
from rocksplicator_b200 import engine, synth
from rocksplicator_b200.write_batch import WriteBatch
engine.SO_PATH = os.environ["RSP_TEST_EMUL_LIB"]
e = engine.Engine(0, arena_bytes=1 << 22)
s = e.open_shard("d")
n = 3000
keys = synth.keys16(7, np.arange(n, dtype=np.uint64))
vals = synth.values(7, 0, np.arange(n, dtype=np.uint64), 0, 64)
for c in range(0, n, 500):
    wb = WriteBatch()
    for i in range(c, c + 500):
        wb.put(bytes(keys[i]), bytes(vals[i]))
    assert s.write(wb.data()) == 0
s.compact()
st = s.stats()
assert st["n_runs"] == 1 and st["run_entries"] == n
assert st["run_bytes"] >= 2 * n * 96, st          # slots at load 0.5: the heap is the table
probe = [bytes(keys[i]) for i in range(0, n, 7)] + [b"\xff" * 16, b"\x00" * 16]
got = s.multi_get(probe, stride=64)
assert got[:-2] == [(0, bytes(vals[i])) for i in range(0, n, 7)] and got[-2:] == [(1, None), (1, None)]
buf = np.zeros(16, dtype=np.uint32)
assert e.lib.rsp_debug_last_pending(e.h, buf.ctypes.data, 16) == 0
want = sorted((bytes(keys[i]), bytes(vals[i])) for i in range(n))
assert s.scan(limit=64) == want[:64]
# batched range scans (Seek + 128 x Next, the bench's scan shape): the gather form of the streaming fast path
starts = [want[i][0] for i in (0, 17, 1500, n - 5)] + [b"\x00" * 16, b"\xff" * 16, want[40][0][:15] + b"\xff"]
res = e.multi_scan([s.index] * len(starts), starts, 128, 128 * 88)
import bisect
keys_sorted = [k for k, _ in want]
for k0, (st_, recs) in zip(starts, res):
    lo = bisect.bisect_left(keys_sorted, k0)
    assert st_ == 0 and recs == want[lo:lo + 128], (k0, len(recs))
assert e.multi_scan([s.index], [want[10][0]], 128, 5 * 88 + 8)[0] == (7, want[10:15])   # output buffer too small: Incomplete
it = s.iterator(); it.seek_to_last(); assert it.key() == want[-1][0]; it.prev(); assert it.key() == want[-2][0]; it.close()
# later writes land in the memtable above the direct run; a second compaction rebuilds it
wb = WriteBatch(); wb.put(bytes(keys[3]), b"x" * 64); wb.delete(bytes(keys[4])); assert s.write(wb.data()) == 0
assert s.multi_get([bytes(keys[3]), bytes(keys[4]), bytes(keys[5])], stride=64) == [(0, b"x" * 64), (1, None), (0, bytes(vals[5]))]
s.compact()
assert s.multi_get([bytes(keys[3]), bytes(keys[4]), bytes(keys[5])], stride=64) == [(0, b"x" * 64), (1, None), (0, bytes(vals[5]))]
assert s.stats()["run_entries"] == n - 1
print("direct-ok")
'''
    env = dict(os.environ)
    env.update({"RSP_DIRECT_RUNS": "1", "RSP_TEST_EMUL_LIB": emul[0], "PYTHONPATH": ROOT})
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    print(p.stdout[-2000:], p.stderr[-2000:])
    assert p.returncode == 0 and "direct-ok" in p.stdout


def test_engine_vs_oracle_fuzz_under_emulation(emul):
    """tests/emul/fuzz_engine_vs_port.py on a few dozen seeds (hundreds run in minutes from the command line)"""
    env = dict(os.environ)
    env.update({"RSP_TEST_EMUL_LIB": emul[0]})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "fuzz_engine_vs_port.py"), "3000", "3060"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    print(p.stdout[-2000:], p.stderr[-2000:])
    assert p.returncode == 0 and "done bad= 0" in p.stdout


@pytest.mark.parametrize("direct", ["0", "1"])
def test_compaction_size_boundaries_under_emulation(emul, direct):
    env = dict(os.environ)
    env.update({"RSP_TEST_EMUL_LIB": emul[0], "RSP_DIRECT_RUNS": direct})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "sweep_sizes.py")], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    print(p.stdout[-1500:], p.stderr[-1500:])
    assert p.returncode == 0 and "SWEEP OK" in p.stdout


def test_parity_suite_under_emulation_fused_decode(emul):
    """RSP_FUSE_DECODE=1 (decode inside the sequencing kernel; excludes RSP_DECODE_THREAD): functional check"""
    out = _pytest_under_emulation(emul[0], {"RSP_FUSE_DECODE": "1", "RSP_MG_PREFETCH": "64", "RSP_MG_MULTIRUN": "1"}, ["tests/test_parity_gpu.py"])
    assert " passed" in out and "failed" not in out


def test_engine_vs_oracle_corruption_fuzz_under_emulation(emul):
    """mutated WriteBatches: the decode kernel's error classes, texts, latch and all-or-nothing effect vs the oracle"""
    env = dict(os.environ)
    env.update({"RSP_TEST_EMUL_LIB": emul[0]})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "fuzz_corrupt_engine_vs_port.py"), "1000", "1030"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    print(p.stdout[-2000:], p.stderr[-2000:])
    assert p.returncode == 0 and "done bad= 0" in p.stdout


@pytest.mark.parametrize("flags", [{}, {"RSP_DIRECT_RUNS": "1", "RSP_MG_PREFETCH": "100", "RSP_FUSE_DECODE": "1"}],
                         ids=["shipped", "experiments"])
def test_bench_control_flow_under_emulation(emul, flags):
    """bench.py end to end at toy size (tests/emul/bench_dryrun.py): every phase, its own full-size parity assertions
    and the JSON contract keys — an edit to bench.py or an API drift shows up here, not on the GPU box"""
    import json
    env = dict(os.environ)
    env.update(flags)
    env["RSP_TEST_EMUL_LIB"] = emul[0]
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emul", "bench_dryrun.py")], cwd=ROOT, env=env,
                       capture_output=True, text=True, timeout=900)
    print(p.stdout[-1500:], p.stderr[-2500:])
    assert p.returncode == 0
    line = json.loads(p.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    seen = {k: v for k, v in line["config"]["flags"].items() if not k.startswith("RSP_TEST_")}
    assert line["metric"] == "multiget_lookups_per_s" and seen == flags
    assert set(line["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert "error" not in line["memtable"] and line["memtable"]["lookups_per_s"] > 0 and line["memtable"]["memtable_entries"] > 0
    assert "error" not in line["two_runs"] and line["two_runs"]["lookups_per_s"] > 0 and "2 sorted runs" in line["two_runs"]["what"]


def test_host_mirror_over_emulated_engine(emul):
    """tests/cpp/host_tests.cpp's GpuDB-backed cases (replication chain, follower == leader, counter_service config 1,
    ApplicationDBManager, SST export / ingest) against the emulated engine"""
    for args in (["gpu-only"], ["gpu", "gpu_export_and_ingest"]):
        p = subprocess.run([emul[1]] + args, capture_output=True, text=True, timeout=900)
        print(p.stdout[-3000:], p.stderr[-1500:])
        assert p.returncode == 0 and " 0 failures" in p.stdout, p.stdout[-3000:]
