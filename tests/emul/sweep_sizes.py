"""TEST INFRASTRUCTURE — shard sizes around the compaction sort's tile and power-of-two boundaries (1 .. 12345 keys
in shuffled order, two flushes, overwrites and deletes left in the memtable, full compaction) on the CPU emulation."""
import os, sys, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from rocksplicator_b200 import engine, synth
from rocksplicator_b200.write_batch import WriteBatch
engine.SO_PATH = os.environ.get("RSP_TEST_EMUL_LIB", os.path.join(ROOT, "tests", "emul", "build", "librsp_b200_emul.so"))
e = engine.Engine(0, arena_bytes=1<<26)
for n in [1, 2, 31, 32, 33, 4095, 4096, 4097, 8191, 8192, 8193, 12345]:
    t0=time.time()
    s = e.open_shard("sw%d"%n, write_buffer_bytes=1<<22)
    rnd = random.Random(n)
    idx = list(range(n)); rnd.shuffle(idx)
    keys = synth.keys16(3, np.array(idx, dtype=np.uint64)); vals = synth.values(3, 0, np.array(idx, dtype=np.uint64), 0, 64)
    want = {}
    cuts = sorted(rnd.sample(range(1, max(2,n)), min(2, max(0,n-1)))) if n > 2 else []
    pos = 0
    for c0 in range(0, n, 1000):
        wb = WriteBatch()
        for i in range(c0, min(n, c0+1000)):
            wb.put(bytes(keys[i]), bytes(vals[i])); want[bytes(keys[i])] = bytes(vals[i])
        assert s.write(wb.data()) == 0
        if cuts and c0 <= cuts[0] < c0+1000: s.flush()
        if len(cuts)>1 and c0 <= cuts[1] < c0+1000: s.flush()
    # overwrite and delete a few, leave them in the memtable
    wb = WriteBatch(); ks = list(want)
    for k in ks[:: max(1, n//7)]:
        wb.put(k, b"y"*64); want[k] = b"y"*64
    for k in ks[1:: max(1, n//5)]:
        wb.delete(k); want.pop(k, None)
    assert s.write(wb.data()) == 0
    exp = sorted(want.items())
    assert s.scan() == exp, ("scan-before", n)
    s.compact()
    assert s.scan() == exp, ("scan-after", n)
    got = s.multi_get(ks, stride=64)
    assert got == [(0, want[k]) if k in want else (1, None) for k in ks], ("mg", n)
    st = s.stats(); assert st["n_runs"] == (1 if exp else 0) and st["run_entries"] == len(exp), (n, st)
    s.close()
    print("n", n, "ok", round(time.time()-t0,1), "s", flush=True)
print("SWEEP OK")
