"""TEST INFRASTRUCTURE — malformed WriteBatches (oracle/fuzz_corrupt_port_vs_ref.py's mutations) through the engine on
the CPU emulation against the oracle port: return code, error text, sequence number, latch, contents.
`python tests/emul/fuzz_corrupt_engine_vs_port.py FIRST LAST`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
from oracle import okv  # noqa: E402
from rocksplicator_b200 import engine  # noqa: E402

spec = importlib.util.spec_from_file_location("fz", os.path.join(ROOT, "oracle", "fuzz_corrupt_port_vs_ref.py"))
fz = importlib.util.module_from_spec(spec)
spec.loader.exec_module(fz)


class EngineDb:
    """the slice of okv.Okv's interface the fuzzer uses, over an engine shard"""
    n = 0

    def __init__(self, eng, mop):
        EngineDb.n += 1
        self.s = eng.open_shard("cz%d" % EngineDb.n, merge_op=mop)
        self.last_error = ""

    def apply(self, bt, ts):
        rc = self.s.apply(bt, ts)
        self.last_error = self.s.last_error if rc else ""
        return rc

    def latest_seq(self): return self.s.latest_seq()
    def scan(self): return self.s.scan()
    def close(self): self.s.close()


if __name__ == "__main__":
    engine.SO_PATH = os.environ.get("RSP_TEST_EMUL_LIB", os.path.join(ROOT, "tests", "emul", "build", "librsp_b200_emul.so"))
    eng = engine.Engine(0, arena_bytes=1 << 24)
    bad = fz.run(int(sys.argv[1]), int(sys.argv[2]), a_lib=None, b_lib=okv.load_port(), make_a=lambda mop: EngineDb(eng, mop))
    eng.close()
    print("done bad=", bad)
