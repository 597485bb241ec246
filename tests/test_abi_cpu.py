"""CPU-only: the C-ABI library builds, loads and exports every symbol include/rsp_b200.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from rocksplicator_b200 import build, engine
    build.build()
    return engine.load_library()


def test_every_declared_symbol_is_exported(lib):
    from rocksplicator_b200 import engine
    hdr = open(os.path.join(ROOT, "include", "rsp_b200.h")).read()
    declared = set(re.findall(r"\b(rsp_[a-z0-9_]+)\s*\(", hdr)) - {"rsp_merge_fn"}
    assert declared, "no declarations parsed"
    missing_in_binding = declared - set(engine.EXPORTS)
    assert not missing_in_binding, missing_in_binding
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.rsp_version().startswith(b"rocksplicator_b200")


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rocksplicator_b200 import engine
    with pytest.raises(RuntimeError):
        engine.Engine(0)


def test_product_does_not_touch_oracle():
    pkg = os.path.join(ROOT, "rocksplicator_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".cc")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "okv_" not in txt and "import oracle" not in txt and "from oracle" not in txt, f
