"""oracle/ — CHECKERS only (test infrastructure).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product package (rocksplicator_b200/) never does.
"""
from .okv import Okv, load_port, load_ref, ref_available, build  # noqa: F401
