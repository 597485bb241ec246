#!/bin/bash
O=gpurun_out/r02_c8c; mkdir -p $O
timeout 900 python -m pytest tests/ -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 1500 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -5 $O/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c8c/bench_n1.json").read().strip().splitlines()[-1])
print(json.dumps({k: d.get(k) for k in ("value", "scans", "config5", "seams")}, indent=None)[:4000])
PY
