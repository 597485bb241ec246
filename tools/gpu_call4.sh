#!/bin/bash
# r02 call 4: parity on merge-path compaction + thread-local stats + 3-buffer stager; bench (full) twice to see
# box-to-box / run-to-run variance of the lookup kernel; launch list of the compaction kernels.
O=gpurun_out/r02_c4; mkdir -p $O
timeout 900 python -m pytest tests/ -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 1200 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -5 $O/bench_n1.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-seams > $O/bench_n1_b.json 2> $O/bench_n1_b.err; echo "bench b rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-seams > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python - <<'PY'
import json
for f in ("bench_n1.json", "bench_n1_b.json"):
    d = json.loads(open("gpurun_out/r02_c4/" + f).read().strip().splitlines()[-1])
    print(f, json.dumps({k: d.get(k) for k in ("value", "roofline", "e2e", "applies", "two_runs", "memtable", "mixed", "seams", "zipf", "scans")}, indent=None)[:3500])
PY
