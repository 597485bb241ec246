#!/bin/bash
# r02 call 8 (1 GPU): parity, compute-sanitizer on smoke, full bench with background merges / fused ticks / single-Put fast path
O=gpurun_out/r02_c8; mkdir -p $O
timeout 900 python -m pytest tests/ -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 $O/sanitizer_racecheck.log
timeout 1500 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -5 $O/bench_n1.err
nvidia-smi --query-gpu=index,clocks.sm,power.draw --format=csv,noheader
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c8/bench_n1.json").read().strip().splitlines()[-1])
print(json.dumps({k: d.get(k) for k in ("value", "applies", "two_runs", "mixed", "config5", "seams")}, indent=None)[:5000])
PY
