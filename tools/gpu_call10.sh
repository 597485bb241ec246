#!/bin/bash
# r02 call 10: HEAD after the re-entry — the -m gpu suite, smoke, the full default bench (N = 1) with wall times
O=gpurun_out/r02_c10; mkdir -p $O
nvidia-smi -L > $O/gpus.txt; nproc >> $O/gpus.txt
T0=$(date +%s)
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - T0 )) s)"; tail -3 $O/pytest_gpu.log
T0=$(date +%s)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$? ($(( $(date +%s) - T0 )) s)"; tail -2 $O/smoke.log
T0=$(date +%s)
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$? ($(( $(date +%s) - T0 )) s)"; tail -3 $O/bench_n1.err
nvidia-smi --query-gpu=name,memory.used,clocks.sm --format=csv > $O/health.txt 2>&1; cat $O/health.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c10/bench_n1.json").read().strip().splitlines()[-1])
s = d.get("seams") or {}
print("lookups/s %.4g frac %.3f | applies/s %.4g e2e %.4g | e2e lookups %.4g | seams applies %.4g mget %.4g get %.4g" % (
    d["value"], d["roofline"]["frac"], d["applies"]["value"], d["applies"]["e2e"]["value"], d["e2e"]["value"],
    s.get("applies_per_s", 0), s.get("multiget_lookups_per_s", 0), s.get("get_per_s", 0)))
PY
