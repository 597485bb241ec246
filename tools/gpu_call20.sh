#!/bin/bash
# r02 call 20: where do 78 ms go in the e2e apply phase of some runs (RSP_TRACE), GPU suite after the pending-list change,
# then the config-2 stretch point (1.1 G KV, > 100 GB resident)
O=gpurun_out/r02_c20; mkdir -p $O
T0=$(date +%s)
timeout 600 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - T0 )) s)"; tail -2 $O/pytest_gpu.log
for i in 1 2; do
  RSP_TRACE=1 timeout 600 python bench.py --no-cpu --no-seams --c5-secs 0 > $O/bench_trace$i.json 2> $O/bench_trace$i.err
  python - $O/bench_trace$i.json $O/bench_trace$i.err <<'PY'
import json, re, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("e2e applies %.4g" % d["applies"]["e2e"]["value"])
lines = [l for l in open(sys.argv[2]) if "rsp trace" in l]
slow = []
for l in lines:
    nums = [float(x) for x in re.findall(r"([0-9]+(?:\.[0-9]+)?) us", l)]
    if nums and max(nums) > 3000 and ("apply_many" in l or "stage" in l): slow.append(l.strip())
print(len(lines), "trace lines; slow apply lines:", len(slow))
for l in slow[:12]: print("  ", l[:300])
PY
done
T0=$(date +%s)
timeout 1500 python tools/stretch.py > $O/stretch.json 2> $O/stretch.err; echo "stretch rc=$? ($(( $(date +%s) - T0 )) s)"; tail -3 $O/stretch.err; cat $O/stretch.json | cut -c1-900
