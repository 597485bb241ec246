#!/bin/bash
# r02 call 22: the coalescing arena — GPU suite, the stretch point, a quick bench (no regression)
O=gpurun_out/r02_c22; mkdir -p $O
T0=$(date +%s)
timeout 600 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - T0 )) s)"; tail -2 $O/pytest_gpu.log
T0=$(date +%s)
timeout 900 python tools/stretch.py > $O/stretch.json 2> $O/stretch.err; echo "stretch rc=$? ($(( $(date +%s) - T0 )) s)"; grep -n "after\|out of memory" $O/stretch.err | cut -c1-400; cat $O/stretch.json | cut -c1-1200
T0=$(date +%s)
timeout 300 python bench.py --no-cpu --no-seams --c5-secs 2 > $O/bench_quick.json 2> $O/bench_quick.err; echo "bench rc=$? ($(( $(date +%s) - T0 )) s)"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c22/bench_quick.json").read().strip().splitlines()[-1])
print("lookups/s %.4g frac %.3f load_s %.2f | applies %.4g e2e %.4g | mixed %.4g | c5 %s" % (d["value"], d["roofline"]["frac"], d["config"]["load_s"], d["applies"]["value"], d["applies"]["e2e"]["value"], d["mixed"]["lookups_per_s"], json.dumps({k: v for k, v in d["config5"].items() if k != "what"})[:400]))
PY
