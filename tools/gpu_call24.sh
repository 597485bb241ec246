#!/bin/bash
# r02 call 24: the default bench line at the round's final sources (N = 1, with the CPU arm)
O=gpurun_out/r02_c24; mkdir -p $O
cat /sys/fs/cgroup/cpu.max > $O/cpu_max.txt 2>&1
T0=$(date +%s)
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$? ($(( $(date +%s) - T0 )) s)"; tail -2 $O/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c24/bench_n1.json").read().strip().splitlines()[-1])
s = d.get("seams") or {}
print("lookups/s %.4g frac %.3f | applies/s %.4g (kernel %.1f us) big frac %.3f e2e %.4g | memtable %.4g mixed %.4g+%.4g | c5 %.4g" % (
    d["value"], d["roofline"]["frac"], d["applies"]["value"], 1e3 * d["applies"]["kernel_ms_last_tick"], d["applies"]["large_ticks"]["hbm_frac_of_peak"], d["applies"]["e2e"]["value"],
    d["memtable"]["lookups_per_s"], d["mixed"]["lookups_per_s"], d["mixed"]["applies_per_s"], d["config5"]["applies_per_s"]))
print("seams load %.4g steady %.4g mget %.4g get %.4g %s mixed %s load500 %.4g" % (s.get("applies_per_s", 0), (s.get("steady") or {}).get("applies_per_s", 0), s.get("multiget_lookups_per_s", 0), s.get("get_per_s", 0), json.dumps(s.get("get_call_us")), json.dumps({k: v for k, v in (s.get("mixed") or {}).items() if k != "what"}), s.get("applies_per_s_at_500_updates_per_response", 0)))
print("cpu", json.dumps(d.get("cpu_baseline"))[:300])
PY
