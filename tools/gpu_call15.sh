#!/bin/bash
# r02 call 15: memtable filter + CPU-time changes: GPU suite, Get / steady probes, bench (no CPU arm), ncu of k_tick_chunks
# and of the flush kernels, launch list
O=gpurun_out/r02_c15; mkdir -p $O
T0=$(date +%s)
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - T0 )) s)"; tail -3 $O/pytest_gpu.log
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; python - "$O/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-16s rc=%d get/s %.4g p50 %.0f us p99 %.0f us read_comb %s | steady %.4g/s trace %s apply_comb %s | cpu_s %s" % (
        sys.argv[2], d["rc"], d["get_per_s"], d["get_p50_us"], d["get_p99_us"], [round(x, 1) for x in d["read_comb"]],
        d["steady_applies_per_s"], [round(x) for x in d["trace_us"]], [round(x, 1) for x in d["apply_comb"]], [round(x, 2) for x in d["cpu_s"]]))
except Exception as ex:
    print(sys.argv[2], "unreadable", ex)
PY
}
G="python tools/seam_probe.py --shards 256 --kv 1000000"
run get64   X=1 $G --get-threads 64
run get256  X=1 $G --get-threads 256
S="python tools/seam_probe.py --shards 1024 --kv 2000000 --get-threads 0 --steady 200"
run steady     X=1 $S
run steady_c16 RSP_COMPLETION_THREADS=16 $S
T0=$(date +%s)
timeout 900 python bench.py --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$? ($(( $(date +%s) - T0 )) s)"; tail -3 $O/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c15/bench_n1.json").read().strip().splitlines()[-1])
s = d.get("seams") or {}
print("lookups/s %.4g frac %.3f | applies/s %.4g (kernel %.1f us) big %.4g (kernel %.1f us, frac %.3f) e2e %.4g" % (
    d["value"], d["roofline"]["frac"], d["applies"]["value"], 1e3 * d["applies"]["kernel_ms_last_tick"],
    d["applies"]["large_ticks"]["applies_per_s"], 1e3 * d["applies"]["large_ticks"]["kernel_ms_per_tick"], d["applies"]["large_ticks"]["hbm_frac_of_peak"],
    d["applies"]["e2e"]["value"]))
print("memtable %.4g two_runs %.4g mixed %.4g + %.4g zipf %.4g scans %.4g" % (d["memtable"]["lookups_per_s"], d["two_runs"]["lookups_per_s"], d["mixed"]["lookups_per_s"], d["mixed"]["applies_per_s"], d["zipf"]["lookups_per_s"], d["scans"]["value"]))
print("config5", json.dumps(d.get("config5"))[:700])
print("seams applies %.4g steady %s" % (s.get("applies_per_s", 0), json.dumps(s.get("steady"))))
print("seams mget %.4g get %.4g %s cpu %s" % (s.get("multiget_lookups_per_s", 0), s.get("get_per_s", 0), json.dumps(s.get("get_call_us")), json.dumps(s.get("cpu_seconds_rank0"))))
print("seams mixed", json.dumps(s.get("mixed")), "load500", s.get("applies_per_s_at_500_updates_per_response"))
PY
T0=$(date +%s)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tick_chunks|k_flush_sort|k_compact_write|k_compact_size" -s 10 -c 7 -o $O/apply_flush \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-seams --c5-secs 0.3 > $O/ncu_bench.log 2>&1; echo "ncu rc=$? ($(( $(date +%s) - T0 )) s)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-seams --c5-secs 0.5 > $O/launch_bench.log 2>&1; echo "launch list rc=$?"
ls -la $O | head -30
