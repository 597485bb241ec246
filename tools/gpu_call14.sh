#!/bin/bash
# r02 call 14: after the CPU-time changes (short spins, futex buffer waits, one completion task per batch, inline
# continuation, executor wake policy): Get vs callers, pull loops steady, then the full bench line (N = 1, with the CPU arm)
O=gpurun_out/r02_c14; mkdir -p $O
thr() { grep -E "nr_throttled|throttled_usec" /sys/fs/cgroup/cpu.stat 2>/dev/null | tr '\n' ' '; }
run() { name=$1; shift; b="$(thr)"; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; python - "$O/$name.json" "$name" "$b" "$(thr)" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-22s rc=%d get/s %.4g p50 %.0f us p99 %.0f us read_comb %s | steady %.4g/s trace %s apply_comb %s | thr %s -> %s" % (
        sys.argv[2], d["rc"], d["get_per_s"], d["get_p50_us"], d["get_p99_us"], [round(x, 1) for x in d["read_comb"]],
        d["steady_applies_per_s"], [round(x) for x in d["trace_us"]], [round(x, 1) for x in d["apply_comb"]], sys.argv[3], sys.argv[4]))
except Exception as ex:
    print(sys.argv[2], "unreadable", ex)
PY
}
G="python tools/seam_probe.py --shards 256 --kv 1000000"
run get16   X=1 $G --get-threads 16
run get64   X=1 $G --get-threads 64
run get256  X=1 $G --get-threads 256
run get1024 X=1 $G --get-threads 1024
S="python tools/seam_probe.py --shards 1024 --kv 2000000 --get-threads 0 --steady 200"
run steady_ex16     X=1 $S --executor 16
run steady_ex32     X=1 $S --executor 32
run steady_c2       RSP_COMPLETION_THREADS=2 $S --executor 16
run steady_upr500   X=1 $S --upr 500 --steady 40 --executor 16
T0=$(date +%s)
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$? ($(( $(date +%s) - T0 )) s)"; tail -3 $O/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c14/bench_n1.json").read().strip().splitlines()[-1])
s = d.get("seams") or {}
print("lookups/s %.4g frac %.3f | applies/s %.4g (kernel %.1f us) big %.4g (kernel %.1f us, frac %.3f) e2e %.4g" % (
    d["value"], d["roofline"]["frac"], d["applies"]["value"], 1e3 * d["applies"]["kernel_ms_last_tick"],
    d["applies"]["large_ticks"]["applies_per_s"], 1e3 * d["applies"]["large_ticks"]["kernel_ms_per_tick"], d["applies"]["large_ticks"]["hbm_frac_of_peak"],
    d["applies"]["e2e"]["value"]))
print("config5", json.dumps(d.get("config5"))[:700])
print("seams applies %.4g steady %s" % (s.get("applies_per_s", 0), json.dumps(s.get("steady"))))
print("seams mget %.4g get %.4g %s" % (s.get("multiget_lookups_per_s", 0), s.get("get_per_s", 0), json.dumps(s.get("get_call_us"))))
print("seams mixed", json.dumps(s.get("mixed")), "load500", s.get("applies_per_s_at_500_updates_per_response"))
print("cpu", json.dumps(d.get("cpu_baseline")))
PY
