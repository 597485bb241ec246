#!/bin/bash
# r02 call 13: is the box's CPU time capped (cgroup quota)?  Get / steady probes with and without spinning; chunked update log
O=gpurun_out/r02_c13; mkdir -p $O
{ echo "== cpu.max"; cat /sys/fs/cgroup/cpu.max 2>&1; echo "== cpu.stat"; cat /sys/fs/cgroup/cpu.stat 2>&1; echo "== cpuset"; cat /sys/fs/cgroup/cpuset.cpus.effective 2>&1; nproc; grep -c processor /proc/cpuinfo; echo "== v1"; cat /sys/fs/cgroup/cpu/cpu.cfs_quota_us /sys/fs/cgroup/cpu/cpu.cfs_period_us 2>&1; lscpu | head -20; cat /sys/devices/system/cpu/cpuidle/current_driver 2>&1; cat /sys/module/intel_idle/parameters/max_cstate 2>&1; } > $O/cpu_limits.txt 2>&1
head -12 $O/cpu_limits.txt
thr() { grep -E "nr_throttled|throttled_usec|usage_usec" /sys/fs/cgroup/cpu.stat 2>/dev/null | tr '\n' ' '; echo; }
run() { name=$1; shift; echo "-- $name before: $(thr)"; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; echo "-- $name after:  $(thr)"; python - "$O/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-28s rc=%d get/s %.4g p50 %.0f us p99 %.0f us read_comb %s | steady %.4g/s trace %s apply_comb %s" % (
        sys.argv[2], d["rc"], d["get_per_s"], d["get_p50_us"], d["get_p99_us"], [round(x, 1) for x in d["read_comb"]],
        d["steady_applies_per_s"], [round(x) for x in d["trace_us"]], [round(x, 1) for x in d["apply_comb"]]))
except Exception as ex:
    print(sys.argv[2], "unreadable", ex)
PY
}
G="python tools/seam_probe.py --shards 256 --kv 1000000"
run get64       X=1 $G --get-threads 64
run get64_nospin RSP_WAIT_SPINS=0 RSP_DISPATCH_SPINS=0 $G --get-threads 64
run get256_nospin RSP_WAIT_SPINS=0 RSP_DISPATCH_SPINS=0 $G --get-threads 256
run get256_longspin RSP_WAIT_SPINS=20000 $G --get-threads 256
S="python tools/seam_probe.py --shards 1024 --kv 2000000 --get-threads 0 --steady 200"
run steady          X=1 $S
run steady_nospin   RSP_WAIT_SPINS=0 RSP_DISPATCH_SPINS=0 $S
run steady_c8       RSP_COMPLETION_THREADS=8 $S
run steady_upr500   X=1 $S --upr 500 --steady 40
