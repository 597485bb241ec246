#!/bin/bash
# r02 call 12: where do the seams lose their time?  Get cycle latency (1 caller), Get throughput vs callers and spin
# length, the pull loops' steady phase with the engine-side completion diagnostics.
O=gpurun_out/r02_c12; mkdir -p $O
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; python - "$O/$name.json" "$name" <<'PY'
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
run get1        X=1 $G --get-threads 1
run get16       X=1 $G --get-threads 16
run get64       X=1 $G --get-threads 64
run get256      X=1 $G --get-threads 256
run get256_s100 RSP_WAIT_SPINS=100 $G --get-threads 256
run get256_s0   RSP_WAIT_SPINS=0 $G --get-threads 256
run get1024_s0  RSP_WAIT_SPINS=0 $G --get-threads 1024
S="python tools/seam_probe.py --shards 1024 --kv 2000000 --get-threads 0 --steady 200"
run steady          X=1 $S
run steady_c64      RSP_COMPLETION_THREADS=64 $S
run steady_ex64     X=1 $S --executor 64
run steady_upr500   X=1 $S --upr 500 --steady 40
