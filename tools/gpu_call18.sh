#!/bin/bash
# r02 call 18 (2 GPUs): bench.py at N = 2 as the driver launches it (after the rank-0-block fix), Get probes with the lock-free stager
O=gpurun_out/r02_c18; mkdir -p $O
T0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench N=2 rc=$? ($(( $(date +%s) - T0 )) s)"; tail -2 $O/bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c18/bench_n2.json").read().strip().splitlines()[-1])
s = d.get("seams") or {}
print("N=2 lookups/s %.4g (per GPU %.4g) frac %.3f | applies/s %.4g big %s e2e %.4g | e2e lookups %.4g | seams steady %.4g mget %.4g get %.4g | c5 %s" % (
    d["value"], d["value"] / 2, d["roofline"]["frac"], d["applies"]["value"], d["applies"]["large_ticks"]["applies_per_s"], d["applies"]["e2e"]["value"], d["e2e"]["value"],
    (s.get("steady") or {}).get("applies_per_s", 0), s.get("multiget_lookups_per_s", 0), s.get("get_per_s", 0), (d.get("config5") or {}).get("applies_per_s")))
PY
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; python - "$O/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-16s rc=%d get/s %.4g p50 %.0f us p99 %.0f us read_comb %s | cpu_s %s" % (
        sys.argv[2], d["rc"], d["get_per_s"], d["get_p50_us"], d["get_p99_us"], [round(x, 1) for x in d["read_comb"]], [round(x, 2) for x in d["cpu_s"]]))
except Exception as ex:
    print(sys.argv[2], "unreadable", ex)
PY
}
G="python tools/seam_probe.py --shards 256 --kv 1000000"
run get64   X=1 $G --get-threads 64
run get256  X=1 $G --get-threads 256
