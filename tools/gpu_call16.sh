#!/bin/bash
# r02 call 16: k_tick_chunks with the poll after the decode, A/B against the CTA-per-group kernel; ncu of it; steady probe
O=gpurun_out/r02_c16; mkdir -p $O
T0=$(date +%s)
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - T0 )) s)"; tail -3 $O/pytest_gpu.log
for v in 1 0; do
  RSP_TICK_CHUNKS=$v timeout 600 python bench.py --no-cpu --no-seams --c5-secs 0 --steps 5 > $O/bench_chunks$v.json 2> $O/bench_chunks$v.err
  python - $O/bench_chunks$v.json $v <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("RSP_TICK_CHUNKS=%s load_s %.2f | small tick kernel %.1f us | big tick kernel %.1f us frac %.3f | e2e applies %.4g | memtable %.4g mixed %.4g+%.4g" % (
    sys.argv[2], d["config"]["load_s"], 1e3 * d["applies"]["kernel_ms_last_tick"], 1e3 * d["applies"]["large_ticks"]["kernel_ms_per_tick"],
    d["applies"]["large_ticks"]["hbm_frac_of_peak"], d["applies"]["e2e"]["value"], d["memtable"]["lookups_per_s"], d["mixed"]["lookups_per_s"], d["mixed"]["applies_per_s"]))
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tick_chunks" -s 3 -c 2 -o $O/tick_chunks \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-seams --c5-secs 0 > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; python - "$O/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-16s rc=%d steady %.4g/s trace %s apply_comb %s | cpu_s %s" % (
        sys.argv[2], d["rc"], d["steady_applies_per_s"], [round(x) for x in d["trace_us"]], [round(x, 1) for x in d["apply_comb"]], [round(x, 2) for x in d["cpu_s"]]))
except Exception as ex:
    print(sys.argv[2], "unreadable", ex)
PY
}
S="python tools/seam_probe.py --shards 1024 --kv 2000000 --get-threads 0 --steady 200"
run steady     X=1 $S
run steady_ex16 X=1 $S --executor 16
ls $O
