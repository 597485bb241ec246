#!/bin/bash
# r02 call 19: the round's final evidence at N = 1 — GPU suite, smoke (plain and under compute-sanitizer memcheck), the
# reference arm, the full bench line, ncu --set full of k_multi_get16 (traffic), launch list
O=gpurun_out/r02_c19; mkdir -p $O
cat /sys/fs/cgroup/cpu.max > $O/cpu_max.txt 2>&1
T0=$(date +%s)
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - T0 )) s)"; tail -2 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
T0=$(date +%s)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_memcheck.log 2>&1; echo "memcheck rc=$? ($(( $(date +%s) - T0 )) s)"; tail -3 $O/smoke_memcheck.log
T0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "reference arm rc=$? ($(( $(date +%s) - T0 )) s)"
T0=$(date +%s)
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$? ($(( $(date +%s) - T0 )) s)"; tail -3 $O/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c19/bench_n1.json").read().strip().splitlines()[-1])
s = d.get("seams") or {}
print("lookups/s %.4g frac %.3f | applies/s %.4g (kernel %.1f us) big %.4g (kernel %.1f us, frac %.3f) e2e %.4g" % (
    d["value"], d["roofline"]["frac"], d["applies"]["value"], 1e3 * d["applies"]["kernel_ms_last_tick"],
    d["applies"]["large_ticks"]["applies_per_s"], 1e3 * d["applies"]["large_ticks"]["kernel_ms_per_tick"], d["applies"]["large_ticks"]["hbm_frac_of_peak"],
    d["applies"]["e2e"]["value"]))
print("memtable %.4g two_runs %.4g mixed %.4g + %.4g zipf %.4g scans %.4g e2e %.4g" % (d["memtable"]["lookups_per_s"], d["two_runs"]["lookups_per_s"], d["mixed"]["lookups_per_s"], d["mixed"]["applies_per_s"], d["zipf"]["lookups_per_s"], d["scans"]["value"], d["e2e"]["value"]))
print("config5", json.dumps(d.get("config5"))[:500])
print("seams applies %.4g steady %s" % (s.get("applies_per_s", 0), json.dumps((s.get("steady") or {}).get("applies_per_s"))))
print("seams mget %.4g get %.4g %s cpu %s" % (s.get("multiget_lookups_per_s", 0), s.get("get_per_s", 0), json.dumps(s.get("get_call_us")), json.dumps(s.get("cpu_seconds_rank0"))))
print("seams mixed", json.dumps(s.get("mixed")), "load500", s.get("applies_per_s_at_500_updates_per_response"))
print("cpu", json.dumps(d.get("cpu_baseline"))[:400], d.get("host"))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_multi_get16" -s 1 -c 1 -o $O/multiget16 \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-seams --c5-secs 0 > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-seams --c5-secs 0.5 > $O/launch_bench.log 2>&1; echo "launch list rc=$?"
nvidia-smi --query-gpu=name,memory.used,clocks.sm --format=csv
