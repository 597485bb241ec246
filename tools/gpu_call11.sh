#!/bin/bash
# r02 call 11: k_tick_fused v2 (conflict-free stage pitch, 64/128-thread shapes), batched flush install, wake fan-out 16,
# seams steady phase with the round-trip trace.  GPU suite, bench (no CPU arm), ncu of the large fused ticks.
O=gpurun_out/r02_c11; mkdir -p $O
T0=$(date +%s)
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$? ($(( $(date +%s) - T0 )) s)"; tail -3 $O/pytest_gpu.log
T0=$(date +%s)
timeout 900 python bench.py --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$? ($(( $(date +%s) - T0 )) s)"; tail -3 $O/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c11/bench_n1.json").read().strip().splitlines()[-1])
s = d.get("seams") or {}
print("lookups/s %.4g frac %.3f | applies/s %.4g (kernel %.1f us) big %.4g (kernel %.1f us, frac %.3f) e2e %.4g" % (
    d["value"], d["roofline"]["frac"], d["applies"]["value"], 1e3 * d["applies"]["kernel_ms_last_tick"],
    d["applies"]["large_ticks"]["applies_per_s"], 1e3 * d["applies"]["large_ticks"]["kernel_ms_per_tick"], d["applies"]["large_ticks"]["hbm_frac_of_peak"],
    d["applies"]["e2e"]["value"]))
print("config5", json.dumps(d.get("config5"))[:600])
print("seams applies %.4g steady %s" % (s.get("applies_per_s", 0), json.dumps(s.get("steady"))))
print("seams mget %.4g get %.4g %s get_comb %s" % (s.get("multiget_lookups_per_s", 0), s.get("get_per_s", 0), json.dumps(s.get("get_call_us")), json.dumps(s.get("get_combiner_rank0"))))
print("seams mixed", json.dumps(s.get("mixed")), "load500", s.get("applies_per_s_at_500_updates_per_response"))
PY
T0=$(date +%s)
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:k_tick_fusedILj128 -s 11 -c 2 -o $O/tick_fused_v2 \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-seams --c5-secs 0 > $O/ncu_bench.log 2>&1; echo "ncu rc=$? ($(( $(date +%s) - T0 )) s)"
ls -la $O
