#!/bin/bash
# r02 call 7: HEAD after the regression fix — parity, the full bench line (CPU arm at the same KV count, seams, miss
# variant, large ticks, config-5 phase), launch list and one full ncu capture of the roofline kernel.
O=gpurun_out/r02_c7; mkdir -p $O
timeout 900 python -m pytest tests/ -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 1500 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -5 $O/bench_n1.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-seams --c5-secs 0.5 > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_multi_get16 -s 1 -c 1 -o $O/multiget16 python bench.py --steps 2 --warmup 1 --no-cpu --no-seams --c5-secs 0 --big-tick 0 > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tick_fused -s 2 -c 2 -o $O/tick_fused python bench.py --steps 2 --warmup 1 --no-cpu --no-seams --c5-secs 0 > $O/ncu_full2.log 2>&1; echo "ncu fused rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c7/bench_n1.json").read().strip().splitlines()[-1])
print(json.dumps({k: d.get(k) for k in ("value", "roofline", "e2e", "applies", "miss10", "two_runs", "memtable", "mixed", "config5", "seams", "cpu_baseline")}, indent=None)[:6000])
PY
