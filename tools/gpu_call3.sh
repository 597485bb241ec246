#!/bin/bash
# r02 call 3: parity on the new code (combiners, multi-run fast path), HBM random-access ceilings, full bench line with
# the seams phase, launch list, one full ncu capture of the roofline kernel.
O=gpurun_out/r02_c3; mkdir -p $O
timeout 900 python -m pytest tests/ -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
for cfg in "960 80 8388608 96" "1280 80 8388608 128" "960 40 8388608 96" "960 20 8388608 96" "1280 20 8388608 128"; do
  echo "== randread $cfg" >> $O/randread.log; timeout 120 tools/randread_bench $cfg >> $O/randread.log 2>&1
done
grep -E "^==|dep=1 lanes=2 tpb=256" $O/randread.log
timeout 1200 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -5 $O/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-seams > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_multi_get16 -s 1 -c 1 -o $O/multiget16 python bench.py --steps 2 --warmup 1 --no-cpu --no-seams > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
nvidia-smi --query-gpu=index,clocks.sm,power.draw --format=csv,noheader
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c3/bench_n1.json").read().strip().splitlines()[-1])
print(json.dumps({k: d[k] for k in ("value", "roofline", "e2e", "applies", "two_runs", "memtable", "mixed", "seams", "cpu_baseline")}, indent=None)[:4000])
PY
