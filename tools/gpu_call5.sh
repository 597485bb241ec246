#!/bin/bash
O=gpurun_out/r02_c5; mkdir -p $O
nvidia-smi --query-gpu=index,name,pci.bus_id,memory.total,ecc.mode.current --format=csv > $O/gpu.txt 2>&1
for i in 1 2; do
  timeout 300 python tools/bisect/mg_micro.py tools/bisect/librsp_b200_r01.so r01 2>&1 | tail -3 | tee -a $O/ab.log
  timeout 300 python tools/bisect/mg_micro.py rocksplicator_b200/librsp_b200.so head 2>&1 | tail -3 | tee -a $O/ab.log
done
timeout 900 python -m pytest tests/ -x -q -m gpu > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c5/bench_n1.json").read().strip().splitlines()[-1])
print(json.dumps({k: d.get(k) for k in ("value", "applies", "two_runs", "memtable", "mixed", "seams")}, indent=None)[:3000])
PY
cat $O/gpu.txt
