#!/bin/bash
O=gpurun_out/r02_c6; mkdir -p $O
for v in r01 07ad77b 4181f40; do timeout 300 python tools/bisect/mg_micro.py tools/bisect/librsp_b200_$v.so $v 2>&1 | tail -2 | head -1 | tee -a $O/ab.log; done
timeout 300 python tools/bisect/mg_micro.py rocksplicator_b200/librsp_b200.so head 2>&1 | tail -2 | head -1 | tee -a $O/ab.log
RSP_DBG_FASTALLOC=1 timeout 300 python tools/bisect/mg_micro.py rocksplicator_b200/librsp_b200.so head_fastalloc 2>&1 | tail -2 | head -1 | tee -a $O/ab.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_c6/bench_n1.json").read().strip().splitlines()[-1])
print(json.dumps({k: d.get(k) for k in ("value", "seams")}, indent=None)[:2500])
PY
