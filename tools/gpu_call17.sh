#!/bin/bash
# r02 call 17 (2 GPUs): the router test on two devices, bench.py at N = 2 exactly as the driver launches it, then N = 1 without the CPU arm
O=gpurun_out/r02_c17; mkdir -p $O
nvidia-smi -L > $O/gpus.txt
T0=$(date +%s)
timeout 600 python -m pytest tests/test_router_gpu.py -x -q -m gpu > $O/pytest_router.log 2>&1; echo "router rc=$? ($(( $(date +%s) - T0 )) s)"; tail -2 $O/pytest_router.log
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench N=2 rc=$? ($(( $(date +%s) - T0 )) s)"; tail -2 $O/bench_n2.err
T0=$(date +%s)
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench N=1 rc=$? ($(( $(date +%s) - T0 )) s)"
python - <<'PY'
import json
for n in (1, 2):
    try:
        d = json.loads(open("gpurun_out/r02_c17/bench_n%d.json" % n).read().strip().splitlines()[-1])
        s = d.get("seams") or {}
        print("N=%d lookups/s %.4g (per GPU %.4g) frac %.3f | applies/s %.4g e2e %.4g | e2e lookups %.4g | seams steady %.4g mget %.4g get %.4g | c5 %s" % (
            n, d["value"], d["value"] / n, d["roofline"]["frac"], d["applies"]["value"], d["applies"]["e2e"]["value"], d["e2e"]["value"],
            (s.get("steady") or {}).get("applies_per_s", 0), s.get("multiget_lookups_per_s", 0), s.get("get_per_s", 0), (d.get("config5") or {}).get("applies_per_s")))
    except Exception as ex:
        print(n, "unreadable", ex)
PY
