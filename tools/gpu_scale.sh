#!/bin/bash
# bench.py at N GPUs exactly as the driver launches it (torchrun); usage: gpurun --gpus N -- bash tools/gpu_scale.sh N
N=$1; O=gpurun_out/r02_scale; mkdir -p $O
cat /sys/fs/cgroup/cpu.max > $O/cpu_max_n$N.txt 2>&1
T0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench N=$N rc=$? ($(( $(date +%s) - T0 )) s)"; tail -2 $O/bench_n$N.err
python - $N <<'PY'
import json, sys
n = int(sys.argv[1])
d = json.loads(open("gpurun_out/r02_scale/bench_n%d.json" % n).read().strip().splitlines()[-1])
s = d.get("seams") or {}
print("N=%d lookups/s %.4g (per GPU %.4g) frac %.3f | applies/s %.4g e2e %.4g | e2e lookups %.4g | seams steady %.4g mget %.4g get %.4g | c5 %s | host %s" % (
    n, d["value"], d["value"] / n, d["roofline"]["frac"], d["applies"]["value"], d["applies"]["e2e"]["value"], d["e2e"]["value"],
    (s.get("steady") or {}).get("applies_per_s", 0), s.get("multiget_lookups_per_s", 0), s.get("get_per_s", 0), (d.get("config5") or {}).get("applies_per_s"), d.get("host")))
PY
