#!/bin/bash
# One gpurun call that measures the experiments written without GPU time:
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/measure_experiments.sh'
# Parity first (shipped configuration, then with the experiments on), then the bench per configuration, then one ncu
# capture of the direct-run lookup kernel.  Everything lands in gpurun_out/exp/.  ~25 GPU-minutes.
set -u
mkdir -p gpurun_out/exp
run() { echo "=== $*"; "$@"; echo "rc=$?"; }
{
  run timeout 600 python -m pytest tests -m gpu -x -q
  RSP_DIRECT_RUNS=1 RSP_DECODE_THREAD=1 run timeout 600 python -m pytest tests -m gpu -x -q
  RSP_MG_PREFETCH=4096 RSP_MG_MULTIRUN=1 run timeout 600 python -m pytest tests -m gpu -x -q
  RSP_FUSE_DECODE=1 run timeout 600 python -m pytest tests -m gpu -x -q
} > gpurun_out/exp/parity.log 2>&1
tail -5 gpurun_out/exp/parity.log
timeout 300 python bench.py --no-cpu > gpurun_out/exp/bench_shipped.json 2> gpurun_out/exp/bench_shipped.err
RSP_DECODE_THREAD=1 timeout 300 python bench.py --no-cpu > gpurun_out/exp/bench_decode_thread.json 2> gpurun_out/exp/bench_decode_thread.err
RSP_FUSE_DECODE=1 timeout 300 python bench.py --no-cpu > gpurun_out/exp/bench_fuse_decode.json 2> gpurun_out/exp/bench_fuse_decode.err
for load in 0.5 0.25 0.75; do
  RSP_DIRECT_RUNS=1 RSP_DIRECT_LOAD=$load timeout 300 python bench.py --no-cpu > gpurun_out/exp/bench_direct_$load.json 2> gpurun_out/exp/bench_direct_$load.err
done
for dist in 131072 262144 1048576; do
  RSP_MG_PREFETCH=$dist timeout 300 python bench.py --no-cpu > gpurun_out/exp/bench_prefetch_$dist.json 2> gpurun_out/exp/bench_prefetch_$dist.err
done
RSP_DIRECT_RUNS=1 RSP_MG_PREFETCH=262144 timeout 300 python bench.py --no-cpu > gpurun_out/exp/bench_direct_prefetch.json 2> gpurun_out/exp/bench_direct_prefetch.err
RSP_DIRECT_RUNS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_multi_get16d -s 1 -c 1 \
  -o gpurun_out/exp/multiget16d python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/exp/ncu.log 2>&1
# the scan fast path with one load in flight per lane (the r01 form) against the default four
RSP_NVCC_EXTRA=-DRSP_SCAN_UNROLL=1 python -m rocksplicator_b200.build --force > gpurun_out/exp/rebuild_unroll1.log 2>&1
timeout 300 python bench.py --no-cpu > gpurun_out/exp/bench_scan_unroll1.json 2> gpurun_out/exp/bench_scan_unroll1.err
python -m rocksplicator_b200.build --force > gpurun_out/exp/rebuild_default.log 2>&1
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/exp/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-40s lookups/s %.3g frac %.3f | applies/s %.3g (tick %.3f ms, kernels %.3f ms) | scans/s %.3g | zipf %.3g" % (
            f.split("/")[-1], d["value"], d["roofline"]["frac"], d["applies"]["value"], d["applies"]["ms_per_tick"],
            d["applies"]["kernel_ms_last_tick"], d["scans"]["value"], d["zipf"]["lookups_per_s"]))
    except Exception as ex:
        print(f, "unreadable:", ex)
PY
