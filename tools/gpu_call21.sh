#!/bin/bash
# r02 call 21: the config-2 stretch point again (1.0 G KV: > 100 GB resident with its indexes) after the arena / batching changes
O=gpurun_out/r02_c21; mkdir -p $O
T0=$(date +%s)
timeout 1200 python tools/stretch.py > $O/stretch.json 2> $O/stretch.err; echo "stretch rc=$? ($(( $(date +%s) - T0 )) s)"; tail -4 $O/stretch.err; cat $O/stretch.json | cut -c1-1200
T0=$(date +%s)
echo skip
