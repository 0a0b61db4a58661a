#!/bin/bash
# engine-level sweep: how search threads / net-frequency knobs fill the device (python tools/engine_bench.py per line)
set -u
for cfg in "--threads 16" "--threads 64" "--threads 128" \
           "--threads 64 --max-outstanding 8 --extra '--eval_thresh 0 --mature_threshold 1'" \
           "--threads 128 --max-outstanding 8 --extra '--eval_thresh 0 --mature_threshold 1'" \
           "--threads 128 --max-outstanding 16 --extra '--eval_thresh 0 --mature_threshold 1 --extra_symmetry 1'"; do
  eval python tools/engine_bench.py --seconds 3 --moves 2 $cfg | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d.get(k) for k in ('threads', 'max_outstanding', 'extra', 'playouts_per_s_mean', 'nn_positions', 'device_batches', 'mean_device_batch', 'rc')})"
done
for t in 64 128 256; do
  python tools/engine_bench.py --netbench --threads $t | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print({k: d.get(k) for k in ('threads', 'netbench', 'mean_device_batch', 'rc')})"
done
