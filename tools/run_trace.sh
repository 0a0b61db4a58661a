#!/bin/bash
set -u
mkdir -p gpurun_out
for rep in 1 2 3; do for l in 0 1; do
  timeout 120 python tools/engine_bench.py --seconds 3 --moves 2 --threads 16 --extra="--queue-linger $l" | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('linger', '$l', {k: d.get(k) for k in ('playouts_per_s_mean', 'mean_device_batch', 'nn_positions', 'rc')})"
done; done 2>&1 | tee gpurun_out/r2h_engine_linger.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "other_shape" 2>&1 | tail -3
