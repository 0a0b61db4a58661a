#!/bin/bash
set -u
mkdir -p gpurun_out
for t in 128 256 512; do timeout 120 python tools/engine_bench.py --netbench --threads $t; done > gpurun_out/r2_netbench_n1_more_threads.json 2>&1
python - <<'PY'
import json
for line in open("gpurun_out/r2_netbench_n1_more_threads.json"):
    if line.startswith("{"):
        d = json.loads(line); print({k: d.get(k) for k in ("threads", "netbench", "mean_device_batch", "rc")})
    else:
        print(line[:200])
PY
timeout 300 python -m pytest tests/test_engine.py -m gpu -x -q 2>&1 | tail -3
