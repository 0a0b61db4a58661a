#!/bin/bash
# multi-GPU checks on one box: N = $1 (every step under its own timeout)
N=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_device or self_test or batched_leaf" 2>&1 | tail -3
timeout 600 python bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/r2_bench_n$N.err | tail -1 > gpurun_out/r2_bench_n$N.json
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_n$N.json").read())
print("N=$N value %.0f e2e %.0f (1 thread %.0f) per-rank e2e %.0f..%.0f in-process %s" % (d["value"], d["e2e"]["value"], d["e2e"]["one_thread"] or 0,
      d["e2e"]["per_rank_min"], d["e2e"]["per_rank_max"], d["config"].get("in_process_e2e", {}).get("value")))
PY
for t in 128 512 1024; do timeout 60 tools/_variants/queue_bench engine/_build/weights_synth.lb2w $t 2 $N 6; done | tee gpurun_out/r2_queue_n$N.json
timeout 120 python tools/engine_bench.py --seconds 3 --moves 2 --threads 128 --gpus $N --max-outstanding 8 --extra='--eval_thresh 0 --mature_threshold 1' | cut -c1-900 | tee gpurun_out/r2_engine_n$N.json
timeout 120 python tools/engine_bench.py --netbench --threads 128 --gpus $N | cut -c1-700 | tee gpurun_out/r2_netbench_n$N.json
