#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 120 python tools/trace_heads.py > gpurun_out/r2d_trace_heads.txt 2>&1; tail -12 gpurun_out/r2d_trace_heads.txt
timeout 120 python tools/trace_timeline.py 256 both > gpurun_out/r2d_timeline_direct.txt 2>&1; mv gpurun_out/trace_both_256.npy gpurun_out/r2d_trace_direct.npy
LB2_LIB=$PWD/tools/_variants/nodirect.so timeout 120 python tools/trace_timeline.py 256 both > gpurun_out/r2d_timeline_nodirect.txt 2>&1; mv gpurun_out/trace_both_256.npy gpurun_out/r2d_trace_nodirect.npy
head -8 gpurun_out/r2d_timeline_direct.txt; head -8 gpurun_out/r2d_timeline_nodirect.txt
for i in 1 2 3; do timeout 60 tools/_variants/queue_bench engine/_build/weights_synth.lb2w 128 2 1 6 0; timeout 60 tools/_variants/queue_bench engine/_build/weights_synth.lb2w 128 2 1 6 1; done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | tail -1 > gpurun_out/r2d_bench_n1_20.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2d_bench_n1_20.json"))
print("value %.0f e2e %.0f 1thr %.0f ms/step %.4f trunk %.1f us frac %.3f exec %.3f" % (d["value"], d["e2e"]["value"], d["e2e"]["one_thread"], d["ms_per_step"], d["roofline"]["launch_ms"] * 1e3, d["roofline"]["frac"], d["roofline"]["executed"]["frac_of_burst"]))
PY
timeout 300 python -m pytest tests -m gpu -x -q -k "not engine" 2>&1 | tail -3
