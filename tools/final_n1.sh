#!/bin/bash
# Round-2 measurement suite on ONE B200 (outputs under gpurun_out/, copied to profiles/ by hand). Every step under its own timeout.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2_bench_reference_arm.json
timeout 400 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2_bench_n1_20.err | tail -1 > gpurun_out/r2_bench_n1_20.json
timeout 300 python bench.py --no-cpu 2>/dev/null | tail -1 > gpurun_out/r2_bench_n1_500.json
timeout 300 python bench.py --no-cpu --steps 3000 2>/dev/null | tail -1 > gpurun_out/r2_bench_n1_3000.json
timeout 300 python bench.py --no-cpu --steps 20 --warmup 5 --value-precision 0 2>/dev/null | tail -1 > gpurun_out/r2_bench_n1_20_fp16.json
timeout 300 python bench.py --no-cpu --steps 20 --warmup 5 --policy-precision 1 2>/dev/null | tail -1 > gpurun_out/r2_bench_n1_20_lite.json
timeout 300 python bench.py --no-cpu --steps 20 --warmup 5 --precise 2>/dev/null | tail -1 > gpurun_out/r2_bench_n1_20_full.json
timeout 300 python bench.py --no-cpu --steps 20 --warmup 5 --no-graphs 2>/dev/null | tail -1 > gpurun_out/r2_bench_n1_20_nographs.json
timeout 300 python tests/parity_report.py gpurun_out/r2_parity.json --all > /dev/null 2>&1
timeout 600 python tools/batch_sweep.py > gpurun_out/r2_batch_sweep.log 2>&1; cp gpurun_out/batch_sweep.json gpurun_out/r2_batch_sweep.json
for t in 1 16 64 128 256 512 1024; do timeout 60 tools/_variants/queue_bench engine/_build/weights_synth.lb2w $t 2 1 6 1; done > gpurun_out/r2_queue_n1.json 2>&1
for t in 16 64 128; do timeout 60 tools/_variants/queue_bench engine/_build/weights_synth.lb2w $t 2 1 6 0; done > gpurun_out/r2_queue_n1_nolinger.json 2>&1
timeout 600 python bench.py --engine > gpurun_out/r2_engine_bench.json 2>gpurun_out/r2_engine_bench.err
for t in 16 64 128; do timeout 120 python tools/engine_bench.py --netbench --threads $t; done > gpurun_out/r2_netbench_n1.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/r2_ncu_bench.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:trunk_kernel -s 4 -c 1 -o gpurun_out/trunk_r2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_full.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:heads_kernel -s 4 -c 1 -o gpurun_out/heads_r2 -f python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 120 python tools/trace_items.py both 256 > gpurun_out/r2_trace_items.txt 2>&1
timeout 120 python tools/trace_heads.py > gpurun_out/r2_trace_heads.txt 2>&1
timeout 120 python tools/trace_timeline.py 256 both > gpurun_out/r2_trace_timeline.txt 2>&1; mv gpurun_out/trace_both_256.npy gpurun_out/r2_trace_raw.npy
for f in gpurun_out/r2_bench_n1_*.json; do python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read())
print(sys.argv[1], "value %.0f e2e %.0f (1 thread %.0f) trunk %.1f us frac %.3f (burst %.3f, executed %.3f) parity %.2e/%.2e" % (d["value"], d["e2e"]["value"], d["e2e"]["one_thread"] or 0,
      d["roofline"]["launch_ms"] * 1e3, d["roofline"]["frac"], d["roofline"]["frac_of_burst"], d["roofline"]["executed"]["frac_of_burst"], d["config"]["parity"]["policy_max"], d["config"]["parity"]["value_max"]))
PY
done
