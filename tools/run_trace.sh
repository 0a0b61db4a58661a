#!/bin/bash
# new heads kernel + queue linger: tests, queue load with/without linger, bench, engine netbench; sustained-regime A/B of the
# L2 measures (position groups, discarding dead tiles). Everything under its own timeout.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for l in 0 1; do for t in 1 4 16 64 128 512; do timeout 60 tools/_variants/queue_bench engine/_build/weights_synth.lb2w $t 2 1 6 $l; done; done > gpurun_out/r2c_queue_linger.json 2>&1
cat gpurun_out/r2c_queue_linger.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu 2>/dev/null | tail -1 > gpurun_out/r2c_bench_n1_20.json
show() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], "value %.0f e2e %.0f 1thr %.0f ms/step %.4f trunk %.1f us frac %.3f exec %.3f parity %.2e %.2e" % (d["value"], d["e2e"]["value"], d["e2e"]["one_thread"], d["ms_per_step"], d["roofline"]["launch_ms"] * 1e3, d["roofline"]["frac"], d["roofline"]["executed"]["frac_of_burst"], d["config"]["parity"]["policy_max"], d["config"]["parity"]["value_max"]))
PY
}
show gpurun_out/r2c_bench_n1_20.json
for t in 16 64 128; do timeout 120 python tools/engine_bench.py --netbench --threads $t; done > gpurun_out/r2c_netbench.json 2>&1
timeout 120 python tools/engine_bench.py --seconds 3 --moves 2 --threads 16 >> gpurun_out/r2c_netbench.json 2>&1
timeout 300 python bench.py --steps 3000 --no-cpu 2>/dev/null | tail -1 > gpurun_out/r2c_sustained_default.json; show gpurun_out/r2c_sustained_default.json
timeout 300 python bench.py --steps 3000 --no-cpu --opt group_positions=128 2>/dev/null | tail -1 > gpurun_out/r2c_sustained_groups128.json; show gpurun_out/r2c_sustained_groups128.json
LB2_LIB=$PWD/tools/_variants/discard.so timeout 300 python bench.py --steps 3000 --no-cpu 2>/dev/null | tail -1 > gpurun_out/r2c_sustained_discard.json; show gpurun_out/r2c_sustained_discard.json
timeout 300 python bench.py --steps 3000 --no-cpu 2>/dev/null | tail -1 > gpurun_out/r2c_sustained_default2.json; show gpurun_out/r2c_sustained_default2.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/r2c_ncu_bench.log 2>&1
grep -E "heads_kernel|expand_planes|trunk_kernel" gpurun_out/r2c_launches.csv | tail -6
