#!/bin/bash
# scratch: whatever the current GPU call needs (every step under its own timeout)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2z_bench_reference_arm.json
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2z_bench_n1_20.json
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --flush-l2 --no-cpu 2>/dev/null | tail -1 > gpurun_out/r2z_bench_n1_20_flush.json
python - <<'PY'
import json
print(open("gpurun_out/r2z_bench_reference_arm.json").read()[:300])
for f in ("r2z_bench_n1_20", "r2z_bench_n1_20_flush"):
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, "value %.0f e2e %.0f 1thr %.0f ms/step %.4f trunk %.1f us frac %.3f exec %.3f traffic %s launches %s" % (d["value"], d["e2e"]["value"], d["e2e"]["one_thread"], d["ms_per_step"], d["roofline"]["launch_ms"] * 1e3, d["roofline"]["frac"], d["roofline"]["executed"]["frac_of_burst"], d["roofline"]["traffic"], d["gpu_launches"]))
PY
