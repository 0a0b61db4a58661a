#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 ncu --set full --clock-control none --import-source on -k regex:trunk_kernel -s 4 -c 1 -o gpurun_out/trunk_r2 -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_ncu_full.log 2>&1
ls -la gpurun_out/trunk_r2.ncu-rep
timeout 400 python bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/r2k_bench_n1_20.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2k_bench_n1_20.json"))
print("value %.0f e2e %.0f 1thr %.0f ms/step %.4f trunk %.1f us frac %.3f exec %.3f traffic %s" % (d["value"], d["e2e"]["value"], d["e2e"]["one_thread"], d["ms_per_step"], d["roofline"]["launch_ms"] * 1e3, d["roofline"]["frac"], d["roofline"]["executed"]["frac_of_burst"], d["roofline"]["traffic"]))
PY
