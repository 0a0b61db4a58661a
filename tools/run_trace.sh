#!/bin/bash
# scratch: whatever the current GPU call needs (every step under its own timeout)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 120 python tools/trace_timeline.py 1 both > gpurun_out/r2_trace_timeline_b1.txt 2>&1; cat gpurun_out/r2_trace_timeline_b1.txt | head -20
timeout 120 python tools/trace_timeline.py 32 both > gpurun_out/r2_trace_timeline_b32.txt 2>&1; head -3 gpurun_out/r2_trace_timeline_b32.txt
