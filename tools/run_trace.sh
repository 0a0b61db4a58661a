#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 500 python tools/ab_variants.py run 3 > gpurun_out/r2f_ab_variants.txt 2>&1; tail -5 gpurun_out/r2f_ab_variants.txt
LB2_LIB=$PWD/tools/_variants/scoutfence.so timeout 120 python tools/trace_timeline.py 256 both > gpurun_out/r2f_timeline_scoutfence.txt 2>&1; mv gpurun_out/trace_both_256.npy gpurun_out/r2f_trace_scoutfence.npy; head -9 gpurun_out/r2f_timeline_scoutfence.txt
LB2_LIB=$PWD/tools/_variants/scoutfence.so timeout 400 python -m pytest tests -m gpu -x -q -k "not engine" 2>&1 | tail -3
