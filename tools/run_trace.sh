#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 500 python tools/ab_variants.py run 4 > gpurun_out/r2l_ab_variants.txt 2>&1; tail -4 gpurun_out/r2l_ab_variants.txt
LB2_LIB=$PWD/tools/_variants/carve.so timeout 120 python tools/trace_heads.py 2>&1 | tail -6
