#!/bin/bash
# scratch: whatever the current GPU call needs (every step under its own timeout)
set -u
mkdir -p gpurun_out
timeout 500 python tools/ab_variants.py run 4 > gpurun_out/r2m_ab_variants.txt 2>&1; tail -4 gpurun_out/r2m_ab_variants.txt
