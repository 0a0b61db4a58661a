#!/bin/bash
# scratch: whatever the current GPU call needs (every step under its own timeout)
set -u
mkdir -p gpurun_out
timeout 400 python tests/soak.py 300 2>&1 | tail -2 | tee gpurun_out/r2_soak_300.txt
