#!/bin/bash
# scratch: whatever the current GPU call needs (every step under its own timeout)
set -u
mkdir -p gpurun_out
timeout 300 python tools/modes_time.py 2>&1 | tail -6
