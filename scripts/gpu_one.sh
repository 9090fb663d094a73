#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest "$@" -m gpu -q --no-header -x 2>&1 | grep -E "^E|passed|failed" | cut -c1-400 | head -20
