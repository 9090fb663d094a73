#!/bin/bash
# full-size parity tests only (fast turn-around), full log back in gpurun_out/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -q --no-header -rA 2>&1 | tail -150 > gpurun_out/pytest_fullsize.log; echo "rc=${PIPESTATUS[0]}"
grep -E "passed|failed|error" gpurun_out/pytest_fullsize.log | tail -5
