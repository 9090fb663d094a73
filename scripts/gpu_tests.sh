#!/bin/bash
# all -m gpu tests, compact failure report back in gpurun_out/
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --no-header -rf "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-700 | tail -40
