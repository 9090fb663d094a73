#!/bin/bash
# tests + microbenchmark ceilings + default bench in one call
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-400 | tail -20
bash scripts/gpu_micro.sh > /dev/null 2>&1; cat gpurun_out/micro_gather.json | cut -c1-600
timeout 1200 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -1 gpurun_out/bench.log | cut -c1-400; grep -E "timed|e2e|eager" gpurun_out/bench.err | cut -c1-400
