#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python scripts/perf_probe.py 8,16 4096 2>&1 | tee gpurun_out/probe.log | grep -E "probe.*split-tc"
timeout 1500 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -1 gpurun_out/bench.log | cut -c1-250; tail -4 gpurun_out/bench.err | cut -c1-300
