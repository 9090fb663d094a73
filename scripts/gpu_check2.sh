#!/bin/bash
mkdir -p gpurun_out
echo "== perf probe"
timeout 900 python scripts/perf_probe.py 8,16 1024,4096 2>&1 | tee gpurun_out/probe.log | grep probe
echo "== pytest gpu (all, no -x)"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
echo "== ncu launch list + full capture (small)"
BENCH_EXTRA="--levels 8" timeout 1800 bash scripts/gpu_profile.sh
