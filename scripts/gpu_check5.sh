#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (tensor-core path first)"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tensor_core or tcgen05" > gpurun_out/pytest_tc.log 2>&1; echo "tc rc=$?"
tail -15 gpurun_out/pytest_tc.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
echo "== perf probe"
timeout 900 python scripts/perf_probe.py 8,16 4096 2>&1 | tee gpurun_out/probe.log | grep probe
