#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tensor_core" > gpurun_out/pytest_tc.log 2>&1; echo "tc rc=$?"
tail -25 gpurun_out/pytest_tc.log | cut -c1-400
timeout 900 python scripts/perf_probe.py 8,16 4096 2>&1 | tee gpurun_out/probe.log | grep -E "probe.*split"
