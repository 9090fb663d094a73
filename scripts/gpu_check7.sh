#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "selftest" > gpurun_out/pytest_tc.log 2>&1; echo "selftests rc=$?"
tail -12 gpurun_out/pytest_tc.log
timeout 1500 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -1 gpurun_out/bench.log | cut -c1-900; tail -6 gpurun_out/bench.err
timeout 900 python bench.py --levels 8 --no-cpu-baseline > gpurun_out/bench_L8.log 2> gpurun_out/bench_L8.err; echo "bench L8 rc=$?"
tail -1 gpurun_out/bench_L8.log | cut -c1-300
