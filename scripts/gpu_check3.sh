#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
echo "== bench (default)"
timeout 1500 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -1 gpurun_out/bench.log; tail -12 gpurun_out/bench.err
echo "== bench reference arm"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
tail -1 gpurun_out/bench_ref.log; tail -5 gpurun_out/bench_ref.err
