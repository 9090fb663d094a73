#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
for v in "" variant_s5 variant_s6; do
  if [ -n "$v" ]; then export L4D_LIB_PATH=$PWD/lidar4d_b200/csrc/$v.so; fi
  echo "== lib ${v:-default}"
  timeout 600 python scripts/perf_probe.py 16 4096 2>&1 | grep -E "probe.*split-tc train"
done
unset L4D_LIB_PATH
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -1 gpurun_out/bench.log | cut -c1-200; tail -2 gpurun_out/bench.err | cut -c1-400
