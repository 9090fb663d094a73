#!/bin/bash
# A/B two builds of the library with the probe (L=16, 4096 rays)
mkdir -p gpurun_out
if [ -z "$SKIP_PYTEST" ]; then timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log; fi
for lib in liblidar4d_b200.so "$@"; do
  echo "== $lib"
  L4D_LIB_PATH=$PWD/lidar4d_b200/csrc/$lib timeout 600 python scripts/perf_probe.py 16 4096 2>&1 | grep -E "probe.*split-tc train"
  L4D_LIB_PATH=$PWD/lidar4d_b200/csrc/$lib timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "kernels" | cut -c1-400
done
