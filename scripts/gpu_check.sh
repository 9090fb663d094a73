#!/bin/bash
# One gpurun call: smoke -> GPU parity tests -> memcheck on one small case -> short bench.
# Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke" | tee gpurun_out/smoke.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
echo "== pytest gpu"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
echo "== memcheck"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -k "test_render_forward_backward and 150" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/memcheck.log
tail -8 gpurun_out/memcheck.log
echo "== bench"
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -2 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
