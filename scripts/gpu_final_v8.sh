#!/bin/bash
# round-2 final validation of HEAD within the remaining GPU budget: gpu tests (6 xdist workers: the oracle side of the parity tests
# is CPU work), the default bench line, smoke(), then the ncu launch list of a short run
mkdir -p gpurun_out
timeout 330 python -m pytest tests -m gpu -q -n 6 --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 300 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -1 gpurun_out/bench.log | cut -c1-250; grep -E "kernels|regimes" gpurun_out/bench.err | cut -c1-520
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee gpurun_out/smoke.log | grep smoke
ARGS="--steps 1 --warmup 1 --rays 8192 --ray-batch 8192 --no-cpu-baseline --eager-rays 0"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py $ARGS > gpurun_out/ncu_launch_run.log 2>&1
echo "launch list rc=$?"
