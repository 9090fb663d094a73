#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_micro.sh > gpurun_out/micro.log 2>&1; cat gpurun_out/micro_plane_stage.json; grep "chamfer\]" gpurun_out/chamfer_probe.log
Q="--no-cpu-baseline --eager-rays 0 --warmup 3"
timeout 600 python bench.py $Q --rays 1024 --steps 20 > gpurun_out/bench_1024.log 2> gpurun_out/bench_1024.err; echo "1024 rays $(grep -E 'timed:|e2e' gpurun_out/bench_1024.err | tail -2 | cut -c1-330)"
timeout 600 python bench.py $Q --rays 4096 --steps 10 > gpurun_out/bench_4096.log 2> gpurun_out/bench_4096.err; echo "4096 rays $(grep -E 'timed:|e2e' gpurun_out/bench_4096.err | tail -2 | cut -c1-330)"
