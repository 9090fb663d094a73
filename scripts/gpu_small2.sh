#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_raydrop_unet.py tests/test_rays_loss.py tests/test_gpu_engine.py tests/test_trainer_dropin.py -m gpu -q --no-header -rf 2>&1 | tail -6 | cut -c1-300
Q="--no-cpu-baseline --eager-rays 0 --warmup 3"
timeout 600 python bench.py $Q --rays 1024 --steps 30 > gpurun_out/bench_1024.log 2> gpurun_out/bench_1024.err; echo "1024 rays $(grep -E 'timed:|e2e' gpurun_out/bench_1024.err | tail -2 | cut -c1-330)"
timeout 600 python bench.py $Q --rays 4096 --steps 10 > gpurun_out/bench_4096.log 2> gpurun_out/bench_4096.err; echo "4096 rays $(grep -E 'timed:' gpurun_out/bench_4096.err | tail -1 | cut -c1-330)"
