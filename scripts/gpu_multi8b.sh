#!/bin/bash
mkdir -p gpurun_out
N=8
Q="--no-cpu-baseline --eager-rays 0 --warmup 3"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N $Q --steps 20 --scaling strong --rays 4096 > gpurun_out/bench_n8_strong4096.log 2> gpurun_out/bench_n8_strong4096.err; echo "strong 4096 rc=$? $(grep -E 'timed:' gpurun_out/bench_n8_strong4096.err | tail -1)"
timeout 600 python bench.py $Q --steps 20 --rays 4096 > gpurun_out/bench_n1_4096.log 2> gpurun_out/bench_n1_4096.err; echo "n1 4096 $(grep -E 'timed:' gpurun_out/bench_n1_4096.err | tail -1)"
timeout 600 $TR bench.py --gpus $N $Q --steps 3 > gpurun_out/bench_n8_weak.log 2> gpurun_out/bench_n8_weak.err; echo "weak rc=$? $(grep -E 'timed:' gpurun_out/bench_n8_weak.err | tail -1)"
