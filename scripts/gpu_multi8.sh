#!/bin/bash
# 8-GPU box: NCCL gradient-equality test, weak and strong scaling lines
mkdir -p gpurun_out
N=${1:-8}
Q="--no-cpu-baseline --eager-rays 0 --warmup 3"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --no-header 2>&1 | tail -3
timeout 600 $TR bench.py --gpus $N $Q --steps 3 > gpurun_out/bench_n${N}_weak.log 2> gpurun_out/bench_n${N}_weak.err; echo "weak rc=$? $(grep -E 'timed:' gpurun_out/bench_n${N}_weak.err | tail -1)"
timeout 600 $TR bench.py --gpus $N $Q --steps 5 --scaling strong > gpurun_out/bench_n${N}_strong65536.log 2> gpurun_out/bench_n${N}_strong65536.err; echo "strong 65536 rc=$? $(grep -E 'timed:' gpurun_out/bench_n${N}_strong65536.err | tail -1)"
timeout 600 $TR bench.py --gpus $N $Q --steps 20 --scaling strong --rays 4096 > gpurun_out/bench_n${N}_strong4096.log 2> gpurun_out/bench_n${N}_strong4096.err; echo "strong 4096 rc=$? $(grep -E 'timed:' gpurun_out/bench_n${N}_strong4096.err | tail -1)"
timeout 600 python bench.py $Q --steps 20 --rays 4096 > gpurun_out/bench_n1_4096.log 2> gpurun_out/bench_n1_4096.err; echo "n1 4096 $(grep -E 'timed:' gpurun_out/bench_n1_4096.err | tail -1)"
for f in gpurun_out/bench_n${N}_*.err; do tail -2 $f | cut -c1-200; done
