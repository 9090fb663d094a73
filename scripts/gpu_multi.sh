#!/bin/bash
# run with: gpurun --gpus 2 -- 'bash scripts/gpu_multi.sh 2'
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; echo "multi rc=$?"
tail -5 gpurun_out/pytest_multi.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29581 \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -1 gpurun_out/bench_n$N.log | cut -c1-300; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29582 \
    bench.py --impl reference --gpus $N --steps 1 --warmup 1 > gpurun_out/bench_ref_n$N.log 2> gpurun_out/bench_ref_n$N.err; echo "ref N=$N rc=$?"
tail -1 gpurun_out/bench_ref_n$N.log | cut -c1-200
