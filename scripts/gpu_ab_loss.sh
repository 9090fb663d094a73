#!/bin/bash
mkdir -p gpurun_out
Q="--no-cpu-baseline --eager-rays 0 --steps 3 --warmup 3"
for v in fused torch fused torch; do
  if [ $v = torch ]; then export BENCH_TORCH_LOSS=1; else unset BENCH_TORCH_LOSS; fi
  timeout 600 python bench.py $Q > gpurun_out/ab_loss_$v.log 2> gpurun_out/ab_loss_$v.err
  echo "$v $(grep -E 'timed:' gpurun_out/ab_loss_$v.err | tail -1) $(grep -oE 'e2e [0-9.]+ ms' gpurun_out/ab_loss_$v.err | tail -1)"
done
