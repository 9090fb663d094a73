#!/bin/bash
mkdir -p gpurun_out
Q="--no-cpu-baseline --eager-rays 0 --warmup 3 --rays 1024 --steps 20"
for g in default 0 2 4; do
  if [ "$g" = default ]; then unset L4D_SCATTER_GRID; else export L4D_SCATTER_GRID=$g; fi
  timeout 600 python bench.py $Q > gpurun_out/ab_sc_$g.log 2> gpurun_out/ab_sc_$g.err
  echo "grid=$g $(grep -E 'timed:' gpurun_out/ab_sc_$g.err | tail -1) $(grep -oE 'k_bwd_scatter [0-9.]+ ms' gpurun_out/ab_sc_$g.err | tail -1)"
done
unset L4D_SCATTER_GRID
timeout 900 python -m pytest tests/test_raydrop_unet.py tests/test_rays_loss.py -m gpu -q --no-header 2>&1 | tail -3
