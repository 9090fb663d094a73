#!/bin/bash
# third A/B: occupancy of the lighter gather (8 | 10 | 12 CTAs/SM) and scatter (4 | 5 CTAs/SM) kernels
mkdir -p gpurun_out
Q="--no-cpu-baseline --eager-rays 0 --steps 3 --warmup 3"
run() { timeout 300 python bench.py $Q > gpurun_out/abe_$1.log 2> gpurun_out/abe_$1.err; echo "$1: $(grep -E 'timed:' gpurun_out/abe_$1.err | tail -1 | cut -c18-) | $(grep -oE 'k_(contract|fwd_gather|bwd_scatter|bwd_scatter_static|fold_dynamic) [0-9.]+ ms' gpurun_out/abe_$1.err | tr '\n' ' ')"; }
for v in g8_sc4 g10_sc4 g12_sc4 g8_sc5; do L4D_LIB_PATH=$PWD/build/lib_$v.so run $v; done
