#!/bin/bash
# second A/B: how the gather reads the contracted tables (L4D_CON_GATHER 1 | 2), gather / scatter occupancy, one RED slice per query
mkdir -p gpurun_out
Q="--no-cpu-baseline --eager-rays 0 --steps 3 --warmup 3"
run() { timeout 300 python bench.py $Q > gpurun_out/abd_$1.log 2> gpurun_out/abd_$1.err; echo "$1: $(grep -E 'timed:' gpurun_out/abd_$1.err | tail -1 | cut -c18-) | $(grep -oE 'k_(contract|fwd_gather|bwd_scatter|bwd_scatter_static|fold_dynamic) [0-9.]+ ms' gpurun_out/abd_$1.err | tr '\n' ' ')"; }
run g1
L4D_LIB_PATH=$PWD/build/lib_g2.so run g2
L4D_LIB_PATH=$PWD/build/lib_g1_8.so run g1_8
L4D_LIB_PATH=$PWD/build/lib_il1.so run il1
L4D_LIB_PATH=$PWD/build/lib_sc4.so run sc4
