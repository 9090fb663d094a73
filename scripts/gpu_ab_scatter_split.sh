#!/bin/bash
# A/B: single scatter kernel (L4D_SCATTER_SPLIT=0) vs time-planes | static-planes+dynamic-hash at different occupancies
mkdir -p gpurun_out
Q="--no-cpu-baseline --eager-rays 0 --steps 3 --warmup 3"
run() { timeout 600 python bench.py $Q > /tmp/ab.log 2> /tmp/ab.err; echo "$1 $(grep -E 'timed:' /tmp/ab.err | tail -1 | cut -c18-) | $(grep -oE 'k_bwd_scatter[a-z_]* [0-9.]+ ms' /tmp/ab.err | tr '\n' ' ')"; }
L4D_SCATTER_SPLIT=0 run single
run split_T1_S1_default_build
for v in T3_S4 T3_S5 T4_S4; do L4D_LIB_PATH=$PWD/build/lib_$v.so run split_$v; done
L4D_LIB_PATH=$PWD/build/lib_T3_S4.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py -m gpu -q --no-header -x -k "render_forward_backward or tensor_core or smooth" 2>&1 | tail -2
