#!/bin/bash
# A/B: run sums of the plane-gradient sinks by redux.sync (fixed point) vs log-depth segmented shuffle scans
mkdir -p gpurun_out
Q="--no-cpu-baseline --eager-rays 0 --steps 3 --warmup 3"
run() { timeout 600 python bench.py $Q > /tmp/ab.log 2> /tmp/ab.err; echo "$1 $(grep -E 'timed:' /tmp/ab.err | tail -1 | cut -c18-) | $(grep -oE 'k_bwd_scatter[a-z_]* [0-9.]+ ms' /tmp/ab.err | tr '\n' ' ')"; }
run redux
L4D_LIB_PATH=$PWD/build/lib_scan.so run scan
run redux
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py -m gpu -q --no-header -x 2>&1 | tail -3 | cut -c1-300
