#!/bin/bash
# A/B of the round-2 dynamic-hash / time-plane work: per-launch contraction (L4D_CONTRACT bit 0 = dynamic tables, bit 1 = time
# rows), RED interleaving (build/lib_il0.so = off) and FFMA scans (build/lib_nofma.so = off); then the parity tests that cover them
mkdir -p gpurun_out
Q="--no-cpu-baseline --eager-rays 0 --steps 3 --warmup 3"
run() { timeout 300 python bench.py $Q > gpurun_out/abc_$1.log 2> gpurun_out/abc_$1.err; echo "$1: $(grep -E 'timed:' gpurun_out/abc_$1.err | tail -1 | cut -c18-) | $(grep -oE 'k_(contract|fwd_gather|bwd_scatter|bwd_scatter_static|fold_dynamic) [0-9.]+ ms' gpurun_out/abc_$1.err | tr '\n' ' ')"; }
L4D_CONTRACT=3 run all
L4D_CONTRACT=0 run c0
L4D_CONTRACT=1 run c1
L4D_CONTRACT=3 L4D_LIB_PATH=$PWD/build/lib_il0.so run il0
L4D_CONTRACT=3 L4D_LIB_PATH=$PWD/build/lib_nofma.so run nofma
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py tests/test_gpu_engine.py -m gpu -q --no-header -x --durations=6 \
  -k "render_forward_backward or reference_golden or full_size_vs_reference or launch_counter or staged_render" 2>&1 | tail -14 | cut -c1-200
