#!/bin/bash
# fourth A/B: private per-CTA weight-gradient copies in k_bwd_dense_tc (L4D_PRIVATE_WGRAD=0: shared copy), warp-aggregated coarse
# levels in k_bwd_scatter_static (build/lib_sagg.so), then the tensor-core parity tests
mkdir -p gpurun_out
Q="--no-cpu-baseline --eager-rays 0 --steps 3 --warmup 3"
run() { timeout 300 python bench.py $Q > gpurun_out/abf_$1.log 2> gpurun_out/abf_$1.err; echo "$1: $(grep -E 'timed:' gpurun_out/abf_$1.err | tail -1 | cut -c18-) | $(grep -oE 'k_(bwd_dense_tc|fwd_gather|bwd_scatter|bwd_scatter_static) [0-9.]+ ms' gpurun_out/abf_$1.err | tr '\n' ' ')"; }
run priv
L4D_PRIVATE_WGRAD=0 run shared
L4D_LIB_PATH=$PWD/build/lib_sagg.so run sagg
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py -m gpu -q --no-header -x -k "tensor_core_path or (full_size_vs_reference and tc) or loss_scaled" 2>&1 | tail -3 | cut -c1-200
