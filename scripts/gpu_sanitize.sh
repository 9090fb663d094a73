#!/bin/bash
# compute-sanitizer memcheck over the small parity cases (ragged tiles, odd level counts, both pipelines, chamfer)
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 1200 compute-sanitizer --tool memcheck --leak-check no --error-exitcode 9 --print-limit 20 \
  python -m pytest tests/test_gpu_parity.py tests/test_chamfer.py tests/test_gpu_engine.py tests/test_rays_loss.py -m gpu -q -x \
  -k "tensor_core_path or (render_forward_backward and 200) or attribute or flow_forward or (chamfer_forward and 777) or (chamfer_forward and 513) or fused_adam or flat_arenas or lidar_rays_kernel or main_loss_kernel" \
  > gpurun_out/sanitize.log 2>&1
echo "sanitizer rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|misaligned" gpurun_out/sanitize.log | head -20
